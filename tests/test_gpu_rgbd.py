"""'Next' row 4, RGB-D half: the monocular and RGB-D Frame constructors (reference src/Frame.cc:226-419):
one extractor, UndistortKeyPoints, ComputeStereoFromRGBD (:1065-1086) or no depth at all, then the same frame grid and
projection search as a stereo frame."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from sbp_check import assert_search_matches
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC
DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)   # Examples/RGB-D/TUM1.yaml


def _depth_image(seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:E["height"], 0:E["width"]].astype(np.float32)
    d = (1.5 + 0.004 * xx + 0.5 * np.sin(yy / 37.0) + rng.uniform(0, 0.05, xx.shape)).astype(np.float32)
    d[rng.random(d.shape) < 0.08] = 0.0            # missing depth
    d[rng.random(d.shape) < 0.01] = -1.0
    return d


@pytest.mark.parametrize("sensor", [2, 1])
def test_mono_and_rgbd_frames(euroc_pair, sensor):
    L, R = euroc_pair
    K = np.array([E["fx"], E["fy"], E["cx"], E["cy"]], np.float32)
    mbf = np.float32(40.0)                                        # Camera.bf of the RGB-D examples
    ctx = ft.Context(E["width"], E["height"], nfeatures=1000, nlevels=8, cam1=list(K), bf=float(mbf))
    ctx.set_sensor(sensor)
    ctx.set_distortion(DIST)
    depth = _depth_image(7) if sensor == 2 else None
    for rep in range(2):                                          # second pass replays the re-captured graph
        ctx.extract_mono(L)
        ctx.depth_from_rgbd(depth)
    g = ctx.download(0, stereo=True)
    ex = oracle.Extractor(1000, 1.2, 8)
    mono, k, d = ex.extract(L)
    assert g["n"] == len(k) and g["mono_index"] == mono
    assert np.array_equal(ft.keypoints_as_array(g["kps"]), k) and np.array_equal(g["desc"], d)
    assert ctx.counts()["n_right"] == 0
    un = oracle.undistort_points(k[:, :2], K, DIST)
    assert np.array_equal(ctx.keypoints_undistorted(), un)
    ur, dp = oracle.stereo_from_rgbd(k[:, :2], un[:, 0], depth, mbf)
    assert np.array_equal(g["u_right"], ur) and np.array_equal(g["depth"], dp)
    if sensor == 2:
        assert (dp > 0).sum() > 0.8 * len(k) and (dp < 0).sum() > 10
    else:
        assert np.all(ur == -1) and np.all(dp == -1)
    # projection search on the frame
    b = oracle.image_bounds(E["width"], E["height"], K, DIST)
    kun = k.copy(); kun[:, :2] = un
    scale = ctx.scale_tables()["scale"]
    F = oracle.Frame(kun, d, scale, E["width"], E["height"], cam1=list(K) + [0, 0, 0, 0], mbf=float(mbf), u_right=ur, bounds=b)
    M = 6000
    mp = synth.mappoints(kun, d, scale, M, seed=81)
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                   mp["holder"], mp["holder_obs"])
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                mp["holder"], mp["holder_obs"])
    gi, gf = ctx.track(M)
    def prefix(k_):
        a = {key: v[:k_] for key, v in mp.items() if key not in ("holder", "holder_obs")}
        o = F.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        g_ = ctx.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        return g_[:3], o[:3]
    assert_search_matches((n_g, h_g, ho_g), (n_o, h_o, ho_o), gi, ti, prefix)
    assert n_o > 100
    # call-order errors
    with pytest.raises(RuntimeError, match="monocular / RGB-D"):
        ctx.stereo_match()
    with pytest.raises(RuntimeError):
        ctx.depth_from_rgbd(None if sensor == 2 else _depth_image(1))
    # back to a stereo rig: the same context extracts both eyes again
    ctx.set_sensor(0); ctx.set_distortion(None)
    l2, r2 = ctx.frame_construct(L, R)
    ref = ft.Context(E["width"], E["height"], nfeatures=1000, nlevels=8, cam1=list(K), bf=float(mbf))
    l3, r3 = ref.frame_construct(L, R)
    for key in ("kps", "desc", "u_right", "depth"):
        assert np.array_equal(l2[key], l3[key])
    assert np.array_equal(r2["kps"], r3["kps"])
    with pytest.raises(RuntimeError, match="stereo rig"):
        ctx.extract_mono(L)
    ctx.close(); ref.close()
