"""'Next' row 2, second half: Frame::UndistortKeyPoints + ComputeImageBounds for a pinhole camera with
distortion coefficients (reference src/Frame.cc:771-835). The oracle's cv::undistortPoints restatement is pinned
against cv2 (tests/golden/cv2_undistort.npz); here the CUDA path must reproduce mvKeysUn, the image bounds, the frame
grid built on mvKeysUn and the projection search that reads them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from sbp_check import assert_search_matches
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC
DIST = {"euroc_mono": [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],       # Examples/Monocular/EuRoC.yaml
        "five": [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]}              # Examples/RGB-D/TUM1.yaml (k3 set)


@pytest.mark.parametrize("which", ["euroc_mono", "five"])
def test_undistorted_keypoints_grid_and_search(euroc_pair, which):
    L, R = euroc_pair
    K = np.array([E["fx"], E["fy"], E["cx"], E["cy"]], np.float32)
    dist = np.array(DIST[which], np.float32)
    mbf = np.float32(E["fx"] * E["baseline"])
    ctx = ft.Context(E["width"], E["height"], nfeatures=1200, nlevels=8, cam1=list(K), bf=float(mbf))
    ctx.set_distortion(dist)
    l, r = ctx.frame_construct(L, R)
    keys = ft.keypoints_as_array(l["kps"])
    un = ctx.keypoints_undistorted()
    exp = oracle.undistort_points(keys[:, :2], K, dist)
    assert np.array_equal(un.view(np.uint32), exp.view(np.uint32))
    assert np.abs(un - keys[:, :2]).max() > 1.0              # the model does move the keypoints
    b = oracle.image_bounds(E["width"], E["height"], K, dist)
    assert np.array_equal(ctx.image_bounds(), b)
    keys_un = keys.copy(); keys_un[:, :2] = exp
    scale = ctx.scale_tables()["scale"]
    F = oracle.Frame(keys_un, l["desc"], scale, E["width"], E["height"], cam1=list(K) + [0, 0, 0, 0], mbf=float(mbf),
                     u_right=l["u_right"], bounds=b)
    co, io = F.grid()
    cg, ig = ctx.grid()
    assert np.array_equal(co, cg) and np.array_equal(io, ig)
    # projection search over a map built around the undistorted keypoints
    M = 7000
    mp = synth.mappoints(keys_un, l["desc"], scale, M, seed=61)
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                   mp["holder"], mp["holder_obs"])
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                mp["holder"], mp["holder_obs"])
    gi, gf = ctx.track(M)
    clean = ti[:, 4] == 0
    assert np.array_equal(gi[clean, 0], ti[clean, 0]) and np.array_equal(gi[clean, 2], ti[clean, 2])
    def prefix(k):
        a = {key: v[:k] for key, v in mp.items() if key not in ("holder", "holder_obs")}
        o = F.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        g = ctx.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        return g[:3], o[:3]
    assert_search_matches((n_g, h_g, ho_g), (n_o, h_o, ho_o), gi, ti, prefix)
    assert n_o > 100
    # k1 == 0 switches it off (Frame.cc:773): mvKeysUn = mvKeys, bounds = the image
    ctx.set_distortion([0.0, 0.1, 0, 0])
    ctx.frame_construct(L, R)
    assert np.array_equal(ctx.keypoints_undistorted(), keys[:, :2])
    assert np.array_equal(ctx.image_bounds(), [0, E["width"], 0, E["height"]])
    ctx.close()


def test_distortion_rejected_on_fisheye_rig():
    T = synth.TUMVI
    Rlr, tlr, _, _ = synth.tumvi_extrinsics()
    ctx = ft.Context(T["width"], T["height"], nfeatures=500, camera_type=1, cam1=T["cam1"], cam2=T["cam2"],
                     lap_left=T["lap"], lap_right=T["lap"], bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    with pytest.raises(RuntimeError, match="KannalaBrandt8"):
        ctx.set_distortion([0.1, 0, 0, 0])
    ctx.close()
