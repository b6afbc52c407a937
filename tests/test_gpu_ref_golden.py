"""GPU: the CUDA extractor against outputs of the REFERENCE's own ORBextractor::operator() (tests/golden/
ref_orbextractor.npz, produced by the reference's src/ORBextractor.cc compiled in the build container; see
tools/make_ref_extractor_golden.py). Bit-exact: keypoints, order, monoIndex and descriptors."""
import os
import sys

import numpy as np
import pytest

import fasttrack_b200 as ft

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_ref_extractor_golden import case_images  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ref_orbextractor.npz")


def test_cuda_extractor_matches_reference_golden():
    g = np.load(GOLD)
    for name, (img, nf, nl, sf, ini, mn, lap) in case_images().items():
        h, w = img.shape
        fisheye = lap != (0, 0)
        kw = dict(camera_type=1, cam1=[190.0, 190.0, w / 2, h / 2, 0, 0, 0, 0], lap_left=lap, lap_right=lap) if fisheye else \
            dict(cam1=[400.0, 400.0, w / 2, h / 2])
        ctx = ft.Context(w, h, nfeatures=nf, nlevels=nl, scale_factor=sf, ini_th=ini, min_th=mn, bf=40.0, **kw)
        ctx.extract_stereo(img, img)
        for eye in (0, 1):
            r = ctx.download(eye)
            assert r["mono_index"] == int(g[name + "_mono"]), name
            assert np.array_equal(ft.keypoints_as_array(r["kps"]), g[name + "_kps"]), name
            assert np.array_equal(r["desc"], g[name + "_desc"]), name
        ctx.close()


# ---- matcher side: outputs of the reference's own Frame / ORBmatcher functions (tests/golden/ref_frame.npz) ----
import oracle  # noqa: E402  (checker only: borderline counts for the last-frame search)
from fasttrack_b200 import synth  # noqa: E402
import make_ref_frame_golden as G  # noqa: E402

GOLD_FRAME = os.path.join(ROOT, "tests", "golden", "ref_frame.npz")
E = synth.EUROC


@pytest.fixture(scope="module")
def euroc_ctx():
    L, R = synth.StereoScene(seed=2).pair()
    ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(G.MBF))
    left, right = ctx.frame_construct(L, R)
    yield ctx, left, right
    ctx.close()


def test_cuda_stereo_matches_reference_golden(euroc_ctx):
    ctx, left, _ = euroc_ctx
    g = np.load(GOLD_FRAME)
    assert np.array_equal(left["u_right"], g["stereo_uRight"]) and np.array_equal(left["depth"], g["stereo_depth"])


def test_cuda_local_map_search_matches_reference_golden(euroc_ctx):
    ctx, left, _ = euroc_ctx
    g = np.load(GOLD_FRAME)
    kL = ft.keypoints_as_array(left["kps"])
    scale = ctx.scale_tables()["scale"]
    ctx.set_pose(np.eye(3), np.zeros(3))
    for M, th, seed in G.LOCAL_CASES:
        mp = synth.mappoints(kL, left["desc"], scale, M, seed=seed)
        n, h, ho, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"],
                                              mp["holder_obs"])
        p = "local_%d_" % M
        assert n == int(g[p + "n"]) and np.array_equal(h, g[p + "holder"]) and np.array_equal(ho, g[p + "holder_obs"])
        ti, tf = ctx.track(M)
        assert np.array_equal(ti, g[p + "track_i"])
        seen = ti[:, 0] > 0
        assert np.array_equal(tf[seen, :5], g[p + "track_f"][seen, :5])


def test_cuda_last_frame_search_matches_reference_golden(euroc_ctx):
    ctx, left, _ = euroc_ctx
    g = np.load(GOLD_FRAME)
    kL, dL = ft.keypoints_as_array(left["kps"]), left["desc"]
    scale = ctx.scale_tables()["scale"]
    N = len(kL)
    for ci, (tz, th, ori) in enumerate(G.LAST_CASES):
        Rcw, tcw, lf = G.last_frame_case(kL, dL, tz)
        F = oracle.Frame(kL, dL, scale, E["width"], E["height"], cam1=G.CAM, mbf=float(G.MBF), u_right=left["u_right"], Rcw=Rcw, tcw=tcw)
        ctx.set_pose(Rcw, tcw, F.Rwc, F.Ow)
        n, h, ho, _ = ctx.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], np.eye(3), np.zeros(3), th,
                                            np.full(N, -1, np.int32), np.zeros(N, np.uint8), b_mono=False, check_ori=ori)
        assert n == int(g["last_%d_n" % ci])
        assert np.array_equal(h, g["last_%d_holder" % ci]) and np.array_equal(ho, g["last_%d_holder_obs" % ci])


def test_cuda_search_by_bow_matches_reference_golden(euroc_ctx):
    ctx, left, _ = euroc_ctx
    g = np.load(GOLD_FRAME)
    kL, dL = ft.keypoints_as_array(left["kps"]), left["desc"]
    voc, kf_desc, kf_angle, kf_has = G.bow_case(dL, np.ascontiguousarray(kL[:, 3]), 5, 1100)
    vg = ft.Vocabulary.from_arrays(10, 3, 0, 0, *voc)
    ctx.compute_bow(vg, 2)
    kf_node = vg.transform(kf_desc, 2)["node"]
    for ci, (ratio, ori) in enumerate(((0.7, True), (0.75, False), (0.9, True))):
        n, m = ctx.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, ratio, ori)
        assert n == int(g["bow_%d_n" % ci]) and np.array_equal(m, g["bow_%d_match" % ci])
    vg.close()


def test_cuda_fisheye_stereo_matches_reference_golden():
    """ComputeStereoFishEyeMatches on the TUM-VI-shaped rig against the reference function's output: match tables exact,
    depth within the tolerance the un-vendored Eigen::JacobiSVD leaves"""
    g = np.load(GOLD_FRAME)
    T = synth.TUMVI
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, _, _ = synth.tumvi_extrinsics()
    ctx = ft.Context(T["width"], T["height"], nfeatures=1000, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=T["lap"],
                     lap_right=T["lap"], bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    r = ctx.download(0, stereo=True)
    assert np.array_equal(r["l2r"], g["fisheye_stereo_l2r"]) and np.array_equal(r["r2l"], g["fisheye_stereo_r2l"])
    assert np.allclose(r["depth"], g["fisheye_stereo_depth"], rtol=1e-4, atol=1e-5)
    ctx.close()
