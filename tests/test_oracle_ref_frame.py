"""The oracle's restatement of the matcher side against the REFERENCE's own functions (text of Frame::ComputeStereoMatches,
AssignFeaturesToGrid, GetFeaturesInArea, isInFrustum[Checks], MapPoint::PredictScale, the camera projections and
ORBmatcher::SearchByProjection (local map / last frame) / SearchByBoW, compiled at build time into
oracle/_ref/libft_ref_frame.so, see oracle/ref_extract_fns.py).

* golden: committed outputs of those functions (tests/golden/ref_frame.npz, tools/make_ref_frame_golden.py)
* live:   the compiled functions themselves on other seeds (skipped where oracle/_ref cannot be built)
Integer results (match tables, levels, in-view flags) are compared exactly; so are the float results, because the oracle
and the reference functions are compiled with the same unfused arithmetic on the same host libm."""
import os
import sys

import numpy as np
import pytest

import oracle
from fasttrack_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_ref_frame_golden as G  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ref_frame.npz")
E, T = synth.EUROC, synth.TUMVI


@pytest.fixture(scope="module")
def euroc():
    exL, exR, kL, dL, kR, dR = G.euroc_frame(lambda: oracle.Extractor())
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(G.MBF), float(G.MB))
    return dict(exL=exL, exR=exR, kL=kL, dL=dL, kR=kR, dR=dR, st=st)


def _direction(F, mb):
    tlc_z = float(F.Ow[2])          # Tlw = identity
    return 1 if tlc_z > mb else (-1 if -tlc_z > mb else 0)


def test_stereo_matches_reference_golden(euroc):
    g = np.load(GOLD)
    assert np.array_equal(euroc["st"]["uRight"], g["stereo_uRight"]) and np.array_equal(euroc["st"]["depth"], g["stereo_depth"])
    assert (g["stereo_depth"] > 0).sum() > 200


def test_local_map_search_matches_reference_golden(euroc):
    g = np.load(GOLD)
    for M, th, seed in G.LOCAL_CASES:
        mp = synth.mappoints(euroc["kL"], euroc["dL"], euroc["exL"].scale, M, seed=seed)
        F = oracle.Frame(euroc["kL"], euroc["dL"], euroc["exL"].scale, E["width"], E["height"], cam1=G.CAM, mbf=float(G.MBF),
                         u_right=euroc["st"]["uRight"])
        n, h, ho, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"],
                                                 mp["holder_obs"])
        p = "local_%d_" % M
        assert n == int(g[p + "n"]) and np.array_equal(h, g[p + "holder"]) and np.array_equal(ho, g[p + "holder_obs"])
        assert np.array_equal(ti[:, :4], g[p + "track_i"])               # mbTrackInView, levels
        seen = ti[:, 0] > 0
        assert np.array_equal(tf[seen, :5], g[p + "track_f"][seen, :5])   # mTrackProjX/Y/XR, mTrackDepth, mTrackViewCos


def test_last_frame_search_matches_reference_golden(euroc):
    g = np.load(GOLD)
    for ci, (tz, th, ori) in enumerate(G.LAST_CASES):
        Rcw, tcw, lf = G.last_frame_case(euroc["kL"], euroc["dL"], tz)
        F = oracle.Frame(euroc["kL"], euroc["dL"], euroc["exL"].scale, E["width"], E["height"], cam1=G.CAM, mbf=float(G.MBF),
                         u_right=euroc["st"]["uRight"], Rcw=Rcw, tcw=tcw)
        N = len(euroc["kL"])
        n, h, ho, bl = F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th, _direction(F, G.MB),
                                           np.full(N, -1, np.int32), np.zeros(N, np.uint8), ori)
        assert n == int(g["last_%d_n" % ci]) and n > 150
        assert np.array_equal(h, g["last_%d_holder" % ci]) and np.array_equal(ho, g["last_%d_holder_obs" % ci])


def test_fisheye_local_map_search_matches_reference_golden():
    g = np.load(GOLD)
    exL, kL, dL, kR, dR, fo, (Rlr, tlr, Rrl, trl) = G.fisheye_frame()
    keys = np.vstack([kL, kR]); desc = np.vstack([dL, dR])
    for all_obs in (True, False):
        mp, holder, hobs = G.fisheye_map(kL, dL, kR, exL.scale, 6000, 13, all_obs)
        F = oracle.Frame(keys, desc, exL.scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"], n_left=len(kL),
                         n_right=len(kR), l2r=fo["l2r"], r2l=fo["r2l"], Rrl=Rrl, trl=trl, tlr=tlr)
        n, h, ho, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
        p = "fisheye_%d_" % int(all_obs)
        assert n == int(g[p + "n"]) and n > 20
        assert np.array_equal(h, g[p + "holder"]) and np.array_equal(ho, g[p + "holder_obs"])
        assert np.array_equal(ti[:, :4], g[p + "track_i"])


def test_fisheye_stereo_matches_reference_golden():
    """ComputeStereoFishEyeMatches + TriangulateMatches: match tables exact; depth and 3-D points within the tolerance of the
    SVD stand-in (the reference's x/w is a float division, the oracle's a double one: one ulp)"""
    g = np.load(GOLD)
    exL, kL, dL, kR, dR, fo, _ = G.fisheye_frame()
    assert np.array_equal(fo["l2r"], g["fisheye_stereo_l2r"]) and np.array_equal(fo["r2l"], g["fisheye_stereo_r2l"])
    acc = fo["l2r"] >= 0
    assert acc.sum() > 100
    assert np.allclose(fo["depth"], g["fisheye_stereo_depth"], rtol=1e-6, atol=0)
    assert np.allclose(fo["p3d"][acc], g["fisheye_stereo_p3d"][acc], rtol=1e-5, atol=1e-6)
    if oracle.ref_frame_lib() is not None:   # live, with a partial lapping area (mono and stereo keypoints mixed)
        L, R = synth.fisheye_pair(seed=8)
        Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
        exL, exR = oracle.Extractor(800), oracle.Extractor(800)
        mL, kL, dL = exL.extract(L, lap=(150, 511)); mR, kR, dR = exR.extract(R, lap=(0, 360))
        a = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
        b = oracle.ref_fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
        assert 0 < mL < len(kL) and (a["l2r"] >= 0).sum() > 50
        assert np.array_equal(a["l2r"], b["l2r"]) and np.array_equal(a["r2l"], b["r2l"])
        assert np.allclose(a["depth"], b["depth"], rtol=1e-6, atol=0)


def test_search_by_bow_matches_reference_golden(euroc):
    g = np.load(GOLD)
    kL, dL = euroc["kL"], euroc["dL"]
    angle = np.ascontiguousarray(kL[:, 3])
    voc, kf_desc, kf_angle, kf_has = G.bow_case(dL, angle, 5, 1100)
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, *voc)
    f_node, kf_node = vo.transform(dL, 2)["node"], vo.transform(kf_desc, 2)["node"]
    for ci, (ratio, ori) in enumerate(((0.7, True), (0.75, False), (0.9, True))):
        n, m = oracle.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, dL, angle, f_node, -1, ratio, ori)
        assert n == int(g["bow_%d_n" % ci]) and n > 100 and np.array_equal(m, g["bow_%d_match" % ci])
    exL, fkL, fdL, fkR, fdR, fo, _ = G.fisheye_frame()
    desc = np.vstack([fdL, fdR]); f_angle = np.concatenate([fkL[:, 3], fkR[:, 3]]).astype(np.float32)
    voc, kf_desc, kf_angle, kf_has = G.bow_case(desc, f_angle, 9, 1500)
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, *voc)
    f_node, kf_node = vo.transform(desc, 2)["node"], vo.transform(kf_desc, 2)["node"]
    for ci, ori in enumerate((True, False)):
        n, m = oracle.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, desc, f_angle, f_node, len(fkL), 0.7, ori)
        assert n == int(g["bow_fisheye_%d_n" % ci]) and np.array_equal(m, g["bow_fisheye_%d_match" % ci])
        assert (m[len(fkL):] >= 0).any()


def test_undistortion_matches_reference_live():
    """Frame::UndistortKeyPoints / ComputeImageBounds (the reference's text over the cv2-pinned undistortPoints)"""
    if oracle.ref_frame_lib() is None:
        pytest.skip("oracle/_ref/libft_ref_frame.so not built and no reference tree here")
    rng = np.random.default_rng(3)
    K = np.array([E["fx"], E["fy"], E["cx"], E["cy"]], np.float32)
    for dist in ([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], [-0.28, 0.07, 0.0002, 1.8e-05, 0.01], [0.0, 0.1, 0, 0]):
        xy = np.stack([rng.uniform(0, E["width"], 3000), rng.uniform(0, E["height"], 3000)], 1).astype(np.float32)
        want = oracle.undistort_points(xy, K, np.array(dist, np.float32)) if dist[0] != 0 else xy   # Frame.cc:773
        got, bounds = oracle.ref_undistort(xy, E["width"], E["height"], K, dist)
        assert np.array_equal(got, want)
        assert np.array_equal(bounds, oracle.image_bounds(E["width"], E["height"], K, np.array(dist, np.float32)))


def test_reference_functions_live():
    """other seeds, sizes and thresholds than the golden, against the compiled reference functions themselves"""
    if oracle.ref_frame_lib() is None:
        pytest.skip("oracle/_ref/libft_ref_frame.so not built and no reference tree here")
    rng = np.random.default_rng(123)
    for trial in range(3):
        L, R = synth.StereoScene(seed=40 + trial).pair()
        nf = int(rng.integers(600, 1800))
        exL, exR = oracle.Extractor(nf), oracle.Extractor(nf)
        _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
        so = oracle.stereo(exL, exR, kL, dL, kR, dR, float(G.MBF), float(G.MB))
        sr = oracle.ref_stereo(exL, exR, kL, dL, kR, dR, float(G.MBF), float(G.MB))
        assert np.array_equal(so["uRight"], sr["uRight"]) and np.array_equal(so["depth"], sr["depth"])
        M, th = int(rng.integers(3000, 15000)), float(rng.choice([1.0, 2.0, 3.0, 10.0]))
        mp = synth.mappoints(kL, dL, exL.scale, M, seed=70 + trial)
        a = 0.01 * trial
        Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        tcw = np.array([0.01 * trial, 0.0, 0.02 * trial], np.float32)
        kw = dict(cam1=G.CAM, mbf=float(G.MBF), u_right=so["uRight"], Rcw=Rcw, tcw=tcw)
        Fo = oracle.Frame(kL, dL, exL.scale, E["width"], E["height"], **kw)
        Fr = oracle.RefFrame(kL, dL, exL.scale, E["width"], E["height"], **kw)
        co, io = Fo.grid(); cr, ir = Fr.grid()
        assert np.array_equal(co, cr) and np.array_equal(io, ir)
        b_far = bool(trial % 2)
        ro = Fo.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"], mp["holder_obs"],
                                    b_far=b_far, th_far=8.0)
        rr = Fr.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"], mp["holder_obs"],
                                    b_far=b_far, th_far=8.0)
        assert ro[0] == rr[0] and np.array_equal(ro[1], rr[1]) and np.array_equal(ro[2], rr[2])
        assert np.array_equal(ro[3][:, :4], rr[3])
        # RGB-D depth lookup
        depth = rng.uniform(-0.5, 8.0, (E["height"], E["width"])).astype(np.float32)
        xy = np.ascontiguousarray(kL[:, :2]); ux = np.ascontiguousarray(kL[:, 0] + rng.normal(0, 0.3, len(kL)).astype(np.float32))
        uo, do = oracle.stereo_from_rgbd(xy, ux, depth, float(G.MBF))
        ur, dr = oracle.ref_stereo_from_rgbd(xy, ux, depth, float(G.MBF))
        assert np.array_equal(uo, ur) and np.array_equal(do, dr)
