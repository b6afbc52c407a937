"""KannalaBrandt8 rig (TUM-VI shaped): lapping-area reorder, brute-force 2-NN + Lowe ratio + triangulation,
and the two-sided projection search with mirrored assignments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
import fasttrack_b200 as ft
from fasttrack_b200 import synth

T = synth.TUMVI


@pytest.fixture(scope="module", params=[1000, 2000])
def tum(request):
    nf = request.param
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    Tlr = np.hstack([Rlr, tlr[:, None]])
    ctx = ft.Context(T["width"], T["height"], nfeatures=nf, camera_type=1, cam1=T["cam1"], cam2=T["cam2"],
                     lap_left=T["lap"], lap_right=T["lap"], bf=T["bf"], Tlr=Tlr)
    ctx.extract_stereo(L, R)
    ctx.stereo_match()
    exL, exR = oracle.Extractor(nf), oracle.Extractor(nf)
    monoL, kL, dL = exL.extract(L, lap=T["lap"]); monoR, kR, dR = exR.extract(R, lap=T["lap"])
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, monoL, kR, dR, monoR)
    yield dict(ctx=ctx, exL=exL, kL=kL, dL=dL, kR=kR, dR=dR, monoL=monoL, monoR=monoR, fo=fo, ext=(Rlr, tlr, Rrl, trl), nf=nf)
    ctx.close()


def test_fisheye_extraction_reversed_order(tum):
    gl, gr = tum["ctx"].download(0), tum["ctx"].download(1)
    assert gl["mono_index"] == tum["monoL"] == 0 and gr["mono_index"] == tum["monoR"] == 0   # everything is in the lapping area
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), tum["kL"]) and np.array_equal(gl["desc"], tum["dL"])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), tum["kR"]) and np.array_equal(gr["desc"], tum["dR"])
    assert np.all(np.diff(tum["kL"][:, 5]) <= 0)        # filled from the back: octaves descend


def test_fisheye_matches(tum):
    g = tum["ctx"].download(0, stereo=True)
    fo = tum["fo"]
    acc = fo["code"] == 1
    assert acc.sum() > 30
    # Match tables are index work: bit-exact. The SVD inside Triangulate has no pinned reference arithmetic
    # (Eigen::JacobiSVD is un-vendored), so depth / 3-D points carry a tolerance.
    assert np.array_equal(g["l2r"], fo["l2r"]) and np.array_equal(g["r2l"], fo["r2l"])
    assert np.allclose(g["depth"], fo["depth"], rtol=1e-4, atol=1e-5)
    assert np.allclose(g["p3d"][acc], fo["p3d"][acc], rtol=1e-3, atol=1e-4)


def test_partial_lapping_area():
    """mono and stereo keypoints mixed: two-ended fill (ORBextractor.cc:1476-1486) and subset matching (Frame.cc:1233-1237)"""
    L, R = synth.fisheye_pair(seed=8)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    lapL, lapR = (150, 511), (0, 360)
    ctx = ft.Context(512, 512, nfeatures=800, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=lapL, lap_right=lapR,
                     bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR = oracle.Extractor(800), oracle.Extractor(800)
    monoL, kL, dL = exL.extract(L, lap=lapL); monoR, kR, dR = exR.extract(R, lap=lapR)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert 0 < monoL < len(kL) and 0 < monoR < len(kR)
    assert gl["mono_index"] == monoL and gr["mono_index"] == monoR
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), kL) and np.array_equal(gl["desc"], dL)
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), kR) and np.array_equal(gr["desc"], dR)
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, monoL, kR, dR, monoR)
    assert np.array_equal(gl["l2r"], fo["l2r"]) and np.array_equal(gl["r2l"], fo["r2l"])
    assert np.all(fo["l2r"][:monoL] == -1)
    ctx.close()


def _fisheye_map(tum, M, seed, all_obs):
    kL, dL = tum["kL"], tum["dL"]
    c1 = T["cam1"]
    mp = synth.mappoints(kL, dL, tum["exL"].scale, M, seed=seed, width=512, height=512, fx=c1[0], fy=c1[1], cx=c1[2], cy=c1[3])
    # the generator back-projects with a pinhole model; under KB8 the points land slightly elsewhere, which is fine:
    # both implementations see the same 3-D points
    N = len(kL) + len(tum["kR"])
    rng = np.random.default_rng(seed)
    holder = np.full(N, -1, np.int32); hobs = np.zeros(N, np.uint8)
    cl = rng.random(N) < 0.2
    holder[cl] = -2; hobs[cl] = 1
    if all_obs:
        mp["flags"] |= 2
    return mp, holder, hobs


@pytest.mark.parametrize("all_obs", [True, False])
def test_fisheye_projection_search(tum, all_obs):
    """all_obs=True runs the fix-point kernel; False forces the in-order kernel (a map point without observations
    can un-block a mirrored keypoint, ORBmatcher.cc:144-148)."""
    ctx = tum["ctx"]
    M = 6000
    mp, holder, hobs = _fisheye_map(tum, M, 13, all_obs)
    Rlr, tlr, Rrl, trl = tum["ext"]
    keys = np.vstack([tum["kL"], tum["kR"]]); desc = np.vstack([tum["dL"], tum["dR"]])
    F = oracle.Frame(keys, desc, tum["exL"].scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"],
                     n_left=len(tum["kL"]), n_right=len(tum["kR"]), l2r=tum["fo"]["l2r"], r2l=tum["fo"]["r2l"],
                     Rrl=Rrl, trl=trl, tlr=tlr)
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
    gi, gf = ctx.track(M)
    # KB8 projection uses atan2f/cosf/sinf: device and glibc may differ in the last ulp -> tolerance on the
    # projected coordinates, exactness on the decisions wherever the oracle is not within 1e-5 of a boundary
    clean = ti[:, 4] == 0
    assert np.array_equal(gi[clean, :2], ti[clean, :2])
    assert np.allclose(gf, tf, rtol=1e-4, atol=2e-3)
    same_inputs = np.array_equal(gi, ti[:, :4]) and np.array_equal(gf, tf)
    if same_inputs:
        assert n_g == n_o and np.array_equal(h_g, h_o) and np.array_equal(ho_g, ho_o)
    else:
        assert abs(n_g - n_o) <= 0.02 * max(n_o, 1) + 2
    assert n_o > 20


def test_fisheye_store_search_equals_snapshot_search(tum):
    """persistent map store on the fisheye path, both resolve kernels"""
    ctx = tum["ctx"]
    ctx.set_pose(np.eye(3), np.zeros(3))
    M, CAP = 5000, 12000
    ctx.map_store_create(CAP)
    for all_obs in (True, False):
        mp, holder, hobs = _fisheye_map(tum, M, 29, all_obs)
        slots = np.random.default_rng(7).permutation(CAP)[:M].astype(np.int32)
        ctx.map_store_update(slots, mp["pos"], mp["normal"], mp["minmax"], mp["desc"])
        ref = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
        got = ctx.search_store(slots, mp["flags"], 3.0, holder, hobs)
        assert ref[0] == got[0] and ref[0] > 20
        for a, b in zip(ref[1:], got[1:]):
            assert np.array_equal(a, b)


def test_fisheye_many_features_search():
    """4000 features per eye on the fisheye rig: the frame-side search structure of both eyes no longer fits the gather
    kernel's shared memory, so the window walks read it from L2 (the kernel's other code path)"""
    nf = 4000
    L, R = synth.fisheye_pair(seed=5)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    ctx = ft.Context(T["width"], T["height"], nfeatures=nf, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=T["lap"],
                     lap_right=T["lap"], bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR = oracle.Extractor(nf), oracle.Extractor(nf)
    mL, kL, dL = exL.extract(L, lap=T["lap"]); mR, kR, dR = exR.extract(R, lap=T["lap"])
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert len(kL) > 3000
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), kL) and np.array_equal(gl["desc"], dL)
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), kR) and np.array_equal(gr["desc"], dR)
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
    # ~600 ratio-test survivors go through TriangulateMatches here; its accept / reject thresholds (parallax, depth signs,
    # chi-square) are evaluated on values that come out of device tanf / atan2f and the Jacobi SVD, so a decision that sits on
    # a threshold may differ from the host's (the non-pinnable residue of DESIGN.md section 2): counted, not tolerated silently
    dl = int((gl["l2r"] != fo["l2r"]).sum()); dr = int((gl["r2l"] != fo["r2l"]).sum())
    assert dl <= 3 and dr <= 3, "fisheye match tables differ in %d / %d entries" % (dl, dr)
    tables_equal = dl == 0 and dr == 0
    c1 = T["cam1"]
    M = 8000
    mp = synth.mappoints(kL, dL, exL.scale, M, seed=17, width=512, height=512, fx=c1[0], fy=c1[1], cx=c1[2], cy=c1[3])
    mp["flags"] |= 2
    N = len(kL) + len(kR)
    holder = np.full(N, -1, np.int32); hobs = np.zeros(N, np.uint8)
    F = oracle.Frame(np.vstack([kL, kR]), np.vstack([dL, dR]), exL.scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"],
                     mbf=T["bf"], n_left=len(kL), n_right=len(kR), l2r=fo["l2r"], r2l=fo["r2l"], Rrl=Rrl, trl=trl, tlr=tlr)
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
    gi, gf = ctx.track(M)
    assert n_o > 100
    if tables_equal and np.array_equal(gi, ti[:, :4]) and np.array_equal(gf, tf):
        assert n_g == n_o and np.array_equal(h_g, h_o) and np.array_equal(ho_g, ho_o)
    else:   # KB8 projection through device atan2f / sinf / cosf: last-ulp differences may move a borderline decision
        assert abs(n_g - n_o) <= 0.02 * n_o + 2
    ctx.close()
