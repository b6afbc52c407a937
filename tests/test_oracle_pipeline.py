"""Operator-level checks of the CPU oracle: committed golden vectors, and properties the reference's
algorithms guarantee (ORBextractor.cc, Frame.cc, ORBmatcher.cc CPU branches)."""
import numpy as np
import pytest

import oracle
from fasttrack_b200 import synth


@pytest.fixture(scope="module")
def mini(mini_golden, mini_cfg):
    c = mini_cfg
    exL = oracle.Extractor(c["nfeatures"], 1.2, c["nlevels"]); exR = oracle.Extractor(c["nfeatures"], 1.2, c["nlevels"])
    monoL, kL, dL = exL.extract(mini_golden["imgL"]); monoR, kR, dR = exR.extract(mini_golden["imgR"])
    return dict(exL=exL, exR=exR, kL=kL, dL=dL, kR=kR, dR=dR, monoL=monoL, monoR=monoR)


def test_scale_tables_are_running_float_products():
    ex = oracle.Extractor()
    s = np.float32(1.0); exp = [s]
    for _ in range(7):
        s = np.float32(np.float64(s) * np.float64(np.float32(1.2))); exp.append(s)
    assert np.array_equal(ex.scale, np.array(exp, np.float32))
    assert list(ex.features_per_level) == [261, 217, 181, 151, 126, 105, 87, 72]     # SURVEY 8 derived sizes
    assert list(ex.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(oracle.Extractor(1000).features_per_level) == [217, 181, 151, 126, 105, 87, 73, 60]


def test_level_sizes_euroc():
    ex = oracle.Extractor()
    ex.extract(np.zeros((480, 752), np.uint8))
    assert [ex.level_dims(l) for l in range(8)] == [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231),
                                                    (302, 193), (252, 161), (210, 134)]


def test_golden_extractor(mini, mini_golden):
    assert np.array_equal(mini["kL"], mini_golden["kL"]) and np.array_equal(mini["dL"], mini_golden["dL"])
    assert np.array_equal(mini["kR"], mini_golden["kR"]) and np.array_equal(mini["dR"], mini_golden["dR"])
    assert mini["monoL"] == int(mini_golden["monoL"]) and mini["monoR"] == int(mini_golden["monoR"])


def test_extractor_invariants(mini, mini_cfg):
    k = mini["kL"]
    ex = mini["exL"]
    assert len(k) <= mini_cfg["nfeatures"] + 3 * mini_cfg["nlevels"]
    assert np.all(np.diff(k[:, 5]) >= 0)                       # level-major order when nothing is in the lapping area
    for l in range(mini_cfg["nlevels"]):
        n = int((k[:, 5] == l).sum())
        assert n <= ex.features_per_level[l] + 3
        lk, _ = ex.level_keys(l)
        assert len(lk) == n
        w, h = ex.level_dims(l)
        assert np.all(lk[:, 0] >= 19) and np.all(lk[:, 0] < w - 19) and np.all(lk[:, 1] >= 19) and np.all(lk[:, 1] < h - 19)
        assert np.all(lk[:, 2] == np.float32(int(np.float32(31) * ex.scale[l])))
    assert np.all((k[:, 3] >= 0) & (k[:, 3] < 360))
    assert np.all(k[:, 4] >= 7)


def test_empty_and_flat_images():
    ex = oracle.Extractor(400, 1.2, 6)
    mono, k, d = ex.extract(np.full((240, 376), 77, np.uint8))
    assert mono == 0 and len(k) == 0 and d.shape == (0, 32)


def test_lapping_area_reorders_from_the_back(mini_golden, mini_cfg):
    """operator() tail (ORBextractor.cc:1408,1476-1486): in-area points fill from the end backwards."""
    c = mini_cfg
    ex = oracle.Extractor(c["nfeatures"], 1.2, c["nlevels"])
    mono0, k0, d0 = ex.extract(mini_golden["imgL"], lap=(0, 0))
    lap = (100, 250)
    mono, k, d = ex.extract(mini_golden["imgL"], lap=lap)
    inl = (k0[:, 0] >= lap[0]) & (k0[:, 0] <= lap[1])
    assert mono == int((~inl).sum())
    assert np.array_equal(k[:mono], k0[~inl]) and np.array_equal(d[:mono], d0[~inl])
    assert np.array_equal(k[mono:], k0[inl][::-1]) and np.array_equal(d[mono:], d0[inl][::-1])
    monoa, ka, _ = ex.extract(mini_golden["imgL"], lap=(0, c["width"]))
    assert monoa == 0 and np.array_equal(ka, k0[::-1])


def test_octree_edge_cases():
    ex = oracle.Extractor()
    assert len(ex.octree(np.zeros((0, 3), np.float32), 16, 736, 16, 464, 100)) == 0
    one = np.array([[10, 20, 30]], np.float32)
    assert np.array_equal(ex.octree(one, 16, 736, 16, 464, 100), one)
    # all points identical position: can never be separated, the best response (first on ties) survives
    same = np.array([[50, 60, 9], [50, 60, 40], [50, 60, 40], [50, 60, 8]], np.float32)
    out = ex.octree(same, 16, 736, 16, 464, 100)
    assert len(out) == 1 and out[0, 2] == 40
    rng = np.random.default_rng(0)
    pts = np.stack([rng.integers(0, 720, 5000), rng.integers(0, 448, 5000), rng.integers(7, 200, 5000)], 1).astype(np.float32)
    for N in (1, 5, 50, 261, 1000):
        out = ex.octree(pts, 16, 736, 16, 464, N)
        assert N <= len(out) <= max(N + 3, 8)    # the first pass always splits both roots (2 -> up to 8 nodes)
        # every output is one of the inputs
        s = {tuple(p) for p in pts}
        assert all(tuple(o) in s for o in out)


def test_golden_stereo(mini, mini_golden):
    st = oracle.stereo(mini["exL"], mini["exR"], mini["kL"], mini["dL"], mini["kR"], mini["dR"], float(mini_golden["mbf"]),
                       float(mini_golden["mb"]))
    assert np.array_equal(st["uRight"], mini_golden["uRight"]) and np.array_equal(st["depth"], mini_golden["depth"])
    assert np.array_equal(st["sad"], mini_golden["sad"])


def test_stereo_recovers_layer_disparities(euroc_pair):
    """The synthetic right image shifts layer k left by d_k: matched disparities must cluster on those values."""
    L, R = euroc_pair
    E = synth.EUROC
    exL, exR = oracle.Extractor(), oracle.Extractor()
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
    ok = st["depth"] > 0
    assert ok.sum() > 200
    disp = kL[ok, 0] - st["uRight"][ok]
    layers = np.exp(np.linspace(np.log(1.5), np.log(64.0), 12))
    err = np.abs(disp[:, None] - layers[None, :]).min(1)
    assert np.mean(err < 0.5 * np.maximum(1.0, kL[ok, 5] * 0 + 1.0) * exL.scale[kL[ok, 5].astype(int)]) > 0.8
    assert np.allclose(st["depth"][ok], mbf / np.maximum(disp, 0.01), rtol=1e-5)
    assert np.all(st["uRight"][~ok] == -1) and np.all(st["depth"][~ok] == -1)


def test_stereo_no_keypoints():
    ex = oracle.Extractor(400, 1.2, 6)
    ex.extract(np.zeros((240, 376), np.uint8))
    st = oracle.stereo(ex, ex, np.zeros((0, 6), np.float32), np.zeros((0, 32), np.uint8), np.zeros((0, 6), np.float32),
                       np.zeros((0, 32), np.uint8), 25.0, 0.11)
    assert len(st["uRight"]) == 0


def test_golden_projection_search(mini, mini_golden, mini_cfg):
    c = mini_cfg
    g = mini_golden
    F = oracle.Frame(mini["kL"], mini["dL"], mini["exL"].scale, c["width"], c["height"],
                     cam1=[c["fx"], c["fy"], c["cx"], c["cy"], 0, 0, 0, 0], mbf=float(g["mbf"]), u_right=g["uRight"])
    n, holder, hobs, ti, tf = F.search_local_points(g["mp_pos"], g["mp_normal"], g["mp_minmax"], g["mp_desc"], g["mp_flags"],
                                                    3.0, g["mp_holder"], g["mp_holder_obs"])
    assert n == int(g["sbp_n"]) and np.array_equal(holder, g["sbp_holder"]) and np.array_equal(hobs, g["sbp_holder_obs"])
    assert np.array_equal(ti, g["track_i"]) and np.array_equal(tf, g["track_f"])
    # properties: pre-claimed blocking keypoints keep their foreign holder; a map point holds at most one keypoint
    pre = (g["mp_holder"] == -2) & (g["mp_holder_obs"] == 1)
    assert np.all(holder[pre] == -2)
    won = holder[holder >= 0]
    assert len(won) == len(set(won.tolist())) <= n     # keypoints won by map points without observations can be re-taken
    assert np.all(ti[won, 0] == 1) and np.all((g["mp_flags"][won] & 1) == 0)


def test_grid_matches_posingrid_rounding(mini, mini_cfg):
    c = mini_cfg
    F = oracle.Frame(mini["kL"], mini["dL"], mini["exL"].scale, c["width"], c["height"], cam1=[c["fx"], c["fy"], c["cx"], c["cy"], 0, 0, 0, 0])
    counts, idx = F.grid()
    assert counts.sum() == len(idx) <= len(mini["kL"])
    k = mini["kL"]
    gx = np.float32(64.0) / np.float32(c["width"]); gy = np.float32(48.0) / np.float32(c["height"])
    px = np.floor(k[:, 0] * gx + np.float32(0.5)).astype(int); py = np.floor(k[:, 1] * gy + np.float32(0.5)).astype(int)
    keep = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    exp = np.bincount((px * 48 + py)[keep], minlength=64 * 48)
    assert np.array_equal(counts, exp)
    start = np.concatenate([[0], np.cumsum(counts)])
    for cidx in np.nonzero(counts > 1)[0][:50]:
        seg = idx[start[cidx]:start[cidx + 1]]
        assert np.all(np.diff(seg) > 0)


def test_kb8_project_unproject_roundtrip():
    cam = np.array(synth.TUMVI["cam1"], np.float32)
    rng = np.random.default_rng(1)
    L = oracle.lib()
    for _ in range(200):
        u, v = rng.uniform(60, 450), rng.uniform(60, 450)
        ray = np.zeros(3, np.float32); uv = np.zeros(2, np.float32)
        L.fto_kb8_unproject(cam, float(u), float(v), ray)
        L.fto_cam_project(1, cam, ray, uv)
        assert abs(uv[0] - u) < 2e-3 and abs(uv[1] - v) < 2e-3


def test_fisheye_triangulation_recovers_depth():
    """Project known 3-D points into both KB8 cameras; matching descriptors must triangulate back to z."""
    cam1 = np.array(synth.TUMVI["cam1"], np.float32); cam2 = np.array(synth.TUMVI["cam2"], np.float32)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    rng = np.random.default_rng(2)
    n = 60
    P1 = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(1.0, 4.0, n)], 1).astype(np.float32)
    P2 = (Rrl @ P1.T).T + trl
    L = oracle.lib()
    kL = np.zeros((n, 6), np.float32); kR = np.zeros((n, 6), np.float32)
    for i in range(n):
        uv = np.zeros(2, np.float32)
        L.fto_cam_project(1, cam1, np.ascontiguousarray(P1[i], np.float32), uv); kL[i, :2] = uv
        L.fto_cam_project(1, cam2, np.ascontiguousarray(P2[i], np.float32), uv); kR[i, :2] = uv
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    sigma2 = oracle.Extractor().sigma2
    out = oracle.fisheye(cam1, cam2, Rlr, tlr, sigma2, kL, d, 0, kR, d, 0)
    assert np.array_equal(out["l2r"], np.arange(n)) and np.array_equal(out["r2l"], np.arange(n))
    assert np.allclose(out["depth"], P1[:, 2], rtol=2e-3)
    assert np.allclose(out["p3d"], P1, atol=2e-2)
