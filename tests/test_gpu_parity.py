"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden vectors. Bit-exact for keypoints, octaves, descriptors, match
indices; stereo uRight/depth compared exactly (north_star tolerance: 1e-3 px)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from sbp_check import assert_search_matches
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC


def _ctx(c, **kw):
    mbf = np.float32(c["fx"] * c["baseline"])
    return ft.Context(c["width"], c["height"], nfeatures=c.get("nfeatures", 1200), nlevels=c.get("nlevels", 8),
                      cam1=[c["fx"], c["fy"], c["cx"], c["cy"]], bf=float(mbf), **kw), mbf, np.float32(mbf / np.float32(c["fx"]))


def _oracle_pair(L, R, nf, nl):
    exL, exR = oracle.Extractor(nf, 1.2, nl), oracle.Extractor(nf, 1.2, nl)
    monoL, kL, dL = exL.extract(L); monoR, kR, dR = exR.extract(R)
    return exL, exR, (monoL, kL, dL), (monoR, kR, dR)


@pytest.fixture(scope="module")
def euroc(euroc_pair):
    L, R = euroc_pair
    ctx, mbf, mb = _ctx(E)
    ctx.extract_stereo(L, R)
    ctx.stereo_match()
    exL, exR, oL, oR = _oracle_pair(L, R, 1200, 8)
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    yield dict(ctx=ctx, mbf=mbf, mb=mb, exL=exL, exR=exR, oL=oL, oR=oR, st=st, L=L, R=R)
    ctx.close()


def test_scale_tables_match_oracle(euroc):
    t = euroc["ctx"].scale_tables()
    ex = euroc["exL"]
    assert np.array_equal(t["scale"], ex.scale) and np.array_equal(t["inv_scale"], ex.inv_scale)
    assert np.array_equal(t["sigma2"], ex.sigma2) and np.array_equal(t["inv_sigma2"], ex.inv_sigma2)
    assert np.array_equal(t["features_per_level"], ex.features_per_level)


def test_pyramid_blur_candidates_bit_exact(euroc):
    ctx = euroc["ctx"]
    for eye, ex in ((0, euroc["exL"]), (1, euroc["exR"])):
        for l in range(8):
            assert np.array_equal(ctx.level_image(eye, l), ex.level_image(l)), "pyramid eye %d level %d" % (eye, l)
            assert np.array_equal(ctx.level_image(eye, l, True), ex.level_image(l, True)), "blur eye %d level %d" % (eye, l)
            assert np.array_equal(ctx.level_candidates(eye, l), ex.level_candidates(l)), "FAST eye %d level %d" % (eye, l)


def test_keypoints_and_descriptors_bit_exact(euroc):
    ctx = euroc["ctx"]
    for eye, (mono, k, d) in ((0, euroc["oL"]), (1, euroc["oR"])):
        g = ctx.download(eye)
        assert g["n"] == len(k) and g["mono_index"] == mono
        assert np.array_equal(ft.keypoints_as_array(g["kps"]), k)
        differing = int((g["desc"] != d).any(axis=1).sum())
        # north_star: descriptors may differ only where a rotated sample sits within 1e-4 of a rounding boundary
        borderline = (euroc["exL"] if eye == 0 else euroc["exR"]).desc_borderline()
        assert differing <= borderline
        assert differing == 0, "descriptor rows differ: %d (borderline samples %d)" % (differing, borderline)


def test_stereo_bit_exact(euroc):
    g = euroc["ctx"].download(0, stereo=True)
    st = euroc["st"]
    assert int((st["depth"] > 0).sum()) > 200
    assert np.array_equal(g["u_right"], st["uRight"])
    assert np.abs(g["u_right"] - st["uRight"]).max() <= 1e-3      # the stated tolerance; exact in practice
    assert np.array_equal(g["depth"], st["depth"])


@pytest.mark.parametrize("M,th", [(5000, 1.0), (10000, 2.0), (20000, 6.0), (20000, 10.0)])
def test_search_by_projection_bit_exact(euroc, M, th):
    ctx = euroc["ctx"]
    _, kL, dL = euroc["oL"]
    mp = synth.mappoints(kL, dL, euroc["exL"].scale, M, seed=4 + M)
    F = oracle.Frame(kL, dL, euroc["exL"].scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                     mbf=float(euroc["mbf"]), u_right=euroc["st"]["uRight"])
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th,
                                                   mp["holder"], mp["holder_obs"])
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, best = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th,
                                                   mp["holder"], mp["holder_obs"])
    gi, gf = ctx.track(M)
    clean = ti[:, 4] == 0          # decisions within 1e-5 of a float boundary are counted, not compared
    assert (~clean).sum() < 0.002 * M + 5
    assert np.array_equal(gi[clean, 0], ti[clean, 0]) and np.array_equal(gi[clean, 2], ti[clean, 2])
    assert np.array_equal(gf[clean, :5], tf[clean, :5])
    def prefix(k):
        a = {key: v[:k] for key, v in mp.items() if key not in ("holder", "holder_obs")}
        o = F.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], th, mp["holder"], mp["holder_obs"])
        g = ctx.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], th, mp["holder"], mp["holder_obs"])
        return g[:3], o[:3]
    assert_search_matches((n_g, h_g, ho_g), (n_o, h_o, ho_o), gi, ti, prefix)
    assert n_o > 100


def test_search_by_projection_moved_pose(euroc):
    """non-identity pose: rotate/translate the camera and the map together, results must not change class"""
    ctx = euroc["ctx"]
    _, kL, dL = euroc["oL"]
    M = 8000
    mp = synth.mappoints(kL, dL, euroc["exL"].scale, M, seed=77)
    a = 0.3
    Rwc = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    Ow = np.array([0.5, -0.2, 1.0], np.float32)
    pos = (Rwc @ mp["pos"].T).T + Ow
    nrm = (Rwc @ mp["normal"].T).T
    Rcw = np.ascontiguousarray(Rwc.T); tcw = -(Rcw @ Ow)
    F = oracle.Frame(kL, dL, euroc["exL"].scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                     mbf=float(euroc["mbf"]), u_right=euroc["st"]["uRight"], Rcw=Rcw, tcw=tcw)
    n_o, h_o, ho_o, ti, tf = F.search_local_points(pos, nrm, mp["minmax"], mp["desc"], mp["flags"], 3.0, mp["holder"], mp["holder_obs"])
    ctx.set_pose(Rcw, tcw, F.Rwc, F.Ow)
    n_g, h_g, ho_g, _ = ctx.search_local_points(pos, nrm, mp["minmax"], mp["desc"], mp["flags"], 3.0, mp["holder"], mp["holder_obs"])
    gi, gf = ctx.track(M)
    clean = ti[:, 4] == 0
    assert np.array_equal(gi[clean, 0], ti[clean, 0]) and np.array_equal(gi[clean, 2], ti[clean, 2])
    assert np.array_equal(gf[clean, :5], tf[clean, :5])
    def prefix(k):
        o = F.search_local_points(pos[:k], nrm[:k], mp["minmax"][:k], mp["desc"][:k], mp["flags"][:k], 3.0, mp["holder"], mp["holder_obs"])
        g = ctx.search_local_points(pos[:k], nrm[:k], mp["minmax"][:k], mp["desc"][:k], mp["flags"][:k], 3.0, mp["holder"], mp["holder_obs"])
        return g[:3], o[:3]
    assert_search_matches((n_g, h_g, ho_g), (n_o, h_o, ho_o), gi, ti, prefix)


def test_grid_matches_oracle(euroc):
    _, kL, dL = euroc["oL"]
    F = oracle.Frame(kL, dL, euroc["exL"].scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0])
    co, io = F.grid()
    cg, ig = euroc["ctx"].grid()
    assert np.array_equal(co, cg) and np.array_equal(io, ig)


def test_golden_mini_pipeline(mini_golden, mini_cfg):
    """CUDA path against committed vectors (no live oracle involved)."""
    g = mini_golden
    ctx, mbf, mb = _ctx(mini_cfg, max_map_points=4000)
    ctx.extract_stereo(g["imgL"], g["imgR"])
    ctx.stereo_match()
    l, r = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(l["kps"]), g["kL"]) and np.array_equal(l["desc"], g["dL"])
    assert np.array_equal(ft.keypoints_as_array(r["kps"]), g["kR"]) and np.array_equal(r["desc"], g["dR"])
    assert np.array_equal(l["u_right"], g["uRight"]) and np.array_equal(l["depth"], g["depth"])
    ctx.set_pose(np.eye(3), np.zeros(3))
    n, holder, hobs, _ = ctx.search_local_points(g["mp_pos"], g["mp_normal"], g["mp_minmax"], g["mp_desc"], g["mp_flags"], 3.0,
                                                 g["mp_holder"], g["mp_holder_obs"])
    assert n == int(g["sbp_n"]) and np.array_equal(holder, g["sbp_holder"]) and np.array_equal(hobs, g["sbp_holder_obs"])
    ctx.close()


def test_graph_and_direct_launch_agree(euroc_pair):
    L, R = euroc_pair
    ctx, _, _ = _ctx(E)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    a = ctx.download(0, stereo=True)
    ctx.set_use_graph(False)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    b = ctx.download(0, stereo=True)
    assert np.array_equal(a["kps"], b["kps"]) and np.array_equal(a["desc"], b["desc"]) and np.array_equal(a["u_right"], b["u_right"])
    ctx.close()


def test_repeated_frames_are_deterministic_and_independent(euroc_pair):
    """frame t+1 must not see state of frame t: extract A, B, A again -> identical results for A"""
    L, R = euroc_pair
    sc = synth.StereoScene(seed=5)
    L2, R2 = sc.pair(pan=(10, 3), noise_seed=1)
    ctx, _, _ = _ctx(E)
    ctx.extract_stereo(L, R); ctx.stereo_match(); a = ctx.download(0, stereo=True)
    ctx.extract_stereo(L2, R2); ctx.stereo_match(); b = ctx.download(0, stereo=True)
    ctx.extract_stereo(L, R); ctx.stereo_match(); c = ctx.download(0, stereo=True)
    assert np.array_equal(a["kps"], c["kps"]) and np.array_equal(a["desc"], c["desc"]) and np.array_equal(a["depth"], c["depth"])
    assert not np.array_equal(a["kps"], b["kps"])
    ctx.close()


@pytest.mark.parametrize("kind", ["flat", "noise", "half_flat", "strided"])
def test_edge_images(kind):
    """empty result, saturated candidate counts, cells that need the minThFAST fallback, non-tight row pitch"""
    c = dict(E, nfeatures=500)
    rng = np.random.default_rng(11)
    if kind == "flat":
        L = np.full((480, 752), 90, np.uint8); R = L.copy()
    elif kind == "noise":
        L = rng.integers(0, 256, (480, 752), dtype=np.uint8); R = rng.integers(0, 256, (480, 752), dtype=np.uint8)
    elif kind == "half_flat":
        L = synth.texture(480, 752, 21); L[:, 376:] = (L[:, 376:] // 16 + 100).astype(np.uint8); R = np.roll(L, -7, axis=1)
    else:
        big = synth.texture(480, 800, 22)
        L = big[:, 13:765]; R = big[:, 5:757]
        assert L.strides[0] == 800
    ctx, mbf, mb = _ctx(c)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR, oL, oR = _oracle_pair(np.ascontiguousarray(L), np.ascontiguousarray(R), 500, 8)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    assert np.array_equal(gl["u_right"], st["uRight"]) and np.array_equal(gl["depth"], st["depth"])
    if kind == "flat":
        assert gl["n"] == 0
    for l in range(8):
        assert np.array_equal(ctx.level_candidates(0, l), exL.level_candidates(l))
    ctx.close()


# (480, 978) and (400, 1083): the last FAST cell column of level 0 is 4 / 1 interior pixels wide (one staged word per row),
# (360, 1344): the same at level 1 (1120 px) -- the degenerate divisors of the per-cell index arithmetic
@pytest.mark.parametrize("shape,nf,nl", [((480, 640), 1000, 8), ((376, 1241), 2000, 8), ((512, 512), 1000, 8), ((240, 376), 300, 5),
                                         ((480, 978), 1500, 8), ((400, 1083), 1500, 6), ((360, 1344), 1500, 4)])
def test_other_resolutions(shape, nf, nl):
    h, w = shape
    sc = synth.StereoScene(seed=31 + h, width=w, height=h, dmin=1.0, dmax=30.0, margin_x=96, margin_y=8)
    L, R = sc.pair()
    c = dict(width=w, height=h, nfeatures=nf, nlevels=nl, fx=400.0, fy=400.0, cx=w / 2.0, cy=h / 2.0, baseline=0.1)
    ctx, mbf, mb = _ctx(c)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR, oL, oR = _oracle_pair(L, R, nf, nl)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    assert np.array_equal(gl["u_right"], st["uRight"]) and np.array_equal(gl["depth"], st["depth"])
    ctx.close()


FUZZ_SHAPES = [(241, 320), (255, 443), (300, 599), (333, 641), (401, 753), (479, 851), (487, 999), (512, 1001), (350, 1119),
               (377, 1153), (600, 800), (721, 1279), (384, 700), (297, 529), (450, 451), (613, 613)]


@pytest.mark.parametrize("shape", FUZZ_SHAPES)
def test_resolution_fuzz(shape):
    """odd widths / heights, pitches that are no multiple of 4, square and wide images, 4-8 levels: every intermediate
    stage of the extractor against the oracle (the FAST cell geometry, the resize tables and the blur tiles all depend on
    the size)"""
    h, w = shape
    nl = 4 + (h + w) % 5
    nf = 300 + (h * 7 + w) % 900
    img = synth.StereoScene(seed=h + w, width=w, height=h, dmin=1.0, dmax=20.0, margin_x=64, margin_y=8).pair()[0]
    ctx = ft.Context(w, h, nfeatures=nf, nlevels=nl, cam1=[400.0, 400.0, w / 2.0, h / 2.0], bf=40.0)
    ctx.extract_stereo(img, img)
    ex = oracle.Extractor(nf, 1.2, nl)
    mono, k, d = ex.extract(img)
    for eye in (0, 1):
        for l in range(nl):
            assert np.array_equal(ctx.level_image(eye, l), ex.level_image(l)), (shape, eye, l, "pyramid")
            assert np.array_equal(ctx.level_image(eye, l, blurred=True), ex.level_image(l, blurred=True)) or \
                ex.level_image(l, blurred=True) is None, (shape, eye, l, "blur")
            assert np.array_equal(ctx.level_candidates(eye, l), ex.level_candidates(l)), (shape, eye, l, "FAST candidates")
        r = ctx.download(eye)
        assert np.array_equal(ft.keypoints_as_array(r["kps"]), k) and np.array_equal(r["desc"], d), (shape, eye)
    ctx.close()


@pytest.mark.parametrize("shape,nl,sf", [((480, 752), 1, 1.2), ((480, 752), 2, 1.2), ((360, 500), 3, 1.5), ((640, 480), 8, 1.2),
                                         ((700, 400), 5, 1.2), ((300, 300), 2, 2.0)])
def test_few_levels_portrait_and_scale_factors(shape, nl, sf):
    """1-3 pyramid levels (the per-level branches of the launch graph), portrait images (one octree root) and other scale
    factors, both eyes against the oracle incl. stereo"""
    h, w = shape
    nf = 600
    L, R = synth.StereoScene(seed=h * 3 + w + nl, width=w, height=h, dmin=1.0, dmax=24.0, margin_x=64, margin_y=8).pair()
    mbf = np.float32(40.0)
    ctx = ft.Context(w, h, nfeatures=nf, nlevels=nl, scale_factor=sf, cam1=[400.0, 400.0, w / 2.0, h / 2.0], bf=float(mbf))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR = oracle.Extractor(nf, sf, nl), oracle.Extractor(nf, sf, nl)
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), kL) and np.array_equal(gl["desc"], dL)
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), kR) and np.array_equal(gr["desc"], dR)
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mbf / np.float32(400.0)))
    assert np.array_equal(gl["u_right"], st["uRight"]) and np.array_equal(gl["depth"], st["depth"])
    ctx.close()


@pytest.mark.parametrize("nf", [5000, 6500, 10000])
def test_many_features(nf):
    """the documented upper end (10000 features on a 1280x720 pair): octree levels with thousands of nodes, the gather
    kernel's path without the shared-memory copy of the frame structure (more than ~7800 keypoints), stereo and a
    25000-point local-map search"""
    w, h = 1280, 720
    L, R = synth.StereoScene(seed=77, width=w, height=h, dmin=1.0, dmax=48.0, margin_x=96, margin_y=8).pair()
    c = dict(width=w, height=h, nfeatures=nf, nlevels=8, fx=700.0, fy=700.0, cx=w / 2.0, cy=h / 2.0, baseline=0.1)
    ctx, mbf, mb = _ctx(c)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR, oL, oR = _oracle_pair(L, R, nf, 8)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert len(oL[1]) > 0.8 * nf
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    assert np.array_equal(gl["u_right"], st["uRight"]) and np.array_equal(gl["depth"], st["depth"])
    M = 25000
    mp = synth.mappoints(oL[1], oL[2], exL.scale, M, seed=5, width=w, height=h, fx=700.0, fy=700.0, cx=w / 2.0, cy=h / 2.0)
    F = oracle.Frame(oL[1], oL[2], exL.scale, w, h, cam1=[700.0, 700.0, w / 2.0, h / 2.0, 0, 0, 0, 0], mbf=float(mbf),
                     u_right=st["uRight"])
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, mp["holder"],
                                                   mp["holder_obs"])
    ctx.set_pose(np.eye(3), np.zeros(3))
    n_g, h_g, ho_g, _ = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, mp["holder"],
                                                mp["holder_obs"])
    gi, gf = ctx.track(M)
    assert n_o > 1000
    def prefix(k):
        a = {key: v[:k] for key, v in mp.items() if key not in ("holder", "holder_obs")}
        o = F.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        g = ctx.search_local_points(a["pos"], a["normal"], a["minmax"], a["desc"], a["flags"], 3.0, mp["holder"], mp["holder_obs"])
        return g[:3], o[:3]
    assert_search_matches((n_g, h_g, ho_g), (n_o, h_o, ho_o), gi, ti, prefix, max_flips=5)
    ctx.close()


@pytest.mark.parametrize("shape,nf", [((1080, 1920), 3000), ((1440, 2560), 2000)])
def test_large_images(shape, nf):
    """1080p / 1440p: more FAST candidates per level than the octree keeps in shared memory (its HBM scratch path), 50+ cell
    columns, pyramid slabs of several MB"""
    h, w = shape
    L, R = synth.StereoScene(seed=h, width=w, height=h, dmin=1.0, dmax=60.0, margin_x=128, margin_y=8).pair()
    c = dict(width=w, height=h, nfeatures=nf, nlevels=8, fx=1000.0, fy=1000.0, cx=w / 2.0, cy=h / 2.0, baseline=0.1)
    ctx, mbf, mb = _ctx(c)
    ctx.extract_stereo(L, R); ctx.stereo_match()
    exL, exR, oL, oR = _oracle_pair(L, R, nf, 8)
    assert len(exL.level_candidates(0)) > 16384
    for l in range(8):
        assert np.array_equal(ctx.level_candidates(0, l), exL.level_candidates(l)), (shape, l)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    assert np.array_equal(gl["u_right"], st["uRight"]) and np.array_equal(gl["depth"], st["depth"])
    ctx.close()


def test_octree_fuzz_many_seeds():
    """the order-sensitive octree (incl. the std::sort tie order) on many different textures / quotas"""
    for seed in range(12):
        nf = [150, 400, 800, 1200, 2000, 3000][seed % 6]
        img = synth.texture(480, 752, 100 + seed, n_rect=2000 + 700 * seed, n_disc=500 * (seed % 4))
        R = np.roll(img, -5, axis=1)
        c = dict(E, nfeatures=nf)
        ctx, _, _ = _ctx(c)
        ctx.extract_stereo(img, R)
        ex = oracle.Extractor(nf, 1.2, 8)
        mono, k, d = ex.extract(img)
        g = ctx.download(0)
        assert np.array_equal(ft.keypoints_as_array(g["kps"]), k), "seed %d nfeatures %d" % (seed, nf)
        assert np.array_equal(g["desc"], d)
        ctx.close()


def test_api_errors():
    ctx, _, _ = _ctx(E)
    with pytest.raises(ft.FtError) as e:
        ctx.stereo_match()
    assert e.value.status == 4
    with pytest.raises(ft.FtError) as e:
        ft.Context(32, 32)
    assert e.value.status == 1
    with pytest.raises(ft.FtError) as e:
        ft.Context(752, 480, nlevels=14)       # top levels smaller than one FAST cell: the reference divides by zero
    assert e.value.status == 1
    L = synth.texture(480, 752, 1)
    ctx.extract_stereo(L, L); ctx.stereo_match()
    n = ctx.counts()["n_left"]
    M = 30000
    z = np.zeros
    with pytest.raises(ft.FtError) as e:       # reference: raise(SIGSEGV) beyond 25000 map points
        ctx.search_local_points(z((M, 3)), z((M, 3)), z((M, 2)), z((M, 32), np.uint8), z(M, np.int32), 1.0, np.full(n, -1), z(n, np.uint8))
    assert e.value.status == 3
    nm, holder, _, _ = ctx.search_local_points(z((0, 3)), z((0, 3)), z((0, 2)), z((0, 32), np.uint8), z(0, np.int32), 1.0,
                                               np.full(n, -1), z(n, np.uint8))
    assert nm == 0 and np.all(holder == -1)
    ctx.close()


def test_frame_construct_one_call(euroc):
    """ft_frame_construct == extract + stereo + downloads"""
    ctx = euroc["ctx"]
    left, right = ctx.frame_construct(euroc["L"], euroc["R"])
    _, kL, dL = euroc["oL"]; _, kR, dR = euroc["oR"]
    assert np.array_equal(ft.keypoints_as_array(left["kps"]), kL) and np.array_equal(left["desc"], dL)
    assert np.array_equal(ft.keypoints_as_array(right["kps"]), kR) and np.array_equal(right["desc"], dR)
    assert np.array_equal(left["u_right"], euroc["st"]["uRight"]) and np.array_equal(left["depth"], euroc["st"]["depth"])


def test_resident_search_matches_host_call(euroc):
    """ft_upload_map_points + ft_upload_holders + ft_search_resident + ft_search_download == ft_search_local_points"""
    ctx = euroc["ctx"]
    _, kL, dL = euroc["oL"]
    M = 7000
    mp = synth.mappoints(kL, dL, euroc["exL"].scale, M, seed=55)
    ctx.set_pose(np.eye(3), np.zeros(3))
    n1, h1, o1, b1 = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                             mp["holder"], mp["holder_obs"])
    ctx.upload_map_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"])
    ctx.upload_holders(mp["holder"], mp["holder_obs"])
    for _ in range(3):          # repeated searches over the resident snapshot start from the same holders
        ctx.search_resident(3.0)
        n2, h2, o2, b2 = ctx.search_download(M)
        assert n1 == n2 and np.array_equal(h1, h2) and np.array_equal(o1, o2) and np.array_equal(b1, b2)


def test_staged_search_matches_host_call(euroc):
    """ft_map_point_staging + ft_search_staged (one H2D, one D2H) == ft_search_local_points"""
    ctx = euroc["ctx"]
    _, kL, dL = euroc["oL"]
    M = 5000
    mp = synth.mappoints(kL, dL, euroc["exL"].scale, M, seed=91)
    ctx.set_pose(np.eye(3), np.zeros(3))
    n1, h1, o1, b1 = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 2.0,
                                             mp["holder"], mp["holder_obs"])
    N = len(kL)
    stg = ctx.map_point_staging(M, N)
    for k in ("pos", "normal", "minmax", "desc", "flags"):
        np.copyto(stg[k], mp[k])
    np.copyto(stg["holder"], mp["holder"]); np.copyto(stg["holder_obs"], mp["holder_obs"])
    n2, h2, o2, b2 = ctx.search_staged(M, N, 2.0)
    assert n1 == n2 and np.array_equal(h1, h2) and np.array_equal(o1, o2) and np.array_equal(b1, b2)


def test_rectification_fused_into_level0(euroc_pair):
    """'next' row 2: cv::remap of System::TrackStereo in front of the extractor (raw 800x500 -> rectified 752x480)"""
    rawL = synth.texture(500, 800, 41); rawR = np.roll(rawL, -9, axis=1)
    yy, xx = np.mgrid[0:480, 0:752].astype(np.float32)
    def maps(cx, cy, k1, dx, dy):
        xn, yn = (xx - cx) / 458.6, (yy - cy) / 457.3
        r2 = xn * xn + yn * yn
        f = 1 + k1 * r2 + 0.074 * r2 * r2
        return (xn * f * 470.0 + cx + 24 + dx).astype(np.float32), (yn * f * 470.0 + cy + 10 + dy).astype(np.float32)
    m1l, m2l = maps(367.2, 248.4, -0.283, 1.3, -0.7)
    m1r, m2r = maps(379.9, 255.2, -0.284, -2.1, 0.4)
    ctx, mbf, mb = _ctx(E)
    ctx.set_rectification(800, 500, m1l, m2l, m1r, m2r)
    ctx.extract_stereo(rawL, rawR); ctx.stereo_match()
    rectL, rectR = oracle.remap(rawL, m1l, m2l), oracle.remap(rawR, m1r, m2r)
    assert np.array_equal(ctx.level_image(0, 0), rectL) and np.array_equal(ctx.level_image(1, 0), rectR)
    exL, exR, oL, oR = _oracle_pair(rectL, rectR, 1200, 8)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    st = oracle.stereo(exL, exR, oL[1], oL[2], oR[1], oR[2], float(mbf), float(mb))
    assert np.array_equal(gl["u_right"], st["uRight"])
    # switching it off again restores the plain path
    ctx.set_rectification(0, 0, None, None, None, None)
    L, R = euroc_pair
    ctx.extract_stereo(L, R)
    assert np.array_equal(ctx.level_image(0, 0), L)
    ctx.close()


@pytest.mark.parametrize("raw", [(1504, 960), (1280, 720), (640, 400)])
def test_input_resize_fused_into_level0(raw):
    """'next' row 2: cv::resize(im, newImSize) of System::TrackStereo (System.cc:282-285) in front of the extractor;
    exact 2x (OpenCV's INTER_AREA shortcut, same bytes), a non-integer downscale and an upscale"""
    rw, rh = raw
    rawL = synth.texture(rh, rw, 43); rawR = np.roll(rawL, -7, axis=1)
    ctx, mbf, mb = _ctx(E)
    ctx.set_input_resize(rw, rh)
    ctx.extract_stereo(rawL, rawR); ctx.stereo_match()
    inL, inR = oracle.resize(rawL, E["width"], E["height"]), oracle.resize(rawR, E["width"], E["height"])
    assert np.array_equal(ctx.level_image(0, 0), inL) and np.array_equal(ctx.level_image(1, 0), inR)
    exL, exR, oL, oR = _oracle_pair(inL, inR, 1200, 8)
    gl, gr = ctx.download(0, stereo=True), ctx.download(1)
    assert np.array_equal(ft.keypoints_as_array(gl["kps"]), oL[1]) and np.array_equal(gl["desc"], oL[2])
    assert np.array_equal(ft.keypoints_as_array(gr["kps"]), oR[1]) and np.array_equal(gr["desc"], oR[2])
    with pytest.raises(RuntimeError, match="one or the other"):
        m = np.zeros((E["height"], E["width"]), np.float32)
        ctx.set_rectification(rw, rh, m, m, m, m)
    ctx.set_input_resize(0, 0)
    L2 = synth.texture(E["height"], E["width"], 44)
    ctx.extract_stereo(L2, L2)
    assert np.array_equal(ctx.level_image(0, 0), L2)
    ctx.close()


def test_two_contexts_pipelined_match_sequential(euroc_pair):
    """bench.py's throughput leg alternates the frames of one sequence between two contexts (2-deep pipeline):
    every frame's results must equal the one-context, one-frame-at-a-time results"""
    import torch
    L, R = euroc_pair
    sc = synth.StereoScene(seed=9)
    frames = [(L, R)] + [sc.pair(pan=(3 * t, t), noise_seed=t) for t in range(1, 6)]
    ctxs = [_ctx(E)[0] for _ in range(2)]
    ref = _ctx(E)[0]
    streams = [torch.cuda.ExternalStream(c.stream()) for c in ctxs]
    dimgs = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in frames]
    exp, maps = [], []
    for (a, b) in frames:
        l, r = ref.frame_construct(a, b)
        mp = synth.mappoints(ft.keypoints_as_array(l["kps"]), l["desc"], ref.scale_tables()["scale"], 3000, seed=len(exp))
        mp["flags"] |= 2
        ref.set_pose(np.eye(3), np.zeros(3))
        n, h, ho, _ = ref.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                              np.full(l["n"], -1, np.int32), np.zeros(l["n"], np.uint8))
        exp.append((l, n, h)); maps.append({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in mp.items() if k in ("pos", "normal", "minmax", "desc", "flags")})
    for c in ctxs:
        c.set_pose(np.eye(3), np.zeros(3)); c.upload_holders(None, None)
    done = [torch.cuda.Event() for _ in frames]
    got = []
    for i, (da, db) in enumerate(dimgs):
        c, s = ctxs[i & 1], streams[i & 1]
        c.frame_enqueue_device(da.data_ptr(), E["width"], db.data_ptr(), E["width"])
        if i > 0:
            s.wait_event(done[i - 1])
        m = maps[i]
        c.bind_map_points_device(3000, m["pos"].data_ptr(), m["normal"].data_ptr(), m["minmax"].data_ptr(), m["desc"].data_ptr(),
                                 m["flags"].data_ptr())
        c.search_resident(3.0)
        done[i].record(s)
        if i >= 1:      # results of frame i-1 are read back while frame i is in flight on the other context
            p = ctxs[(i - 1) & 1]
            got.append((p.download(0, stereo=True), p.search_download(3000)))
    got.append((ctxs[(len(frames) - 1) & 1].download(0, stereo=True), ctxs[(len(frames) - 1) & 1].search_download(3000)))
    for (l, n, h), (gl, (gn, gh, _, _)) in zip(exp, got):
        assert np.array_equal(gl["kps"], l["kps"]) and np.array_equal(gl["desc"], l["desc"]) and np.array_equal(gl["u_right"], l["u_right"])
        assert gn == n and np.array_equal(gh, h)
    for c in ctxs + [ref]:
        c.close()


def test_frame_submit_collect_two_contexts(euroc_pair):
    """ft_frame_submit / ft_frame_collect on two contexts (frame t+1 in flight while frame t is collected) return what
    ft_frame_construct returns; collect without a pending frame is FT_ERR_STATE"""
    L, R = euroc_pair
    sc = synth.StereoScene(seed=11)
    frames = [(L, R)] + [sc.pair(pan=(2 * t, -t), noise_seed=20 + t) for t in range(1, 5)]
    ref = _ctx(E)[0]
    exp = [ref.frame_construct(a, b) for a, b in frames]
    ctxs = [_ctx(E)[0] for _ in range(2)]
    with pytest.raises(RuntimeError):
        ctxs[0].frame_collect()
    ctxs[0].frame_submit(*frames[0])
    for i in range(len(frames)):
        if i + 1 < len(frames):
            ctxs[(i + 1) & 1].frame_submit(*frames[i + 1])
        l, r = ctxs[i & 1].frame_collect()
        el, er = exp[i]
        for k in ("kps", "desc", "u_right", "depth"):
            assert np.array_equal(l[k], el[k]), (i, k)
        assert np.array_equal(r["kps"], er["kps"]) and np.array_equal(r["desc"], er["desc"])
    for c in ctxs + [ref]:
        c.close()
