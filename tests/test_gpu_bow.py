"""GPU parity of the bag-of-words row: ft_vocabulary_* / ft_compute_bow / ft_search_by_bow against
the oracle restatement of DBoW2's transform (itself pinned against the reference's own DBoW2 code, tests/test_oracle_bow.py)
and of ORBmatcher::SearchByBoW(KeyFrame*, Frame&). All comparisons are exact (integers; the BowVector doubles bit for bit)."""
import os

import numpy as np
import pytest

import fasttrack_b200 as ft
import oracle
from fasttrack_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dbow2_ref.npz")

FX, FY, CX, CY, BASE = 458.654, 457.296, 367.215, 248.375, 0.110074


def _same_transform(a, b, words=True):
    assert np.array_equal(a["node"], b["node"])
    if words:
        live = b["node"] >= 0
        assert np.array_equal(a["word"][live], b["word"][live])
    assert np.array_equal(a["bow_ids"], b["bow_ids"])
    assert np.array_equal(a["bow_vals"], b["bow_vals"])


def test_transform_matches_reference_golden(tmp_path):
    """the CUDA transform against outputs of the reference's own DBoW2 code (committed fixture), via both loaders"""
    g = np.load(GOLD)
    for ci in range(len([k for k in g.files if k.endswith("_cfg")])):
        p = "c%d_" % ci
        k, L, sc, wt, tn, lu, n_words = [int(x) for x in g[p + "cfg"]]
        path = str(tmp_path / ("voc%d.txt" % ci))
        synth.write_vocabulary_text(path, k, L, g[p + "parent"], g[p + "leaf"], g[p + "desc"], g[p + "weight"], scoring=sc,
                                    weighting=wt, trailing_newline=bool(tn))
        voc = ft.Vocabulary.load_text(path)
        assert (voc.k, voc.depth, voc.scoring, voc.weighting) == (k, L, sc, wt)
        r = voc.transform(g[p + "query"], lu)
        ref = dict(node=g[p + "node"], bow_ids=g[p + "bow_ids"], bow_vals=g[p + "bow_vals"])
        _same_transform(r, ref, words=False)
        if not tn:
            assert voc.n_words == n_words
            va = ft.Vocabulary.from_arrays(k, L, sc, wt, g[p + "parent"], g[p + "leaf"], g[p + "desc"], g[p + "weight"])
            _same_transform(va.transform(g[p + "query"], lu), ref, words=False)
            va.close()
        voc.close()


@pytest.mark.parametrize("k,L,sc,wt", [(10, 4, 0, 0), (3, 6, 1, 1), (20, 2, 5, 0), (7, 3, 0, 3), (33, 2, 3, 2)])
def test_transform_matches_oracle(k, L, sc, wt):
    parent, leaf, desc, weight = synth.make_vocabulary(k, L, seed=k * 10 + L, stop_fraction=0.05)
    vo = oracle.Vocabulary.from_arrays(k, L, sc, wt, parent, leaf, desc, weight)
    vg = ft.Vocabulary.from_arrays(k, L, sc, wt, parent, leaf, desc, weight)
    assert vg.n_words == vo.n_words and vg.n_nodes == vo.n_nodes
    rng = np.random.default_rng(k)
    for n in (1, 31, 1200, 5000):
        q = np.vstack([synth.vocabulary_like_descriptors(desc, n, seed=n, flips=int(rng.integers(0, 50))),
                       rng.integers(0, 256, (n // 3, 32), dtype=np.uint8)])
        for lu in (0, 2, L, L + 3):
            _same_transform(vg.transform(q, lu), vo.transform(q, lu))
    # ties: many identical descriptors, and descriptors equidistant from several children
    q = np.repeat(desc[rng.integers(0, len(desc), 40)], 25, axis=0)
    _same_transform(vg.transform(q, 1), vo.transform(q, 1))
    r0 = vg.transform(np.zeros((0, 32), np.uint8), 2)
    assert len(r0["node"]) == 0 and len(r0["bow_ids"]) == 0
    vg.close()


def test_full_size_vocabulary_descents():
    """ORBvoc's shape (k = 10, L = 6: 1,111,110 nodes, 35.5 MB of descriptors) -- descents against the oracle"""
    parent, leaf, desc, weight = synth.make_vocabulary_bfs(10, 6, seed=1, stop_fraction=0.01)
    assert len(parent) == 1111110
    vo = oracle.Vocabulary.from_arrays(10, 6, 0, 0, parent, leaf, desc, weight)
    vg = ft.Vocabulary.from_arrays(10, 6, 0, 0, parent, leaf, desc, weight)
    q = np.vstack([synth.vocabulary_like_descriptors(desc, 2400, seed=2, flips=25),
                   np.random.default_rng(3).integers(0, 256, (600, 32), dtype=np.uint8)])
    _same_transform(vg.transform(q, 4), vo.transform(q, 4))
    vg.close()


def _vocabulary_for(descs, seed=5):
    """a 10-ary depth-3 vocabulary grown around the frame's own descriptors so that groups are populated"""
    parent, leaf, desc, weight = synth.make_vocabulary(10, 3, seed=seed, stop_fraction=0.03)
    rng = np.random.default_rng(seed)
    # plant frame descriptors as leaf centres so that KeyFrame / Frame features share words
    leaves = np.nonzero(leaf)[0]
    take = rng.choice(len(descs), size=min(len(descs), len(leaves) // 2), replace=False)
    desc[leaves[: len(take)]] = descs[take]
    return parent, leaf, desc, weight


def _keyframe_from(frame_desc, frame_angle, rng, n_kf):
    pick = rng.integers(0, len(frame_desc), n_kf)
    bits = np.unpackbits(frame_desc[pick], axis=1)
    for i in range(n_kf):
        bits[i, rng.choice(256, size=int(rng.integers(0, 45)), replace=False)] ^= 1
    kf_desc = np.packbits(bits, axis=1)
    kf_angle = (frame_angle[pick] + np.where(rng.random(n_kf) < 0.75, rng.normal(20, 4, n_kf), rng.uniform(0, 360, n_kf))
                ).astype(np.float32) % np.float32(360)
    return kf_desc, kf_angle, (rng.random(n_kf) < 0.8).astype(np.uint8)


def test_compute_bow_and_search_by_bow_pinhole():
    sc = synth.StereoScene(seed=21)
    L, R = sc.pair()
    mbf = np.float32(FX * BASE)
    ctx = ft.Context(752, 480, cam1=[FX, FY, CX, CY], bf=float(mbf))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    gl = ctx.download(0)
    parent, leaf, desc, weight = _vocabulary_for(gl["desc"])
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, parent, leaf, desc, weight)
    vg = ft.Vocabulary.from_arrays(10, 3, 0, 0, parent, leaf, desc, weight)
    with pytest.raises(ft.FtError):
        ctx.bow_download()                     # ComputeBoW has not run on this frame
    ctx.compute_bow(vg, 2)
    fb = ctx.bow_download()
    fo = vo.transform(gl["desc"], 2)
    _same_transform(fb, fo)
    assert len(fb["node"]) == gl["n"]
    rng = np.random.default_rng(4)
    kf_desc, kf_angle, kf_has = _keyframe_from(gl["desc"], gl["kps"]["angle"], rng, 1100)
    kf_node = vg.transform(kf_desc, 2)["node"]
    assert np.array_equal(kf_node, vo.transform(kf_desc, 2)["node"])
    for nnratio, ori in ((0.7, True), (0.75, False), (0.9, True)):
        nm_o, m_o = oracle.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, gl["desc"], gl["kps"]["angle"], fo["node"], -1,
                                         nnratio, ori)
        nm_g, m_g = ctx.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, nnratio, ori)
        assert nm_o > 100
        assert nm_g == nm_o and np.array_equal(m_g, m_o)
    # empty KeyFrame, KeyFrame without map points
    assert ctx.search_by_bow(kf_desc[:0], kf_angle[:0], kf_node[:0], kf_has[:0])[0] == 0
    assert ctx.search_by_bow(kf_desc, kf_angle, kf_node, np.zeros_like(kf_has))[0] == 0
    # a new frame invalidates the BoW state
    ctx.extract_stereo(L, R)
    with pytest.raises(ft.FtError):
        ctx.search_by_bow(kf_desc, kf_angle, kf_node, kf_has)
    ctx.close(); vg.close()


def test_search_by_bow_fisheye_rig():
    """F.Nleft != -1: left and right keypoints are matched separately (the `|| true` branch of ORBmatcher.cc:451)"""
    T = synth.TUMVI                                            # the TUM-VI-shaped rig of tests/test_gpu_fisheye.py
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, _, _ = synth.tumvi_extrinsics()
    ctx = ft.Context(T["width"], T["height"], nfeatures=1000, camera_type=1, cam1=T["cam1"], cam2=T["cam2"],
                     lap_left=T["lap"], lap_right=T["lap"], bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    ctx.extract_stereo(L, R); ctx.stereo_match()
    gl, gr = ctx.download(0), ctx.download(1)
    f_desc = np.vstack([gl["desc"], gr["desc"]]); f_angle = np.concatenate([gl["kps"]["angle"], gr["kps"]["angle"]])
    parent, leaf, desc, weight = _vocabulary_for(f_desc, seed=9)
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, parent, leaf, desc, weight)
    vg = ft.Vocabulary.from_arrays(10, 3, 0, 0, parent, leaf, desc, weight)
    ctx.compute_bow(vg, 2)
    fb, fo = ctx.bow_download(), vo.transform(f_desc, 2)
    _same_transform(fb, fo)
    rng = np.random.default_rng(8)
    kf_desc, kf_angle, kf_has = _keyframe_from(f_desc, f_angle, rng, 1500)
    kf_node = vo.transform(kf_desc, 2)["node"]
    for ori in (True, False):
        nm_o, m_o = oracle.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, f_desc, f_angle, fo["node"], gl["n"], 0.7, ori)
        nm_g, m_g = ctx.search_by_bow(kf_desc, kf_angle, kf_node, kf_has, 0.7, ori)
        assert nm_o > 100 and (m_o[gl["n"]:] >= 0).any()
        assert nm_g == nm_o and np.array_equal(m_g, m_o)
    ctx.close(); vg.close()
