"""CPU: the oracle's bag-of-words restatement (oracle/ft_oracle_bow.cpp) against the REFERENCE's own DBoW2.

* test_transform_matches_reference_golden: committed outputs of the reference's Thirdparty/DBoW2 code
  (tests/golden/dbow2_ref.npz, made by tools/make_dbow2_golden.py through oracle/_ref/libft_ref_dbow2.so).
* test_transform_matches_reference_live: the same comparison against the compiled reference code itself, on fresh
  inputs, whenever oracle/_ref is built or can be built (skipped on a machine with neither the .so nor /root/reference).
* test_search_by_bow_*: properties of the SearchByBoW restatement (its comparison with the reference's own function text is in
  tests/test_oracle_ref_frame.py).
"""
import os

import numpy as np
import pytest

import oracle
from fasttrack_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dbow2_ref.npz")


def _check(vo, q, lu, node, bow_ids, bow_vals, order=None):
    r = vo.transform(q, lu)
    assert np.array_equal(r["node"], node)
    assert np.array_equal(r["bow_ids"], bow_ids)
    assert np.array_equal(r["bow_vals"], bow_vals)      # doubles, bit for bit
    if order is not None:   # FeatureVector map order: nodes ascending, feature indices ascending inside a node
        keep = np.nonzero(r["node"] >= 0)[0]
        mine = keep[np.argsort(r["node"][keep], kind="stable")]
        assert np.array_equal(mine, order)


def test_transform_matches_reference_golden(tmp_path):
    g = np.load(GOLD)
    n_cases = len([k for k in g.files if k.endswith("_cfg")])
    assert n_cases >= 6
    for ci in range(n_cases):
        p = "c%d_" % ci
        k, L, sc, wt, tn, lu, n_words = [int(x) for x in g[p + "cfg"]]
        path = str(tmp_path / ("voc%d.txt" % ci))
        synth.write_vocabulary_text(path, k, L, g[p + "parent"], g[p + "leaf"], g[p + "desc"], g[p + "weight"], scoring=sc,
                                    weighting=wt, trailing_newline=bool(tn))
        vo = oracle.Vocabulary.load_text(path)
        assert vo.n_words == n_words and (vo.k, vo.L, vo.scoring, vo.weighting) == (k, L, sc, wt)
        _check(vo, g[p + "query"], lu, g[p + "node"], g[p + "bow_ids"], g[p + "bow_vals"], g[p + "featvec_order"])
        if not tn:   # without the phantom node the array constructor describes the same tree
            va = oracle.Vocabulary.from_arrays(k, L, sc, wt, g[p + "parent"], g[p + "leaf"], g[p + "desc"], g[p + "weight"])
            _check(va, g[p + "query"], lu, g[p + "node"], g[p + "bow_ids"], g[p + "bow_vals"])


def test_transform_matches_reference_live(tmp_path):
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref/libft_ref_dbow2.so not built and no reference tree here")
    rng = np.random.default_rng(5)
    for trial in range(6):
        k, L = int(rng.integers(2, 11)), int(rng.integers(2, 5))
        sc, wt, tn = int(rng.integers(0, 6)), int(rng.integers(0, 4)), bool(rng.integers(0, 2))
        parent, leaf, desc, weight = synth.make_vocabulary(k, L, seed=40 + trial, stop_fraction=0.1)
        path = str(tmp_path / ("v%d.txt" % trial))
        synth.write_vocabulary_text(path, k, L, parent, leaf, desc, weight, scoring=sc, weighting=wt, trailing_newline=tn)
        vo, vr = oracle.Vocabulary.load_text(path), oracle.RefVocabulary(path)
        assert vo.n_words == vr.n_words
        q = np.vstack([synth.vocabulary_like_descriptors(desc, 900, seed=trial, flips=int(rng.integers(0, 60))),
                       rng.integers(0, 256, (300, 32), dtype=np.uint8)])
        for lu in (0, 1, L - 1, L, L + 2):
            r = vr.transform(q, lu)
            _check(vo, q, lu, r["node"], r["bow_ids"], r["bow_vals"], r["featvec_order"])


def test_load_text_round_trip_of_arrays(tmp_path):
    parent, leaf, desc, weight = synth.make_vocabulary(5, 3, seed=3)
    path = str(tmp_path / "v.txt")
    synth.write_vocabulary_text(path, 5, 3, parent, leaf, desc, weight, trailing_newline=False)
    vo = oracle.Vocabulary.load_text(path)
    p2, l2, d2, w2 = vo.arrays()
    assert np.array_equal(p2, parent) and np.array_equal(l2, leaf) and np.array_equal(d2, desc) and np.array_equal(w2, weight)
    with open(path, "a") as f:
        f.write("\n")
    vp = oracle.Vocabulary.load_text(path)      # the phantom node of the reference's eof loop
    assert vp.n_nodes == vo.n_nodes + 1 and vp.n_words == vo.n_words + 1
    p3, l3, d3, w3 = vp.arrays()
    assert p3[-1] == parent[-1] and l3[-1] == 1 and w3[-1] == 0.0 and not d3[-1].any()


def _bow_case(seed, n_kf=900, n_f=1000, stereo=False):
    """a KeyFrame / Frame pair that share most descriptors (with bit noise) over a small vocabulary"""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = synth.make_vocabulary(6, 3, seed=seed)
    vo = oracle.Vocabulary.from_arrays(6, 3, 0, 0, parent, leaf, desc, weight)
    f_desc = synth.vocabulary_like_descriptors(desc, n_f, seed=seed + 1, flips=10)
    pick = rng.integers(0, n_f, n_kf)
    bits = np.unpackbits(f_desc[pick], axis=1)
    for i in range(n_kf):
        bits[i, rng.choice(256, size=int(rng.integers(0, 40)), replace=False)] ^= 1
    kf_desc = np.packbits(bits, axis=1)
    f_angle = rng.uniform(0, 360, n_f).astype(np.float32)
    kf_angle = (f_angle[pick] + np.where(rng.random(n_kf) < 0.8, rng.normal(12, 3, n_kf), rng.uniform(0, 360, n_kf))
                ).astype(np.float32) % np.float32(360)
    kf_has = (rng.random(n_kf) < 0.7).astype(np.uint8)
    f_node = vo.transform(f_desc, 2)["node"]; kf_node = vo.transform(kf_desc, 2)["node"]
    return dict(kf_desc=kf_desc, kf_angle=kf_angle, kf_node=kf_node, kf_has_mp=kf_has, f_desc=f_desc, f_angle=f_angle,
                f_node=f_node, f_nleft=(n_f * 3 // 5 if stereo else -1))


@pytest.mark.parametrize("stereo", [False, True])
def test_search_by_bow_properties(stereo):
    c = _bow_case(11, stereo=stereo)
    nm, match = oracle.search_by_bow(nnratio=0.7, check_ori=True, **c)
    assert nm == int((match >= 0).sum()) and nm > 50
    m = np.nonzero(match >= 0)[0]
    # a match joins features of the same vocabulary node, within TH_LOW, whose KeyFrame side holds a map point
    assert np.array_equal(c["f_node"][m], c["kf_node"][match[m]])
    assert c["kf_has_mp"][match[m]].all()
    d = np.unpackbits(c["f_desc"][m] ^ c["kf_desc"][match[m]], axis=1).sum(1)
    assert d.max() <= 50
    # without the orientation check nothing is withdrawn afterwards: a superset
    nm2, match2 = oracle.search_by_bow(nnratio=0.7, check_ori=False, **c)
    assert nm2 >= nm and np.array_equal(match2[m], match[m])
    if not stereo:   # monocular frames: a KeyFrame feature gives its map point to at most one frame keypoint
        assert len(np.unique(match[m])) == len(m)
    # nobody holds a map point: nothing to match
    c0 = dict(c, kf_has_mp=np.zeros_like(c["kf_has_mp"]))
    assert oracle.search_by_bow(**c0)[0] == 0
