"""The oracle's extractor restatement against the REFERENCE's own ORBextractor::operator() (CPU branch of
src/ORBextractor.cc, compiled where it lies against the OpenCV stand-in whose image primitives are the cv2-pinned ones).

* golden: committed outputs of the reference code (tests/golden/ref_orbextractor.npz, tools/make_ref_extractor_golden.py)
* live:   the compiled reference code itself on fresh images (skipped where oracle/_ref cannot be built)
Everything is compared exactly: keypoint coordinates, size, angle, response, octave, order, monoIndex, descriptors."""
import os
import sys

import numpy as np
import pytest

import oracle
from fasttrack_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_ref_extractor_golden import case_images  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "ref_orbextractor.npz")


def test_oracle_extractor_matches_reference_golden():
    g = np.load(GOLD)
    cases = case_images()
    assert len(cases) >= 10
    for name, (img, nf, nl, sf, ini, mn, lap) in cases.items():
        ex = oracle.Extractor(nf, sf, nl, ini, mn)
        mono, k, d = ex.extract(img, lap=lap)
        assert mono == int(g[name + "_mono"]), name
        assert np.array_equal(k, g[name + "_kps"]), name
        assert np.array_equal(d, g[name + "_desc"]), name


def test_oracle_extractor_matches_reference_live():
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref/libft_ref_orbextractor.so not built and no reference tree here")
    rng = np.random.default_rng(77)
    for trial in range(5):
        w, h = int(rng.integers(320, 900)), int(rng.integers(240, 600))
        nf, nl = int(rng.integers(200, 2500)), int(rng.integers(3, 9))
        sf = float(rng.choice([1.2, 1.3, 1.5]))
        img = synth.StereoScene(seed=100 + trial, width=w, height=h, margin_x=64, margin_y=8).pair()[trial % 2]
        lap = (0, 0) if trial % 2 else (int(w * 0.3), int(w * 0.8))
        ref = oracle.RefExtractor(nf, sf, nl, 20, 7, w, h)
        ex = oracle.Extractor(nf, sf, nl, 20, 7)
        mr, kr, dr = ref.extract(img, lap)
        mo, ko, do = ex.extract(img, lap=lap)
        assert mr == mo and np.array_equal(kr, ko) and np.array_equal(dr, do), (trial, w, h, nf, nl, sf)
        t = ref.scale_tables()
        assert np.array_equal(t["scale"], ex.scale) and np.array_equal(t["inv_scale"], ex.inv_scale)
        assert np.array_equal(t["sigma2"], ex.sigma2) and np.array_equal(t["inv_sigma2"], ex.inv_sigma2)
        for lvl in range(nl):   # mvImagePyramid
            assert np.array_equal(ref.level_image(lvl), ex.level_image(lvl))
