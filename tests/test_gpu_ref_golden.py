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
