"""Writes tests/golden/ref_orbextractor.npz: keypoints and descriptors produced by the REFERENCE's own
ORBextractor::operator() (CPU branch of src/ORBextractor.cc compiled where it lies into
oracle/_ref/libft_ref_orbextractor.so by `make -C oracle ref`) on seeded synthetic images. The images are regenerated
from the seeds by the tests (fasttrack_b200/synth.py), only the reference's outputs are stored.
Run in the build container (needs /root/reference); the fixture travels, the reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fasttrack_b200 import synth  # noqa: E402


def case_images():
    """name -> (image, nfeatures, nlevels, scale, iniTh, minTh, lap)"""
    out = {}
    L, R = synth.StereoScene(seed=2).pair()
    out["euroc_left"] = (L, 1200, 8, 1.2, 20, 7, (0, 0))
    out["euroc_right"] = (R, 1200, 8, 1.2, 20, 7, (0, 0))
    out["euroc_2000"] = (L, 2000, 8, 1.2, 20, 7, (0, 0))
    out["euroc_300"] = (R, 300, 8, 1.2, 20, 7, (0, 0))
    fl, fr = synth.fisheye_pair(seed=3)
    out["tumvi_left_lap"] = (fl, 1000, 8, 1.2, 20, 7, (0, 511))
    out["tumvi_right_partial"] = (fr, 1500, 8, 1.2, 20, 7, (150, 400))
    small = synth.StereoScene(seed=7, width=376, height=240, dmin=1.0, dmax=32.0, margin_x=64, margin_y=8).pair()[0]
    out["small_6_levels"] = (small, 400, 6, 1.2, 20, 7, (0, 0))
    rng = np.random.default_rng(12)
    out["noise"] = (rng.integers(0, 256, (480, 640), dtype=np.uint8), 1000, 8, 1.2, 20, 7, (0, 0))
    flat = np.full((480, 752), 120, np.uint8)
    flat[::16, ::24] += 9                      # weak corners only: the minThFAST fallback and sparse octrees
    flat[200:260, 300:420] = L[200:260, 300:420]
    out["low_texture"] = (flat, 1200, 8, 1.2, 20, 7, (0, 0))
    out["scale_1p5_4_levels"] = (L, 800, 4, 1.5, 25, 10, (0, 0))
    return out


def main():
    gold = {}
    for name, (img, nf, nl, sf, ini, mn, lap) in case_images().items():
        ref = oracle.RefExtractor(nf, sf, nl, ini, mn, img.shape[1], img.shape[0])
        mono, k, d = ref.extract(img, lap)
        gold[name + "_mono"] = np.int32(mono)
        gold[name + "_kps"] = k
        gold[name + "_desc"] = d
        print("%-22s n=%d mono=%d" % (name, len(k), mono))
    dst = os.path.join(ROOT, "tests", "golden", "ref_orbextractor.npz")
    np.savez_compressed(dst, **gold)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
