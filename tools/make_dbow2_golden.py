"""Writes tests/golden/dbow2_ref.npz: outputs of the REFERENCE's own DBoW2 (oracle/_ref/libft_ref_dbow2.so, the
reference's Thirdparty/DBoW2 sources compiled where they lie by `make -C oracle ref`) on synthetic vocabularies.
Run in the build container (needs /root/reference); the fixture travels, the reference does not."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fasttrack_b200 import synth  # noqa: E402

CASES = [  # k, L, scoring, weighting, trailing newline, levelsup
    (10, 3, 0, 0, True, 1), (10, 3, 0, 0, False, 2), (6, 4, 1, 0, True, 2), (5, 4, 5, 1, True, 4), (8, 3, 0, 2, False, 0),
    (9, 3, 2, 3, True, 3),
]


def main():
    out = {}
    tmp = tempfile.mkdtemp()
    for ci, (k, L, sc, wt, tn, lu) in enumerate(CASES):
        parent, leaf, desc, weight = synth.make_vocabulary(k, L, seed=100 + ci)
        path = os.path.join(tmp, "voc%d.txt" % ci)
        synth.write_vocabulary_text(path, k, L, parent, leaf, desc, weight, scoring=sc, weighting=wt, trailing_newline=tn)
        ref = oracle.RefVocabulary(path)
        q = np.vstack([synth.vocabulary_like_descriptors(desc, 700, seed=200 + ci, flips=28),
                       np.random.default_rng(300 + ci).integers(0, 256, (250, 32), dtype=np.uint8),
                       np.zeros((2, 32), np.uint8)])
        r = ref.transform(q, lu)
        p = "c%d_" % ci
        out[p + "cfg"] = np.array([k, L, sc, wt, int(tn), lu, ref.n_words], np.int32)
        out[p + "parent"], out[p + "leaf"], out[p + "desc"], out[p + "weight"] = parent, leaf, desc, weight
        out[p + "query"] = q
        out[p + "node"], out[p + "bow_ids"], out[p + "bow_vals"] = r["node"], r["bow_ids"], r["bow_vals"]
        out[p + "featvec_order"] = r["featvec_order"]
    dst = os.path.join(ROOT, "tests", "golden", "dbow2_ref.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
