"""Unconditional comparison of a projection search (CUDA path) with the oracle.

isInFrustum evaluates float expressions whose last-ulp behaviour (Eigen expression order, device vs host libm in the KB8
model) can flip a decision that sits within 1e-5 of a boundary; the oracle flags those map points (track_i[:, 4] != 0).
A flipped map point changes which keypoints later map points may claim, so the tables cannot be compared blindly -- but
skipping the comparison whenever anything flipped (what these tests used to do) hides real bugs. The rule here:

  * frustum outcomes (in view left / right, predicted levels) may differ ONLY on flagged map points, at most `max_flips`;
  * no flip: match count and both holder tables are equal, exactly;
  * flips: the search is loop-carried in map-point order, so everything in front of the first flipped map point is
    untouched by it: both sides are re-run on that prefix and must be equal, exactly.
"""
import numpy as np


def assert_search_matches(gpu, orc, gi, ti, rerun_prefix, max_flips=3, cols=(0, 2)):
    """gpu / orc: (nmatches, holder, holder_obs); gi / ti: integer track scratch [M, >=4] (+ the oracle's borderline flag in
    ti[:, 4]); rerun_prefix(k) -> (gpu, orc) for the first k map points; cols: track columns to compare (0 / 2 = in view /
    level of the left camera, 1 / 3 the right camera's on fisheye rigs)."""
    cols = list(cols)
    differ = np.any(gi[:, cols] != ti[:, cols], axis=1)
    flips = np.nonzero(differ)[0]
    flagged = ti[:, 4] != 0
    assert np.all(flagged[flips]), "frustum decision differs on a map point that is not near a boundary: %s" % flips[~flagged[flips]][:5]
    assert len(flips) <= max_flips, "%d borderline frustum decisions flipped (bound %d)" % (len(flips), max_flips)
    if len(flips) == 0:
        assert gpu[0] == orc[0], "nmatches %d vs %d" % (gpu[0], orc[0])
        assert np.array_equal(gpu[1], orc[1]), "holder tables differ in %d slots" % int((np.asarray(gpu[1]) != np.asarray(orc[1])).sum())
        assert np.array_equal(gpu[2], orc[2])
        return 0
    k = int(flips[0])
    if k > 0:
        g2, o2 = rerun_prefix(k)
        assert g2[0] == o2[0] and np.array_equal(g2[1], o2[1]) and np.array_equal(g2[2], o2[2]), \
            "searches differ on the %d map points in front of the first borderline flip" % k
    return len(flips)
