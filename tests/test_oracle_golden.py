"""Pins the CPU oracle's primitives against vectors produced by OpenCV (tests/golden/cv2_primitives.npz,
made by tools/make_cv2_golden.py with cv2 4.13.0). These are the un-vendored third-party calls on the
reference's path: cv::resize (ORBextractor.cc:1508), GaussianBlur (:1457), cv::FAST (:1157,1176),
fastAtan2 (:65), BFMatcher::knnMatch (Frame.cc:1249)."""
import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("name", ["tex", "noise"])
def test_resize_chain_matches_cv2(cv2_golden, name):
    cur = cv2_golden["img_" + name]
    for l in range(3):
        ref = cv2_golden["resize_%s_%d" % (name, l)]
        mine = oracle.resize(cur, ref.shape[1], ref.shape[0])
        assert np.array_equal(mine, ref), "level %d" % l
        cur = ref


def test_resize_odd_shape(cv2_golden):
    ref = cv2_golden["resize_odd"]
    assert np.array_equal(oracle.resize(cv2_golden["img_odd"], ref.shape[1], ref.shape[0]), ref)


@pytest.mark.parametrize("name", ["tex", "noise"])
def test_blur_matches_cv2(cv2_golden, name):
    assert np.array_equal(oracle.blur(cv2_golden["img_" + name]), cv2_golden["blur_" + name])


@pytest.mark.parametrize("name", ["tex", "noise"])
@pytest.mark.parametrize("th", [20, 7])
def test_fast_matches_cv2(cv2_golden, name, th):
    img = cv2_golden["img_" + name]
    assert np.array_equal(oracle.fast(img, th), cv2_golden["fast_%s_%d" % (name, th)])
    roi = img[10:52, 20:64]   # strided view, like the reference's rowRange/colRange cells
    assert np.array_equal(oracle.fast(roi, th), cv2_golden["fastroi_%s_%d" % (name, th)])


def test_fast_atan2_matches_cv2(cv2_golden):
    yx, ref = cv2_golden["atan_yx"], cv2_golden["atan_deg"]
    mine = np.array([oracle.fast_atan2(y, x) for y, x in yx], np.float32)
    assert np.array_equal(mine, ref)


def test_knn_matches_cv2(cv2_golden):
    idx, dist = oracle.knn2(cv2_golden["knn_q"], cv2_golden["knn_t"])
    assert np.array_equal(idx, cv2_golden["knn_idx"])
    assert np.array_equal(dist, cv2_golden["knn_dist"])


def test_cv_round_half_even():
    for v, r in [(0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (-2.5, -2), (3.4999, 3), (3.5001, 4)]:
        assert oracle.cv_round(v) == r


def test_remap_matches_cv2():
    """cv::remap INTER_LINEAR, CV_32F maps, constant-0 border (System.cc:279-280), incl. out-of-range and integer coordinates"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_remap.npz"))
    assert np.array_equal(oracle.remap(g["img"], g["mx"], g["my"]), g["ref"])
    assert np.array_equal(oracle.remap(g["big"], g["m1l"], g["m2l"]), g["ref_big"])


def test_undistort_points_matches_cv2():
    """cv::undistortPoints(pts, K, distCoef, Mat(), K) of Frame::UndistortKeyPoints / ComputeImageBounds
    (Frame.cc:771-835): bit-exact incl. the image corners, far outliers and the icdist < 0 bail-out"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_undistort.npz"))
    for name in ("tum1", "euroc_mono", "strong"):
        K, dist, pts, ref = g[name + "_K"], g[name + "_dist"], g[name + "_pts"], g[name + "_ref"]
        got = oracle.undistort_points(pts, K, dist)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), name
        w, h = g[name + "_wh"]
        b = oracle.image_bounds(int(w), int(h), K, dist)
        assert b[0] == min(ref[0, 0], ref[2, 0]) and b[1] == max(ref[1, 0], ref[3, 0])
        assert b[2] == min(ref[0, 1], ref[1, 1]) and b[3] == max(ref[2, 1], ref[3, 1])
    assert np.array_equal(oracle.image_bounds(640, 480, g["tum1_K"], np.zeros(5, np.float32)), [0, 640, 0, 480])


def test_stereo_from_rgbd_restatement():
    """Frame::ComputeStereoFromRGBD (Frame.cc:1065-1086): float coordinates truncate, depth <= 0 leaves -1, uRight
    comes from the undistorted x; depth=None is the monocular frame"""
    rng = np.random.default_rng(2)
    depth = rng.uniform(-0.5, 6.0, (60, 80)).astype(np.float32)
    xy = np.stack([rng.uniform(0, 79.99, 500), rng.uniform(0, 59.99, 500)], 1).astype(np.float32)
    unx = (xy[:, 0] + rng.uniform(-3, 3, 500)).astype(np.float32)
    mbf = np.float32(40.0)
    ur, dp = oracle.stereo_from_rgbd(xy, unx, depth, mbf)
    d = depth[xy[:, 1].astype(np.int32), xy[:, 0].astype(np.int32)]
    ok = d > 0
    assert np.array_equal(dp, np.where(ok, d, np.float32(-1))) and ok.sum() > 300 and (~ok).sum() > 10
    exp = np.where(ok, unx - mbf / np.where(ok, d, np.float32(1)), np.float32(-1)).astype(np.float32)
    assert np.array_equal(ur, exp)
    ur0, dp0 = oracle.stereo_from_rgbd(xy, unx, None, mbf)
    assert np.all(ur0 == -1) and np.all(dp0 == -1)


def test_orientation_and_descriptor_match_opencv_orb():
    """IC_Angle and computeOrbDescriptor (ORBextractor.cc:39-108) against OpenCV's own ORB (ICAngles / computeOrbDescriptors,
    from which ORB-SLAM copied them): angles bit-exact on the raw image, descriptors bit-exact on the blurred image that
    cv::ORB sampled (tools/make_cv2_golden_orb.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_orb.npz"))
    ex = oracle.Extractor(1200, 1.2, 8)
    assert len(g["xy"]) > 400
    assert np.array_equal(ex.ic_angles(g["img"], g["xy"]).view(np.uint32), g["angle"].view(np.uint32))
    assert np.array_equal(oracle.orb_descriptors(g["blurred"], g["xy"], g["angle"]), g["desc"])
