"""Differential tests oracle-vs-cv2 on fresh random inputs. Skipped where cv2 is not importable; the committed
vectors in tests/golden/ carry the same pin everywhere else."""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

import oracle
from fasttrack_b200 import synth


@pytest.mark.parametrize("shape", [(480, 752), (512, 512), (376, 1241)])
def test_pyramid_chain_and_blur(shape):
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    ex = oracle.Extractor()
    cur = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for l in range(1, 8):
        dw = oracle.cv_round(np.float32(w) * ex.inv_scale[l]); dh = oracle.cv_round(np.float32(h) * ex.inv_scale[l])
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize(cur, dw, dh), ref)
        assert np.array_equal(oracle.blur(ref), cv2.GaussianBlur(ref, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        cur = ref


def test_extractor_stages_against_cv2_driven_pipeline():
    """ComputePyramid + per-cell FAST with threshold fallback (ORBextractor.cc:1112-1203) rebuilt from cv2 calls."""
    img = synth.texture(240, 376, 5)
    ex = oracle.Extractor(400, 1.2, 6)
    ex.extract(img)
    for l in range(6):
        im = ex.level_image(l)
        w, h = ex.level_dims(l)
        if l > 0:
            assert np.array_equal(cv2.resize(ex.level_image(l - 1), (w, h), interpolation=cv2.INTER_LINEAR), im)
        mb, mxx, mxy = 16, w - 16, h - 16
        width, height = np.float32(mxx - mb), np.float32(mxy - mb)
        ncols, nrows = int(width / np.float32(35)), int(height / np.float32(35))
        wc, hc = math.ceil(width / ncols), math.ceil(height / nrows)
        out = []
        for i in range(nrows):
            iy, my = mb + i * hc, min(mb + i * hc + hc + 6, mxy)
            if iy >= mxy - 3:
                continue
            for j in range(ncols):
                ix, mx = mb + j * wc, min(mb + j * wc + wc + 6, mxx)
                if ix >= mxx - 6:
                    continue
                roi = np.ascontiguousarray(im[iy:my, ix:mx])
                k = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True).detect(roi)
                if not k:
                    k = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True).detect(roi)
                out += [(p.pt[0] + j * wc, p.pt[1] + i * hc, p.response) for p in k]
        assert np.array_equal(np.array(out, np.float32).reshape(-1, 3), ex.level_candidates(l))
        b = ex.level_image(l, True)
        if b is not None:
            assert np.array_equal(b, cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


def test_fast_random_rois():
    rng = np.random.default_rng(3)
    img = synth.texture(200, 300, 9)
    for _ in range(30):
        x0, y0 = int(rng.integers(0, 250)), int(rng.integers(0, 150))
        w, h = int(rng.integers(7, 50)), int(rng.integers(7, 50))
        roi = img[y0:y0 + h, x0:x0 + w]
        th = int(rng.choice([7, 20, 35]))
        k = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True).detect(np.ascontiguousarray(roi))
        ref = np.array([[p.pt[0], p.pt[1], p.response] for p in k], np.float32).reshape(-1, 3)
        assert np.array_equal(oracle.fast(roi, th), ref)


def test_remap_live():
    rng = np.random.default_rng(5)
    img = synth.texture(480, 752, 3)
    yy, xx = np.mgrid[0:480, 0:752].astype(np.float32)
    xn, yn = (xx - 367.2) / 458.6, (yy - 248.4) / 457.3
    r2 = xn * xn + yn * yn
    f = 1 - 0.283 * r2 + 0.074 * r2 * r2
    mx = (xn * f * 458.6 + 367.2 + 1.3).astype(np.float32); my = (yn * f * 457.3 + 248.4 - 0.7).astype(np.float32)
    assert np.array_equal(oracle.remap(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    mx2 = rng.uniform(-8, 760, (100, 120)).astype(np.float32); my2 = rng.uniform(-8, 488, (100, 120)).astype(np.float32)
    assert np.array_equal(oracle.remap(img, mx2, my2), cv2.remap(img, mx2, my2, cv2.INTER_LINEAR))


def test_undistort_points_live():
    rng = np.random.default_rng(8)
    for K, dist in (([517.3, 516.5, 318.6, 255.3], [0.2624, -0.9531, -0.0054, 0.0026, 1.1633]),
                    ([458.654, 457.296, 367.215, 248.375], [-0.2834, 0.0740, 0.00019, 1.76e-05]),
                    ([300.0, 310.0, 320.0, 240.0], [0.1, 0.0, 0.0, 0.0])):
        K = np.array(K, np.float32); dist = np.array(dist, np.float32)
        Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
        pts = (rng.random((20000, 2)) * [900, 700] - [80, 80]).astype(np.float32)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, dist, None, Km).reshape(-1, 2)
        assert np.array_equal(oracle.undistort_points(pts, K, dist).view(np.uint32), ref.view(np.uint32))


def test_input_resize_shapes_live():
    """cv::resize(im, newImSize) of System::TrackStereo (System.cc:282-285): exact 2x (OpenCV switches to its INTER_AREA
    shortcut, which yields the same bytes), 3x, non-integer down- and up-scaling"""
    rng = np.random.default_rng(12)
    for (sh, sw), (dh, dw) in [((960, 1504), (480, 752)), ((720, 1280), (480, 752)), ((400, 640), (480, 752)),
                               ((1440, 2256), (480, 752)), ((1024, 1024), (512, 512))]:
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        assert np.array_equal(oracle.resize(img, dw, dh), cv2.resize(img, (dw, dh)))


def test_orientation_and_descriptor_against_opencv_orb_live():
    """oracle IC_Angle / computeOrbDescriptor vs cv2.ORB (nlevels=1) on two images; the descriptor is sampled from the
    blur cv::ORB applies internally (classic sepFilter2D path on a sub-matrix), reproduced here with sepFilter2D"""
    ex = oracle.Extractor(1200, 1.2, 8)
    kx = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    for seed, shape in ((5, (480, 752)), (6, (300, 401))):
        img = synth.texture(shape[0], shape[1], seed)
        orb = cv2.ORB_create(nfeatures=3000, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2, patchSize=31,
                             fastThreshold=20)
        kps = orb.detect(img)
        xy = np.array([p.pt for p in kps], np.float32); ang = np.array([p.angle for p in kps], np.float32)
        assert len(kps) > 500 and np.array_equal(xy, np.rint(xy))
        assert np.array_equal(ex.ic_angles(img, xy).view(np.uint32), ang.view(np.uint32))
        kps2, desc = orb.compute(img, kps)
        blurred = cv2.sepFilter2D(img, -1, kx, kx, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(oracle.orb_descriptors(blurred, xy, ang), desc)


def test_kannala_brandt_model_against_cv2_fisheye():
    """KannalaBrandt8::project / unproject (KannalaBrandt8.cpp:67-84, 116-143) implement the equidistant model of
    cv2.fisheye (theta_d = theta (1 + k1 theta^2 + ... + k4 theta^8)); the reference evaluates it in float with atan2 / cos /
    sin, cv2 in double, so this is a tolerance check of the model and its Newton inverse, not a bit-exact pin."""
    cam = np.array(synth.TUMVI["cam1"], np.float32)
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float64)
    D = cam[4:8].astype(np.float64)
    rng = np.random.default_rng(4)
    P = np.stack([rng.uniform(-3, 3, 4000), rng.uniform(-3, 3, 4000), rng.uniform(0.5, 6, 4000)], 1).astype(np.float32)
    uv = oracle.cam_project(1, cam, P)
    ref, _ = cv2.fisheye.projectPoints(P.astype(np.float64).reshape(1, -1, 3), np.zeros(3), np.zeros(3), K, D)
    assert np.abs(uv - ref.reshape(-1, 2)).max() < 2e-3            # pixels (float32 evaluation)
    inside = (uv[:, 0] > 0) & (uv[:, 0] < 512) & (uv[:, 1] > 0) & (uv[:, 1] < 512)
    ray = oracle.kb8_unproject(cam, uv[inside])
    und = cv2.fisheye.undistortPoints(uv[inside].astype(np.float64).reshape(1, -1, 2), K, D).reshape(-1, 2)
    assert inside.sum() > 500
    assert np.allclose(ray[:, :2], und, rtol=2e-4, atol=2e-4) and np.all(ray[:, 2] == 1)
    # and the inverse really inverts: unproject(project(P)) is P / z
    assert np.allclose(ray[:, :2], (P[inside, :2] / P[inside, 2:3]), rtol=2e-4, atol=2e-4)
