"""Generates tests/golden/cv2_remap.npz and tests/golden/cv2_undistort.npz with the in-container cv2:
  * cv::remap(src, M1, M2, INTER_LINEAR) as System::TrackStereo calls it (reference src/System.cc:279-280), with
    CV_32F maps incl. out-of-range, integer and half-pixel coordinates;
  * cv::undistortPoints(pts, K, distCoef, Mat(), K) as Frame::UndistortKeyPoints / ComputeImageBounds call it
    (reference src/Frame.cc:771-835) for the distorted pinhole cameras of the reference's example settings.
The oracle is pinned against these vectors in tests/test_oracle_golden.py (and live in tests/test_oracle_cv2_live.py)."""
import os
import numpy as np
import cv2

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(4321)

# ---- remap ----
img = rng.integers(0, 256, (97, 131), dtype=np.uint8)
mx = rng.uniform(-6, 137, (60, 80)).astype(np.float32); my = rng.uniform(-6, 103, (60, 80)).astype(np.float32)
mx[0, :10] = np.arange(10); my[0, :10] = 5                       # integer coordinates (the saturating table entry)
mx[1, :10] = np.arange(10) + 0.5; my[1, :10] = 7.5               # half pixels
mx[2, :6] = [-1, -0.5, 130, 130.5, 131, 1e6]; my[2, :6] = [0, 0, 96, 96.5, 97, -1e6]   # border cases
ref = cv2.remap(img, mx, my, cv2.INTER_LINEAR)
# a rectification-style map (smooth radial model) on a larger image
big = rng.integers(0, 256, (120, 188), dtype=np.uint8)
yy, xx = np.mgrid[0:120, 0:188].astype(np.float32)
xn, yn = (xx - 91.8) / 114.6, (yy - 62.1) / 114.3
r2 = xn * xn + yn * yn
f = 1 - 0.283 * r2 + 0.074 * r2 * r2
m1l = (xn * f * 114.6 + 91.8 + 1.3).astype(np.float32); m2l = (yn * f * 114.3 + 62.1 - 0.7).astype(np.float32)
dst = os.path.join(root, "tests", "golden", "cv2_remap.npz")
np.savez_compressed(dst, img=img, mx=mx, my=my, ref=ref, big=big, m1l=m1l, m2l=m2l, ref_big=cv2.remap(big, m1l, m2l, cv2.INTER_LINEAR),
                    note=np.array("cv2 %s remap INTER_LINEAR CV_32F maps" % cv2.__version__))
print("wrote", dst, os.path.getsize(dst), "bytes")

# ---- undistortPoints ----
cams = {
    # Examples/RGB-D/TUM1.yaml (k1 k2 p1 p2 k3), Examples/Monocular/EuRoC.yaml (k1 k2 p1 p2), a strong synthetic one
    "tum1": ([517.306408, 516.469215, 318.643040, 255.313989], [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], (640, 480)),
    "euroc_mono": ([458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], (752, 480)),
    "strong": ([190.978, 190.973, 254.932, 256.897], [-0.5, 0.3, 0.01, -0.02, 0.1], (512, 512)),
}
out = {"note": np.array("cv2 %s undistortPoints(pts, K, dist, None, K)" % cv2.__version__)}
for name, (K, dist, (w, h)) in cams.items():
    K = np.array(K, np.float32); dist = np.array(dist, np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    pts = (rng.random((3000, 2)) * [w + 60, h + 60] - [30, 30]).astype(np.float32)
    pts[:4] = [[0, 0], [w, 0], [0, h], [w, h]]                   # the corners ComputeImageBounds undistorts
    pts[4:8] = [[K[2], K[3]], [K[2] + 0.5, K[3]], [1e4, 1e4], [-1e4, 3]]   # principal point, far outliers (icdist < 0 path)
    out[name + "_K"], out[name + "_dist"], out[name + "_wh"] = K, dist, np.array([w, h], np.int32)
    out[name + "_pts"] = pts
    out[name + "_ref"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, dist, None, Km).reshape(-1, 2)
dst = os.path.join(root, "tests", "golden", "cv2_undistort.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")
