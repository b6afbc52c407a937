"""Generates tests/golden/cv2_primitives.npz: outputs of the OpenCV primitives the reference calls
(cv::resize INTER_LINEAR, GaussianBlur 7x7 s2, FAST-9/16 + NMS, fastAtan2, BFMatcher knn k=2), produced by the
in-container cv2 (run where cv2 is importable). The oracle is pinned against these vectors in
tests/test_oracle_golden.py, so the pin also holds on machines without cv2."""
import os, sys
import numpy as np
import cv2

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasttrack_b200 import synth

out = {}
out["cv2_version"] = np.array(cv2.__version__)
rng = np.random.default_rng(1234)
img_tex = synth.texture(150, 200, 11)
img_noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
out["img_tex"], out["img_noise"] = img_tex, img_noise
for name, img in (("tex", img_tex), ("noise", img_noise)):
    h, w = img.shape
    cur = img
    for l, s in enumerate((1.2, 1.2, 1.2)):
        dw, dh = int(round(cur.shape[1] / s)), int(round(cur.shape[0] / s))
        cur = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        out["resize_%s_%d" % (name, l)] = cur
    out["blur_%s" % name] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        k = det.detect(img)
        out["fast_%s_%d" % (name, th)] = np.array([[p.pt[0], p.pt[1], p.response] for p in k], np.float32).reshape(-1, 3)
        roi = np.ascontiguousarray(img[10:52, 20:64])
        k = det.detect(roi)
        out["fastroi_%s_%d" % (name, th)] = np.array([[p.pt[0], p.pt[1], p.response] for p in k], np.float32).reshape(-1, 3)
# odd shapes for resize (upper clamp paths)
odd = rng.integers(0, 256, (33, 47), dtype=np.uint8)
out["img_odd"] = odd
out["resize_odd"] = cv2.resize(odd, (39, 28), interpolation=cv2.INTER_LINEAR)
yx = rng.integers(-2_000_000, 2_000_000, (4000, 2)).astype(np.float32)
yx[:9] = [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (5, 5), (-5, 5), (5, -5), (-5, -5)]
out["atan_yx"] = yx
out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
q = rng.integers(0, 256, (120, 32), dtype=np.uint8)
t = (rng.integers(0, 4, (200, 32), dtype=np.uint8) * 85).astype(np.uint8)
q[:30] = t[:30]
m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
out["knn_q"], out["knn_t"] = q, t
out["knn_idx"] = np.array([[x.trainIdx for x in mm] for mm in m], np.int32)
out["knn_dist"] = np.array([[int(x.distance) for x in mm] for mm in m], np.int32)
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_primitives.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")
