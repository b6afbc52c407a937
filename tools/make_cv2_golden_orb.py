"""Generates tests/golden/cv2_orb.npz with the in-container cv2: orientation and rBRIEF descriptors of OpenCV's own ORB
(cv::ORB with nlevels = 1, edgeThreshold 19, patchSize 31, WTA_K 2) on a synthetic image. ORB-SLAM's IC_Angle and
computeOrbDescriptor (reference src/ORBextractor.cc:39-108) are copies of the two OpenCV routines behind
ORB::detect (ICAngles) and ORB::compute (computeOrbDescriptors), so these vectors pin the oracle's restatement of both
against an independent implementation. cv::ORB blurs a sub-matrix of its pyramid buffer, which takes OpenCV's classic
sepFilter2D path (not the bit-exact fixed-point GaussianBlur that ORB-SLAM's clone() gets); the blurred image the
descriptors were sampled from is therefore stored alongside (it is reproduced by sepFilter2D with the 7x7 sigma-2
kernel, which the script asserts)."""
import os, sys
import numpy as np
import cv2

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from fasttrack_b200 import synth

img = synth.texture(240, 320, 77)
orb = cv2.ORB_create(nfeatures=1500, scaleFactor=1.2, nlevels=1, edgeThreshold=19, firstLevel=0, WTA_K=2, patchSize=31, fastThreshold=20)
kps = orb.detect(img)
xy = np.array([p.pt for p in kps], np.float32); ang = np.array([p.angle for p in kps], np.float32)
assert np.array_equal(xy, np.rint(xy)) and len(kps) > 400, len(kps)
kps2, desc = orb.compute(img, kps)
assert len(kps2) == len(kps)
kx = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
blurred = cv2.sepFilter2D(img, -1, kx, kx, borderType=cv2.BORDER_REFLECT_101)
dst = os.path.join(root, "tests", "golden", "cv2_orb.npz")
np.savez_compressed(dst, img=img, blurred=blurred, xy=xy, angle=ang, desc=desc, note=np.array("cv2 %s ORB nlevels=1" % cv2.__version__))
print("wrote", dst, os.path.getsize(dst), "bytes,", len(kps), "keypoints")
