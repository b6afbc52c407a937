"""Diagnostic: scans a synthetic sequence for descriptor rows that differ between the CUDA path and the oracle and prints
the libm vs correctly-rounded cos/sin of the keypoint angle (the one known cause: glibc sinf/cosf are not always
correctly rounded and a rotated sample can land exactly on a .5 rounding boundary). Needs a GPU."""
import sys, numpy as np, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, fasttrack_b200 as ft
from fasttrack_b200 import synth
E = synth.EUROC
libm = C.CDLL("libm.so.6"); libm.cosf.restype = C.c_float; libm.cosf.argtypes = [C.c_float]; libm.sinf.restype = C.c_float; libm.sinf.argtypes = [C.c_float]
sc = synth.StereoScene(seed=5)
mbf = np.float32(E["fx"] * E["baseline"])
ctx = ft.Context(E["width"], E["height"], nfeatures=1200, nlevels=8, cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
exL, exR = oracle.Extractor(), oracle.Extractor()
tot = 0
dump = []
for t in range(0, 2000, 50):
    L, R = sc.pair(pan=sc.sequence_pan(t), noise_seed=900000 + t)
    l, r = ctx.frame_construct(L, R)
    for name, g, ex, img in (("L", l, exL, L), ("R", r, exR, R)):
        _, k, d = ex.extract(img)
        bl = ex.desc_borderline()
        assert np.array_equal(ft.keypoints_as_array(g["kps"]), k)
        rows = np.nonzero((g["desc"] != d).any(axis=1))[0]
        tot += len(k)
        for i in rows:
            bits = int(np.unpackbits(g["desc"][i] ^ d[i]).sum())
            ang = np.float32(k[i, 3]); a32 = np.float32(ang * np.float32(np.pi / 180.0))
            fa = np.float32(np.float32(ang) * np.float32(0.017453292519943295))
            dump.append((t, name, i, g["desc"][i].copy(), d[i].copy(), k[i].copy()))
            print("t", t, name, "row", i, "bits", bits, "borderline", bl, "angle", float(ang), "oct", k[i, 5],
                  "glibc cos/sin", libm.cosf(float(fa)), libm.sinf(float(fa)), "cr", np.float32(np.cos(np.float64(fa))), np.float32(np.sin(np.float64(fa))))
print("keypoints checked", tot)

np.savez("gpurun_out/diag_desc.npz", gpu=np.array([x[3] for x in dump]), cpu=np.array([x[4] for x in dump]), kp=np.array([x[5] for x in dump]), row=np.array([x[2] for x in dump]))
