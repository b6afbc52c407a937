"""Debug: per-phase clock64 deltas of the octree kernel (library must be built with FT_EXTRA_NVCC_FLAGS=-DFT_OCT_CLOCK)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench
E = synth.EUROC
ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=50.0)
(L, R), = bench.make_frames(5, 1)
for _ in range(3):
    ctx.extract_stereo(L, R); ctx.synchronize()
buf = np.zeros((2, 16, 64), np.int64)
ctx.L.ft_debug_oct_clock.argtypes = [C.c_void_p, C.c_void_p]
ctx.L.ft_debug_oct_clock(ctx.h, buf.ctypes.data)
cand, kp = ctx.level_counts(0)
for lvl in range(8):
    t = buf[0, lvl]
    ticks = t[:40]; n = int((ticks > 0).sum())
    d = np.diff(ticks[:n])
    modes = t[40:40 + max(n - 5, 0)]
    sub = t[20:27]; print("   careful sub-phases", np.diff(sub).tolist(), "m", int(t[63]))
    print("   last normal pass sub-phases (cand loop, node loop, scan, placement)", np.diff(t[27:32]).tolist())
    print("level", lvl, "C", cand[lvl], "K", kp[lvl], "total cycles", int(ticks[n - 1] - ticks[0]), "phases", d.tolist(), "mode*1e5+n", modes.tolist())
