"""Debug: per-phase clock64 deltas of the octree kernel (library must be built with FT_EXTRA_NVCC_FLAGS=-DFT_OCT_CLOCK).
Slots: 0 start, 1 cell scan, 2 copy + keys + bin histogram, 3 bin scan + scatter, 4 rank, 5 neighbour stats + decision,
6 list construction, 7.. careful passes, 20 careful done, 21 end."""
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
names = {0: "start", 1: "cellscan|dense-load", 2: "copy+key|pyramid", 3: "binscan+scatter", 4: "rank", 5: "stats", 6: "list", 20: "careful_done", 21: "end"}
for lvl in range(8):
    t = buf[0, lvl]
    slots = [s for s in range(30) if t[s] > 0]
    prev = None
    out = []
    for s in slots:
        if prev is not None:
            out.append("%s:%d" % (names.get(s, "careful%d" % (s - 7)), t[s] - t[prev]))
        prev = s
    sub = [s for s in range(30, 36) if t[s] > 0]
    if len(sub) > 1:
        out.append("| first careful pass (sort loop, rank, splits, scan+P, children): " + " ".join(str(int(t[sub[i + 1]] - t[sub[i]])) for i in range(len(sub) - 1)))
    if t[40]: out.append("| FELL BACK to the general path")
    print("level", lvl, "C", cand[lvl], "K", kp[lvl], "total", int(t[slots[-1]] - t[slots[0]]), " ".join(out))
