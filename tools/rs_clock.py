"""Debug: per-phase clock64 deltas of k_resolve for every CTA of the cluster (library must be built with
FT_EXTRA_NVCC_FLAGS=-DFT_RS_CLOCK). Slots: 0 start, 1 after the wait for k_gather + first loads, 2 prologue done,
3 first cluster barrier passed, per round r: 4+4r cleared, 5+4r scans + scatter done (thread 0), 6+4r barrier passed,
7+4r table copied; 40 final scatter, 41 barrier, 42 end."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench
E = synth.EUROC
mbf = np.float32(E["fx"] * E["baseline"])
ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
(L, R), = bench.make_frames(5, 1)
dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
ctx.extract_stereo(L, R)
g = ctx.download(0)
mp = bench.fast_mappoints(ft.keypoints_as_array(g["kps"]), g["desc"], ctx.scale_tables()["scale"], bench.M_POINTS, 1)
ctx.set_pose(np.eye(3), np.zeros(3))
ctx.upload_map_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"])
ctx.upload_holders(None, None)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(6):
    ctx.frame_enqueue_device(dL.data_ptr(), E["width"], dR.data_ptr(), E["width"])
    ctx.synchronize()
    if os.environ.get("RS_FLUSH", "1") == "1": flush.fill_(i); torch.cuda.synchronize()
    ctx.search_resident(bench.TH)
    ctx.synchronize()
print(ctx.stats())
buf = np.zeros((16, 64), np.int64)
ctx.L.ft_debug_rs_clock.argtypes = [C.c_void_p, C.c_void_p]
ctx.L.ft_debug_rs_clock(ctx.h, buf.ctypes.data)
names = {1: "wait+args", 2: "prologue", 3: "bar0", 40: "final-scatter", 41: "bar", 42: "holders"}
for r in range(8):
    names[4 + 4 * r] = "r%d:clear" % r; names[5 + 4 * r] = "scan"; names[6 + 4 * r] = "bar"; names[7 + 4 * r] = "copy"
for cta in (0, 1, 7, 15):
    t = buf[cta]
    slots = [s for s in range(64) if t[s] > 0]
    out = ["%s %d" % (names.get(b_, str(b_)), t[b_] - t[a_]) for a_, b_ in zip(slots[:-1], slots[1:])]
    print("cta", cta, "total", int(t[slots[-1]] - t[slots[0]]), "cycles |", " ".join(out))
