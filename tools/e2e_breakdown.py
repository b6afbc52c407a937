"""Wall-clock breakdown of the end-to-end (host buffers) call sequence, per call."""
import os, sys, time, ctypes as C
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
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hL, hR = pin(L), pin(R)
ctx.extract_stereo(L, R); g = ctx.download(0)
mp = bench.fast_mappoints(ft.keypoints_as_array(g["kps"]), g["desc"], ctx.scale_tables()["scale"], bench.M_POINTS, 1)
ctx.set_pose(np.eye(3), np.zeros(3))
cap = ctx.cap
ok = [torch.empty(cap * 24, dtype=torch.uint8).pin_memory() for _ in range(2)]
od = [torch.empty(cap * 32, dtype=torch.uint8).pin_memory() for _ in range(2)]
ur = torch.empty(cap, dtype=torch.float32).pin_memory(); dp = torch.empty(cap, dtype=torch.float32).pin_memory()
c4 = torch.zeros(4, dtype=torch.int32).pin_memory()
L_ = ctx.L
M = bench.M_POINTS
stg = ctx.map_point_staging(M, cap)
T = {k: [] for k in ("frame_construct", "marshal", "search_staged", "upload_only", "frame_enqueue+sync", "search_resident+sync")}
dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
for it in range(60):
    t0 = time.perf_counter()
    ctx._ck(L_.ft_frame_construct(ctx.h, hL.data_ptr(), E["width"], hR.data_ptr(), E["width"], ok[0].data_ptr(), od[0].data_ptr(),
                                  ok[1].data_ptr(), od[1].data_ptr(), c4.data_ptr(), ur.data_ptr(), dp.data_ptr(), None, None, None))
    t1 = time.perf_counter()
    for key in ("pos", "normal", "minmax", "desc", "flags"):
        np.copyto(stg[key], mp[key])
    stg["holder"].fill(-1); stg["holder_obs"].fill(0)
    t2 = time.perf_counter()
    nm, h, ho, b = ctx.search_staged(M, int(c4[0]), bench.TH)
    t3 = time.perf_counter()
    # pieces
    ctx._ck(L_.ft_extract_stereo(ctx.h, hL.data_ptr(), E["width"], hR.data_ptr(), E["width"])) if False else None
    t4 = time.perf_counter()
    ctx.frame_enqueue_device(dL.data_ptr(), E["width"], dR.data_ptr(), E["width"]); ctx.synchronize()
    t5 = time.perf_counter()
    ctx.search_resident(bench.TH); ctx.synchronize()
    t6 = time.perf_counter()
    if it >= 10:
        T["frame_construct"].append(t1 - t0); T["marshal"].append(t2 - t1); T["search_staged"].append(t3 - t2)
        T["frame_enqueue+sync"].append(t5 - t4); T["search_resident+sync"].append(t6 - t5)
for k, v in T.items():
    if v: print("%-24s median %.1f us" % (k, 1e6 * float(np.median(v))))
