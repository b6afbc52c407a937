"""Where does the pipelined throughput go? Times (a) extraction + stereo only on D contexts, (b) the search chain only on
one context, (c) both as bench.py's throughput leg does. Device-resident inputs, CUDA events."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench
E = synth.EUROC
D = int(os.environ.get("FT_DEPTH", "4")); N = 400
mbf = np.float32(E["fx"] * E["baseline"])
ctxs = [ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf)) for _ in range(D)]
streams = [torch.cuda.ExternalStream(c.stream()) for c in ctxs]
frames = bench.make_frames(5, 16)
dL = [torch.from_numpy(a).cuda() for a, _ in frames]; dR = [torch.from_numpy(b).cuda() for _, b in frames]
ctxs[0].extract_stereo(*frames[0]); g = ctxs[0].download(0)
mp = bench.fast_mappoints(ft.keypoints_as_array(g["kps"]), g["desc"], ctxs[0].scale_tables()["scale"], bench.M_POINTS, 1)
for c in ctxs:
    c.set_pose(np.eye(3), np.zeros(3)); c.upload_map_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"]); c.upload_holders(None, None)
    c.frame_enqueue_device(dL[0].data_ptr(), E["width"], dR[0].data_ptr(), E["width"]); c.search_resident(3.0); c.synchronize()
def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(streams[0])
    for s in streams[1:]: s.wait_event(a)
    fn()
    ends = [torch.cuda.Event() for _ in streams]
    for e, s in zip(ends, streams): e.record(s)
    for e in ends: streams[0].wait_event(e)
    b.record(streams[0]); torch.cuda.synchronize()
    return a.elapsed_time(b) / N * 1e3
def extract_only():
    for i in range(N):
        ctxs[i % D].frame_enqueue_device(dL[i % 16].data_ptr(), E["width"], dR[i % 16].data_ptr(), E["width"])
def search_only():
    for i in range(N):
        ctxs[0].search_resident(3.0)
def both():
    done = [torch.cuda.Event() for _ in range(N)]
    for i in range(N):
        c, s = ctxs[i % D], streams[i % D]
        c.frame_enqueue_device(dL[i % 16].data_ptr(), E["width"], dR[i % 16].data_ptr(), E["width"])
        if i: s.wait_event(done[i - 1])
        c.search_resident(3.0); done[i].record(s)
for name, fn in (("extract+stereo only", extract_only), ("search chain only", search_only), ("both (bench leg)", both)):
    fn(); print("%-22s %.1f us/frame (depth %d)" % (name, timed(fn), D))

# how long do the two search kernels take while later frames are being extracted? (events inside the pipelined leg)
def both_timed():
    done = [torch.cuda.Event() for _ in range(N)]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
    for i in range(N):
        c, s = ctxs[i % D], streams[i % D]
        c.frame_enqueue_device(dL[i % 16].data_ptr(), E["width"], dR[i % 16].data_ptr(), E["width"])
        if i: s.wait_event(done[i - 1])
        ev[i][0].record(s); c.search_resident(3.0); ev[i][1].record(s); done[i].record(s)
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev[20:]]) * 1e3
    print("search (gather + resolve) under load: mean %.1f us, p50 %.1f, p95 %.1f" % (t.mean(), np.percentile(t, 50), np.percentile(t, 95)))
both_timed()
