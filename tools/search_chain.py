"""Where does the period of the pipelined sequence go? Searches only (gather -> resolve), 2000 in a row:
(a) on ONE context / one stream, (b) alternating over D contexts whose streams are chained by events (what bench.py's
throughput leg does), (c) = (b) with the extraction of the next frames running beside it."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench
E = synth.EUROC
D = int(os.environ.get("D", "4")); N = 2000
mbf = np.float32(E["fx"] * E["baseline"])
(L, R), = bench.make_frames(5, 1)
dL = torch.from_numpy(L).cuda(); dR = torch.from_numpy(R).cuda()
ctxs = []
for d in range(D):
    c = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
    c.extract_stereo(L, R)
    if d == 0:
        g = c.download(0)
        mp = bench.fast_mappoints(ft.keypoints_as_array(g["kps"]), g["desc"], c.scale_tables()["scale"], bench.M_POINTS, 1)
    c.set_pose(np.eye(3), np.zeros(3))
    c.upload_map_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"])
    c.upload_holders(None, None)
    c.frame_enqueue_device(dL.data_ptr(), E["width"], dR.data_ptr(), E["width"]); c.search_resident(bench.TH); c.synchronize()
    ctxs.append(c)
streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", 0)) for c in ctxs]
def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(a, b)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / N * 1e3
def one(a, b):
    a.record(streams[0])
    for i in range(N): ctxs[0].search_resident(bench.TH)
    b.record(streams[0])
def chain(extract):
    def f(a, b):
        done = [torch.cuda.Event() for _ in range(N)]
        a.record(streams[0])
        for s_ in streams[1:]: s_.wait_event(a)
        for i in range(N):
            c_, s_ = ctxs[i % D], streams[i % D]
            if extract: c_.frame_enqueue_device(dL.data_ptr(), E["width"], dR.data_ptr(), E["width"])
            if i > 0: s_.wait_event(done[i - 1])
            c_.search_resident(bench.TH)
            done[i].record(s_)
        for j in range(1, D + 1): streams[0].wait_event(done[N - j])
        b.record(streams[0])
    return f
for rep in range(2):
    print("one stream %.1f us | event chain over %d streams %.1f us | with extraction beside it %.1f us" % (timed(one), D, timed(chain(False)), timed(chain(True))))
