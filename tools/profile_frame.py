"""Small driver for ncu: W warm-up + K timed-shape frames of the bench workload (device-resident inputs).
Kernel launches per frame: 14 (extract) + 2 (stereo + grid) + 2 (search) = 18; setup issues 14 more before the loop."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench

E = synth.EUROC
W = int(os.environ.get("FT_PROF_WARMUP", "3")); K = int(os.environ.get("FT_PROF_STEPS", "5"))
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
for i in range(W + K):
    ctx.frame_enqueue_device(dL.data_ptr(), E["width"], dR.data_ptr(), E["width"])
    ctx.search_resident(bench.TH)
    ctx.synchronize()
print("profiled frames:", K, "launch counts", ctx.launch_counts(), ctx.stats())
