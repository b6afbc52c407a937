"""Small driver for ncu: Frame::ComputeBoW + ORBmatcher::SearchByBoW on the bench's ORBvoc-shaped synthetic vocabulary
(k = 10, L = 6), W warm-up + K profiled repetitions."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasttrack_b200 as ft
from fasttrack_b200 import synth
import bench

E = synth.EUROC
W = int(os.environ.get("FT_PROF_WARMUP", "2")); K = int(os.environ.get("FT_PROF_STEPS", "4"))
mbf = np.float32(E["fx"] * E["baseline"])
ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
frames = bench.make_frames(5, 2)
parent, leaf, vdesc, weight = synth.make_vocabulary_bfs(10, 6, seed=11)
ctx.extract_stereo(*frames[0]); ctx.stereo_match()
kf = ctx.download(0)
spread = np.nonzero(leaf)[0][::max(1, int(leaf.sum()) // kf["n"])][:kf["n"]]
vdesc[spread] = kf["desc"][:len(spread)]
voc = ft.Vocabulary.from_arrays(10, 6, 0, 0, parent, leaf, vdesc, weight)
kf_node = voc.transform(kf["desc"], 4)["node"]
ctx.extract_stereo(*frames[1]); ctx.stereo_match()
for i in range(W + K):
    ctx.compute_bow(voc, 4)
    nm, m = ctx.search_by_bow(kf["desc"], kf["kps"]["angle"], kf_node, np.ones(kf["n"], np.uint8), 0.7, True)
print("profiled repetitions:", K, "matches", nm)
