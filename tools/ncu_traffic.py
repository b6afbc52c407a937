"""profiles/r1_traffic.json from one `ncu --set full` capture of a frame: dram__bytes_read.sum + dram__bytes_write.sum per
launch, keyed by the stage names bench.py uses (sums where a stage is several launches)."""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]; units = rr[1]
ci = {k: h.index(k) for k in ("Kernel Name", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
def val(r, k):
    return float(r[ci[k]].replace(",", "")) * mult.get(units[ci[k]], 1.0)
rows = rr[2:2 + 30]          # one frame = 30 launches
acc = {}
octs = []
for r in rows:
    name = r[ci["Kernel Name"]].split("(")[0]; grid = int(r[ci["launch__grid_size"]].replace(",", ""))
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    if name == "k_octree": octs.append(b); continue
    key = {"k_resize": "resize", "k_orient_desc": "orient_desc", "k_grid_build": "grid", "k_stereo_match": "stereo_match",
           "k_gather": "gather", "k_resolve": "resolve"}.get(name)
    if name == "k_fast_cells": key = "fast_cells_l0" if grid == 480 else "fast_cells"
    if name == "k_blur": key = "blur_l0" if grid == 360 else "blur"
    if key: acc[key] = acc.get(key, 0.0) + b
if octs:
    acc["octree_l0"] = max(octs); acc["octree"] = sum(octs) - max(octs)
json.dump({"source": "profiles/r1_ncu_summary.md capture (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                     "one frame of the bench workload, cold L2; stages that are several launches are summed)",
           "dram_bytes_per_launch": acc}, open(out, "w"), indent=1)
print(json.dumps(acc, indent=1))
