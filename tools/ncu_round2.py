"""profiles/r2_ncu_summary.md + profiles/r2_traffic.json from an ncu launch list (csv) and one full capture (.ncu-rep, taken
with `--set full --metrics lts__t_bytes.sum,lts__t_sectors.sum --clock-control none --import-source on`).

Per kernel (first launch of each (kernel, grid) in the capture): duration, DRAM read + write bytes, L2 bytes, achieved HBM
and L2 GB/s against the measured HBM peak (MEASURED_PEAKS.json), warp instructions, warps-active, registers.
usage: python tools/ncu_round2.py launches.csv prof.ncu-rep frames out_md out_traffic_json [git-head]"""
import collections, csv, json, os, subprocess, sys

launches_csv, rep, frames, out_md, out_json = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
head = sys.argv[6] if len(sys.argv) > 6 else "?"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6553.3
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = json.load(open(pp))["hbm_gbs"]

rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[1:]:
    acc.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in acc.values())
frames = max(1, len(acc.get("k_resolve", [])) or frames)     # one claim resolution per frame: the list's own frame count

raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
cols = {k: (h.index(k) if k in h else None) for k in want}
MULT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3,
        "sector": 1.0, "inst": 1.0, "%": 1.0}


def val(r, name):
    c = cols.get(name)
    if c is None or r[c] in ("", "n/a"):
        return float("nan")
    return float(r[c].replace(",", "")) * MULT.get(units[c], 1.0)


def l2_bytes(r):
    b = val(r, "lts__t_bytes.sum")
    if b != b:   # nan: fall back to sectors x 32 B
        b = val(r, "lts__t_sectors.sum") * 32.0
    return b


full = collections.OrderedDict()
for r in rr[2:]:
    key = (r[cols["Kernel Name"]].split("(")[0], r[cols["launch__grid_size"]])
    full.setdefault(key, r)

STAGE = {"k_resize": "resize", "k_orient_desc": "orient_desc", "k_grid_build": "grid", "k_stereo_match": "stereo_match",
         "k_gather": "gather", "k_resolve": "resolve"}
traffic, l2t = {}, {}
octs, octl2 = [], []
per_frame = rr[2:]
first_frame = []
seen_orient = 0
for r in per_frame:          # one frame = everything up to and including the first k_resolve
    first_frame.append(r)
    if r[cols["Kernel Name"]].startswith("k_resolve"):
        break
fast_grids = sorted({int(r[cols["launch__grid_size"]].replace(",", "")) for r in first_frame if r[cols["Kernel Name"]].startswith("k_fast_cells")})
blur_grids = sorted({int(r[cols["launch__grid_size"]].replace(",", "")) for r in first_frame if r[cols["Kernel Name"]].startswith("k_blur")})
for r in first_frame:
    name = r[cols["Kernel Name"]].split("(")[0]; grid = int(r[cols["launch__grid_size"]].replace(",", ""))
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    l2 = l2_bytes(r)
    if name == "k_octree":
        octs.append(b); octl2.append(l2); continue
    key = STAGE.get(name)
    if name == "k_fast_cells":
        key = "fast_cells_l0" if grid == max(fast_grids) else "fast_cells"
    if name == "k_blur":
        key = "blur_l0" if (len(blur_grids) > 1 and grid == min(blur_grids)) else "blur"
    if key:
        traffic[key] = traffic.get(key, 0.0) + b; l2t[key] = l2t.get(key, 0.0) + l2
if octs:
    i = max(range(len(octs)), key=lambda j: octs[j])
    traffic["octree_l0"] = octs[i]; traffic["octree"] = sum(octs) - octs[i]
    l2t["octree_l0"] = octl2[i]; l2t["octree"] = sum(octl2) - octl2[i]
json.dump({"source": "%s capture at %s (ncu --set full + lts__t_bytes.sum, --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per "
                     "launch, first frame of the capture, cold L2; stages that are several launches are summed)" % (os.path.basename(out_md), head),
           "dram_bytes_per_launch": traffic, "l2_bytes_per_launch": l2t}, open(out_json, "w"), indent=1)

with open(out_md, "w") as f:
    f.write("# ncu summary, round 2 (code at %s; %d frames of the bench workload, device-resident inputs)\n\n" % (head, frames))
    f.write("Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_` (cold-cache, serialised: compare shares).\n\n")
    f.write("| kernel | launches/frame | mean us | share of kernel time |\n|---|---|---|---|\n")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        f.write("| %s | %.0f | %.2f | %.1f%% |\n" % (k, len(v) / frames, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    f.write("\nSum of kernel time per frame: %.1f us\n\n" % (tot / frames / 1e3))
    f.write("Full capture (`ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct --clock-control none`, one frame), "
            "first launch of each (kernel, grid). HBM GB/s = (DRAM read + write) / duration, L2 GB/s = lts__t_bytes / "
            "duration, both against the measured HBM copy peak of %.1f GB/s (MEASURED_PEAKS.json); under ncu every launch runs alone with a "
            "cold L2, so these are per-launch figures, not the pipelined frame's.\n\n" % peak)
    f.write("| kernel | grid x block | regs | us | DRAM rd B | DRAM wr B | HBM GB/s | % of HBM peak | L2 bytes | L2 GB/s | L2 hit % | warp insts | warps active % |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for (k, g), r in full.items():
        us = val(r, "gpu__time_duration.sum")
        rd, wr, l2 = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), l2_bytes(r)
        hbm = (rd + wr) / (us * 1e-6) / 1e9 if us == us and us > 0 else float("nan")
        l2g = l2 / (us * 1e-6) / 1e9 if us == us and us > 0 else float("nan")
        f.write("| %s | %s x %s | %s | %.2f | %.0f | %.0f | %.1f | %.2f | %.0f | %.1f | %.1f | %.0f | %.1f |\n" % (
            k, g, r[cols["launch__block_size"]], r[cols["launch__registers_per_thread"]], us, rd, wr, hbm, 100 * hbm / peak, l2, l2g,
            val(r, "lts__t_sector_hit_rate.pct"), val(r, "smsp__inst_executed.sum"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
print(open(out_md).read())
