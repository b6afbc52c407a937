"""Summarise an ncu launch list (csv) and a full capture (.ncu-rep) into a markdown table for profiles/."""
import collections, csv, subprocess, sys

launches_csv, rep, out_md, frames = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[1:]:
    acc.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in acc.values())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
def col(name):
    return h.index(name) if name in h else None
cols = {k: col(k) for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                            "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
                            "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum"]}
units = rr[1]
full = collections.OrderedDict()
for r in rr[2:]:
    key = (r[cols["Kernel Name"]].split("(")[0], r[cols["launch__grid_size"]])
    if key not in full:
        full[key] = r
def val(r, name, scale=1.0):
    c = cols[name]
    if c is None or r[c] in ("", "n/a"):
        return float("nan")
    v = float(r[c].replace(",", ""))
    u = units[c]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u, 1.0)
    return v * mult * scale
with open(out_md, "w") as f:
    f.write("# ncu summary (%d frames of the bench workload, device-resident inputs)\n\n" % frames)
    f.write("Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_` (cold-cache, serialised: compare shares).\n\n")
    f.write("| kernel | launches/frame | mean us | share of frame |\n|---|---|---|---|\n")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        f.write("| %s | %.0f | %.2f | %.1f%% |\n" % (k, len(v) / frames, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    f.write("\nSum of kernel time per frame: %.1f us\n\n" % (tot / frames / 1e3))
    f.write("Full capture (`ncu --set full --clock-control none --import-source on`), first launch of each (kernel, grid):\n\n")
    f.write("| kernel | grid x block | regs | us | DRAM read B | DRAM write B | L2 bytes | warp insts | warps active % |\n|---|---|---|---|---|---|---|---|---|\n")
    for (k, g), r in full.items():
        f.write("| %s | %s x %s | %s | %.2f | %.0f | %.0f | %.0f | %.0f | %.1f |\n" % (
            k, g, r[cols["launch__block_size"]], r[cols["launch__registers_per_thread"]], val(r, "gpu__time_duration.sum"),
            val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "lts__t_bytes.sum"),
            val(r, "smsp__inst_executed.sum"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
print(open(out_md).read())
