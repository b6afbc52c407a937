"""Per-CUDA-source-line stall sample summary of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
idx = sys.argv[4] if len(sys.argv) > 4 else "0"
metric = sys.argv[5] if len(sys.argv) > 5 else "# Samples"   # e.g. "Instructions Executed"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel,
                      "--launch-skip", idx, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) >= 7 and r[0].isdigit(): lines.append(r)
si = hdr.index(metric) if hdr and metric in hdr else 6
tot = sum(int(r[si]) for r in lines if r[si].isdigit()) or 1
print("kernel", kernel, "total", metric, tot)
for r in sorted(lines, key=lambda r: -(int(r[si]) if r[si].isdigit() else 0))[:top]:
    print("%5s %5s %5.1f%%  %s" % (r[0], r[si], 100.0 * int(r[si]) / tot, r[1].strip()[:130]))
