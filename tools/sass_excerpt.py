"""profiles/r2_sass.md: per kernel of the built library, the SASS instruction count, the mnemonic histogram and the lines that
show which hardware features the kernel uses (cuobjdump -sass on fasttrack_b200/_build/libfasttrack_b200.so, sm_100a)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "fasttrack_b200", "_build", "libfasttrack_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass.md")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if m and cur:
        kern[cur].append(m.group(1).strip())
FEATURE = ("ATOMS", "ATOMG", "RED.", "REDUX", "SHFL", "VOTE", "MATCH", "BAR.", "UCGABAR", "ACQBULK", "PREEXIT", "LDS", "STS", "LDG", "STG",
           "VIMNMX", "VABSDIFF", "POPC", "LDSM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "CCTL", "ERRBAR", "MEMBAR", "DEPBAR", "NANOSLEEP")
with open(out, "w") as f:
    f.write("# SASS of the built kernels (cuobjdump -sass, sm_100a)\n\n")
    f.write("No dense contraction exists on this path, so there is no UTCMMA / tcgen05 instruction; image-tile movement is LDG.32/64/128 -> STS\n"
            "(the tiles are 1-4 KB per CTA with reflected borders, see DESIGN.md section 4); the 40 KB search structure of k_gather is moved by\n"
            "the TMA engine as 1-D bulk copies (UBLKCP = cp.async.bulk global -> shared, completion on an mbarrier: SYNCS.*), and k_resolve\n"
            "keeps its stamp tables in distributed shared memory (red / ld / st .shared::cluster, no L2 round trip). Programmatic dependent launch shows up as\n"
            "ACQBULK (griddepcontrol.wait) / PREEXIT-class instructions, the claim-resolution cluster barrier as UCGABAR_ARV / UCGABAR_WAIT.\n\n")
    f.write("| kernel | SASS instructions | " + " | ".join(FEATURE) + " |\n|---|---|" + "---|" * len(FEATURE) + "\n")
    for k, ins in kern.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        if not name.startswith("k_"):
            continue
        cnt = [sum(1 for i in ins if re.search(r"(^|\s)@?!?P?\d*\s*" + re.escape(ft), i) or i.startswith(ft) or (" " + ft) in (" " + i)) for ft in FEATURE]
        f.write("| %s | %d | %s |\n" % (name, len(ins), " | ".join(str(c) for c in cnt)))
    f.write("\n## Excerpts\n")
    for k, ins in kern.items():
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        if name not in ("k_octree", "k_fast_cells", "k_resolve", "k_gather", "k_stereo_match", "k_blur", "k_resize", "k_orient_desc"):
            continue
        hist = collections.Counter(i.split()[0] if not i.startswith("@") else i.split()[1] for i in ins)
        f.write("\n### %s (%d instructions)\n\nMost frequent opcodes: %s\n\n```\n" % (name, len(ins), ", ".join("%s x%d" % kv for kv in hist.most_common(14))))
        interesting = [i for i in ins if any(t in i for t in ("ACQBULK", "UCGABAR", "REDUX", "ATOMS", "MATCH", "VIMNMX", "POPC", "PREEXIT", "ATOMG", "RED.E"))][:14]
        f.write("\n".join(interesting) + "\n```\n")
print(open(out).read()[:3000])
