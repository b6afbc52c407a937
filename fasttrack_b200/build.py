"""Builds fasttrack_b200/_build/libfasttrack_b200.so with nvcc for sm_100a (in-tree, no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libfasttrack_b200.so")
DRIVER_SRC = os.path.join(HERE, "host", "ft_sequence_driver.cpp")
DRIVER_SO = os.path.join(OUT_DIR, "libft_sequence_driver.so")
SOURCES = ["ft_context.cu", "ft_extract.cu", "ft_octree.cu", "ft_stereo.cu", "ft_sbp.cu", "ft_bow.cu"]
HEADERS = ["ft_device.cuh", "ft_camera.cuh", "ft_internal.h", "ft_sort.h",
           os.path.join("..", "..", "include", "fasttrack_b200.h"),
           os.path.join("..", "..", "include", "ft_orb_pattern.inc")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              # every float op on this path must round like the reference's scalar C++: no FMA contraction
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("FT_EXTRA_NVCC_FLAGS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(SO) or not os.path.exists(DRIVER_SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [DRIVER_SRC]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([_nvcc(), "-Wno-deprecated-gpu-targets", "-shared", "-o", SO] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    # host-side C++ sequence loop over the public C ABI (used by bench.py's end-to-end legs)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", DRIVER_SRC, "-o", DRIVER_SO, "-L" + OUT_DIR,
                           "-lfasttrack_b200", "-Wl,-rpath,$ORIGIN"])
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
