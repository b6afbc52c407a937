"""The C-ABI library must load and export every symbol include/fasttrack_b200.h declares (no compute here)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from fasttrack_b200 import build as b
    if b.needs_build():
        g.build()
    import fasttrack_b200
    return fasttrack_b200.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "fasttrack_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ft_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n


def test_python_binding_lists_the_same_exports(lib):
    import fasttrack_b200
    hdr = open(os.path.join(ROOT, "include", "fasttrack_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ft_[a-z0-9_]+)\s*\(", hdr))
    assert names == set(fasttrack_b200.EXPORTS)


def test_config_struct_layout_matches_header():
    import fasttrack_b200
    # 9 ints + 1 float + ... : the ctypes mirror must have the C struct's size (all 4-byte members, no padding)
    assert ctypes.sizeof(fasttrack_b200.Config) == 4 * (9 + 8 + 8 + 2 + 2 + 1 + 12 + 1)
    assert fasttrack_b200.KEYPOINT_DTYPE.itemsize == 24


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.ft_version()
    assert isinstance(lib.ft_last_error(), bytes)


def test_create_without_gpu_fails_loudly(lib):
    """No CPU fallback: on a box without a CUDA device context creation returns FT_ERR_CUDA."""
    import fasttrack_b200
    from conftest import _has_gpu
    if _has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(fasttrack_b200.FtError) as e:
        fasttrack_b200.Context(752, 480)
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)


def test_vocabulary_host_side_errors(lib, tmp_path):
    """ft_vocabulary_load_text / ft_vocabulary_create validate on the host before touching the device; with a valid tree
    and no GPU they fail with FT_ERR_CUDA (no CPU fallback for the bag-of-words path either)."""
    import numpy as np
    import fasttrack_b200 as ft
    from fasttrack_b200 import synth
    from conftest import _has_gpu
    with pytest.raises(ft.FtError) as e:
        ft.Vocabulary.load_text(str(tmp_path / "missing.txt"))
    assert e.value.status == 1
    bad = tmp_path / "bad.txt"
    bad.write_text("this is not a vocabulary\n1 2 3\n")
    with pytest.raises(ft.FtError) as e:
        ft.Vocabulary.load_text(str(bad))
    assert e.value.status == 1 and "DBoW2" in str(e.value)
    parent, leaf, desc, weight = synth.make_vocabulary(3, 2, seed=1)
    broken = parent.copy(); broken[2] = 7                      # a parent that does not exist yet
    with pytest.raises(ft.FtError) as e:
        ft.Vocabulary.from_arrays(3, 2, 0, 0, broken, leaf, desc, weight)
    assert e.value.status == 1
    fwd = tmp_path / "fwd.txt"
    synth.write_vocabulary_text(str(fwd), 3, 2, broken, leaf, desc, weight)
    with pytest.raises(ft.FtError) as e:
        ft.Vocabulary.load_text(str(fwd))
    assert e.value.status == 1
    if not _has_gpu():
        good = tmp_path / "good.txt"
        synth.write_vocabulary_text(str(good), 3, 2, parent, leaf, desc, weight)
        with pytest.raises(ft.FtError) as e:
            ft.Vocabulary.load_text(str(good))
        assert e.value.status == 2


def test_pattern_table_pinned():
    txt = open(os.path.join(ROOT, "include", "ft_orb_pattern.inc")).read()
    nums = [int(x) for x in re.findall(r"-?\d+", re.sub(r"//.*", "", txt))]
    assert len(nums) == 1024
    digest = hashlib.sha256(bytes([(n + 256) % 256 for n in nums])).hexdigest()
    assert digest == "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"
    ref = "/root/reference/src/ORBextractor.cc"
    if os.path.exists(ref):   # only in the build container; the GPU box has no reference tree
        src = "\n".join(open(ref).read().split("\n")[133:391])
        refnums = [int(x) for x in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", src))]
        assert refnums == nums


def test_descriptor_sincos_matches_host_libm(lib):
    """The descriptor kernel rotates the rBRIEF pattern with cosf/sinf of the keypoint angle (reference
    src/ORBextractor.cc:74). glibc's sinf/cosf are not always correctly rounded, and a rotated sample can land exactly
    on a .5 rounding boundary, so the kernel evaluates glibc's own polynomial; this pins that restatement (same code,
    compiled for the host) against the libm of the machine the oracle runs on: 2e6 angles, every degree-grid value the
    orientation can produce near quadrant boundaries, and the one input that exposed the difference."""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    rng = np.random.default_rng(3)
    deg = np.concatenate([rng.uniform(0, 360, 2_000_000), np.arange(0, 360.5, 0.5), [41.9094352722168, 360.0, 0.0, 1e-5]]).astype(np.float32)
    ang = (deg * np.float32(np.pi / np.float32(180.0))).astype(np.float32)      # float angle * (float)(CV_PI/180.f)
    s = np.zeros_like(ang); c = np.zeros_like(ang)
    lib.ft_debug_sincosf.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ft_debug_sincosf.restype = None
    lib.ft_debug_sincosf(len(ang), ang.ctypes.data, s.ctypes.data, c.ctypes.data)
    import oracle
    rs, rc = oracle.libm_sincosf(ang)                  # the host libm, as the oracle's computeOrbDescriptor calls it
    assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), rc.view(np.uint32))
    libm.sinf.restype = C.c_float; libm.sinf.argtypes = [C.c_float]
    # the exposed case: libm's sinf(0.73145765f) is 1 ulp above the correctly rounded value
    a0 = np.float32(0.73145765)
    assert np.float32(libm.sinf(float(a0))) != np.float32(np.sin(np.float64(a0)))
    s0 = np.zeros(1, np.float32); c0 = np.zeros(1, np.float32); a1 = np.array([a0], np.float32)
    lib.ft_debug_sincosf(1, a1.ctypes.data, s0.ctypes.data, c0.ctypes.data)
    assert s0[0] == np.float32(libm.sinf(float(a0)))


def test_host_mirror_and_driver_compile_and_link(lib, tmp_path):
    """The C++ host side above the C ABI (fasttrack_b200/host/ft_shim.h via tests/native/shim_demo.cpp, and the sequence
    driver) compiles with g++ and links against the library here, without a GPU."""
    import subprocess
    import fasttrack_b200 as ft
    libdir = os.path.dirname(ft.library_path())
    exe = str(tmp_path / "shim_demo")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-Wall", os.path.join(ROOT, "tests", "native", "shim_demo.cpp"),
                           "-o", exe, "-L" + libdir, "-lfasttrack_b200", "-Wl,-rpath," + libdir])
    assert os.path.exists(os.path.join(libdir, "libft_sequence_driver.so"))
    drv = ctypes.CDLL(os.path.join(libdir, "libft_sequence_driver.so"))
    for name in ("ftd_run_serial", "ftd_run_pipelined"):
        getattr(drv, name)
    # without a GPU the demo must fail loudly through the mirror's exception path, not crash or fall back
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode != 0


def test_sequence_driver_exports_and_rejects_bad_arguments(lib):
    """The host-side C++ loop bench.py times (fasttrack_b200/host/ft_sequence_driver.cpp) is built next to the library, binds to
    it, and exports the three entry points; without contexts it fails with a negative return instead of touching anything."""
    import fasttrack_b200
    drv = ctypes.CDLL(os.path.join(os.path.dirname(fasttrack_b200.library_path()), "libft_sequence_driver.so"))
    for name in ("ftd_run_serial", "ftd_run_pipelined", "ftd_phase_seconds"):
        assert hasattr(drv, name), name
    drv.ftd_run_pipelined.restype = ctypes.c_double
    nm = ctypes.c_longlong(7)
    assert drv.ftd_run_pipelined(None, 0, None, 0, ctypes.c_float(3.0), 0, 0, ctypes.byref(nm)) < 0   # D < 1
    buf = (ctypes.c_double * 7)()
    drv.ftd_phase_seconds.restype = None
    drv.ftd_phase_seconds(buf)
    assert all(v >= 0 for v in buf)


def test_split_search_entry_points_check_their_arguments(lib):
    """ft_search_store_submit / ft_search_collect (the asynchronous halves of ft_search_store) reject null arguments before any
    CUDA call, like the rest of the ABI."""
    lib.ft_search_store_submit.restype = ctypes.c_int
    lib.ft_search_collect.restype = ctypes.c_int
    vp = ctypes.c_void_p
    assert lib.ft_search_store_submit(vp(), 0, vp(), vp(), ctypes.c_float(3.0), 0, ctypes.c_float(50.0), ctypes.c_float(0.8), vp(), vp(), 1) == 1
    assert b"null argument" in lib.ft_last_error()
    assert lib.ft_search_collect(vp(), vp(), vp(), vp(), vp()) == 1
    assert b"null context" in lib.ft_last_error()
