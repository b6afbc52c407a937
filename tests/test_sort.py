"""The device introsort (fasttrack_b200/csrc/ft_sort.h) must move elements exactly like libstdc++ std::sort:
the octree's careful mode depends on the order of equivalent elements (ORBextractor.cc:805-852)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sort") / "sort_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "sort_harness.cpp")])
    L = ctypes.CDLL(so)
    p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    L.harness_ftsort.argtypes = [p, ctypes.c_int]
    L.harness_stdsort.argtypes = [p, ctypes.c_int]
    return L


def _same(L, keys):
    n = len(keys)
    a = (np.asarray(keys).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    b = a.copy()
    L.harness_ftsort(a, n)
    L.harness_stdsort(b, n)
    return np.array_equal(a, b)


def test_random_with_ties(harness):
    rng = np.random.default_rng(0)
    for _ in range(4000):
        n = int(rng.integers(0, 400))
        kmax = int(rng.choice([1, 2, 3, 5, 10, 100, 100000]))
        assert _same(harness, rng.integers(0, kmax, n))


@pytest.mark.parametrize("n", [17, 33, 100, 257, 1000, 5000])
def test_structured(harness, n):
    assert _same(harness, np.arange(n))
    assert _same(harness, np.arange(n)[::-1].copy())
    assert _same(harness, np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]))
    assert _same(harness, np.zeros(n, np.int64))


def test_median_of_three_killer_reaches_heapsort(harness):
    # Musser's adversary drives introsort past its depth limit, exercising the heapsort branch
    for k in (64, 256, 1024, 4096):
        a = np.zeros(2 * k, np.int64)
        for i in range(k):
            a[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 == 0 else 0)
            a[k + i] = 2 * (i + 1)
        assert _same(harness, a)
