"""The warp-cooperative device introsort (ft_sort.h: warp_partition + stable_rank) against libstdc++ std::sort,
on tie-heavy inputs: the octree's careful mode observes the order of equivalent elements."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import fasttrack_b200 as ft

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def stdsort(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sort") / "sort_harness.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "sort_harness.cpp")])
    L = ctypes.CDLL(so)
    L.harness_stdsort.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), ctypes.c_int]
    def f(words):
        a = np.ascontiguousarray(words, np.uint64).copy()
        L.harness_stdsort(a, len(a))
        return a
    return f


def _words(keys):
    keys = np.asarray(keys)
    return (keys.astype(np.uint64) << np.uint64(32)) | np.arange(len(keys), dtype=np.uint64)


def test_device_sort_random_ties(stdsort):
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(0, 700))
        kmax = int(rng.choice([1, 2, 3, 5, 10, 100, 100000]))
        w = _words(rng.integers(0, kmax, n))
        assert np.array_equal(ft.device_sort(w), stdsort(w)), (n, kmax)


@pytest.mark.parametrize("n", [16, 17, 33, 100, 257, 1000, 4096])
def test_device_sort_structured(stdsort, n):
    for keys in (np.arange(n), np.arange(n)[::-1].copy(), np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]),
                 np.zeros(n, np.int64), np.arange(n) % 3):
        w = _words(keys)
        assert np.array_equal(ft.device_sort(w), stdsort(w))


def test_device_sort_heapsort_branch(stdsort):
    for k in (64, 256, 1024, 2048):
        a = np.zeros(2 * k, np.int64)
        for i in range(k):
            a[i] = i + 1 if i % 2 == 0 else k + i + (1 if k % 2 == 0 else 0)
            a[k + i] = 2 * (i + 1)
        w = _words(a)
        assert np.array_equal(ft.device_sort(w), stdsort(w))
