import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, "/root/repo")
os.environ["FT_SORT_CLOCK"] = "1"
import fasttrack_b200 as ft
L = ft.load_library()
L.ft_debug_sort.argtypes = [C.c_void_p, C.c_int]
rng = np.random.default_rng(0)
for n in (17, 32, 64, 128, 128, 256, 512):
    size = rng.integers(2, 40, n).astype(np.uint64); ulx = (rng.integers(0, 16, n) * 22).astype(np.uint64)
    keys = (((size << np.uint64(12)) | ulx) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    a = np.ascontiguousarray(keys)
    L.ft_debug_sort(a.ctypes.data, n)
