"""The host-side C++ sequence loop (fasttrack_b200/host/ft_sequence_driver.cpp, public C ABI only) that bench.py times
for its end-to-end legs: one frame at a time, several frames in flight, and several frames in flight over the
persistent map store must all find the matches the Python-driven calls find."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC


def test_native_sequence_loops_agree(euroc_pair):
    L0, R0 = euroc_pair
    sc = synth.StereoScene(seed=13)
    frames = [(L0, R0)] + [sc.pair(pan=(4 * t, -t), noise_seed=40 + t) for t in range(1, 5)]
    M, D, steps = 4000, 3, 11
    mbf = np.float32(E["fx"] * E["baseline"])
    mk = lambda: ft.Context(E["width"], E["height"], nfeatures=1200, nlevels=8, cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
    ctxs = [mk() for _ in range(D)]
    for c in ctxs:
        c.set_pose(np.eye(3), np.zeros(3))
    scale = ctxs[0].scale_tables()["scale"]
    maps, expect = [], []
    for k, (a, b) in enumerate(frames):
        l, r = ctxs[0].frame_construct(a, b)
        mp = synth.mappoints(ft.keypoints_as_array(l["kps"]), l["desc"], scale, M, seed=500 + k)
        mp = {key: np.ascontiguousarray(v) for key, v in mp.items()}
        mp["flags"] = mp["flags"].astype(np.int32)
        maps.append(mp)
        n = l["n"]
        expect.append(ctxs[0].search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                  np.full(n, -1, np.int32), np.zeros(n, np.uint8))[0])
    total = sum(expect[i % len(frames)] for i in range(steps))
    ctxs[0].map_store_create(len(frames) * M)
    for c in ctxs[1:]:
        c.map_store_attach(ctxs[0])
    rows = [np.arange(k * M, (k + 1) * M, dtype=np.int32) for k in range(len(frames))]
    for k, mp in enumerate(maps):
        ctxs[0].map_store_update(rows[k], mp["pos"], mp["normal"], mp["minmax"], mp["desc"])

    drv = C.CDLL(os.path.join(os.path.dirname(ft.library_path()), "libft_sequence_driver.so"))
    names = ("imgL", "imgR", "pos", "normal", "minmax", "desc", "flags", "rows")

    class Seq(C.Structure):
        _fields_ = [("n_frames", C.c_int), ("width", C.c_int), ("height", C.c_int), ("M", C.c_int)] + [(k, C.POINTER(C.c_void_p)) for k in names]
    arr = lambda ptrs: (C.c_void_p * len(ptrs))(*ptrs)
    keep = dict(imgL=arr([a.ctypes.data for a, _ in frames]), imgR=arr([b.ctypes.data for _, b in frames]), rows=arr([r.ctypes.data for r in rows]))
    for key in ("pos", "normal", "minmax", "desc", "flags"):
        keep[key] = arr([m[key].ctypes.data for m in maps])
    seq = Seq(len(frames), E["width"], E["height"], M, *[C.cast(keep[k], C.POINTER(C.c_void_p)) for k in names])
    drv.ftd_run_serial.restype = C.c_double
    drv.ftd_run_serial.argtypes = [C.c_void_p, C.POINTER(Seq), C.c_int, C.c_float, C.POINTER(C.c_longlong)]
    drv.ftd_run_pipelined.restype = C.c_double
    drv.ftd_run_pipelined.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Seq), C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    hctx = (C.c_void_p * D)(*[c.h for c in ctxs])
    nm = C.c_longlong()
    assert drv.ftd_run_serial(ctxs[0].h, C.byref(seq), steps, 3.0, C.byref(nm)) > 0 and nm.value == total
    assert drv.ftd_run_pipelined(hctx, D, C.byref(seq), steps, 3.0, 0, 0, C.byref(nm)) > 0 and nm.value == total
    assert drv.ftd_run_pipelined(hctx, D, C.byref(seq), steps, 3.0, 1, 300, C.byref(nm)) > 0 and nm.value == total
    assert drv.ftd_run_pipelined(hctx, 1, C.byref(seq), steps, 3.0, 0, 0, C.byref(nm)) > 0 and nm.value == total
    # the search as ft_search_store_submit / ft_search_collect with the next frame handed over between the halves (2), and with
    # frame t+1 collected + upserted in the shadow of the search of frame t as well (3); fewer contexts fall back
    for mode in (2, 3):
        for d in (3, 2, 1):
            assert drv.ftd_run_pipelined(hctx, d, C.byref(seq), steps, 3.0, mode, 300, C.byref(nm)) > 0 and nm.value == total, (mode, d)
    assert drv.ftd_run_pipelined(hctx, D, C.byref(seq), 1, 3.0, 3, 300, C.byref(nm)) > 0 and nm.value == expect[0]
    assert drv.ftd_run_pipelined(hctx, D, C.byref(seq), 2, 3.0, 3, 300, C.byref(nm)) > 0 and nm.value == expect[0] + expect[1]
    for c in ctxs:
        c.close()
