"""The C++ host mirror (fasttrack_b200/host/ft_shim.h): ORBextractor::operator() from two threads,
Frame::ComputeStereoMatches and ORBmatcher::SearchByProjection, compiled with g++ against the C-ABI library and
compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
import fasttrack_b200 as ft
from fasttrack_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = synth.EUROC


def test_shim_matches_oracle(tmp_path, euroc_pair):
    exe = str(tmp_path / "shim_demo")
    libdir = os.path.dirname(ft.library_path())
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "native", "shim_demo.cpp"),
                           "-o", exe, "-L" + libdir, "-lfasttrack_b200", "-Wl,-rpath," + libdir])
    L, R = euroc_pair
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    exL, exR = oracle.Extractor(), oracle.Extractor()
    monoL, kL, dL = exL.extract(L); monoR, kR, dR = exR.extract(R)
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
    mp = synth.mappoints(kL, dL, exL.scale, 6000, seed=21)
    mp["flags"] = mp["flags"].astype(np.int32)
    d = str(tmp_path)
    L.tofile(d + "/L.bin"); R.tofile(d + "/R.bin")
    np.array([E["fx"], E["fy"], E["cx"], E["cy"], mbf], np.float32).tofile(d + "/cam.bin")
    for k in ("pos", "normal", "minmax", "desc", "flags"):
        mp[k].tofile(d + "/mp_%s.bin" % k)
    out = subprocess.run([exe, d], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    gk = np.fromfile(d + "/out_kL.bin", np.float32).reshape(-1, 6)
    gd = np.fromfile(d + "/out_dL.bin", np.uint8).reshape(-1, 32)
    meta = np.fromfile(d + "/out_meta.bin", np.int32)
    assert np.array_equal(gk, kL) and np.array_equal(gd, dL)
    assert meta[0] == monoL and meta[1] == monoR and meta[2] == len(kL) and meta[3] == len(kR)
    assert np.array_equal(np.fromfile(d + "/out_uRight.bin", np.float32), st["uRight"])
    assert np.array_equal(np.fromfile(d + "/out_depth.bin", np.float32), st["depth"])
    F = oracle.Frame(kL, dL, exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                     mbf=float(mbf), u_right=st["uRight"])
    holder0 = np.full(len(kL), -1, np.int32); hobs0 = np.zeros(len(kL), np.uint8)
    n_o, h_o, _, ti, _ = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder0, hobs0)
    assert meta[4] == n_o
    assert np.array_equal(np.fromfile(d + "/out_holder.bin", np.int32), h_o)
