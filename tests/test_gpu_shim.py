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
    # bag of words through the mirror: vocabulary text file + a KeyFrame derived from the left frame
    parent, leaf, vdesc, weight = synth.make_vocabulary(10, 5, seed=31)
    leaves = np.nonzero(leaf)[0]
    vdesc[leaves[:len(dL)]] = dL
    synth.write_vocabulary_text(d + "/voc.txt", 10, 5, parent, leaf, vdesc, weight)
    rng = np.random.default_rng(32)
    pick = rng.integers(0, len(dL), 1000)
    bits = np.unpackbits(dL[pick], axis=1)
    for i in range(len(pick)):
        bits[i, rng.choice(256, size=int(rng.integers(0, 30)), replace=False)] ^= 1
    kf_desc = np.packbits(bits, axis=1)
    kf_angle = ((kL[pick, 3] + rng.normal(15, 3, len(pick))) % 360).astype(np.float32)
    kf_has = (rng.random(len(pick)) < 0.8).astype(np.uint8)
    kf_desc.tofile(d + "/kf_desc.bin"); kf_angle.tofile(d + "/kf_angle.bin"); kf_has.tofile(d + "/kf_has.bin")
    out = subprocess.run([exe, d], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    vo = oracle.Vocabulary.load_text(d + "/voc.txt")
    fo, ko = vo.transform(dL, 4), vo.transform(kf_desc, 4)
    nm_bow, m_bow = oracle.search_by_bow(kf_desc, kf_angle, ko["node"], kf_has, dL, np.ascontiguousarray(kL[:, 3]), fo["node"],
                                         -1, 0.7, True)
    got = np.fromfile(d + "/out_bow_match.bin", np.int32)
    assert nm_bow > 100 and got[-1] == nm_bow and np.array_equal(got[:-1], m_bow)
    bv = np.fromfile(d + "/out_bow_vec.bin", np.float64).reshape(-1, 2)
    assert np.array_equal(bv[:, 0].astype(np.uint32), fo["bow_ids"]) and np.array_equal(bv[:, 1], fo["bow_vals"])
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
