"""Randomised sweep of the oracle against the REFERENCE's own compiled code (oracle/_ref, needs /root/reference or the prebuilt
libraries): extractor over random sizes / scale factors / thresholds / image kinds / lapping areas, local-map and last-frame
searches over random poses, radii, far-point limits and holder patterns, fisheye searches. Prints the number of mismatches
(round 1: 0 in 55 extractor configurations, 50 pinhole searches and 10 fisheye searches). Longer than the test suite wants to
be; the suite runs a fixed subset (tests/test_oracle_ref_extractor.py, tests/test_oracle_ref_frame.py).
Usage: python tools/ref_sweep.py [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle  # noqa: E402
import make_ref_frame_golden as G  # noqa: E402
from fasttrack_b200 import synth  # noqa: E402

E, T = synth.EUROC, synth.TUMVI


def sweep_extractor(rng, trials=60):
    bad = 0
    for trial in range(trials):
        w = int(rng.integers(200, 1400)); h = int(rng.integers(max(120, w // 2 - 50), min(w + 1, 900)))
        nl = int(rng.integers(1, 9)); sf = float(rng.choice([1.1, 1.2, 1.25, 1.4, 2.0]))
        nf = int(rng.choice([30, 100, 500, 1000, 1200, 3000, 6000]))
        ini = int(rng.choice([20, 12, 30])); mn = int(rng.choice([7, 5, 10]))
        if min(w, h) / (sf ** (nl - 1)) < 32 + 36:      # the smallest level must hold one 35-px cell
            continue
        kind = trial % 3
        if kind == 0:
            img = synth.StereoScene(seed=500 + trial, width=w, height=h, margin_x=64, margin_y=8).pair()[0]
        elif kind == 1:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            img = np.full((h, w), 100, np.uint8)
            img[rng.integers(0, h, 200), rng.integers(0, w, 200)] = rng.integers(0, 256, 200)
            img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4), dtype=np.uint8)
        lap = (0, 0) if trial % 4 else (int(w * 0.2), int(w * 0.7))
        ref = oracle.RefExtractor(nf, sf, nl, ini, mn, w, h); ex = oracle.Extractor(nf, sf, nl, ini, mn)
        mr, kr, dr = ref.extract(img, lap); mo, ko, do = ex.extract(img, lap=lap)
        if not (mr == mo and np.array_equal(kr, ko) and np.array_equal(dr, do)):
            bad += 1; print("MISMATCH extractor", (w, h, nl, sf, nf, ini, mn, kind, lap))
    return bad


def sweep_searches(rng, trials=25):
    bad = 0
    L, R = synth.StereoScene(seed=61).pair()
    exL, exR = oracle.Extractor(1500), oracle.Extractor(1500)
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    so = oracle.stereo(exL, exR, kL, dL, kR, dR, float(G.MBF), float(G.MB))
    sr = oracle.ref_stereo(exL, exR, kL, dL, kR, dR, float(G.MBF), float(G.MB))
    bad += not (np.array_equal(so["uRight"], sr["uRight"]) and np.array_equal(so["depth"], sr["depth"]))
    N = len(kL)
    for trial in range(trials):
        a, b = float(rng.normal(0, 0.03)), float(rng.normal(0, 0.02))
        Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        Rcw = (Ry @ Rx).astype(np.float32); tcw = rng.normal(0, 0.05, 3).astype(np.float32)
        M = int(rng.integers(100, 12000)); th = float(rng.choice([1.0, 1.5, 3.0, 5.0, 15.0]))
        mp = synth.mappoints(kL, dL, exL.scale, M, seed=300 + trial, claimed_frac=float(rng.choice([0, 0.25, 0.6])))
        if trial % 3 == 0:
            mp["flags"] = (mp["flags"] & ~2) | ((rng.random(M) < 0.5) * 2).astype(mp["flags"].dtype)
        kw = dict(cam1=G.CAM, mbf=float(G.MBF), u_right=so["uRight"], Rcw=Rcw, tcw=tcw)
        Fo = oracle.Frame(kL, dL, exL.scale, E["width"], E["height"], **kw)
        Fr = oracle.RefFrame(kL, dL, exL.scale, E["width"], E["height"], **kw)
        args = (mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"], mp["holder_obs"])
        opt = dict(b_far=bool(trial % 2), th_far=float(rng.choice([3.0, 8.0, 20.0])), nnratio=float(rng.choice([0.8, 0.6, 0.9])))
        ro, rr = Fo.search_local_points(*args, **opt), Fr.search_local_points(*args, **opt)
        if not (ro[0] == rr[0] and np.array_equal(ro[1], rr[1]) and np.array_equal(ro[2], rr[2]) and np.array_equal(ro[3][:, :4], rr[3])):
            bad += 1; print("MISMATCH local map", trial, M, th)
        tz = float(rng.choice([0, 0.3, -0.3, 0.1])); Rc, tc, lf = G.last_frame_case(kL, dL, tz)
        kw2 = dict(cam1=G.CAM, mbf=float(G.MBF), u_right=so["uRight"], Rcw=Rc, tcw=tc)
        Fo = oracle.Frame(kL, dL, exL.scale, E["width"], E["height"], **kw2)
        Fr = oracle.RefFrame(kL, dL, exL.scale, E["width"], E["height"], **kw2)
        th2 = float(rng.choice([7, 15, 3])); ori = bool(trial % 2); mono = trial % 5 == 0
        mb = float(G.MB); tlc = float(Fo.Ow[2])
        direction = 0 if mono else (1 if tlc > mb else (-1 if -tlc > mb else 0))
        h0, o0 = np.full(N, -1, np.int32), np.zeros(N, np.uint8)
        a1 = Fo.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th2, direction, h0, o0, ori)
        a2 = Fr.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th2, np.eye(3), np.zeros(3), mb, h0, o0,
                                  mono, ori)
        if not (a1[0] == a2[0] and np.array_equal(a1[1], a2[1]) and np.array_equal(a1[2], a2[2])):
            bad += 1; print("MISMATCH last frame", trial, tz, th2, ori, mono)
    fexL, fkL, fdL, fkR, fdR, fo, (Rlr, tlr, Rrl, trl) = G.fisheye_frame()
    keys = np.vstack([fkL, fkR]); desc = np.vstack([fdL, fdR])
    for trial in range(10):
        mp, holder, hobs = G.fisheye_map(fkL, fdL, fkR, fexL.scale, int(rng.integers(500, 9000)), 40 + trial, bool(trial % 2))
        a = float(rng.normal(0, 0.02))
        Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        kw = dict(cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"], n_left=len(fkL), n_right=len(fkR), l2r=fo["l2r"], r2l=fo["r2l"],
                  Rrl=Rrl, trl=trl, tlr=tlr, Rcw=Rcw, tcw=rng.normal(0, 0.03, 3).astype(np.float32))
        Fo = oracle.Frame(keys, desc, fexL.scale, 512, 512, **kw); Fr = oracle.RefFrame(keys, desc, fexL.scale, 512, 512, **kw)
        th = float(rng.choice([1.0, 3.0, 6.0]))
        args = (mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, holder, hobs)
        ro, rr = Fo.search_local_points(*args), Fr.search_local_points(*args)
        if not (ro[0] == rr[0] and np.array_equal(ro[1], rr[1]) and np.array_equal(ro[2], rr[2]) and np.array_equal(ro[3][:, :4], rr[3])):
            bad += 1; print("MISMATCH fisheye", trial)
    return bad


if __name__ == "__main__":
    if oracle.ref_frame_lib() is None:
        sys.exit("oracle/_ref is not built and the reference tree is absent")
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 2024)
    t0 = time.time()
    b1 = sweep_extractor(rng); b2 = sweep_searches(rng)
    print("mismatches: extractor %d, searches %d (%.0f s)" % (b1, b2, time.time() - t0))
