"""Ad-hoc end-to-end parity check on a GPU box: CUDA path vs CPU oracle, stage by stage."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import fasttrack_b200 as ft
from fasttrack_b200 import synth

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print("  %-28s SHAPE MISMATCH %s vs %s" % (name, a.shape, b.shape)); return False
    bad = int((a != b).sum())
    print("  %-28s %s (%d/%d differ)" % (name, "OK" if bad == 0 else "MISMATCH", bad, a.size))
    return bad == 0

def main():
    E = synth.EUROC
    t0 = time.time()
    sc = synth.StereoScene(seed=2)
    L, R = sc.pair()
    print("scene %.1fs" % (time.time() - t0))
    exL, exR = oracle.Extractor(), oracle.Extractor()
    t0 = time.time()
    monoL, kL, dL = exL.extract(L); monoR, kR, dR = exR.extract(R)
    print("oracle extract 2 imgs %.1f ms, nL=%d nR=%d" % ((time.time() - t0) * 1e3, len(kL), len(kR)))
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
    ctx.extract_stereo(L, R)
    c = ctx.counts()
    print("gpu counts", c)
    ok = True
    for eye, ex in ((0, exL), (1, exR)):
        print("eye", eye)
        for l in range(8):
            ok &= cmp("pyr L%d" % l, ctx.level_image(eye, l), ex.level_image(l))
            ob = ex.level_image(l, True)
            if ob is not None: ok &= cmp("blur L%d" % l, ctx.level_image(eye, l, True), ob)
            ok &= cmp("cand L%d" % l, ctx.level_candidates(eye, l), ex.level_candidates(l))
    for eye, (k, d, mono) in ((0, (kL, dL, monoL)), (1, (kR, dR, monoR))):
        g = ctx.download(eye)
        gk = ft.keypoints_as_array(g["kps"])
        print("eye", eye, "final n", g["n"], "oracle", len(k), "mono", g["mono_index"], mono)
        ok &= cmp("kps", gk, k)
        if gk.shape == k.shape:
            for ci, nm in enumerate(["x", "y", "size", "angle", "resp", "oct"]):
                bad = int((gk[:, ci] != k[:, ci]).sum())
                if bad: print("     col", nm, "bad", bad, "first", np.nonzero(gk[:, ci] != k[:, ci])[0][:5])
            dd = (g["desc"] != d).any(axis=1)
            print("  desc rows differing: %d of %d (oracle borderline samples: %d)" % (dd.sum(), len(d), exL.desc_borderline() if eye == 0 else exR.desc_borderline()))
    # stereo
    ctx.stereo_match()
    g = ctx.download(0, stereo=True)
    o = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
    print("stereo: oracle matched", int((o["depth"] > 0).sum()), "gpu", int((g["depth"] > 0).sum()))
    if len(g["u_right"]) == len(o["uRight"]):
        ok &= cmp("uRight", g["u_right"], o["uRight"])
        ok &= cmp("depth", g["depth"], o["depth"])
        print("  max |duR|", float(np.abs(g["u_right"] - o["uRight"]).max()))
    # projection search
    sf = exL.scale
    for M, th in ((5000, 1.0), (20000, 6.0)):
        mp = synth.mappoints(kL, dL, sf, M, seed=4)
        F = oracle.Frame(kL, dL, sf, E["width"], E["height"], cam_type=0, cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                         mbf=float(mbf), u_right=o["uRight"])
        t0 = time.time()
        n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"], mp["holder_obs"])
        t_or = time.time() - t0
        ctx.set_pose(np.eye(3), np.zeros(3))
        t0 = time.time()
        n_g, h_g, ho_g, best = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"], mp["holder_obs"])
        t_g = time.time() - t0
        gi, gf = ctx.track(M)
        print("SBP M=%d th=%g: oracle nmatches %d (%.1f ms), gpu %d (%.1f ms), borderline %d, stats %s" % (M, th, n_o, t_or * 1e3, n_g, t_g * 1e3, int(ti[:, 4].sum()), ctx.stats()))
        ok &= cmp("inView", gi[:, 0], ti[:, 0]); ok &= cmp("level", gi[:, 2], ti[:, 2])
        ok &= cmp("trackF", gf[:, :5], tf[:, :5])
        ok &= cmp("holder", h_g, h_o); ok &= cmp("holderObs", ho_g, ho_o)
        ok &= (n_o == n_g)
    print("ALL OK" if ok else "SOME MISMATCH")

if __name__ == "__main__":
    main()
