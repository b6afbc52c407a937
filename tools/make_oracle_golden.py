"""Generates tests/golden/pipeline_mini.npz: outputs of the CPU oracle on a small seeded stereo pair
(376x240, 400 features, 6 levels) and a 2000-point local map. These freeze the operator-level results so
that (a) a change to the oracle is visible in review and (b) the CUDA path can be checked against committed
vectors as well as against the live oracle."""
import os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from fasttrack_b200 import synth

MINI = dict(width=376, height=240, nfeatures=400, nlevels=6, fx=229.327, fy=228.648, cx=183.6, cy=124.2, baseline=0.110074)


def main():
    sc = synth.StereoScene(seed=7, width=MINI["width"], height=MINI["height"], dmin=1.0, dmax=32.0, margin_x=64, margin_y=8)
    L, R = sc.pair()
    exL = oracle.Extractor(MINI["nfeatures"], 1.2, MINI["nlevels"]); exR = oracle.Extractor(MINI["nfeatures"], 1.2, MINI["nlevels"])
    monoL, kL, dL = exL.extract(L); monoR, kR, dR = exR.extract(R)
    mbf = np.float32(MINI["fx"] * MINI["baseline"]); mb = np.float32(mbf / np.float32(MINI["fx"]))
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
    mp = synth.mappoints(kL, dL, exL.scale, 2000, seed=9, width=MINI["width"], height=MINI["height"], fx=MINI["fx"],
                         fy=MINI["fy"], cx=MINI["cx"], cy=MINI["cy"])
    F = oracle.Frame(kL, dL, exL.scale, MINI["width"], MINI["height"], cam1=[MINI["fx"], MINI["fy"], MINI["cx"], MINI["cy"], 0, 0, 0, 0],
                     mbf=float(mbf), u_right=st["uRight"])
    n, holder, hobs, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                    mp["holder"], mp["holder_obs"])
    out = dict(imgL=L, imgR=R, kL=kL, dL=dL, kR=kR, dR=dR, monoL=monoL, monoR=monoR, uRight=st["uRight"], depth=st["depth"],
               sad=st["sad"], mbf=mbf, mb=mb, sbp_n=n, sbp_holder=holder, sbp_holder_obs=hobs, track_i=ti, track_f=tf,
               cand_counts=np.array([len(exL.level_candidates(l)) for l in range(MINI["nlevels"])]))
    for k, v in mp.items():
        out["mp_" + k] = v
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pipeline_mini.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes; nL", len(kL), "nR", len(kR), "stereo", int((st["depth"] > 0).sum()), "sbp", n)


if __name__ == "__main__":
    main()
