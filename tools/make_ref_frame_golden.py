"""Writes tests/golden/ref_frame.npz: outputs of the REFERENCE's own Frame / ORBmatcher functions (their text compiled at
build time into oracle/_ref/libft_ref_frame.so, see oracle/ref_extract_fns.py) on seeded synthetic inputs:
ComputeStereoMatches, the isInFrustum loop + SearchByProjection over a local map (pinhole and fisheye rigs),
SearchByProjection(CurrentFrame, LastFrame), ComputeStereoFromRGBD, SearchByBoW. Inputs are regenerated from the seeds by
the tests; keypoints / descriptors come from the extractor, whose reference outputs are in ref_orbextractor.npz.
Run in the build container (needs /root/reference); the fixture travels, the reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fasttrack_b200 import synth  # noqa: E402

E, T = synth.EUROC, synth.TUMVI
CAM = [E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0]
MBF = np.float32(E["fx"] * E["baseline"])
MB = np.float32(MBF / np.float32(E["fx"]))
LOCAL_CASES = [(5000, 3.0, 4), (20000, 1.0, 5), (10000, 6.0, 6)]          # M, th, seed
LAST_CASES = [(0.0, 7.0, True), (0.4, 7.0, True), (-0.4, 15.0, True), (0.05, 14.0, False)]   # tz, th, check_ori


def euroc_frame(make_extractor):
    """(exL, exR, kL, dL, kR, dR) of the EuRoC-shaped pair, through `make_extractor` (oracle.Extractor or RefExtractor)"""
    L, R = synth.StereoScene(seed=2).pair()
    exL, exR = make_extractor(), make_extractor()
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    return exL, exR, kL, dL, kR, dR


def last_frame_case(kL, dL, tz):
    a = 0.02
    Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    tcw = np.array([0.03, -0.02, -tz], np.float32)
    lf = synth.last_frame_points(kL, dL, 1000, seed=int(100 + 10 * tz), Rcw=Rcw, tcw=tcw, fx=E["fx"], fy=E["fy"], cx=E["cx"], cy=E["cy"])
    return Rcw, tcw, lf


def fisheye_frame():
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    exL, exR = oracle.Extractor(1000), oracle.Extractor(1000)
    mL, kL, dL = exL.extract(L, lap=T["lap"]); mR, kR, dR = exR.extract(R, lap=T["lap"])
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
    return exL, kL, dL, kR, dR, fo, (Rlr, tlr, Rrl, trl)


def fisheye_map(kL, dL, kR, scale, M, seed, all_obs):
    c1 = T["cam1"]
    mp = synth.mappoints(kL, dL, scale, M, seed=seed, width=512, height=512, fx=c1[0], fy=c1[1], cx=c1[2], cy=c1[3])
    N = len(kL) + len(kR)
    rng = np.random.default_rng(seed)
    holder = np.full(N, -1, np.int32); hobs = np.zeros(N, np.uint8)
    cl = rng.random(N) < 0.2
    holder[cl] = -2; hobs[cl] = 1
    if all_obs:
        mp["flags"] |= 2
    return mp, holder, hobs


def bow_case(desc, angle, seed, n_kf):
    """vocabulary + KeyFrame around a frame's descriptors: returns (oracle vocabulary arrays, kf_desc, kf_angle, kf_has)"""
    parent, leaf, vdesc, weight = synth.make_vocabulary(10, 3, seed=seed, stop_fraction=0.03)
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(leaf)[0]
    take = rng.choice(len(desc), size=min(len(desc), len(leaves) // 2), replace=False)
    vdesc[leaves[: len(take)]] = desc[take]
    pick = rng.integers(0, len(desc), n_kf)
    bits = np.unpackbits(desc[pick], axis=1)
    for i in range(n_kf):
        bits[i, rng.choice(256, size=int(rng.integers(0, 45)), replace=False)] ^= 1
    kf_desc = np.packbits(bits, axis=1)
    kf_angle = (angle[pick] + np.where(rng.random(n_kf) < 0.75, rng.normal(20, 4, n_kf), rng.uniform(0, 360, n_kf))
                ).astype(np.float32) % np.float32(360)
    return (parent, leaf, vdesc, weight), kf_desc, kf_angle, (rng.random(n_kf) < 0.8).astype(np.uint8)


def main():
    g = {}
    exL, exR, kL, dL, kR, dR = euroc_frame(lambda: oracle.RefExtractor())
    st = oracle.ref_stereo(exL, exR, kL, dL, kR, dR, float(MBF), float(MB))
    g["stereo_uRight"], g["stereo_depth"] = st["uRight"], st["depth"]
    scale = exL.scale_tables()["scale"]
    for M, th, seed in LOCAL_CASES:
        mp = synth.mappoints(kL, dL, scale, M, seed=seed)
        F = oracle.RefFrame(kL, dL, scale, E["width"], E["height"], cam1=CAM, mbf=float(MBF), u_right=st["uRight"])
        n, h, ho, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th, mp["holder"],
                                                 mp["holder_obs"])
        p = "local_%d_" % M
        g[p + "n"], g[p + "holder"], g[p + "holder_obs"], g[p + "track_i"], g[p + "track_f"] = np.int32(n), h, ho, ti, tf
        print("local map M=%d th=%.0f: %d matches" % (M, th, n))
    for ci, (tz, th, ori) in enumerate(LAST_CASES):
        Rcw, tcw, lf = last_frame_case(kL, dL, tz)
        F = oracle.RefFrame(kL, dL, scale, E["width"], E["height"], cam1=CAM, mbf=float(MBF), u_right=st["uRight"], Rcw=Rcw, tcw=tcw)
        N = len(kL)
        n, h, ho = F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th, np.eye(3), np.zeros(3),
                                       float(MB), np.full(N, -1, np.int32), np.zeros(N, np.uint8), False, ori)
        g["last_%d_n" % ci], g["last_%d_holder" % ci], g["last_%d_holder_obs" % ci] = np.int32(n), h, ho
        print("last frame tz=%.2f th=%.0f: %d matches" % (tz, th, n))
    # fisheye rig: local-map search with and without observation-less map points
    fexL, fkL, fdL, fkR, fdR, fo, (Rlr, tlr, Rrl, trl) = fisheye_frame()
    keys = np.vstack([fkL, fkR]); desc = np.vstack([fdL, fdR])
    for all_obs in (True, False):
        mp, holder, hobs = fisheye_map(fkL, fdL, fkR, fexL.scale, 6000, 13, all_obs)
        F = oracle.RefFrame(keys, desc, fexL.scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"],
                            n_left=len(fkL), n_right=len(fkR), l2r=fo["l2r"], r2l=fo["r2l"], Rrl=Rrl, trl=trl, tlr=tlr)
        n, h, ho, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, holder, hobs)
        p = "fisheye_%d_" % int(all_obs)
        g[p + "n"], g[p + "holder"], g[p + "holder_obs"], g[p + "track_i"], g[p + "track_f"] = np.int32(n), h, ho, ti, tf
        print("fisheye local map all_obs=%s: %d matches" % (all_obs, n))
    # Frame::ComputeStereoFishEyeMatches on the same rig (tables exact; depth / 3-D points carry the SVD stand-in's tolerance)
    fr = oracle.ref_fisheye(T["cam1"], T["cam2"], Rlr, tlr, fexL.sigma2, fkL, fdL, 0, fkR, fdR, 0)
    g["fisheye_stereo_l2r"], g["fisheye_stereo_r2l"] = fr["l2r"], fr["r2l"]
    g["fisheye_stereo_depth"], g["fisheye_stereo_p3d"] = fr["depth"], fr["p3d"]
    print("fisheye stereo: %d matches" % int((fr["l2r"] >= 0).sum()))
    # SearchByBoW, monocular-style and two-camera frames
    angle = np.ascontiguousarray(kL[:, 3])
    voc, kf_desc, kf_angle, kf_has = bow_case(dL, angle, 5, 1100)
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, *voc)
    f_node, kf_node = vo.transform(dL, 2)["node"], vo.transform(kf_desc, 2)["node"]
    for ci, (ratio, ori) in enumerate(((0.7, True), (0.75, False), (0.9, True))):
        n, m = oracle.ref_search_by_bow(kf_desc, kf_angle, kf_node, kf_has, dL, angle, f_node, -1, ratio, ori)
        g["bow_%d_n" % ci], g["bow_%d_match" % ci] = np.int32(n), m
        print("SearchByBoW ratio=%.2f ori=%s: %d matches" % (ratio, ori, n))
    f_angle = np.concatenate([fkL[:, 3], fkR[:, 3]]).astype(np.float32)
    voc, kf_desc, kf_angle, kf_has = bow_case(desc, f_angle, 9, 1500)
    vo = oracle.Vocabulary.from_arrays(10, 3, 0, 0, *voc)
    f_node, kf_node = vo.transform(desc, 2)["node"], vo.transform(kf_desc, 2)["node"]
    for ci, ori in enumerate((True, False)):
        n, m = oracle.ref_search_by_bow(kf_desc, kf_angle, kf_node, kf_has, desc, f_angle, f_node, len(fkL), 0.7, ori)
        g["bow_fisheye_%d_n" % ci], g["bow_fisheye_%d_match" % ci] = np.int32(n), m
        print("SearchByBoW two cameras ori=%s: %d matches (%d on the right)" % (ori, n, int((m[len(fkL):] >= 0).sum())))
    dst = os.path.join(ROOT, "tests", "golden", "ref_frame.npz")
    np.savez_compressed(dst, **g)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
