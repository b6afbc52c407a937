"""'Next' row 1 of SURVEY.md section 8f: ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), the
frame-to-last-frame search of TrackWithMotionModel, incl. the rotation-histogram check (ComputeThreeMaxima)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC


@pytest.fixture(scope="module")
def cur(euroc_pair):
    L, R = euroc_pair
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    ctx = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))
    left, right = ctx.frame_construct(L, R)
    exL, exR = oracle.Extractor(), oracle.Extractor()
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
    yield dict(ctx=ctx, kL=kL, dL=dL, scale=exL.scale, st=st, mbf=mbf)
    ctx.close()


@pytest.mark.parametrize("tz,th,check_ori", [(0.0, 7.0, True), (0.4, 7.0, True), (-0.4, 15.0, True), (0.05, 14.0, False)])
def test_last_frame_search_bit_exact(cur, tz, th, check_ori):
    ctx, kL, dL = cur["ctx"], cur["kL"], cur["dL"]
    a = 0.02
    Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    tcw = np.array([0.03, -0.02, -tz], np.float32)       # camera centre moved by +tz along the optical axis
    lf = synth.last_frame_points(kL, dL, 1000, seed=int(100 + 10 * tz), Rcw=Rcw, tcw=tcw, fx=E["fx"], fy=E["fy"], cx=E["cx"], cy=E["cy"])
    F = oracle.Frame(kL, dL, cur["scale"], E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                     mbf=float(cur["mbf"]), u_right=cur["st"]["uRight"], Rcw=Rcw, tcw=tcw)
    Rlw = np.eye(3, dtype=np.float32); tlw = np.zeros(3, np.float32)
    mb = np.float32(cur["mbf"] / np.float32(E["fx"]))
    tlc_z = float((Rlw @ F.Ow + tlw)[2])
    direction = 1 if tlc_z > mb else (-1 if -tlc_z > mb else 0)
    assert direction == (1 if tz > 0.2 else (-1 if tz < -0.2 else 0))
    N = len(kL)
    holder0 = np.full(N, -1, np.int32); hobs0 = np.zeros(N, np.uint8)
    n_o, h_o, ho_o, bl = F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th, direction,
                                             holder0, hobs0, check_ori)
    ctx.set_pose(Rcw, tcw, F.Rwc, F.Ow)
    n_g, h_g, ho_g, best = ctx.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], Rlw, tlw, th,
                                                 holder0, hobs0, b_mono=False, check_ori=check_ori)
    assert n_o > 150
    assert bl.sum() < 5
    if bl.sum() == 0:
        assert n_g == n_o and np.array_equal(h_g, h_o) and np.array_equal(ho_g, ho_o)
    else:
        assert abs(n_g - n_o) <= 3


def test_last_frame_search_fisheye():
    T = synth.TUMVI
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    ctx = ft.Context(512, 512, nfeatures=1000, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=T["lap"], lap_right=T["lap"],
                     bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    left, right = ctx.frame_construct(L, R)
    exL, exR = oracle.Extractor(1000), oracle.Extractor(1000)
    mL, kL, dL = exL.extract(L, lap=T["lap"]); mR, kR, dR = exR.extract(R, lap=T["lap"])
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
    keys = np.vstack([kL, kR]); desc = np.vstack([dL, dR])
    Rcw = np.eye(3, dtype=np.float32); tcw = np.array([0.01, 0.0, 0.02], np.float32)
    c1 = T["cam1"]
    lf = synth.last_frame_points(kL, dL, 800, seed=5, Rcw=Rcw, tcw=tcw, fx=c1[0], fy=c1[1], cx=c1[2], cy=c1[3], kb8=c1)
    lf["flags"] |= 2
    F = oracle.Frame(keys, desc, exL.scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"], n_left=len(kL),
                     n_right=len(kR), l2r=fo["l2r"], r2l=fo["r2l"], Rcw=Rcw, tcw=tcw, Rrl=Rrl, trl=trl, tlr=tlr)
    N = len(keys)
    holder0 = np.full(N, -1, np.int32); hobs0 = np.zeros(N, np.uint8)
    n_o, h_o, ho_o, bl = F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], 7.0, 0, holder0, hobs0, True)
    ctx.set_pose(Rcw, tcw, F.Rwc, F.Ow)
    n_g, h_g, ho_g, _ = ctx.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], np.eye(3), np.zeros(3), 7.0,
                                              holder0, hobs0, b_mono=False, check_ori=True)
    assert n_o > 100
    # KB8 projection goes through atan2f/cosf/sinf (device vs glibc last-ulp differences): windows are radius >= 7 px,
    # so decisions agree unless a keypoint sits within ~1e-4 px of a window edge
    same = np.array_equal(h_g, h_o)
    assert same or (np.mean(h_g != h_o) < 0.01 and abs(n_g - n_o) <= 3)
    ctx.close()


def test_last_frame_fisheye_empty_left_window_skips_right_eye():
    """`if(vIndices2.empty()) continue;` (ORBmatcher.cc:1836-1837) stands in front of the right-eye block (:1915): a
    last-frame point whose LEFT window holds no keypoint is never searched in the right image, even when right keypoints
    lie inside its right window. Points are planted where the left image has no keypoints (the 19-px border) but whose
    right projection lands on right keypoints, carrying those keypoints' descriptors; holders must be equal, exactly."""
    T = synth.TUMVI
    L, R = synth.fisheye_pair(seed=3)
    Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
    ctx = ft.Context(512, 512, nfeatures=1000, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=T["lap"], lap_right=T["lap"],
                     bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]))
    ctx.frame_construct(L, R)
    exL, exR = oracle.Extractor(1000), oracle.Extractor(1000)
    mL, kL, dL = exL.extract(L, lap=T["lap"]); mR, kR, dR = exR.extract(R, lap=T["lap"])
    fo = oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL, mL, kR, dR, mR)
    keys = np.vstack([kL, kR]); desc = np.vstack([dL, dR])
    Rcw = np.eye(3, dtype=np.float32); tcw = np.zeros(3, np.float32)
    c1 = np.asarray(T["cam1"], np.float64)
    th = 7.0
    # for every right keypoint: the camera-1 point (several depths) whose right-eye projection (Trl * x, LEFT camera model,
    # as the reference does) is that keypoint; keep those whose LEFT projection falls where no left keypoint is within reach
    pts, dsc, octs, angs = [], [], [], []
    for z in (0.6, 1.0, 2.0, 5.0):
        ray = synth.kb8_unproject(c1, kR[:, 0].astype(np.float64), kR[:, 1].astype(np.float64))
        Pr = ray * z / np.maximum(ray[:, 2:3], 1e-9)
        Pc = (Pr - trl.astype(np.float64)) @ Rrl.astype(np.float64)       # x3Dc = Rrl^T (x3Dr - trl)
        ul, vl = synth.kb8_project(c1, Pc)
        inside = (ul > 0.5) & (ul < 511.5) & (vl > 0.5) & (vl < 511.5) & (Pc[:, 2] > 0)
        for j in np.nonzero(inside)[0]:
            o = int(kR[j, 5])
            rad = th * float(exL.scale[o]) + 1.0
            near = (np.abs(kL[:, 0] - ul[j]) < rad) & (np.abs(kL[:, 1] - vl[j]) < rad)
            if not near.any():
                pts.append(Pc[j]); dsc.append(dR[j]); octs.append(o); angs.append(kR[j, 3])
    assert len(pts) >= 5, "the synthetic pair offers too few right keypoints whose left window is empty"
    n = len(pts)
    lf = dict(pos=np.ascontiguousarray(pts, np.float32), desc=np.ascontiguousarray(dsc, np.uint8), octave=np.asarray(octs, np.int32),
              angle=np.asarray(angs, np.float32), flags=np.full(n, 2, np.int32))
    F = oracle.Frame(keys, desc, exL.scale, 512, 512, cam_type=1, cam1=T["cam1"], cam2=T["cam2"], mbf=T["bf"], n_left=len(kL),
                     n_right=len(kR), l2r=fo["l2r"], r2l=fo["r2l"], Rcw=Rcw, tcw=tcw, Rrl=Rrl, trl=trl, tlr=tlr)
    N = len(keys)
    holder0 = np.full(N, -1, np.int32); hobs0 = np.zeros(N, np.uint8)
    n_o, h_o, ho_o, bl = F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], th, 0, holder0, hobs0, False)
    ctx.set_pose(Rcw, tcw, F.Rwc, F.Ow)
    n_g, h_g, ho_g, _ = ctx.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], np.eye(3), np.zeros(3), th,
                                              holder0, hobs0, b_mono=False, check_ori=False)
    assert n_o == 0 and not (h_o >= 0).any(), "the planted points must stay unmatched in the reference's order of tests"
    assert n_g == n_o and np.array_equal(h_g, h_o) and np.array_equal(ho_g, ho_o)
    ctx.close()
