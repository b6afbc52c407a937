"""BASELINE config 5 at full length: a 2000-frame synthetic stereo sequence (SURVEY.md 8d config 5 shape: panned
12-layer scene, per-frame noise reseed, extract x2 -> stereo -> SearchLocalPoints per frame). The oracle cannot run
2000 frames in seconds, so the full length is covered by size-independent properties:
  * a checksum of every frame's results is identical between (a) one context, one frame at a time, host snapshot
    search and (b) three frames in flight over three contexts with the local map named as rows of the persistent
    store (rows recycled as a rolling window, upserted incrementally);
  * structural invariants of every frame (counts, index ranges, one keypoint per map point, stereo consistency);
  * every 250th frame is compared bit-exactly with the oracle (extraction, stereo, projection search)."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from sbp_check import assert_search_matches
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC
N_FRAMES = 2000
M = 3000
WINDOW = 10            # frames whose map rows stay in the store (rolling pool)


def _ctx():
    mbf = np.float32(E["fx"] * E["baseline"])
    return ft.Context(E["width"], E["height"], nfeatures=1200, nlevels=8, cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))


def _frame(sc, t):
    return sc.pair(pan=sc.sequence_pan(t), noise_seed=900000 + t)


def _local_map(keys, desc, scale, t):
    """light vectorised local map anchored on the frame's keypoints (pose = identity)"""
    rng = np.random.default_rng(77000 + t)
    N = len(keys)
    k = rng.integers(0, N, M)
    anchored = rng.random(M) < 0.6
    z = rng.uniform(0.5, 20.0, M)
    lvl = np.where(anchored, keys[k, 5].astype(np.int64), rng.integers(0, 8, M))
    u = np.where(anchored, keys[k, 0] + rng.uniform(-2, 2, M), rng.uniform(-50, E["width"] + 50, M))
    v = np.where(anchored, keys[k, 1] + rng.uniform(-2, 2, M), rng.uniform(-50, E["height"] + 50, M))
    d = np.where(anchored[:, None], desc[k] ^ np.packbits(rng.random((M, 256)) < 0.06, axis=1), rng.integers(0, 256, (M, 32), dtype=np.uint8))
    P = np.stack([(u - E["cx"]) * z / E["fx"], (v - E["cy"]) * z / E["fy"], z], 1)
    dist = np.linalg.norm(P, axis=1)
    n = -P / dist[:, None]
    n = -n                                             # normal along the viewing ray: viewCos = 1
    maxd = dist * scale[lvl] * 0.95
    flags = np.full(M, 2, np.int32)
    flags[rng.random(M) < 0.03] |= 1
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(pos=f32(P), normal=f32(n), minmax=f32(np.stack([maxd / scale[7], maxd], 1)), desc=np.ascontiguousarray(d.astype(np.uint8)), flags=flags)


def _digest(l, r, res):
    h = 0
    for a in (l["kps"], l["desc"], l["u_right"], l["depth"], r["kps"], r["desc"], np.int32(res[0]), res[1], res[2], res[3]):
        h = zlib.crc32(np.ascontiguousarray(a).tobytes(), h)
    return h


def _invariants(l, r, res, mp):
    nm, holder, hobs, best = res
    n = l["n"]
    assert 0 < n <= 1200 + 24 and 0 < r["n"] <= 1200 + 24
    ur, dp = l["u_right"], l["depth"]
    ok = ur >= 0
    assert np.all(dp[~ok] == -1) and np.all(dp[ok] > 0)
    x = l["kps"]["x"]
    assert np.all(ur[ok] <= x[ok] + 1e-3)              # disparity >= 0 (Frame.cc:975-985)
    won = holder[holder >= 0]
    assert len(won) == nm and len(np.unique(won)) == len(won) and (len(won) == 0 or won.max() < M)
    assert np.all((mp["flags"][won] & 1) == 0)         # skipped map points never win a keypoint
    assert np.all(hobs[holder >= 0] == 1)


def test_sequence_2000_frames():
    sc = synth.StereoScene(seed=5)
    seq = _ctx()
    seq.set_pose(np.eye(3), np.zeros(3))
    scale = seq.scale_tables()["scale"]
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    D = 3
    pipe = [_ctx() for _ in range(D)]
    for c in pipe:
        c.set_pose(np.eye(3), np.zeros(3))
    pipe[0].map_store_create(WINDOW * M)
    for c in pipe[1:]:
        c.map_store_attach(pipe[0])
    digests_seq, digests_pipe = [], []
    maps = {}
    total_matches = 0
    exL = exR = None
    borderline_rows, checked_exact = [], []

    def collect(t):
        c = pipe[t % D]
        l, r = c.frame_collect()
        mp = maps.pop(t)
        rows = ((t % WINDOW) * M + np.arange(M)).astype(np.int32)      # the rows of frame t-WINDOW are recycled
        pipe[(t + 1) % D].map_store_update(rows, mp["pos"], mp["normal"], mp["minmax"], mp["desc"])   # any context may upsert
        res = c.search_store(rows, mp["flags"], 3.0, np.full(l["n"], -1, np.int32), np.zeros(l["n"], np.uint8))
        digests_pipe.append(_digest(l, r, res))

    for t in range(N_FRAMES):
        L, R = _frame(sc, t)
        # (a) sequential reference run
        l, r = seq.frame_construct(L, R)
        keys = ft.keypoints_as_array(l["kps"])
        mp = _local_map(keys, l["desc"], scale, t)
        res = seq.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                      np.full(l["n"], -1, np.int32), np.zeros(l["n"], np.uint8))
        digests_seq.append(_digest(l, r, res))
        _invariants(l, r, res, mp)
        total_matches += res[0]
        if t % 250 == 0:                                # oracle on a sample of the sequence
            if exL is None:
                exL, exR = oracle.Extractor(), oracle.Extractor()
            _, kL, dL = exL.extract(L); blL = exL.desc_borderline()
            _, kR, dR = exR.extract(R); blR = exR.desc_borderline()
            assert np.array_equal(keys, kL) and np.array_equal(ft.keypoints_as_array(r["kps"]), kR)
            # north_star: descriptor rows may differ only where a rotated sample sits on a rounding boundary (counted).
            # Seen once in ~10^5 keypoints: glibc's sinf is not correctly rounded for that angle (frame 1250)
            exact = True
            for got, exp, bl in ((l["desc"], dL, blL), (r["desc"], dR, blR)):
                rows = np.nonzero((got != exp).any(axis=1))[0]
                borderline_rows.append(len(rows))
                assert len(rows) <= min(bl, 2)
                for i in rows:
                    assert int(np.unpackbits(got[i] ^ exp[i]).sum()) <= 2
                exact &= len(rows) == 0
            if exact:
                st = oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))
                assert np.array_equal(l["u_right"], st["uRight"]) and np.array_equal(l["depth"], st["depth"])
                F = oracle.Frame(kL, dL, scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                                 mbf=float(mbf), u_right=st["uRight"])
                n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                               np.full(len(kL), -1, np.int32), np.zeros(len(kL), np.uint8))
                gi, gf = seq.track(M)
                def prefix(k_):
                    h0, o0 = np.full(len(kL), -1, np.int32), np.zeros(len(kL), np.uint8)
                    o = F.search_local_points(mp["pos"][:k_], mp["normal"][:k_], mp["minmax"][:k_], mp["desc"][:k_], mp["flags"][:k_], 3.0, h0, o0)
                    g_ = seq.search_local_points(mp["pos"][:k_], mp["normal"][:k_], mp["minmax"][:k_], mp["desc"][:k_], mp["flags"][:k_], 3.0, h0, o0)
                    return g_[:3], o[:3]
                assert_search_matches(res[:3], (n_o, h_o, ho_o), gi, ti, prefix)
                checked_exact.append(t)
        # (b) pipelined run: frame t is submitted, frame t-(D-1) is collected and searched
        maps[t] = mp
        pipe[t % D].frame_submit(L, R)
        if t >= D - 1:
            collect(t - (D - 1))
    for t in range(N_FRAMES - (D - 1), N_FRAMES):
        collect(t)
    assert len(digests_pipe) == N_FRAMES and digests_pipe == digests_seq
    assert zlib.crc32(np.array(digests_seq, np.uint32).tobytes()) == zlib.crc32(np.array(digests_pipe, np.uint32).tobytes())
    assert len(checked_exact) >= 6 and sum(borderline_rows) <= 2, (checked_exact, borderline_rows)
    assert total_matches > 300 * N_FRAMES              # the search is doing real work on every frame
    for c in pipe + [seq]:
        c.close()
