"""'Next' row 3: the persistent device-side MapPoint store. A search that names its local map as rows of
the store must return exactly what ft_search_local_points returns on the gathered arrays (which test_gpu_parity.py
pins against the oracle), before and after incremental updates, and from a second context attached to the store."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle
from sbp_check import assert_search_matches
import fasttrack_b200 as ft
from fasttrack_b200 import synth

E = synth.EUROC


def _ctx():
    mbf = np.float32(E["fx"] * E["baseline"])
    return ft.Context(E["width"], E["height"], nfeatures=1200, nlevels=8, cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf))


@pytest.fixture(scope="module")
def frame(euroc_pair):
    L, R = euroc_pair
    ctx = _ctx()
    l, r = ctx.frame_construct(L, R)
    ctx.set_pose(np.eye(3), np.zeros(3))
    yield dict(ctx=ctx, L=L, R=R, kL=ft.keypoints_as_array(l["kps"]), dL=l["desc"], scale=ctx.scale_tables()["scale"], left=l)
    ctx.close()


def _same(a, b):
    return a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


def test_store_search_equals_snapshot_search_and_oracle(frame):
    ctx = frame["ctx"]
    M, CAP = 9000, 20000
    mp = synth.mappoints(frame["kL"], frame["dL"], frame["scale"], M, seed=301)
    rng = np.random.default_rng(5)
    slots = rng.permutation(CAP)[:M].astype(np.int32)          # rows scattered over the store, local-map order kept
    ctx.map_store_create(CAP)
    for a in range(0, M, 2500):                                # the mapping side upserts in batches
        b = min(M, a + 2500)
        ctx.map_store_update(slots[a:b], mp["pos"][a:b], mp["normal"][a:b], mp["minmax"][a:b], mp["desc"][a:b])
    ref = ctx.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, mp["holder"], mp["holder_obs"])
    got = ctx.search_store(slots, mp["flags"], 3.0, mp["holder"], mp["holder_obs"])
    assert _same(ref, got) and ref[0] > 100
    # against the oracle directly
    F = oracle.Frame(frame["kL"], frame["dL"], frame["scale"], E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                     mbf=float(np.float32(E["fx"] * E["baseline"])), u_right=frame["left"]["u_right"])
    n_o, h_o, ho_o, ti, tf = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0,
                                                   mp["holder"], mp["holder_obs"])
    gi, gf = ctx.track(M)
    def prefix(k):
        o = F.search_local_points(mp["pos"][:k], mp["normal"][:k], mp["minmax"][:k], mp["desc"][:k], mp["flags"][:k], 3.0,
                                  mp["holder"], mp["holder_obs"])
        g = ctx.search_store(slots[:k], mp["flags"][:k], 3.0, mp["holder"], mp["holder_obs"])
        return g[:3], o[:3]
    assert_search_matches(got[:3], (n_o, h_o, ho_o), gi, ti, prefix)

    # incremental update: LocalMapping moved / re-described 700 of the points; a sub-list of the map is searched
    ch = rng.choice(M, 700, replace=False)
    mp2 = {k: v.copy() for k, v in mp.items()}
    alt = synth.mappoints(frame["kL"], frame["dL"], frame["scale"], M, seed=302)
    for k in ("pos", "normal", "minmax", "desc"):
        mp2[k][ch] = alt[k][ch]
    ctx.map_store_update(slots[ch], mp2["pos"][ch], mp2["normal"][ch], mp2["minmax"][ch], mp2["desc"][ch])
    sub = np.sort(rng.choice(M, 6000, replace=False))
    ref2 = ctx.search_local_points(mp2["pos"][sub], mp2["normal"][sub], mp2["minmax"][sub], mp2["desc"][sub], mp2["flags"][sub], 2.0,
                                   mp["holder"], mp["holder_obs"])
    got2 = ctx.search_store(slots[sub], mp2["flags"][sub], 2.0, mp["holder"], mp["holder_obs"])
    assert _same(ref2, got2)
    assert not np.array_equal(ref2[1], ref[1])

    # a second context of the sequence attached to the same rows (two frames in flight)
    other = _ctx()
    sc = synth.StereoScene(seed=21)
    L2, R2 = sc.pair(pan=(5, 2), noise_seed=3)
    other.frame_construct(L2, R2)
    other.set_pose(np.eye(3), np.zeros(3))
    other.map_store_attach(ctx)
    ref3 = other.search_local_points(mp2["pos"], mp2["normal"], mp2["minmax"], mp2["desc"], mp2["flags"], 3.0,
                                     np.full(other.cap, -1, np.int32)[:other.counts()["n_left"]], np.zeros(other.counts()["n_left"], np.uint8))
    n0 = other.counts()["n_left"]
    got3 = other.search_store(slots, mp2["flags"], 3.0, np.full(n0, -1, np.int32), np.zeros(n0, np.uint8))
    assert _same(ref3, got3)
    # updates from the owner while the other context keeps searching: ordering is the library's job
    for it in range(5):
        ch = rng.choice(M, 300, replace=False)
        alt = synth.mappoints(frame["kL"], frame["dL"], frame["scale"], M, seed=310 + it)
        for k in ("pos", "normal", "minmax", "desc"):
            mp2[k][ch] = alt[k][ch]
        ctx.map_store_update(slots[ch], mp2["pos"][ch], mp2["normal"][ch], mp2["minmax"][ch], mp2["desc"][ch])
        g = other.search_store(slots, mp2["flags"], 3.0, np.full(n0, -1, np.int32), np.zeros(n0, np.uint8))
        r = other.search_local_points(mp2["pos"], mp2["normal"], mp2["minmax"], mp2["desc"], mp2["flags"], 3.0,
                                      np.full(n0, -1, np.int32), np.zeros(n0, np.uint8))
        assert _same(r, g), it
    # the store outlives its creator as long as a context is attached
    other.close()


def test_store_errors(frame):
    c = _ctx()
    c.frame_construct(frame["L"], frame["R"])
    n = c.counts()["n_left"]
    h, ho = np.full(n, -1, np.int32), np.zeros(n, np.uint8)
    with pytest.raises(RuntimeError, match="no map store"):
        c.search_store(np.zeros(4, np.int32), np.zeros(4, np.int32), 3.0, h, ho)
    with pytest.raises(RuntimeError, match="no map store"):
        c.map_store_update(np.zeros(1, np.int32), np.zeros((1, 3)), np.zeros((1, 3)), np.zeros((1, 2)), np.zeros((1, 32), np.uint8))
    c.map_store_create(100)
    with pytest.raises(RuntimeError, match="already has"):
        c.map_store_create(100)
    with pytest.raises(RuntimeError, match="capacity"):
        c.map_store_update(np.array([100], np.int32), np.zeros((1, 3)), np.zeros((1, 3)), np.zeros((1, 2)), np.zeros((1, 32), np.uint8))
    with pytest.raises(RuntimeError, match="capacity"):
        c.search_store(np.array([3, -1], np.int32), np.zeros(2, np.int32), 3.0, h, ho)
    # an empty local map leaves the holders alone
    nm, h2, ho2, _ = c.search_store(np.zeros(0, np.int32), np.zeros(0, np.int32), 3.0, h, ho)
    assert nm == 0 and np.array_equal(h2, h)
    c.close()


def test_split_search_equals_synchronous_search(frame):
    """ft_search_store_submit / ft_search_collect (the asynchronous halves of ft_search_store) return exactly what the
    synchronous call returns, also with the next frame submitted on another context of the sequence between the two halves
    and with upserts enqueued in front of the search (the order of the end-to-end loop in ft_sequence_driver.cpp)."""
    M, CAP = 8000, 12000
    mp = synth.mappoints(frame["kL"], frame["dL"], frame["scale"], M, seed=401)
    rng = np.random.default_rng(9)
    slots = rng.permutation(CAP)[:M].astype(np.int32)
    a, b = _ctx(), _ctx()
    try:
        for c in (a, b):
            c.set_pose(np.eye(3), np.zeros(3))
        a.frame_construct(frame["L"], frame["R"])
        a.map_store_create(CAP)
        b.map_store_attach(a)
        a.map_store_update(slots, mp["pos"], mp["normal"], mp["minmax"], mp["desc"])
        n = a.counts()["n_left"]
        h, ho = np.full(n, -1, np.int32), np.zeros(n, np.uint8)
        ref = a.search_store(slots, mp["flags"], 3.0, h, ho)
        assert ref[0] > 100
        sc = synth.StereoScene(seed=21)
        for it in range(4):
            # upserts that do not change the data, in front of the search on the same stream
            ch = rng.choice(M, 400, replace=False)
            a.map_store_update(slots[ch], mp["pos"][ch], mp["normal"][ch], mp["minmax"][ch], mp["desc"][ch])
            a.search_store_submit(slots, mp["flags"], 3.0, h, ho)
            # a second search on the same context cannot be issued before the first one has been collected
            with pytest.raises(RuntimeError, match="not been collected"):
                a.search_store_submit(slots, mp["flags"], 3.0, h, ho)
            with pytest.raises(RuntimeError, match="not been collected"):
                a.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, h, ho)
            L2, R2 = sc.pair(pan=(3 * it, it), noise_seed=it)
            b.frame_submit(L2, R2)                       # the next camera frame, on the other context
            got = a.search_collect()
            assert _same(ref, got), it
            b.frame_collect()
        with pytest.raises(RuntimeError, match="no submitted search"):
            a.search_collect()
        # without the per-point selections
        a.search_store_submit(slots, mp["flags"], 3.0, h, ho, want_best=False)
        got = a.search_collect()
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
        # empty local map: nothing is enqueued, the collect reports no matches
        a.search_store_submit(np.zeros(0, np.int32), np.zeros(0, np.int32), 3.0, h, ho)
        nm, h2, _, _ = a.search_collect()
        assert nm == 0 and np.array_equal(h2, h)
        # the synchronous call still works afterwards and a snapshot search after a split one sees a free staging buffer
        assert _same(ref, a.search_store(slots, mp["flags"], 3.0, h, ho))
        assert _same(ref, a.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], 3.0, h, ho))
        a.search_store_submit(slots, mp["flags"], 3.0, h, ho)
        assert _same(ref, a.search_collect())
    finally:
        b.close(); a.close()
