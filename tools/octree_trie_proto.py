"""Prototype of the sort-once ("trie") formulation of DistributeOctTree (reference src/ORBextractor.cc:660-884) that
k_octree implements, checked here against the oracle's list-based restatement on random inputs.

Facts used (all follow from DivideNode being a fixed spatial quadtree, ORBextractor.cc:510-566):
  * the quadrant a keypoint falls into at depth d depends only on its pixel and its root -> every candidate has a
    path key (root, q1, q2, ..., qD) that can be computed up front;
  * a normal pass splits EVERY node holding more than one keypoint, so after pass k the list holds the non-empty
    depth-k nodes plus the single-keypoint nodes that settled earlier; sizes and nToExpand per pass follow from the
    common-prefix lengths of neighbouring sorted keys;
  * push_front of n1..n4 while walking the list front to back reverses the order at every depth:
    order_k = (reverse order_{k-1} of the parent, quadrant descending). With Q = path key whose even-depth digits are
    complemented, the list order of the depth-k nodes is ascending Q for even k and descending Q for odd k;
  * the careful phase (size + 3*nToExpand > N) works on node records (lo, hi, depth) into the Q-sorted array.
Run: python tools/octree_trie_proto.py [trials]
"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

D = 12


def _stdsort():
    so = os.path.join(tempfile.gettempdir(), "ft_sort_harness.so")
    if not os.path.exists(so):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "native", "sort_harness.cpp")])
    L = ctypes.CDLL(so)
    L.harness_stdsort.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS"), ctypes.c_int]
    return L


def path_key(x, y, W, H, nIni, hX):
    r = int(np.float32(x) / hX)
    if r >= nIni:
        r = nIni - 1
    x0 = int(np.float32(hX * np.float32(r))); x1 = int(np.float32(hX * np.float32(r + 1))); y0 = 0; y1 = H
    key = r
    for d in range(1, D + 1):
        mx = x0 + ((x1 - x0 + 1) >> 1); my = y0 + ((y1 - y0 + 1) >> 1)
        qx = 1 if x >= mx else 0; qy = 1 if y >= my else 0
        q = qx | (qy << 1)
        if qx: x0 = mx
        else: x1 = mx
        if qy: y0 = my
        else: y1 = my
        key = (key << 2) | ((3 - q) if d % 2 == 0 else q)
    return key


def node_bounds(key_prefix, depth, W, H, nIni, hX):
    """bounds of the node whose Q-prefix (root + depth digits) is key_prefix"""
    r = key_prefix >> (2 * depth)
    x0 = int(np.float32(hX * np.float32(r))); x1 = int(np.float32(hX * np.float32(r + 1))); y0 = 0; y1 = H
    for d in range(1, depth + 1):
        dq = (key_prefix >> (2 * (depth - d))) & 3
        q = (3 - dq) if d % 2 == 0 else dq
        mx = x0 + ((x1 - x0 + 1) >> 1); my = y0 + ((y1 - y0 + 1) >> 1)
        if q & 1: x0 = mx
        else: x1 = mx
        if q & 2: y0 = my
        else: y1 = my
    return x0, y0, x1, y1


def octree_trie(xyr, minX, maxX, minY, maxY, N, stdsort):
    C = len(xyr)
    if C == 0:
        return np.zeros((0, 3), np.float32)
    W = maxX - minX; H = maxY - minY
    nIni = int(np.round(np.float32(W) / np.float32(H)))   # std::round == np.round except at .5; checked below
    f = float(np.float32(W) / np.float32(H))
    nIni = int(np.floor(f + 0.5))
    hX = np.float32(W) / np.float32(nIni)
    keys = np.array([path_key(int(p[0]), int(p[1]), W, H, nIni, hX) for p in xyr], np.int64)
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    # common leading digits (root counts as one digit) between neighbours: cl in [0, D+1]
    def common(a, b):
        c = 0
        for j in range(D + 1):
            sh = 2 * (D - j)
            if (a >> sh) != (b >> sh):
                break
            c += 1
        return c
    cl = np.array([common(int(sk[i]), int(sk[i + 1])) for i in range(C - 1)], np.int64)
    clL = np.concatenate([[0], cl]); clR = np.concatenate([cl, [0]])     # neighbours to the left / right of element i
    settle = np.maximum(clL, clR)                                           # depth at which element i is alone
    def size(k):
        return 1 + int((cl <= k).sum())
    def nexp(k):   # depth-k runs with >= 2 elements
        starts = (clL <= k) & (clR >= k + 1)
        return int(starts.sum())
    # ---- normal passes, closed form ----
    K = 1
    mode = None
    while True:
        if K > D:
            mode = "finish"; K = D; break
        prev, cur = size(K - 1), size(K)
        if cur >= N or cur == prev:
            mode = "finish"; break
        if cur + 3 * nexp(K) > N:
            mode = "careful"; break
        K += 1
    # ---- list after pass K: nodes as (lo, hi, depth) over the sorted array ----
    def group_nodes(k):
        """depth-k nodes created in pass k (elements not settled before k), in list order"""
        heads = [i for i in range(C) if clL[i] <= k and settle[i] >= k] if k > 0 else [i for i in range(C) if clL[i] <= 0]
        nodes = []
        for h in heads:
            e = h + 1
            while e < C and clL[e] >= k + 1:
                e += 1
            nodes.append((h, e, k))
        if k % 2 == 1:
            nodes.reverse()
        return nodes
    nodes = group_nodes(K)
    for j in range(K - 1, -1, -1):
        singles = [(i, i + 1, j) for i in range(C) if settle[i] == j]
        if j % 2 == 1:
            singles.reverse()
        nodes += singles
    assert len(nodes) == size(K), (len(nodes), size(K))
    if mode == "careful":
        # vec: depth-K nodes with more than one keypoint in creation order = reverse list order
        vec = [n for n in nodes if n[2] == K and n[1] - n[0] > 1][::-1]
        finish = False
        while not finish:
            prevSize = len(nodes)
            m = len(vec)
            # std::sort by (size, UL.x)
            arr = np.zeros(m, np.uint64)
            for i, (lo, hi, d) in enumerate(vec):
                x0, _, _, _ = node_bounds(int(sk[lo]) >> (2 * (D - d)), d, W, H, nIni, hX)
                arr[i] = (np.uint64(((hi - lo) << 12) | x0) << np.uint64(32)) | np.uint64(i)
            stdsort.harness_stdsort(arr, m)
            newvec = []
            front = []          # groups of children, later processed in front
            removed = set()
            n = len(nodes)
            for r in range(m):
                lo, hi, d = vec[int(arr[m - 1 - r]) & 0xFFFFFFFF]
                removed.add((lo, hi, d))
                # children by actual quadrant 0..3
                sh = 2 * (D - (d + 1))
                ch = []
                for q in range(4):
                    dq = (3 - q) if (d + 1) % 2 == 0 else q
                    idx = [i for i in range(lo, hi) if ((int(sk[i]) >> sh) & 3) == dq]
                    if idx:
                        ch.append((idx[0], idx[-1] + 1, d + 1))
                        assert idx[-1] + 1 - idx[0] == len(idx)
                for c in ch:
                    if c[1] - c[0] > 1:
                        newvec.append(c)
                front = ch[::-1] + front
                n += len(ch) - 1
                if n >= N:
                    break
            nodes = front + [nd for nd in nodes if nd not in removed]
            assert len(nodes) == n
            vec = newvec
            if n >= N or n == prevSize:
                finish = True
    out = []
    for lo, hi, d in nodes:
        best = None
        for i in range(lo, hi):
            c = int(order[i])
            kk = (float(xyr[c][2]), -c)
            if best is None or kk > best[0]:
                best = (kk, c)
        out.append(xyr[best[1]])
    return np.array(out, np.float32).reshape(-1, 3)


def main():
    import oracle
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    ex = oracle.Extractor()
    S = _stdsort()
    rng = np.random.default_rng(1)
    bad = 0
    for t in range(trials):
        w = int(rng.integers(60, 1300)); h = int(rng.integers(60, 800))
        if w / h < 0.5:
            continue
        C = int(rng.choice([1, 2, 3, 10, 100, 500, 2000, 5000]))
        mode = rng.integers(0, 3)
        if mode == 0:
            xs = rng.integers(0, w, C); ys = rng.integers(0, h, C)
        elif mode == 1:   # clustered
            cx = rng.integers(0, w, 6); cy = rng.integers(0, h, 6)
            k = rng.integers(0, 6, C)
            xs = np.clip(cx[k] + rng.normal(0, 12, C).astype(int), 0, w - 1); ys = np.clip(cy[k] + rng.normal(0, 12, C).astype(int), 0, h - 1)
        else:             # lattice (many equal sizes -> sort ties)
            step = int(rng.integers(2, 9))
            gx, gy = np.meshgrid(np.arange(0, w, step), np.arange(0, h, step))
            sel = rng.permutation(gx.size)[:C]
            xs = gx.ravel()[sel]; ys = gy.ravel()[sel]
        pix = np.unique(np.stack([xs, ys], 1), axis=0)
        pix = pix[rng.permutation(len(pix))]
        resp = rng.integers(7, 40 if rng.random() < 0.5 else 255, len(pix))
        xyr = np.concatenate([pix, resp[:, None]], 1).astype(np.float32)
        N = int(rng.choice([1, 5, 30, 72, 105, 261, 700, 3000]))
        ref = ex.octree(xyr, 16, 16 + w, 16, 16 + h, N)
        got = octree_trie(xyr, 16, 16 + w, 16, 16 + h, N, S)
        ok = ref.shape == got.shape and np.array_equal(ref, got)
        if not ok:
            bad += 1
            print("MISMATCH trial", t, "w,h", w, h, "C", len(xyr), "N", N, "mode", mode, ref.shape, got.shape)
    print("trials", trials, "mismatches", bad)


if __name__ == "__main__":
    main()
