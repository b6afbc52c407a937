"""Deterministic synthetic inputs for the tracking front-end.

All generators are seeded numpy; no files, no network. Shapes follow the reference's
dataset configs: EuRoC pinhole 752x480 (Examples/Stereo/EuRoC.yaml:23-26,43-44) and
TUM-VI KannalaBrandt8 512x512 (Examples/Stereo-Inertial/TUM-VI.yaml:11-49).
"""
import numpy as np

EUROC = dict(width=752, height=480, fx=458.654, fy=457.296, cx=367.215, cy=248.375, baseline=0.110074,
             nfeatures=1200, nlevels=8, scale=1.2, ini_th=20, min_th=7)

# Examples/Stereo-Inertial/TUM-VI.yaml (camera 1 / camera 2 intrinsics, T_c1_c2, overlap)
TUMVI = dict(
    width=512, height=512,
    cam1=[190.978477, 190.973307, 254.931706, 256.897442, 0.0034823894022493434, 0.0007150348452162257,
          -0.0020532361418706202, 0.00020293673591811182],
    cam2=[190.442369, 190.434438, 252.597254, 254.917234, 0.0034003170790442797, 0.001766278153469831,
          -0.00266312569781606, 0.0003299517423931039],
    T_c1_c2=[[0.9999994317, 0.0008361597, 0.0006758557, -0.1010596120],
             [-0.0008042732, 0.9989843, -0.0450513, -0.0019463],
             [-0.0007128, 0.0450507, 0.9989844, -0.0015185]],
    lap=(0, 511), nfeatures=1000, nlevels=8, scale=1.2, ini_th=20, min_th=7, bf=19.3079)


def _orthonormalize(R):
    u, _, vt = np.linalg.svd(np.asarray(R, np.float64))
    return u @ vt


def tumvi_extrinsics():
    """returns Rlr(3x3), tlr(3), Rrl, trl as float32 (Tlr = T_c1_c2, a proper SE3)"""
    T = np.asarray(TUMVI["T_c1_c2"], np.float64)
    R = _orthonormalize(T[:, :3]); t = T[:, 3]
    Rrl = R.T; trl = -Rrl @ t
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return f(R), f(t), f(Rrl), f(trl)


def texture(h, w, seed, n_rect=6000, n_disc=3000, noise=3):
    """Config-1 texture: random rectangles + discs (painter's order), 3x3 box blur, uniform noise."""
    rng = np.random.default_rng(seed)
    area = (w * h) / (752.0 * 480.0)
    n_rect = max(8, int(n_rect * area)); n_disc = max(4, int(n_disc * area))
    img = np.full((h, w), 128, np.float32)
    n = n_rect + n_disc
    kind = np.zeros(n, np.int32); kind[n_rect:] = 1
    rng.shuffle(kind)
    cx = rng.integers(0, w, n); cy = rng.integers(0, h, n)
    sx = rng.integers(3, 41, n); sy = rng.integers(3, 41, n)
    gray = rng.integers(0, 256, n)
    for i in range(n):
        if kind[i] == 0:
            x0, x1 = max(cx[i] - sx[i] // 2, 0), min(cx[i] + sx[i] // 2 + 1, w)
            y0, y1 = max(cy[i] - sy[i] // 2, 0), min(cy[i] + sy[i] // 2 + 1, h)
            img[y0:y1, x0:x1] = gray[i]
        else:
            r = sx[i] // 2 + 1
            x0, x1 = max(cx[i] - r, 0), min(cx[i] + r + 1, w)
            y0, y1 = max(cy[i] - r, 0), min(cy[i] + r + 1, h)
            if x1 <= x0 or y1 <= y0:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1]
            m = (xx - cx[i]) ** 2 + (yy - cy[i]) ** 2 <= r * r
            img[y0:y1, x0:x1][m] = gray[i]
    p = np.pad(img, 1, mode="edge")
    img = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9.0
    img = img + rng.integers(-noise, noise + 1, (h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _shift_left(a, d):
    """resample a[y, x + d] with bilinear interpolation along x (fractional d >= 0), edge-clamped"""
    h, w = a.shape
    x = np.arange(w, dtype=np.float64) + d
    x0 = np.floor(x).astype(np.int64); f = (x - x0).astype(np.float32)
    x0c = np.clip(x0, 0, w - 1); x1c = np.clip(x0 + 1, 0, w - 1)
    return a[:, x0c] * (1 - f) + a[:, x1c] * f


class StereoScene:
    """Config-2/5 scene: 12 fronto-parallel textured layers with log-spaced disparities."""

    def __init__(self, seed=2, width=752, height=480, n_layers=12, dmin=1.5, dmax=64.0, margin_x=128, margin_y=16):
        self.w, self.h = width, height
        self.mx, self.my = margin_x, margin_y
        cw, ch = width + 2 * margin_x, height + 2 * margin_y
        rng = np.random.default_rng(seed)
        self.disp = np.exp(np.linspace(np.log(dmin), np.log(dmax), n_layers))
        left = np.zeros((ch, cw), np.float32); right = np.zeros((ch, cw), np.float32)
        for k in range(n_layers):
            tex = texture(ch, cw, seed * 1000 + k).astype(np.float32)
            if k == 0:
                mask = np.ones((ch, cw), np.float32)
            else:
                mask = np.zeros((ch, cw), np.float32)
                target = cw * ch / n_layers
                while mask.sum() < target:
                    bw, bh = rng.integers(40, 200), rng.integers(40, 160)
                    x0, y0 = rng.integers(0, cw - 20), rng.integers(0, ch - 20)
                    mask[y0:y0 + bh, x0:x0 + bw] = 1
            left = np.where(mask > 0, tex, left)
            ts, ms = _shift_left(tex, self.disp[k]), _shift_left(mask, self.disp[k])
            right = right * (1 - ms) + ts * ms
        self.left, self.right = left, right

    def pair(self, pan=(0, 0), noise_seed=None, noise=2):
        px = int(np.clip(round(pan[0]), -self.mx + 64, self.mx - 64)) + self.mx
        py = int(np.clip(round(pan[1]), -self.my, self.my)) + self.my
        L = self.left[py:py + self.h, px:px + self.w]
        R = self.right[py:py + self.h, px:px + self.w]
        if noise_seed is not None:
            rng = np.random.default_rng(noise_seed)
            L = L + rng.integers(-noise, noise + 1, L.shape)
            R = R + rng.integers(-noise, noise + 1, R.shape)
        f = lambda a: np.ascontiguousarray(np.clip(np.rint(a), 0, 255).astype(np.uint8))
        return f(L), f(R)

    def sequence_pan(self, t):
        return (40.0 * np.sin(2 * np.pi * t / 200.0), 12.0 * np.sin(2 * np.pi * t / 130.0))


def kb8_project(cam, P):
    """KannalaBrandt8::project, vectorised (float64)"""
    x, y, z = P[..., 0], P[..., 1], P[..., 2]
    theta = np.arctan2(np.sqrt(x * x + y * y), z)
    psi = np.arctan2(y, x)
    t2 = theta * theta
    r = theta * (1 + t2 * (cam[4] + t2 * (cam[5] + t2 * (cam[6] + t2 * cam[7]))))
    return cam[0] * r * np.cos(psi) + cam[2], cam[1] * r * np.sin(psi) + cam[3]


def kb8_unproject(cam, u, v):
    """KannalaBrandt8::unproject, vectorised Newton iteration (float64); returns rays with z = 1"""
    px, py = (u - cam[2]) / cam[0], (v - cam[3]) / cam[1]
    td = np.clip(np.sqrt(px * px + py * py), 1e-9, np.pi / 2)
    th = td.copy()
    for _ in range(12):
        t2 = th * th
        f = th * (1 + t2 * (cam[4] + t2 * (cam[5] + t2 * (cam[6] + t2 * cam[7])))) - td
        df = 1 + t2 * (3 * cam[4] + t2 * (5 * cam[5] + t2 * (7 * cam[6] + t2 * 9 * cam[7])))
        th = th - f / df
    sc = np.tan(th) / td
    return np.stack([px * sc, py * sc, np.ones_like(px)], -1)


def fisheye_pair(seed=3, size=512):
    """Config-3 pair, geometrically consistent with the TUM-VI KannalaBrandt8 rig: the left image is a config-1
    texture; the right image is rendered by casting every right pixel's ray onto a piecewise-constant depth map
    (0.6-5 m blocks, defined in the right camera frame), moving the point into the left camera with T_c1_c2 and
    sampling the left image bilinearly at its KB8 projection."""
    rng = np.random.default_rng(seed)
    left = texture(size, size, seed * 1000 + 1).astype(np.float64)
    depth = np.full((size, size), 3.0)
    for _ in range(40):
        bw, bh = rng.integers(60, 200), rng.integers(60, 200)
        x0, y0 = rng.integers(0, size - 30), rng.integers(0, size - 30)
        depth[y0:y0 + bh, x0:x0 + bw] = rng.uniform(0.6, 5.0)
    cam1 = np.asarray(TUMVI["cam1"], np.float64); cam2 = np.asarray(TUMVI["cam2"], np.float64)
    Rlr, tlr, _, _ = tumvi_extrinsics()
    vv, uu = np.mgrid[0:size, 0:size].astype(np.float64)
    ray2 = kb8_unproject(cam2, uu, vv)
    P2 = ray2 * depth[..., None]
    P1 = P2 @ Rlr.astype(np.float64).T + tlr.astype(np.float64)
    u1, v1 = kb8_project(cam1, P1)
    x0 = np.clip(np.floor(u1).astype(np.int64), 0, size - 2); y0 = np.clip(np.floor(v1).astype(np.int64), 0, size - 2)
    fx = np.clip(u1 - x0, 0, 1); fy = np.clip(v1 - y0, 0, 1)
    right = (left[y0, x0] * (1 - fx) * (1 - fy) + left[y0, x0 + 1] * fx * (1 - fy) + left[y0 + 1, x0] * (1 - fx) * fy +
             left[y0 + 1, x0 + 1] * fx * fy)
    right = right + rng.integers(-2, 3, right.shape)
    f = lambda a: np.ascontiguousarray(np.clip(np.rint(a), 0, 255).astype(np.uint8))
    return f(left), f(right)


def mappoints(keys, desc, scale_factors, M, seed=4, width=752, height=480, fx=458.654, fy=457.296, cx=367.215,
              cy=248.375, frac_inside=0.7, claimed_frac=0.25):
    """Config-4 local map: M MapPoints against a frame at identity pose.

    returns dict(pos[M,3], normal[M,3], minmax[M,2], desc[M,32], flags[M], holder[N], holder_obs[N])
    flags: bit0 = skip (bad / already matched), bit1 = Observations() > 0.
    """
    rng = np.random.default_rng(seed)
    N = len(keys)
    nl = len(scale_factors)
    pos = np.zeros((M, 3), np.float32); normal = np.zeros((M, 3), np.float32)
    minmax = np.zeros((M, 2), np.float32); d = np.zeros((M, 32), np.uint8)
    inside = rng.random(M) < frac_inside
    anchored = rng.random(M) < 0.5
    for i in range(M):
        z = rng.uniform(0.4, 25.0)
        lvl = int(rng.integers(0, nl))
        if inside[i] and anchored[i] and N > 0:
            k = int(rng.integers(0, N))
            lvl = int(keys[k, 5])
            jit = 1.5 * scale_factors[lvl]
            u = keys[k, 0] + rng.uniform(-jit, jit); v = keys[k, 1] + rng.uniform(-jit, jit)
            bits = np.unpackbits(desc[k])
            nflip = int(rng.integers(0, 49))
            flip = rng.choice(256, nflip, replace=False)
            bits[flip] ^= 1
            d[i] = np.packbits(bits)
        else:
            u = rng.uniform(0, width); v = rng.uniform(0, height)
            d[i] = rng.integers(0, 256, 32, dtype=np.uint8)
        P = np.array([(u - cx) * z / fx, (v - cy) * z / fy, z])
        dist = np.linalg.norm(P)
        n = P / dist + 0.26 * rng.standard_normal(3)
        n /= np.linalg.norm(n)
        maxd = dist * scale_factors[lvl] * 0.95
        mind = maxd / scale_factors[nl - 1]
        if not inside[i]:
            mode = int(rng.integers(0, 4))
            if mode == 0:
                P[2] = -P[2]
            elif mode == 1:
                P[0] += (2.0 * width / fx) * z * (1 if rng.random() < 0.5 else -1)
            elif mode == 2:
                maxd = dist * 0.3; mind = maxd / scale_factors[nl - 1]
            else:
                n = -n
        pos[i] = P; normal[i] = n; minmax[i] = (mind, maxd)
    flags = np.full(M, 2, np.int32)
    flags[rng.random(M) < 0.03] = 0          # Observations() == 0
    flags[rng.random(M) < 0.02] |= 1         # skipped (bad / matched already)
    holder = np.full(N, -1, np.int32); holder_obs = np.zeros(N, np.uint8)
    cl = rng.random(N) < claimed_frac
    holder[cl] = -2
    holder_obs[cl] = (rng.random(int(cl.sum())) < 0.9).astype(np.uint8)
    return dict(pos=pos, normal=normal, minmax=minmax, desc=d, flags=flags, holder=holder, holder_obs=holder_obs)


def last_frame_points(keys, desc, n, seed, Rcw, tcw, fx, fy, cx, cy, nlevels=8, kb8=None):
    """Synthetic input of the frame-to-last-frame search (TrackWithMotionModel): n last-frame keypoints holding map
    points that re-project (under the CURRENT pose Rcw, tcw) close to current keypoints.

    returns dict(pos[n,3], desc[n,32], octave[n], angle[n], flags[n])"""
    rng = np.random.default_rng(seed)
    N = len(keys)
    k = rng.integers(0, N, n)
    z = rng.uniform(1.0, 20.0, n)
    jit = rng.uniform(-2.0, 2.0, (n, 2)) * (1.0 + keys[k, 5:6] * 0.3)
    u = keys[k, 0] + jit[:, 0]; v = keys[k, 1] + jit[:, 1]
    if kb8 is None:
        Pc = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    else:
        ray = kb8_unproject(np.asarray(kb8, np.float64), u.astype(np.float64), v.astype(np.float64))
        Pc = ray * z[:, None]
    Rwc = np.asarray(Rcw, np.float64).T
    P = (Pc - np.asarray(tcw, np.float64)) @ Rwc.T
    bits = np.unpackbits(desc[k], axis=1)
    nflip = rng.integers(0, 41, n)
    flip = rng.random((n, 256)).argsort(axis=1) < nflip[:, None]
    d = np.packbits(bits ^ flip.astype(np.uint8), axis=1)
    rnd = rng.random(n) < 0.15
    d[rnd] = rng.integers(0, 256, (int(rnd.sum()), 32), dtype=np.uint8)
    octave = np.clip(keys[k, 5].astype(np.int32) + rng.choice([0, 0, 0, 1, -1], n), 0, nlevels - 1).astype(np.int32)
    angle = (keys[k, 3] + rng.normal(0, 3.0, n)).astype(np.float32)
    wild = rng.random(n) < 0.2
    angle[wild] = rng.uniform(0, 360, int(wild.sum()))
    angle = np.mod(angle, 360.0).astype(np.float32)
    flags = np.full(n, 2, np.int32)
    flags[rng.random(n) < 0.05] = 0
    flags[rng.random(n) < 0.05] |= 1
    return dict(pos=np.ascontiguousarray(P, np.float32), desc=np.ascontiguousarray(d), octave=octave, angle=angle, flags=flags)


# ---- synthetic DBoW2 vocabulary (the real ORBvoc.txt is a 145 MB file that is not in the image) ----
def make_vocabulary(k=10, L=4, seed=7, stop_fraction=0.02):
    """A k-ary, depth-L vocabulary tree in DBoW2's text-file node order (depth first, as saveToTextFile writes a tree
    built by HKmeansStep): returns (parent[n], is_leaf[n], desc[n,32], weight[n]) for nodes 1..n (node 0 = root).
    A child's descriptor is its parent's with 96 >> level random bits flipped, so descents are decided by real
    Hamming margins and ties occur; weights look like idf values, a few words are stopped (weight 0)."""
    rng = np.random.default_rng(seed)
    parent, leaf, desc, weight = [], [], [], []

    def grow(pid, pdesc, level):
        for _ in range(k):
            d = pdesc.copy()
            flips = rng.choice(256, size=max(96 >> (level - 1), 3), replace=False)
            bits = np.unpackbits(d)
            bits[flips] ^= 1
            d = np.packbits(bits)
            parent.append(pid); desc.append(d)
            nid = len(parent)
            if level == L:
                leaf.append(1)
                weight.append(0.0 if rng.random() < stop_fraction else float(rng.uniform(0.5, 9.0)))
            else:
                leaf.append(0); weight.append(0.0)
                grow(nid, d, level + 1)

    grow(0, rng.integers(0, 256, 32, dtype=np.uint8), 1)
    return (np.asarray(parent, np.int32), np.asarray(leaf, np.uint8), np.asarray(desc, np.uint8).reshape(-1, 32),
            np.asarray(weight, np.float64))


def make_vocabulary_bfs(k=10, L=6, seed=7, stop_fraction=0.02):
    """Same kind of tree as make_vocabulary, generated level by level with numpy (seconds for ORBvoc's k = 10, L = 6) and
    numbered breadth first (parents still precede their children, which is all the array constructor requires)."""
    rng = np.random.default_rng(seed)
    parents, leaves, descs, weights = [], [], [], []
    prev_desc = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    prev_first = 0                                   # node id of the first node of the previous level (root = 0)
    next_id = 1
    for level in range(1, L + 1):
        n = len(prev_desc) * k
        d = np.repeat(prev_desc, k, axis=0)
        nflip = max(96 >> (level - 1), 3)
        pos = rng.integers(0, 256, (n, nflip))
        rows = np.arange(n)
        for j in range(nflip):
            d[rows, pos[:, j] >> 3] ^= (1 << (pos[:, j] & 7)).astype(np.uint8)
        parents.append(prev_first + np.arange(n) // k)
        is_leaf = level == L
        leaves.append(np.full(n, 1 if is_leaf else 0, np.uint8))
        w = np.zeros(n)
        if is_leaf:
            w = rng.uniform(0.5, 9.0, n)
            w[rng.random(n) < stop_fraction] = 0.0
        weights.append(w); descs.append(d)
        prev_desc, prev_first, next_id = d, next_id, next_id + n
    return (np.concatenate(parents).astype(np.int32), np.concatenate(leaves), np.ascontiguousarray(np.concatenate(descs)),
            np.concatenate(weights).astype(np.float64))


def write_vocabulary_text(path, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0, trailing_newline=True):
    """DBoW2 text format (TemplatedVocabulary::saveToTextFile): header `k L  scoring weighting`, then one line per node:
    `parent isLeaf d0 .. d31 weight`."""
    lines = ["%d %d  %d %d" % (k, L, scoring, weighting)]
    for i in range(len(parent)):
        lines.append("%d %d %s %r" % (parent[i], is_leaf[i], " ".join(str(int(b)) for b in desc[i]), float(weight[i])))
    with open(path, "w") as f:
        f.write("\n".join(lines))
        if trailing_newline:
            f.write("\n")


def vocabulary_like_descriptors(desc_nodes, n, seed=8, flips=20):
    """n query descriptors: random leaves of the tree with `flips` random bit flips (so they descend non-trivially)"""
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(desc_nodes), n)
    bits = np.unpackbits(desc_nodes[pick], axis=1)
    for i in range(n):
        bits[i, rng.choice(256, size=flips, replace=False)] ^= 1
    return np.packbits(bits, axis=1)
