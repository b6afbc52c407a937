"""ctypes bindings over the CPU ORACLE (oracle/ft_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (fasttrack_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libft_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ft_oracle.cpp", "ft_oracle_bow.cpp", "ft_oracle_capi.cpp", "ft_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "ft_orb_pattern.inc"))
    if not force and os.path.exists(_SO) and all(
        (not os.path.exists(s)) or os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs
    ):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


class FrameDesc(C.Structure):
    _fields_ = [
        ("Nleft", C.c_int), ("Nright", C.c_int), ("N", C.c_int),
        ("keys6", C.c_void_p), ("desc", C.c_void_p), ("uRight", C.c_void_p),
        ("l2r", C.c_void_p), ("r2l", C.c_void_p),
        ("minX", C.c_float), ("maxX", C.c_float), ("minY", C.c_float), ("maxY", C.c_float),
        ("nlevels", C.c_int), ("scale", C.c_void_p), ("logScale", C.c_float),
        ("camType", C.c_int), ("cam1", C.c_float * 8), ("cam2", C.c_float * 8),
        ("mbf", C.c_float),
        ("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Rwc", C.c_float * 9), ("Ow", C.c_float * 3),
        ("Rrl", C.c_float * 9), ("trl", C.c_float * 3), ("tlr", C.c_float * 3),
    ]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.fto_extractor_create.restype = C.c_void_p
    L.fto_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.fto_extractor_destroy.argtypes = [C.c_void_p]
    L.fto_extractor_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
    L.fto_extract.restype = C.c_int
    L.fto_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u8p,
                              C.POINTER(C.c_int)]
    L.fto_level_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fto_level_image.restype = C.c_int
    L.fto_level_image.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p]
    L.fto_level_candidates.restype = C.c_int
    L.fto_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p]
    L.fto_level_keys.restype = C.c_int
    L.fto_level_keys.argtypes = [C.c_void_p, C.c_int, C.c_int, f32p, u8p]
    L.fto_desc_borderline.restype = C.c_long
    L.fto_desc_borderline.argtypes = [C.c_void_p]
    L.fto_resize.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
    L.fto_blur.argtypes = [u8p, C.c_int, C.c_int, u8p]
    L.fto_remap.argtypes = [u8p, C.c_int, C.c_int, f32p, f32p, C.c_int, C.c_int, u8p]
    L.fto_undistort_points.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, f32p]
    L.fto_libm_sincosf.argtypes = [C.c_int, f32p, f32p, f32p]
    L.fto_ic_angles.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, f32p, C.c_int, f32p]
    L.fto_orb_descriptors.argtypes = [u8p, C.c_int, C.c_int, f32p, f32p, C.c_int, u8p]
    L.fto_stereo_from_rgbd.argtypes = [f32p, f32p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, f32p, f32p]
    L.fto_image_bounds.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_int, f32p]
    L.fto_fast.restype = C.c_int
    L.fto_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    L.fto_fast_atan2.restype = C.c_float
    L.fto_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.fto_cv_round.restype = C.c_int
    L.fto_cv_round.argtypes = [C.c_float]
    L.fto_knn2.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p]
    L.fto_descriptor_distance.restype = C.c_int
    L.fto_descriptor_distance.argtypes = [u8p, u8p]
    L.fto_octree.restype = C.c_int
    L.fto_octree.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    L.fto_stereo.argtypes = [C.c_void_p, C.c_void_p, f32p, C.c_int, u8p, f32p, C.c_int, u8p, C.c_float, C.c_float,
                             f32p, f32p, i32p, i32p]
    L.fto_fisheye.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, f32p, C.c_int, u8p, C.c_int, f32p, C.c_int, u8p,
                              C.c_int, i32p, i32p, f32p, f32p, i32p]
    L.fto_cam_project.argtypes = [C.c_int, f32p, f32p, f32p]
    L.fto_kb8_unproject.argtypes = [f32p, C.c_float, C.c_float, f32p]
    L.fto_frame_create.restype = C.c_void_p
    L.fto_frame_create.argtypes = [C.POINTER(FrameDesc)]
    L.fto_frame_destroy.argtypes = [C.c_void_p]
    L.fto_frame_grid.restype = C.c_int
    L.fto_frame_grid.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
    L.fto_frustum.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, u8p, i32p, C.c_float, i32p, f32p]
    L.fto_search_local_points.restype = C.c_int
    L.fto_search_local_points.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, u8p, i32p, C.c_float, C.c_int,
                                          C.c_float, C.c_float, i32p, u8p, C.c_void_p, C.c_void_p]
    L.fto_search_last_frame.restype = C.c_int
    L.fto_search_last_frame.argtypes = [C.c_void_p, C.c_int, f32p, u8p, i32p, f32p, i32p, C.c_float, C.c_int, C.c_int, i32p, u8p, i32p]
    L.fto_time_stereo_frame.restype = C.c_double
    L.fto_time_stereo_frame.argtypes = [C.c_void_p, C.c_void_p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L


class Extractor:
    """Oracle mirror of ORB_SLAM3::ORBextractor (CPU branch)."""

    def __init__(self, nfeatures=1200, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = C.c_void_p(self.L.fto_extractor_create(nfeatures, scale, nlevels, ini_th, min_th))
        self.scale = np.zeros(nlevels, np.float32)
        self.inv_scale = np.zeros(nlevels, np.float32)
        self.sigma2 = np.zeros(nlevels, np.float32)
        self.inv_sigma2 = np.zeros(nlevels, np.float32)
        self.features_per_level = np.zeros(nlevels, np.int32)
        self.umax = np.zeros(16, np.int32)
        self.L.fto_extractor_tables(self.h, self.scale, self.inv_scale, self.sigma2, self.inv_sigma2,
                                    self.features_per_level, self.umax)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fto_extractor_destroy(self.h)
            self.h = None

    def extract(self, img, lap=(0, 0)):
        """returns (monoIndex, kps[n,6] float32, desc[n,32] uint8)"""
        img = np.ascontiguousarray(img, np.uint8)
        hgt, wid = img.shape
        cap = self.nfeatures + 64
        kps = np.zeros((cap, 6), np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.fto_extract(self.h, img.ctypes.data, wid, hgt, img.strides[0], lap[0], lap[1], cap, kps, desc,
                                  C.byref(n))
        assert n.value <= cap
        return mono, kps[: n.value].copy(), desc[: n.value].copy()

    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        self.L.fto_level_dims(self.h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_image(self, level, blurred=False):
        w, h = self.level_dims(level)
        out = np.zeros((h, w), np.uint8)
        ok = self.L.fto_level_image(self.h, level, int(blurred), out)
        return out if ok else None

    def level_candidates(self, level):
        n = self.L.fto_level_candidates(self.h, level, 0, np.zeros((1, 3), np.float32))
        out = np.zeros((max(n, 1), 3), np.float32)
        self.L.fto_level_candidates(self.h, level, n, out)
        return out[:n]

    def level_keys(self, level):
        cap = self.nfeatures + 64
        kps = np.zeros((cap, 6), np.float32)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.fto_level_keys(self.h, level, cap, kps, desc)
        return kps[:n].copy(), desc[:n].copy()

    def ic_angles(self, img, xy):
        """IC_Angle (ORBextractor.cc:39-66) of the given integer keypoint positions on `img`"""
        img = np.ascontiguousarray(img, np.uint8); xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros(len(xy), np.float32)
        self.L.fto_ic_angles(self.h, img, img.shape[1], img.shape[0], xy, len(xy), out)
        return out

    def desc_borderline(self):
        return int(self.L.fto_desc_borderline(self.h))

    def octree(self, xyr, min_x, max_x, min_y, max_y, n_target):
        xyr = np.ascontiguousarray(xyr, np.float32)
        out = np.zeros((max(len(xyr), 1), 3), np.float32)
        n = self.L.fto_octree(self.h, xyr, len(xyr), min_x, max_x, min_y, max_y, n_target, out)
        return out[:n].copy()


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().fto_resize(src, src.shape[1], src.shape[0], dst, dw, dh)
    return dst


def blur(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().fto_blur(src, src.shape[1], src.shape[0], dst)
    return dst


def remap(src, mapx, mapy):
    """cv::remap(src, map1=mapx, map2=mapy, INTER_LINEAR) with the default constant-0 border"""
    src = np.ascontiguousarray(src, np.uint8)
    mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
    dh, dw = mapx.shape
    dst = np.zeros((dh, dw), np.uint8)
    lib().fto_remap(src, src.shape[1], src.shape[0], mapx, mapy, dw, dh, dst)
    return dst


def orb_descriptors(blurred, xy, angles):
    """computeOrbDescriptor (ORBextractor.cc:68-108) of the given keypoints on an already blurred image -> [n,32]"""
    blurred = np.ascontiguousarray(blurred, np.uint8); xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros((len(xy), 32), np.uint8)
    lib().fto_orb_descriptors(blurred, blurred.shape[1], blurred.shape[0], xy, angles, len(xy), out)
    return out


def cam_project(cam_type, params8, P):
    """GeometricCamera::project (Pinhole.cpp:43-49 / KannalaBrandt8.cpp:67-84) of [n,3] points -> [n,2]"""
    p8 = np.ascontiguousarray(params8, np.float32); P = np.ascontiguousarray(P, np.float32).reshape(-1, 3)
    out = np.zeros((len(P), 2), np.float32)
    L = lib()
    for i in range(len(P)):
        L.fto_cam_project(int(cam_type), p8, P[i], out[i])
    return out


def kb8_unproject(params8, uv):
    """KannalaBrandt8::unproject (KannalaBrandt8.cpp:116-143) of [n,2] pixels -> [n,3] rays (z = 1)"""
    p8 = np.ascontiguousarray(params8, np.float32); uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    out = np.zeros((len(uv), 3), np.float32)
    L = lib()
    for i in range(len(uv)):
        L.fto_kb8_unproject(p8, float(uv[i, 0]), float(uv[i, 1]), out[i])
    return out


def libm_sincosf(angles):
    """host libm sinf / cosf of a float32 array (the calls of computeOrbDescriptor, ORBextractor.cc:74)"""
    a = np.ascontiguousarray(angles, np.float32)
    s = np.zeros_like(a); c = np.zeros_like(a)
    lib().fto_libm_sincosf(len(a), a, s, c)
    return s, c


def stereo_from_rgbd(keys_xy, keys_un_x, depth, mbf):
    """Frame::ComputeStereoFromRGBD -> (uRight, depth); depth=None is the monocular frame (all -1)"""
    xy = np.ascontiguousarray(keys_xy, np.float32).reshape(-1, 2)
    unx = np.ascontiguousarray(keys_un_x, np.float32)
    ur = np.zeros(len(xy), np.float32); dp = np.zeros(len(xy), np.float32)
    if depth is None:
        lib().fto_stereo_from_rgbd(xy, unx, len(xy), None, 0, 0, float(mbf), ur, dp)
    else:
        d = np.ascontiguousarray(depth, np.float32)
        lib().fto_stereo_from_rgbd(xy, unx, len(xy), d.ctypes.data, d.shape[1], d.shape[0], float(mbf), ur, dp)
    return ur, dp


def undistort_points(xy, K, dist):
    """cv::undistortPoints(xy, K, dist, R=None, P=K) -> [n,2] float32 (Frame::UndistortKeyPoints)"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(xy)
    lib().fto_undistort_points(xy, len(xy), K, dist, len(dist), out)
    return out


def image_bounds(cols, rows, K, dist):
    """Frame::ComputeImageBounds -> (minX, maxX, minY, maxY)"""
    K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.zeros(4, np.float32)
    lib().fto_image_bounds(cols, rows, K, dist, len(dist), out)
    return out


def fast(img, th):
    """cv::FAST(img, th, nonmax=True) on a (possibly strided) 2-D uint8 view; returns [n,3] x,y,response"""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = max(1, (w * h) // 4 + 16)
    out = np.zeros((cap, 3), np.float32)
    n = lib().fto_fast(img.ctypes.data, img.strides[0], w, h, th, cap, out)
    return out[:n].copy()


def fast_atan2(y, x):
    return float(lib().fto_fast_atan2(float(y), float(x)))


def cv_round(v):
    return int(lib().fto_cv_round(float(v)))


def knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.zeros((len(q), 2), np.int32)
    dist = np.zeros((len(q), 2), np.int32)
    lib().fto_knn2(q, len(q), t, len(t), idx, dist)
    return idx, dist


def stereo(exL, exR, kL, dL, kR, dR, mbf, mb):
    kL = np.ascontiguousarray(kL, np.float32); kR = np.ascontiguousarray(kR, np.float32)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    n = len(kL)
    u = np.zeros(n, np.float32); d = np.zeros(n, np.float32)
    b = np.zeros(n, np.int32); s = np.zeros(n, np.int32)
    lib().fto_stereo(exL.h, exR.h, kL, n, dL, kR, len(kR), dR, mbf, mb, u, d, b, s)
    return dict(uRight=u, depth=d, bestIdxR=b, sad=s)


def fisheye(cam1, cam2, Rlr, tlr, sigma2, kL, dL, mono_left, kR, dR, mono_right):
    f = lambda a: np.ascontiguousarray(a, np.float32)
    kL, kR = f(kL), f(kR)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    nL, nR = len(kL), len(kR)
    l2r = np.zeros(max(nL, 1), np.int32); r2l = np.zeros(max(nR, 1), np.int32)
    depth = np.zeros(max(nL, 1), np.float32); p3d = np.zeros((max(nL, 1), 3), np.float32)
    code = np.zeros(max(nL, 1), np.int32)
    lib().fto_fisheye(f(cam1), f(cam2), f(Rlr).reshape(-1), f(tlr), f(sigma2), len(sigma2), kL, nL, dL, mono_left, kR,
                      nR, dR, mono_right, l2r, r2l, depth, p3d, code)
    return dict(l2r=l2r[:nL], r2l=r2l[:nR], depth=depth[:nL], p3d=p3d[:nL], code=code[:nL])


class Frame:
    """Oracle mirror of the parts of ORB_SLAM3::Frame that the projection search reads."""

    def __init__(self, keys, desc, scale, width, height, cam_type=0, cam1=None, cam2=None, mbf=0.0, u_right=None,
                 n_left=-1, n_right=-1, l2r=None, r2l=None, Rcw=None, tcw=None, Rrl=None, trl=None, tlr=None, bounds=None):
        self.L = lib()
        f = lambda a: np.ascontiguousarray(a, np.float32)
        self._keep = dict(keys=f(keys), desc=np.ascontiguousarray(desc, np.uint8), scale=f(scale))
        d = FrameDesc()
        d.Nleft, d.Nright, d.N = n_left, n_right, len(keys)
        d.keys6 = self._keep["keys"].ctypes.data
        d.desc = self._keep["desc"].ctypes.data
        if u_right is not None:
            self._keep["ur"] = f(u_right); d.uRight = self._keep["ur"].ctypes.data
        if l2r is not None:
            self._keep["l2r"] = np.ascontiguousarray(l2r, np.int32); d.l2r = self._keep["l2r"].ctypes.data
        if r2l is not None:
            self._keep["r2l"] = np.ascontiguousarray(r2l, np.int32); d.r2l = self._keep["r2l"].ctypes.data
        d.minX, d.maxX, d.minY, d.maxY = 0.0, float(width), 0.0, float(height)
        if bounds is not None:   # Frame::ComputeImageBounds with a distorted pinhole camera
            d.minX, d.maxX, d.minY, d.maxY = [float(b) for b in bounds]
        d.nlevels = len(scale)
        d.scale = self._keep["scale"].ctypes.data
        d.logScale = float(np.log(np.float32(scale[1]))) if len(scale) > 1 else 1.0
        d.logScale = float(np.float32(np.log(np.float32(scale[1])))) if len(scale) > 1 else 1.0
        d.camType = cam_type
        cam1 = f(cam1 if cam1 is not None else np.zeros(8)); cam2 = f(cam2 if cam2 is not None else cam1)
        for i in range(8):
            d.cam1[i] = cam1[i]; d.cam2[i] = cam2[i]
        d.mbf = mbf
        Rcw = f(np.eye(3) if Rcw is None else Rcw).reshape(3, 3); tcw = f(np.zeros(3) if tcw is None else tcw)
        Rwc = np.ascontiguousarray(Rcw.T); Ow = (-(Rwc @ tcw)).astype(np.float32)
        Rrl = f(np.eye(3) if Rrl is None else Rrl).reshape(3, 3)
        trl = f(np.zeros(3) if trl is None else trl); tlr = f(np.zeros(3) if tlr is None else tlr)
        for i in range(9):
            d.Rcw[i] = Rcw.reshape(-1)[i]; d.Rwc[i] = Rwc.reshape(-1)[i]; d.Rrl[i] = Rrl.reshape(-1)[i]
        for i in range(3):
            d.tcw[i] = tcw[i]; d.Ow[i] = Ow[i]; d.trl[i] = trl[i]; d.tlr[i] = tlr[i]
        self.Rwc, self.Ow = Rwc, Ow
        self.N = len(keys)
        self._create(d)

    def _create(self, d):
        self.h = C.c_void_p(self.L.fto_frame_create(C.byref(d)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fto_frame_destroy(self.h)
            self.h = None

    def grid(self, right=False):
        counts = np.zeros(64 * 48, np.int32)
        idx = np.zeros(max(self.N, 1), np.int32)
        n = self.L.fto_frame_grid(self.h, int(right), counts, idx)
        return counts, idx[:n].copy()

    @staticmethod
    def _mp(pos, normal, minmax, desc, flags):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        return f(pos), f(normal), f(minmax), np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(flags, np.int32)

    def frustum(self, pos, normal, minmax, desc, flags, view_cos_limit=0.5):
        pos, normal, minmax, desc, flags = self._mp(pos, normal, minmax, desc, flags)
        M = len(pos)
        ti = np.zeros((M, 5), np.int32); tf = np.zeros((M, 9), np.float32)
        self.L.fto_frustum(self.h, M, pos, normal, minmax, desc, flags, view_cos_limit, ti, tf)
        return ti, tf

    def search_local_points(self, pos, normal, minmax, desc, flags, th, holder, holder_obs, b_far=False, th_far=50.0,
                            nnratio=0.8):
        pos, normal, minmax, desc, flags = self._mp(pos, normal, minmax, desc, flags)
        M = len(pos)
        holder = np.ascontiguousarray(holder, np.int32).copy()
        holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        ti = np.zeros((M, 5), np.int32); tf = np.zeros((M, 9), np.float32)
        n = self.L.fto_search_local_points(self.h, M, pos, normal, minmax, desc, flags, th, int(b_far), th_far, nnratio,
                                           holder, holder_obs, ti.ctypes.data, tf.ctypes.data)
        return n, holder, holder_obs, ti, tf


def _frame_search_last_frame(self, pos, desc, octave, angle, flags, th, direction, holder, holder_obs, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on this (current) frame"""
    pos = np.ascontiguousarray(pos, np.float32); desc = np.ascontiguousarray(desc, np.uint8)
    octave = np.ascontiguousarray(octave, np.int32); angle = np.ascontiguousarray(angle, np.float32)
    flags = np.ascontiguousarray(flags, np.int32)
    holder = np.ascontiguousarray(holder, np.int32).copy(); holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
    bl = np.zeros(max(len(pos), 1), np.int32)
    n = self.L.fto_search_last_frame(self.h, len(pos), pos, desc, octave, angle, flags, th, direction, int(check_ori), holder,
                                     holder_obs, bl)
    return n, holder, holder_obs, bl[: len(pos)]


Frame.search_last_frame = _frame_search_last_frame


def time_stereo_frame(exL, exR, imgL, imgR, mbf, mb, two_threads=True):
    imgL = np.ascontiguousarray(imgL, np.uint8); imgR = np.ascontiguousarray(imgR, np.uint8)
    h, w = imgL.shape
    nl, nr, ns = C.c_int(), C.c_int(), C.c_int()
    ms = lib().fto_time_stereo_frame(exL.h, exR.h, imgL, imgR, w, h, w, mbf, mb, int(two_threads), C.byref(nl),
                                     C.byref(nr), C.byref(ns))
    return ms, nl.value, nr.value, ns.value


# ---- bag of words (oracle/ft_oracle_bow.cpp) ----
_REF_SO = os.path.join(_HERE, "_ref", "libft_ref_dbow2.so")
_REFERENCE = os.environ.get("FT_REFERENCE_DIR", "/root/reference")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


_REF_ORB_SO = os.path.join(_HERE, "_ref", "libft_ref_orbextractor.so")
_REF_GPU_SO = os.path.join(_HERE, "_ref", "libft_ref_orbextractor_gpu.so")


def build_ref():
    """oracle/_ref/*.so: pieces of the REFERENCE itself compiled where they lie (`make ref`): its vendored DBoW2 and the CPU
    branch of src/ORBextractor.cc. Returns the directory, or None when neither the prebuilt libraries nor the reference
    tree are available (e.g. on the GPU box before a snapshot carried the files)."""
    srcs = [os.path.join(_HERE, f) for f in ("ref_dbow2_capi.cpp", "ref_orbextractor_capi.cpp", "ref_frame_capi.cpp", "ref_gpu_capi.cpp",
                                             "ref_extract_fns.py", os.path.join("ref_stubs", "ref_frame_shim.h"),
                                             "ft_oracle.cpp", "ft_oracle.h",
                                             os.path.join("ref_stubs", "ft_cv_standin.cpp"),
                                             os.path.join("ref_stubs", "opencv2", "opencv.hpp"))]
    outs = [_REF_SO, _REF_ORB_SO, os.path.join(_HERE, "_ref", "libft_ref_frame.so"), _REF_GPU_SO]
    have_ref = os.path.isdir(os.path.join(_REFERENCE, "Thirdparty", "DBoW2", "DBoW2"))
    built = all(os.path.exists(o) for o in outs)
    fresh = built and all(os.path.getmtime(o) >= os.path.getmtime(x) for o in outs for x in srcs)
    if fresh or (built and not have_ref):
        return os.path.dirname(_REF_SO)
    if not have_ref:
        return None
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + _REFERENCE])
    return os.path.dirname(_REF_SO)


_ref_gpu = None


def ref_gpu_lib():
    """oracle/_ref/libft_ref_orbextractor_gpu.so: the reference's OWN CUDA kernels (src/{resize,gaussian_blur,fast,orientation,
    descriptor}.cu, src/Kernels/StereoMatchKernel.cu) compiled with nvcc from where they lie, plus src/ORBextractor.cc in GPU run
    mode. Bench infrastructure only (the performance bar); None when the library did not travel with the snapshot."""
    global _ref_gpu
    if _ref_gpu is None:
        if build_ref() is None or not os.path.exists(_REF_GPU_SO):
            return None
        R = C.CDLL(_REF_GPU_SO)
        R.ftrefgpu_extractor_create.restype = C.c_void_p
        R.ftrefgpu_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        R.ftrefgpu_extractor_destroy.argtypes = [C.c_void_p]
        R.ftrefgpu_extract.restype = C.c_int
        R.ftrefgpu_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int]
        R.ftrefgpu_stage_times.restype = C.c_int
        R.ftrefgpu_stage_times.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, f32p]
        R.ftrefgpu_stereo_time.restype = C.c_int
        R.ftrefgpu_stereo_time.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, C.c_int, u8p, f32p, C.c_int,
                                           u8p, C.c_float, C.c_float, C.c_int, f32p]
        _ref_gpu = R
    return _ref_gpu


def ref_gpu_stage_times(img, nlevels=8, scale_factor=1.2, ini_th=20, min_th=7, reps=20):
    """CUDA-event ms of the reference's five extractor launchers on one image (both resize..descriptor chains as
    ComputePyramidGPU / ComputeKeyPointsOctTreeGPU issue them): dict(resize, gaussian_blur, fast_extract, compute_orientation,
    compute_descriptor, corners)"""
    R = ref_gpu_lib()
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(8, np.float32)
    rc = R.ftrefgpu_stage_times(img, img.shape[1], img.shape[0], img.strides[0], nlevels, scale_factor, ini_th, min_th, reps, out)
    if rc != 0:
        raise RuntimeError("ftrefgpu_stage_times failed (%d)" % rc)
    names = ("resize", "gaussian_blur", "fast_extract", "compute_orientation", "compute_descriptor")
    d = {k: float(out[i]) for i, k in enumerate(names)}
    d["corners"] = int(out[5])
    return d


def ref_gpu_extract_ms(imgs, nfeatures=1200, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    """wall-clock ms per image of the reference's ORBextractor::operator() in GPU run mode (H2D, its kernels, D2H of all
    corners, DistributeOctTreeGPU on the host); first image is a warm-up. Returns (ms_per_image, keypoints of the last)"""
    import time
    R = ref_gpu_lib()
    h, w = imgs[0].shape
    ex = R.ftrefgpu_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th, w, h)
    t = []
    n = 0
    for i, im in enumerate(imgs):
        im = np.ascontiguousarray(im, np.uint8)
        t0 = time.perf_counter()
        n = R.ftrefgpu_extract(C.c_void_p(ex), im, w, h, im.strides[0])
        if i:
            t.append((time.perf_counter() - t0) * 1e3)
    R.ftrefgpu_extractor_destroy(C.c_void_p(ex))
    return float(np.mean(t)), int(n)


def ref_gpu_stereo_ms(imgL, imgR, kL, dL, kR, dR, mbf, mb, nlevels=8, scale_factor=1.2, reps=10):
    """wall-clock ms per call of the reference's GPU stereo matching (Frame::ComputeStereoMatchesGPU -> StereoMatchKernel::launch)
    on an extracted pair (kL / kR: oracle keypoint rows x, y, size, angle, response, octave). Returns (ms, matches kept)"""
    R = ref_gpu_lib()
    imgL = np.ascontiguousarray(imgL, np.uint8); imgR = np.ascontiguousarray(imgR, np.uint8)
    a = np.ascontiguousarray(kL[:, [0, 1, 5]], np.float32); b = np.ascontiguousarray(kR[:, [0, 1, 5]], np.float32)
    out = np.zeros(4, np.float32)
    rc = R.ftrefgpu_stereo_time(imgL, imgR, imgL.shape[1], imgL.shape[0], imgL.strides[0], nlevels, scale_factor, a, len(a),
                                np.ascontiguousarray(dL, np.uint8), b, len(b), np.ascontiguousarray(dR, np.uint8), float(mbf), float(mb),
                                reps, out)
    if rc != 0:
        raise RuntimeError("ftrefgpu_stereo_time failed (%d)" % rc)
    return float(out[0]), int(out[1])


class RefExtractor:
    """The reference's own ORBextractor (CPU branch of src/ORBextractor.cc, oracle/_ref/libft_ref_orbextractor.so): used
    ONLY to pin the oracle's restatement and to make golden vectors."""

    def __init__(self, nfeatures=1200, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, width=752, height=480):
        if build_ref() is None:
            raise FileNotFoundError("oracle/_ref/libft_ref_orbextractor.so is not built and the reference tree is absent")
        R = C.CDLL(_REF_ORB_SO)
        R.ftref_extractor_create.restype = C.c_void_p
        R.ftref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        R.ftref_extractor_destroy.argtypes = [C.c_void_p]
        R.ftref_extract.restype = C.c_int
        R.ftref_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u8p, C.c_int,
                                    C.POINTER(C.c_int)]
        R.ftref_level_image.restype = C.c_int
        R.ftref_level_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        R.ftref_scale_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
        self.R, self.nlevels, self.cap = R, nlevels, nfeatures + 64 * nlevels
        self.h = R.ftref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height)

    def extract(self, img, lap=(0, 0)):
        """ORBextractor::operator(): returns (monoIndex, kps[n,6], desc[n,32])"""
        img = np.ascontiguousarray(img, np.uint8)
        k = np.zeros((self.cap, 6), np.float32); d = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int()
        mono = self.R.ftref_extract(C.c_void_p(self.h), img, img.shape[1], img.shape[0], img.strides[0], int(lap[0]), int(lap[1]),
                                    k, d, self.cap, C.byref(n))
        assert n.value <= self.cap
        return mono, k[:n.value].copy(), d[:n.value].copy()

    def level_image(self, level):
        w, h = C.c_int(), C.c_int()
        if not self.R.ftref_level_image(C.c_void_p(self.h), level, None, C.byref(w), C.byref(h)):
            return None
        out = np.zeros((h.value, w.value), np.uint8)
        self.R.ftref_level_image(C.c_void_p(self.h), level, out.ctypes.data, C.byref(w), C.byref(h))
        return out

    def scale_tables(self):
        a = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.R.ftref_scale_tables(C.c_void_p(self.h), *a)
        return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3])

    def __del__(self):
        try:
            self.R.ftref_extractor_destroy(C.c_void_p(self.h))
        except Exception:
            pass


class Vocabulary:
    """Oracle restatement of ORBVocabulary (DBoW2::TemplatedVocabulary<FORB>)."""

    def __init__(self, handle):
        self.h = handle
        L = lib()
        info = np.zeros(6, np.int32)
        L.fto_voc_info(C.c_void_p(self.h), info)
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [int(x) for x in info]

    @classmethod
    def load_text(cls, path):
        L = _voc_lib()
        h = L.fto_voc_load_text(path.encode())
        if not h:
            raise IOError("not a DBoW2 text vocabulary: %s" % path)
        return cls(h)

    @classmethod
    def from_arrays(cls, k, Ldepth, scoring, weighting, parent, is_leaf, desc, weight):
        L = _voc_lib()
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        return cls(L.fto_voc_from_arrays(k, Ldepth, scoring, weighting, len(parent), parent, is_leaf, desc, weight))

    def arrays(self):
        """(parent, is_leaf, desc, weight) without the root: the arguments of ft_vocabulary_create"""
        n = self.n_nodes - 1
        parent = np.zeros(n, np.int32); leaf = np.zeros(n, np.uint8); desc = np.zeros((n, 32), np.uint8)
        weight = np.zeros(n, np.float64)
        lib().fto_voc_arrays(C.c_void_p(self.h), parent, leaf, desc, weight)
        return parent, leaf, desc, weight

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW: returns dict(node[n] (-1 = stopped word), word[n], bow_ids, bow_vals)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        node = np.zeros(max(n, 1), np.int32); word = np.zeros(max(n, 1), np.int32)
        ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64)
        m = lib().fto_voc_transform(C.c_void_p(self.h), desc if n else np.zeros((1, 32), np.uint8), n, levelsup, node, word,
                                    ids, vals, max(n, 1))
        return dict(node=node[:n], word=word[:n], bow_ids=ids[:m].copy(), bow_vals=vals[:m].copy())

    def __del__(self):
        try:
            lib().fto_voc_free(C.c_void_p(self.h))
        except Exception:
            pass


_voc_bound = False


def _voc_lib():
    global _voc_bound
    L = lib()
    if not _voc_bound:
        L.fto_voc_load_text.restype = C.c_void_p
        L.fto_voc_load_text.argtypes = [C.c_char_p]
        L.fto_voc_from_arrays.restype = C.c_void_p
        L.fto_voc_from_arrays.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, u8p, u8p, f64p]
        L.fto_voc_free.argtypes = [C.c_void_p]
        L.fto_voc_info.argtypes = [C.c_void_p, i32p]
        L.fto_voc_arrays.argtypes = [C.c_void_p, i32p, u8p, u8p, f64p]
        L.fto_voc_transform.restype = C.c_int
        L.fto_voc_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, i32p, i32p, u32p, f64p, C.c_int]
        L.fto_search_by_bow.restype = C.c_int
        L.fto_search_by_bow.argtypes = [C.c_int, u8p, f32p, i32p, u8p, C.c_int, u8p, f32p, i32p, C.c_int, C.c_float, C.c_int,
                                        i32p]
        _voc_bound = True
    return L


def search_by_bow(kf_desc, kf_angle, kf_node, kf_has_mp, f_desc, f_angle, f_node, f_nleft=-1, nnratio=0.7, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches): returns (nmatches, match[nF])"""
    L = _voc_lib()
    c8 = lambda a: np.ascontiguousarray(a, np.uint8)
    kf_desc, f_desc, kf_has_mp = c8(kf_desc).reshape(-1, 32), c8(f_desc).reshape(-1, 32), c8(kf_has_mp)
    kf_angle, f_angle = np.ascontiguousarray(kf_angle, np.float32), np.ascontiguousarray(f_angle, np.float32)
    kf_node, f_node = np.ascontiguousarray(kf_node, np.int32), np.ascontiguousarray(f_node, np.int32)
    match = np.full(max(len(f_desc), 1), -1, np.int32)
    pad = lambda a, dt, shape: a if len(a) else np.zeros(shape, dt)
    nm = L.fto_search_by_bow(len(kf_desc), pad(kf_desc, np.uint8, (1, 32)), pad(kf_angle, np.float32, 1),
                             pad(kf_node, np.int32, 1), pad(kf_has_mp, np.uint8, 1), len(f_desc),
                             pad(f_desc, np.uint8, (1, 32)), pad(f_angle, np.float32, 1), pad(f_node, np.int32, 1),
                             int(f_nleft), float(nnratio), int(check_ori), match)
    return nm, match[:len(f_desc)]


class RefVocabulary:
    """The reference's own DBoW2 (oracle/_ref/libft_ref_dbow2.so): used ONLY to pin the oracle's restatement."""

    def __init__(self, path):
        if build_ref() is None:
            raise FileNotFoundError("oracle/_ref/libft_ref_dbow2.so is not built and the reference tree is absent")
        R = C.CDLL(_REF_SO)
        R.ftref_voc_load_text.restype = C.c_void_p
        R.ftref_voc_load_text.argtypes = [C.c_char_p]
        R.ftref_voc_free.argtypes = [C.c_void_p]
        R.ftref_voc_words.argtypes = [C.c_void_p]
        R.ftref_voc_transform.restype = C.c_int
        R.ftref_voc_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, i32p, u32p, f64p, C.c_int, i32p,
                                          C.POINTER(C.c_int)]
        self.R = R
        self.h = R.ftref_voc_load_text(path.encode())
        if not self.h:
            raise IOError("reference loadFromTextFile failed: %s" % path)
        self.n_words = R.ftref_voc_words(C.c_void_p(self.h))

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        node = np.zeros(max(n, 1), np.int32); order = np.zeros(max(n, 1), np.int32)
        ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64)
        nf = C.c_int()
        m = self.R.ftref_voc_transform(C.c_void_p(self.h), desc if n else np.zeros((1, 32), np.uint8), n, levelsup, node, ids,
                                       vals, max(n, 1), order, C.byref(nf))
        return dict(node=node[:n], bow_ids=ids[:m].copy(), bow_vals=vals[:m].copy(), featvec_order=order[:nf.value].copy())

    def __del__(self):
        try:
            self.R.ftref_voc_free(C.c_void_p(self.h))
        except Exception:
            pass


# ---- functions of the reference's Frame.cc / ORBmatcher.cc compiled from their own text (oracle/_ref/libft_ref_frame.so) ----
_REF_FRAME_SO = os.path.join(_HERE, "_ref", "libft_ref_frame.so")


class RefFrameDesc(C.Structure):
    _fields_ = FrameDesc._fields_ + [("keysUn6", C.c_void_p)]


_ref_frame_lib = None


def ref_frame_lib():
    """dlopen oracle/_ref/libft_ref_frame.so; None when it is not built and cannot be built here"""
    global _ref_frame_lib
    if _ref_frame_lib is not None:
        return _ref_frame_lib
    if build_ref() is None or not os.path.exists(_REF_FRAME_SO):
        return None
    R = C.CDLL(_REF_FRAME_SO)
    vp = C.c_void_p
    R.ftref_frame_create.restype = vp
    R.ftref_frame_create.argtypes = [C.POINTER(RefFrameDesc)]
    R.ftref_frame_destroy.argtypes = [vp]
    R.ftref_frame_grid.argtypes = [vp, C.c_int, i32p, i32p]
    R.ftref_features_in_area.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
    R.ftref_search_local_points.argtypes = [vp, C.c_int, f32p, f32p, f32p, u8p, i32p, C.c_float, C.c_int, C.c_float, C.c_float,
                                            i32p, u8p, vp, vp]
    R.ftref_search_last_frame.argtypes = [vp, C.c_int, f32p, u8p, i32p, f32p, i32p, f32p, f32p, C.c_float, C.c_int, C.c_int,
                                          C.c_float, i32p, u8p]
    R.ftref_stereo_matches.argtypes = [C.c_int, i32p, i32p, u8p, u8p, f32p, f32p, u8p, C.c_int, f32p, u8p, C.c_int, C.c_float,
                                       C.c_float, f32p, f32p]
    R.ftref_stereo_matches.restype = None
    R.ftref_stereo_fisheye.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, f32p, u8p, C.c_int, C.c_int, f32p, u8p, C.c_int, C.c_int,
                                       i32p, i32p, f32p, f32p]
    R.ftref_stereo_fisheye.restype = None
    R.ftref_undistort.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_int, f32p, f32p]
    R.ftref_undistort.restype = None
    R.ftref_stereo_from_rgbd.argtypes = [f32p, f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_float, f32p, f32p]
    R.ftref_stereo_from_rgbd.restype = None
    R.ftref_search_by_bow.argtypes = [C.c_int, u8p, f32p, i32p, u8p, C.c_int, u8p, f32p, i32p, C.c_int, C.c_float, C.c_int, i32p]
    _ref_frame_lib = R
    return R


class RefFrame(Frame):
    """The reference's own Frame::{AssignFeaturesToGrid, GetFeaturesInArea, isInFrustum[Checks]} and
    ORBmatcher::SearchByProjection functions (text taken from the reference files at build time): same constructor and
    methods as the oracle's Frame. Used ONLY to pin the oracle's restatement and to make golden vectors."""

    def _create(self, d):
        self.R = ref_frame_lib()
        if self.R is None:
            raise FileNotFoundError("oracle/_ref/libft_ref_frame.so is not built and the reference tree is absent")
        rd = RefFrameDesc()
        for name, _ in FrameDesc._fields_:
            setattr(rd, name, getattr(d, name))
        rd.keysUn6 = None
        self.h = C.c_void_p(self.R.ftref_frame_create(C.byref(rd)))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self, "R", None):
            self.R.ftref_frame_destroy(self.h)
            self.h = None

    def grid(self, right=False):
        counts = np.zeros(64 * 48, np.int32)
        idx = np.zeros(max(self.N, 1), np.int32)
        n = self.R.ftref_frame_grid(self.h, int(right), counts, idx)
        return counts, idx[:n].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1, right=False):
        out = np.zeros(max(self.N, 1), np.int32)
        n = self.R.ftref_features_in_area(self.h, x, y, r, min_level, max_level, int(right), out, len(out))
        return out[:n].copy()

    def search_local_points(self, pos, normal, minmax, desc, flags, th, holder, holder_obs, b_far=False, th_far=50.0,
                            nnratio=0.8):
        pos, normal, minmax, desc, flags = self._mp(pos, normal, minmax, desc, flags)
        M = len(pos)
        holder = np.ascontiguousarray(holder, np.int32).copy()
        holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        ti = np.zeros((M, 4), np.int32); tf = np.zeros((M, 9), np.float32)
        n = self.R.ftref_search_local_points(self.h, M, pos, normal, minmax, desc, flags, th, int(b_far), th_far, nnratio,
                                             holder, holder_obs, ti.ctypes.data, tf.ctypes.data)
        return n, holder, holder_obs, ti, tf

    def search_last_frame(self, pos, desc, octave, angle, flags, th, Rlw, tlw, mb, holder, holder_obs, b_mono=False,
                          check_ori=True):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        pos, angle, Rlw, tlw = f(pos), f(angle), f(Rlw).reshape(-1), f(tlw)
        desc = np.ascontiguousarray(desc, np.uint8)
        octave = np.ascontiguousarray(octave, np.int32); flags = np.ascontiguousarray(flags, np.int32)
        holder = np.ascontiguousarray(holder, np.int32).copy(); holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        n = self.R.ftref_search_last_frame(self.h, len(pos), pos, desc, octave, angle, flags, Rlw, tlw, th, int(b_mono),
                                           int(check_ori), mb, holder, holder_obs)
        return n, holder, holder_obs


def ref_stereo(exL, exR, kL, dL, kR, dR, mbf, mb):
    """The reference's own Frame::ComputeStereoMatches on the pyramids of two extractors (oracle.Extractor or
    RefExtractor objects, after extract()): dict(uRight, depth)"""
    R = ref_frame_lib()
    nl = exL.nlevels
    lv = [exL.level_image(l) for l in range(nl)]; rv = [exR.level_image(l) for l in range(nl)]
    lw = np.array([a.shape[1] for a in lv], np.int32); lh = np.array([a.shape[0] for a in lv], np.int32)
    pl = np.concatenate([a.reshape(-1) for a in lv]); pr = np.concatenate([a.reshape(-1) for a in rv])
    scale = np.ascontiguousarray(exL.scale if hasattr(exL, "scale") else exL.scale_tables()["scale"], np.float32)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    ur = np.zeros(max(len(kL), 1), np.float32); dp = np.zeros(max(len(kL), 1), np.float32)
    R.ftref_stereo_matches(nl, lw, lh, np.ascontiguousarray(pl), np.ascontiguousarray(pr), scale, f(kL),
                           np.ascontiguousarray(dL, np.uint8), len(kL), f(kR), np.ascontiguousarray(dR, np.uint8), len(kR),
                           mbf, mb, ur, dp)
    return dict(uRight=ur[:len(kL)], depth=dp[:len(kL)])


def ref_fisheye(cam1, cam2, Rlr, tlr, sigma2, kL, dL, mono_left, kR, dR, mono_right):
    """The reference's own Frame::ComputeStereoFishEyeMatches + KannalaBrandt8::TriangulateMatches (BFMatcher = the oracle's
    cv2-pinned knn, JacobiSVD = the oracle's Jacobi routine): dict(l2r, r2l, depth, p3d)"""
    R = ref_frame_lib()
    f = lambda a: np.ascontiguousarray(a, np.float32)
    nL, nR = len(kL), len(kR)
    l2r = np.zeros(max(nL, 1), np.int32); r2l = np.zeros(max(nR, 1), np.int32)
    depth = np.zeros(max(nL, 1), np.float32); p3d = np.zeros((max(nL, 1), 3), np.float32)
    sigma2 = f(sigma2)
    R.ftref_stereo_fisheye(f(cam1), f(cam2), f(Rlr).reshape(-1), f(tlr), sigma2, len(sigma2), f(kL), np.ascontiguousarray(dL, np.uint8),
                           nL, int(mono_left), f(kR), np.ascontiguousarray(dR, np.uint8), nR, int(mono_right), l2r, r2l, depth, p3d)
    return dict(l2r=l2r[:nL], r2l=r2l[:nR], depth=depth[:nL], p3d=p3d[:nL])


def ref_undistort(xy, cols, rows, K, dist):
    """The reference's own Frame::UndistortKeyPoints + ComputeImageBounds (cv::undistortPoints = the oracle's cv2-pinned one):
    returns (mvKeysUn xy [n,2], bounds [mnMinX, mnMaxX, mnMinY, mnMaxY])"""
    R = ref_frame_lib()
    f = lambda a: np.ascontiguousarray(a, np.float32)
    xy, K, dist = f(xy).reshape(-1, 2), f(K), f(dist)
    out = np.zeros((max(len(xy), 1), 2), np.float32); b = np.zeros(4, np.float32)
    R.ftref_undistort(xy if len(xy) else np.zeros((1, 2), np.float32), len(xy), int(cols), int(rows), K,
                      dist if len(dist) else np.zeros(1, np.float32), len(dist), out, b)
    return out[:len(xy)], b


def ref_stereo_from_rgbd(keys_xy, keys_un_x, depth, mbf):
    R = ref_frame_lib()
    f = lambda a: np.ascontiguousarray(a, np.float32)
    keys_xy, keys_un_x, depth = f(keys_xy), f(keys_un_x), f(depth)
    n = len(keys_un_x)
    ur = np.zeros(max(n, 1), np.float32); dp = np.zeros(max(n, 1), np.float32)
    R.ftref_stereo_from_rgbd(keys_xy, keys_un_x, n, depth, depth.shape[1], depth.shape[0], mbf, ur, dp)
    return ur[:n], dp[:n]


def ref_search_by_bow(kf_desc, kf_angle, kf_node, kf_has_mp, f_desc, f_angle, f_node, f_nleft=-1, nnratio=0.7, check_ori=True):
    """the reference's own ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) over its own DBoW2::FeatureVector"""
    R = ref_frame_lib()
    c8 = lambda a: np.ascontiguousarray(a, np.uint8)
    kf_desc, f_desc, kf_has_mp = c8(kf_desc).reshape(-1, 32), c8(f_desc).reshape(-1, 32), c8(kf_has_mp)
    kf_angle, f_angle = np.ascontiguousarray(kf_angle, np.float32), np.ascontiguousarray(f_angle, np.float32)
    kf_node, f_node = np.ascontiguousarray(kf_node, np.int32), np.ascontiguousarray(f_node, np.int32)
    match = np.full(max(len(f_desc), 1), -1, np.int32)
    nm = R.ftref_search_by_bow(len(kf_desc), kf_desc, kf_angle, kf_node, kf_has_mp, len(f_desc), f_desc, f_angle, f_node,
                               int(f_nleft), float(nnratio), int(check_ori), match)
    return nm, match[:len(f_desc)]
