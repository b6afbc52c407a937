"""fasttrack_b200 -- B200-native stereo tracking front-end (ORB extract -> stereo match -> SearchByProjection).

The product is the C-ABI shared library built from fasttrack_b200/csrc (see include/fasttrack_b200.h);
this module is a thin ctypes binding used by the tests and bench.py. There is no CPU fallback: creating a
Context without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = ["Config", "Context", "Vocabulary", "FtError", "load_library", "library_path", "KEYPOINT_DTYPE"]

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4")])

EXPORTS = [
    "ft_host_alloc", "ft_host_free", "ft_host_register", "ft_host_unregister",
    "ft_last_error", "ft_version", "ft_context_create", "ft_context_destroy", "ft_get_scale_tables",
    "ft_extract_stereo", "ft_extract_stereo_device", "ft_stereo_match", "ft_stereo_match_fisheye", "ft_frame_counts",
    "ft_frame_download", "ft_set_pose", "ft_search_local_points", "ft_synchronize", "ft_debug_level_dims",
    "ft_debug_level_image", "ft_debug_level_candidates", "ft_debug_track", "ft_debug_grid", "ft_debug_stats",
    "ft_context_stream", "ft_set_use_graph", "ft_launch_counts", "ft_upload_map_points", "ft_upload_holders",
    "ft_search_resident", "ft_search_download", "ft_set_stage_timing", "ft_get_stage_times", "ft_debug_level_counts", "ft_debug_sort", "ft_max_keypoints", "ft_frame_construct", "ft_frame_enqueue_device", "ft_map_point_staging", "ft_search_staged", "ft_search_last_frame", "ft_set_rectification", "ft_bind_map_points_device", "ft_frame_submit", "ft_frame_collect", "ft_set_sensor", "ft_extract_mono", "ft_depth_from_rgbd", "ft_debug_sincosf", "ft_set_input_resize", "ft_set_distortion", "ft_image_bounds", "ft_frame_keypoints_undistorted", "ft_map_store_create", "ft_map_store_attach", "ft_map_store_update", "ft_search_store", "ft_search_store_submit", "ft_search_collect",
    "ft_vocabulary_load_text", "ft_vocabulary_create", "ft_vocabulary_destroy", "ft_vocabulary_info", "ft_vocabulary_transform",
    "ft_compute_bow", "ft_bow_download", "ft_search_by_bow",
]

STAGES = ["copy_level0", "resize", "blur", "fast_cells", "octree", "orient_desc", "grid", "stereo_match",
          "stereo_outliers", "frustum", "gather", "resolve", "fast_cells_l0", "octree_l0", "blur_l0"]


class FtError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("fasttrack_b200 status %d: %s" % (status, msg))
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("device_id", C.c_int), ("width", C.c_int), ("height", C.c_int), ("nfeatures", C.c_int), ("nlevels", C.c_int),
        ("scale_factor", C.c_float), ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int), ("camera_type", C.c_int),
        ("cam1", C.c_float * 8), ("cam2", C.c_float * 8), ("lap_left", C.c_int * 2), ("lap_right", C.c_int * 2),
        ("bf", C.c_float), ("Tlr", C.c_float * 12), ("max_map_points", C.c_int),
    ]


def library_path():
    return _build.SO


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.SO):
        raise FtError(-1, "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                      % _build.SO)
    L = C.CDLL(_build.SO)
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.ft_last_error.restype = C.c_char_p
    L.ft_version.restype = C.c_char_p
    L.ft_context_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ft_context_destroy.argtypes = [vp]
    L.ft_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.ft_host_free.argtypes = [vp]
    L.ft_host_register.argtypes = [vp, C.c_size_t]
    L.ft_host_unregister.argtypes = [vp]
    L.ft_get_scale_tables.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ft_extract_stereo.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.ft_extract_stereo_device.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.ft_stereo_match.argtypes = [vp]
    L.ft_stereo_match_fisheye.argtypes = [vp]
    L.ft_frame_counts.argtypes = [vp, ip, ip, ip, ip]
    L.ft_frame_download.argtypes = [vp, C.c_int, C.c_int, vp, vp, ip, ip, vp, vp, vp, vp, vp]
    L.ft_set_pose.argtypes = [vp, vp, vp, vp, vp]
    L.ft_search_local_points.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_float, C.c_float,
                                         vp, vp, vp, ip]
    L.ft_synchronize.argtypes = [vp]
    L.ft_debug_level_dims.argtypes = [vp, C.c_int, ip, ip]
    L.ft_debug_level_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.ft_debug_level_candidates.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, ip]
    L.ft_debug_track.argtypes = [vp, C.c_int, vp, vp]
    L.ft_debug_grid.argtypes = [vp, C.c_int, vp, vp, ip]
    L.ft_debug_stats.argtypes = [vp, vp, C.c_int]
    L.ft_context_stream.restype = vp
    L.ft_context_stream.argtypes = [vp]
    L.ft_set_use_graph.argtypes = [vp, C.c_int]
    L.ft_launch_counts.argtypes = [vp, ip, ip, ip]
    L.ft_upload_map_points.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.ft_upload_holders.argtypes = [vp, C.c_int, vp, vp]
    L.ft_search_resident.argtypes = [vp, C.c_float, C.c_int, C.c_float, C.c_float]
    L.ft_search_download.argtypes = [vp, vp, vp, vp, ip]
    L.ft_set_stage_timing.argtypes = [vp, C.c_int]
    L.ft_get_stage_times.argtypes = [vp, vp, C.c_int]
    L.ft_debug_level_counts.argtypes = [vp, C.c_int, vp, vp]
    L.ft_debug_sort.argtypes = [vp, C.c_int]
    L.ft_max_keypoints.argtypes = [vp]
    L.ft_bind_map_points_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.ft_set_rectification.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
    L.ft_search_last_frame.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp, vp, vp, ip]
    L.ft_map_point_staging.argtypes = [vp, C.c_int] + [C.POINTER(vp)] * 7
    L.ft_search_staged.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), ip]
    L.ft_frame_enqueue_device.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.ft_frame_construct.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ft_frame_submit.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.ft_map_store_create.argtypes = [vp, C.c_int]
    L.ft_set_distortion.argtypes = [vp, vp, C.c_int]
    L.ft_set_input_resize.argtypes = [vp, C.c_int, C.c_int]
    L.ft_set_sensor.argtypes = [vp, C.c_int]
    L.ft_extract_mono.argtypes = [vp, vp, C.c_int]
    L.ft_depth_from_rgbd.argtypes = [vp, vp, C.c_int]
    L.ft_debug_sincosf.argtypes = [C.c_int, vp, vp, vp]
    L.ft_debug_sincosf.restype = None
    L.ft_image_bounds.argtypes = [vp, vp]
    L.ft_frame_keypoints_undistorted.argtypes = [vp, C.c_int, vp, vp]
    L.ft_map_store_attach.argtypes = [vp, vp]
    L.ft_map_store_update.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.ft_search_store.argtypes = [vp, C.c_int, vp, vp, C.c_float, C.c_int, C.c_float, C.c_float, vp, vp, vp, vp]
    L.ft_search_store_submit.argtypes = [vp, C.c_int, vp, vp, C.c_float, C.c_int, C.c_float, C.c_float, vp, vp, C.c_int]
    L.ft_search_collect.argtypes = [vp, vp, vp, vp, vp]
    L.ft_frame_collect.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ft_vocabulary_load_text.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.ft_vocabulary_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(vp)]
    L.ft_vocabulary_destroy.argtypes = [vp]
    L.ft_vocabulary_info.argtypes = [vp, ip, ip, ip, ip, ip, ip]
    L.ft_vocabulary_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, ip]
    L.ft_compute_bow.argtypes = [vp, vp, C.c_int]
    L.ft_bow_download.argtypes = [vp, C.c_int, vp, vp, vp, vp, ip, ip]
    L.ft_search_by_bow.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_float, C.c_int, vp, ip]
    for name in EXPORTS:
        if name not in ("ft_last_error", "ft_version", "ft_context_stream"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


class Vocabulary:
    """ORBVocabulary on one device (ft_vocabulary_*): shared read-only by the contexts of that device."""

    def __init__(self, handle):
        self.L = load_library()
        self.h = handle
        v = [C.c_int() for _ in range(6)]
        self._ck(self.L.ft_vocabulary_info(self.h, *[C.byref(x) for x in v]))
        self.k, self.depth, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in v]

    def _ck(self, st):
        if st != 0:
            raise FtError(st, self.L.ft_last_error().decode())

    @classmethod
    def load_text(cls, path, device_id=0):
        L = load_library()
        h = C.c_void_p()
        st = L.ft_vocabulary_load_text(device_id, str(path).encode(), C.byref(h))
        if st != 0:
            raise FtError(st, L.ft_last_error().decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, k, depth, scoring, weighting, parent, is_leaf, desc, weight, device_id=0):
        L = load_library()
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        st = L.ft_vocabulary_create(device_id, k, depth, scoring, weighting, len(parent), _ptr(parent), _ptr(is_leaf), _ptr(desc),
                                    _ptr(weight), C.byref(h))
        if st != 0:
            raise FtError(st, L.ft_last_error().decode())
        return cls(h)

    def transform(self, desc, levelsup=4):
        """ORBVocabulary::transform on host descriptors: dict(word[n], node[n] (-1 = stopped), bow_ids, bow_vals)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(max(n, 1), np.int32); node = np.zeros(max(n, 1), np.int32)
        ids = np.zeros(max(n, 1), np.uint32); vals = np.zeros(max(n, 1), np.float64)
        nb = C.c_int()
        self._ck(self.L.ft_vocabulary_transform(self.h, _ptr(desc) if n else None, n, levelsup, _ptr(word), _ptr(node), _ptr(ids),
                                                _ptr(vals), max(n, 1), C.byref(nb)))
        return dict(word=word[:n], node=node[:n], bow_ids=ids[:nb.value].copy(), bow_vals=vals[:nb.value].copy())

    def close(self):
        if getattr(self, "h", None):
            self.L.ft_vocabulary_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One (GPU, rig, sequence) front-end context; mirrors the C ABI one to one."""

    def __init__(self, width, height, nfeatures=1200, nlevels=8, scale_factor=1.2, ini_th=20, min_th=7,
                 camera_type=0, cam1=None, cam2=None, lap_left=(0, 0), lap_right=(0, 0), bf=0.0, Tlr=None,
                 max_map_points=25000, device_id=0):
        self.L = load_library()
        cfg = Config()
        cfg.device_id = device_id
        cfg.width, cfg.height, cfg.nfeatures, cfg.nlevels = width, height, nfeatures, nlevels
        cfg.scale_factor, cfg.ini_th_fast, cfg.min_th_fast, cfg.camera_type = scale_factor, ini_th, min_th, camera_type
        cam1 = np.zeros(8, np.float32) if cam1 is None else np.asarray(cam1, np.float32)
        cam2 = cam1 if cam2 is None else np.asarray(cam2, np.float32)
        for i in range(8):
            cfg.cam1[i] = float(cam1[i]) if i < len(cam1) else 0.0
            cfg.cam2[i] = float(cam2[i]) if i < len(cam2) else 0.0
        cfg.lap_left[0], cfg.lap_left[1] = lap_left
        cfg.lap_right[0], cfg.lap_right[1] = lap_right
        cfg.bf = bf
        T = np.hstack([np.eye(3), np.zeros((3, 1))]) if Tlr is None else np.asarray(Tlr, np.float64).reshape(3, 4)
        for i in range(12):
            cfg.Tlr[i] = float(T.reshape(-1)[i])
        cfg.max_map_points = max_map_points
        self.cfg = cfg
        self.width, self.height, self.nlevels, self.nfeatures = width, height, nlevels, nfeatures
        self.fisheye = camera_type == 1
        h = C.c_void_p()
        self._ck(self.L.ft_context_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.cap = max(nfeatures + 64 * nlevels, int(self.L.ft_max_keypoints(self.h)))

    def _ck(self, st):
        if st != 0:
            raise FtError(st, self.L.ft_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.ft_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- tables ----
    def scale_tables(self):
        n = self.nlevels
        a = [np.zeros(n, np.float32) for _ in range(4)]
        q = np.zeros(n, np.int32)
        self._ck(self.L.ft_get_scale_tables(self.h, *[_ptr(x) for x in a], _ptr(q)))
        return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3], features_per_level=q)

    def set_rectification(self, raw_w, raw_h, m1l, m2l, m1r, m2r):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        maps = [f(m1l), f(m2l), f(m1r), f(m2r)]
        self._ck(self.L.ft_set_rectification(self.h, raw_w, raw_h, *[_ptr(m) for m in maps]))
        self._raw = (raw_h, raw_w) if maps[0] is not None else None

    # ---- per-frame operators ----
    def extract_stereo(self, imgL, imgR):
        assert imgL.dtype == np.uint8 and imgR.dtype == np.uint8 and imgL.strides[1] == 1 and imgR.strides[1] == 1
        shape = getattr(self, "_raw", None) or (self.height, self.width)
        assert imgL.shape == shape and imgR.shape == shape
        self._ck(self.L.ft_extract_stereo(self.h, imgL.ctypes.data, imgL.strides[0], imgR.ctypes.data, imgR.strides[0]))

    def extract_stereo_ptr(self, ptrL, stepL, ptrR, stepR, device=False):
        f = self.L.ft_extract_stereo_device if device else self.L.ft_extract_stereo
        self._ck(f(self.h, ptrL, stepL, ptrR, stepR))

    def frame_enqueue_device(self, ptrL, stepL, ptrR, stepR):
        self._ck(self.L.ft_frame_enqueue_device(self.h, ptrL, stepL, ptrR, stepR))

    def stereo_match(self):
        self._ck(self.L.ft_stereo_match_fisheye(self.h) if self.fisheye else self.L.ft_stereo_match(self.h))

    def counts(self):
        v = [C.c_int() for _ in range(4)]
        self._ck(self.L.ft_frame_counts(self.h, *[C.byref(x) for x in v]))
        return dict(n_left=v[0].value, n_right=v[1].value, mono_left=v[2].value, mono_right=v[3].value)

    def download(self, eye, stereo=False):
        """returns dict(kps structured array, desc[n,32], mono_index, and stereo outputs when requested)"""
        kps = np.zeros(self.cap, KEYPOINT_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n, mono = C.c_int(), C.c_int()
        out = {}
        ur = dp = l2r = r2l = p3d = None
        if stereo:
            ur = np.zeros(self.cap, np.float32); dp = np.zeros(self.cap, np.float32)
            if self.fisheye:
                l2r = np.zeros(self.cap, np.int32); r2l = np.zeros(self.cap, np.int32)
                p3d = np.zeros((self.cap, 3), np.float32)
        self._ck(self.L.ft_frame_download(self.h, eye, self.cap, _ptr(kps), _ptr(desc), C.byref(n), C.byref(mono),
                                          _ptr(ur), _ptr(dp), _ptr(l2r), _ptr(r2l), _ptr(p3d)))
        n = n.value
        out.update(kps=kps[:n].copy(), desc=desc[:n].copy(), n=n, mono_index=mono.value)
        if stereo:
            c = self.counts()
            nl, nr = c["n_left"], c["n_right"]
            out.update(u_right=ur[:nl].copy(), depth=dp[:nl].copy())
            if self.fisheye:
                out.update(l2r=l2r[:nl].copy(), r2l=r2l[:nr].copy(), p3d=p3d[:nl].copy())
        return out

    def frame_construct(self, imgL, imgR):
        """Frame constructor in one call: returns (left dict, right dict) like download()"""
        self.frame_submit(imgL, imgR)
        return self.frame_collect()

    def frame_submit(self, imgL, imgR):
        """asynchronous half of frame_construct: upload + extract + stereo + result download enqueued; the images are
        kept alive until frame_collect()"""
        self._keep_frame = (imgL, imgR)
        self._ck(self.L.ft_frame_submit(self.h, imgL.ctypes.data, imgL.strides[0], imgR.ctypes.data, imgR.strides[0]))

    def frame_collect(self):
        """waits for the submitted frame; returns (left dict, right dict) like download()"""
        cap = self.cap
        kL = np.zeros(cap, KEYPOINT_DTYPE); kR = np.zeros(cap, KEYPOINT_DTYPE)
        dL = np.zeros((cap, 32), np.uint8); dR = np.zeros((cap, 32), np.uint8)
        ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32)
        cnt = np.zeros(4, np.int32)
        l2r = r2l = p3d = None
        if self.fisheye:
            l2r = np.zeros(cap, np.int32); r2l = np.zeros(cap, np.int32); p3d = np.zeros((cap, 3), np.float32)
        self._ck(self.L.ft_frame_collect(self.h, _ptr(kL), _ptr(dL), _ptr(kR), _ptr(dR), _ptr(cnt), _ptr(ur), _ptr(dp),
                                         _ptr(l2r), _ptr(r2l), _ptr(p3d)))
        self._keep_frame = None
        nl, ml, nr, mr = [int(x) for x in cnt]
        left = dict(kps=kL[:nl].copy(), desc=dL[:nl].copy(), n=nl, mono_index=ml, u_right=ur[:nl].copy(), depth=dp[:nl].copy())
        right = dict(kps=kR[:nr].copy(), desc=dR[:nr].copy(), n=nr, mono_index=mr)
        if self.fisheye:
            left.update(l2r=l2r[:nl].copy(), r2l=r2l[:nr].copy(), p3d=p3d[:nl].copy())
        return left, right

    # ---- monocular / RGB-D ----
    def set_sensor(self, sensor):
        """0 stereo (default), 1 monocular, 2 RGB-D"""
        self._ck(self.L.ft_set_sensor(self.h, int(sensor)))

    def extract_mono(self, img):
        assert img.dtype == np.uint8 and img.strides[1] == 1
        self._ck(self.L.ft_extract_mono(self.h, img.ctypes.data, img.strides[0]))

    def depth_from_rgbd(self, depth=None):
        """Frame::ComputeStereoFromRGBD with a float32 depth image (None on a monocular context); builds the frame grid"""
        if depth is None:
            self._ck(self.L.ft_depth_from_rgbd(self.h, None, 0))
            return
        assert depth.dtype == np.float32 and depth.strides[1] == 4
        self._keep_depth = depth
        self._ck(self.L.ft_depth_from_rgbd(self.h, depth.ctypes.data, depth.strides[0]))

    def set_input_resize(self, raw_width, raw_height):
        """cv::resize of the raw input to the camera size in front of the extractor (System.cc:282-285); 0 = off"""
        self._ck(self.L.ft_set_input_resize(self.h, int(raw_width), int(raw_height)))
        self._raw = (int(raw_height), int(raw_width)) if raw_width and raw_height else None

    # ---- pinhole distortion (Frame::UndistortKeyPoints / ComputeImageBounds) ----
    def set_distortion(self, dist_coef):
        d = np.ascontiguousarray([] if dist_coef is None else dist_coef, np.float32)
        self._ck(self.L.ft_set_distortion(self.h, _ptr(d) if len(d) else None, len(d)))

    def image_bounds(self):
        out = np.zeros(4, np.float32)
        self._ck(self.L.ft_image_bounds(self.h, _ptr(out)))
        return out

    def keypoints_undistorted(self):
        xy = np.zeros((self.cap, 2), np.float32)
        n = C.c_int()
        self._ck(self.L.ft_frame_keypoints_undistorted(self.h, self.cap, _ptr(xy), C.byref(n)))
        return xy[:n.value].copy()

    # ---- persistent map store ----
    def map_store_create(self, capacity):
        self._ck(self.L.ft_map_store_create(self.h, int(capacity)))

    def map_store_attach(self, owner):
        self._ck(self.L.ft_map_store_attach(self.h, owner.h))

    def map_store_update(self, slots, pos, normal, minmax, desc):
        slots = np.ascontiguousarray(slots, np.int32)
        f = lambda a: np.ascontiguousarray(a, np.float32)
        pos, normal, minmax = f(pos), f(normal), f(minmax)
        desc = np.ascontiguousarray(desc, np.uint8)
        self._ck(self.L.ft_map_store_update(self.h, len(slots), _ptr(slots), _ptr(pos), _ptr(normal), _ptr(minmax), _ptr(desc)))

    def search_store(self, slots, flags, th, holder, holder_obs, b_far=False, th_far=50.0, nnratio=0.8):
        slots = np.ascontiguousarray(slots, np.int32); flags = np.ascontiguousarray(flags, np.int32)
        holder = np.ascontiguousarray(holder, np.int32).copy()
        holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        M = len(slots)
        best = np.full((max(M, 1), 2), -1, np.int32)
        nm = C.c_int()
        self._ck(self.L.ft_search_store(self.h, M, _ptr(slots), _ptr(flags), th, int(b_far), th_far, nnratio,
                                        _ptr(holder), _ptr(holder_obs), _ptr(best), C.byref(nm)))
        return nm.value, holder, holder_obs, best[:M]

    def search_store_submit(self, slots, flags, th, holder, holder_obs, b_far=False, th_far=50.0, nnratio=0.8, want_best=True):
        """first half of search_store: everything is enqueued, nothing is waited for (ft_search_store_submit)"""
        slots = np.ascontiguousarray(slots, np.int32); flags = np.ascontiguousarray(flags, np.int32)
        holder = np.ascontiguousarray(holder, np.int32).copy()
        holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        self._ck(self.L.ft_search_store_submit(self.h, len(slots), _ptr(slots), _ptr(flags), th, int(b_far), th_far, nnratio,
                                               _ptr(holder), _ptr(holder_obs), int(want_best)))
        self._pending_search = (len(slots), holder, holder_obs)

    def search_collect(self):
        """second half: waits for the submitted search; returns what search_store returns (ft_search_collect)"""
        M, holder, holder_obs = self._pending_search
        best = np.full((max(M, 1), 2), -1, np.int32)
        nm = C.c_int()
        self._ck(self.L.ft_search_collect(self.h, _ptr(holder), _ptr(holder_obs), _ptr(best), C.byref(nm)))
        return nm.value, holder, holder_obs, best[:M]

    def set_pose(self, Rcw, tcw, Rwc=None, Ow=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32).reshape(-1)
        Rcw, tcw, Rwc, Ow = f(Rcw), f(tcw), f(Rwc), f(Ow)
        self._keep_pose = (Rcw, tcw, Rwc, Ow)
        self._ck(self.L.ft_set_pose(self.h, _ptr(Rcw), _ptr(tcw), _ptr(Rwc), _ptr(Ow)))

    def search_local_points(self, pos, normal, minmax, desc, flags, th, holder, holder_obs, b_far=False, th_far=50.0,
                            nnratio=0.8):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        pos, normal, minmax = f(pos), f(normal), f(minmax)
        desc = np.ascontiguousarray(desc, np.uint8); flags = np.ascontiguousarray(flags, np.int32)
        holder = np.ascontiguousarray(holder, np.int32).copy()
        holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        M = len(pos)
        best = np.full((max(M, 1), 2), -1, np.int32)
        nm = C.c_int()
        self._ck(self.L.ft_search_local_points(self.h, M, _ptr(pos), _ptr(normal), _ptr(minmax), _ptr(desc), _ptr(flags),
                                               th, int(b_far), th_far, nnratio, _ptr(holder), _ptr(holder_obs),
                                               _ptr(best), C.byref(nm)))
        return nm.value, holder, holder_obs, best[:M]

    def search_local_points_raw(self, M, pos, normal, minmax, desc, flags, th, holder, holder_obs, best, b_far=False,
                                th_far=50.0, nnratio=0.8):
        """pointer-level call for bench.py (no numpy copies); arguments are integer addresses"""
        nm = C.c_int()
        self._ck(self.L.ft_search_local_points(self.h, M, pos, normal, minmax, desc, flags, th, int(b_far), th_far,
                                               nnratio, holder, holder_obs, best, C.byref(nm)))
        return nm.value

    # ---- resident projection search (map-point snapshot stays on the device) ----
    def upload_map_points(self, pos, normal, minmax, desc, flags):
        f = lambda a: np.ascontiguousarray(a, np.float32)
        self._mp_keep = (f(pos), f(normal), f(minmax), np.ascontiguousarray(desc, np.uint8),
                         np.ascontiguousarray(flags, np.int32))
        self._ck(self.L.ft_upload_map_points(self.h, len(self._mp_keep[0]), *[_ptr(a) for a in self._mp_keep]))

    def bind_map_points_device(self, M, d_pos, d_normal, d_minmax, d_desc, d_flags):
        """device pointers (ints) of a snapshot that already lives in HBM; used in place"""
        self._ck(self.L.ft_bind_map_points_device(self.h, M, d_pos, d_normal, d_minmax, d_desc, d_flags))

    def upload_holders(self, holder=None, holder_obs=None):
        if holder is None:
            self._ck(self.L.ft_upload_holders(self.h, 0, None, None))
            return
        self._h_keep = (np.ascontiguousarray(holder, np.int32), np.ascontiguousarray(holder_obs, np.uint8))
        self._ck(self.L.ft_upload_holders(self.h, len(self._h_keep[0]), _ptr(self._h_keep[0]), _ptr(self._h_keep[1])))

    def search_resident(self, th, b_far=False, th_far=50.0, nnratio=0.8):
        self._ck(self.L.ft_search_resident(self.h, th, int(b_far), th_far, nnratio))

    def search_download(self, M):
        c = self.counts()
        N = c["n_left"] + (c["n_right"] if self.fisheye else 0)
        holder = np.zeros(max(N, 1), np.int32); hobs = np.zeros(max(N, 1), np.uint8)
        best = np.full((max(M, 1), 2), -1, np.int32)
        nm = C.c_int()
        self._ck(self.L.ft_search_download(self.h, _ptr(holder), _ptr(hobs), _ptr(best), C.byref(nm)))
        return nm.value, holder[:N], hobs[:N], best[:M]

    def map_point_staging(self, M, N):
        """numpy views over the context's pinned staging buffers for M map points and N frame keypoints"""
        p = [C.c_void_p() for _ in range(7)]
        self._ck(self.L.ft_map_point_staging(self.h, M, *[C.byref(x) for x in p]))
        def view(ptr, ctype, count, dtype, shape):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).view(dtype).reshape(shape)
        return dict(pos=view(p[0], C.c_float, 3 * M, np.float32, (M, 3)), normal=view(p[1], C.c_float, 3 * M, np.float32, (M, 3)),
                    minmax=view(p[2], C.c_float, 2 * M, np.float32, (M, 2)), desc=view(p[3], C.c_uint8, 32 * M, np.uint8, (M, 32)),
                    flags=view(p[4], C.c_int, M, np.int32, (M,)), holder=view(p[5], C.c_int, N, np.int32, (N,)),
                    holder_obs=view(p[6], C.c_uint8, N, np.uint8, (N,)))

    def search_staged(self, M, N, th, b_far=False, th_far=50.0, nnratio=0.8):
        """search over the snapshot written into map_point_staging(); returns views into pinned result memory"""
        ho, hb, bi = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nm = C.c_int()
        self._ck(self.L.ft_search_staged(self.h, M, th, int(b_far), th_far, nnratio, C.byref(ho), C.byref(hb), C.byref(bi),
                                         C.byref(nm)))
        holder = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_int)), shape=(max(N, 1),))[:N]
        hobs = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_uint8)), shape=(max(N, 1),))[:N]
        best = np.ctypeslib.as_array(C.cast(bi, C.POINTER(C.c_int)), shape=(max(2 * M, 1),))[:2 * M].reshape(M, 2)
        return nm.value, holder, hobs, best

    def search_last_frame(self, pos, desc, octave, angle, flags, Rlw, tlw, th, holder, holder_obs, b_mono=False,
                          check_ori=True):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)"""
        f = lambda a: np.ascontiguousarray(a, np.float32)
        pos, angle, Rlw, tlw = f(pos), f(angle), f(Rlw).reshape(-1), f(tlw)
        desc = np.ascontiguousarray(desc, np.uint8)
        octave = np.ascontiguousarray(octave, np.int32); flags = np.ascontiguousarray(flags, np.int32)
        holder = np.ascontiguousarray(holder, np.int32).copy(); holder_obs = np.ascontiguousarray(holder_obs, np.uint8).copy()
        n = len(pos)
        best = np.full((max(n, 1), 2), -1, np.int32)
        nm = C.c_int()
        self._ck(self.L.ft_search_last_frame(self.h, n, _ptr(pos), _ptr(desc), _ptr(octave), _ptr(angle), _ptr(flags), _ptr(Rlw),
                                             _ptr(tlw), th, int(b_mono), int(check_ori), _ptr(holder), _ptr(holder_obs),
                                             _ptr(best), C.byref(nm)))
        return nm.value, holder, holder_obs, best[:n]

    # ---- bag of words ----
    def compute_bow(self, voc, levelsup=4):
        """Frame::ComputeBoW on the device-resident descriptors (asynchronous)"""
        self._ck(self.L.ft_compute_bow(self.h, voc.h, int(levelsup)))

    def bow_download(self):
        """mBowVec / mFeatVec of the frame: dict(word[N], node[N], bow_ids, bow_vals)"""
        cap = 2 * self.cap
        word = np.zeros(cap, np.int32); node = np.zeros(cap, np.int32)
        ids = np.zeros(cap, np.uint32); vals = np.zeros(cap, np.float64)
        nb, n = C.c_int(), C.c_int()
        self._ck(self.L.ft_bow_download(self.h, cap, _ptr(word), _ptr(node), _ptr(ids), _ptr(vals), C.byref(nb), C.byref(n)))
        return dict(word=word[:n.value].copy(), node=node[:n.value].copy(), bow_ids=ids[:nb.value].copy(),
                    bow_vals=vals[:nb.value].copy())

    def search_by_bow(self, kf_desc, kf_angle, kf_node, kf_has_mp, nnratio=0.7, check_ori=True):
        """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches): returns (nmatches, match[N])"""
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32)
        kf_angle = np.ascontiguousarray(kf_angle, np.float32); kf_node = np.ascontiguousarray(kf_node, np.int32)
        kf_has_mp = np.ascontiguousarray(kf_has_mp, np.uint8)
        n = len(kf_desc)
        match = np.full(2 * self.cap, -1, np.int32)
        nm = C.c_int()
        a = (lambda x: _ptr(x) if n else None)
        self._ck(self.L.ft_search_by_bow(self.h, n, a(kf_desc), a(kf_angle), a(kf_node), a(kf_has_mp), float(nnratio),
                                         int(check_ori), _ptr(match), C.byref(nm)))
        c = self.counts()
        N = c["n_left"] + (c["n_right"] if self.fisheye else 0)
        return nm.value, match[:N].copy()

    def set_stage_timing(self, enable):
        self._ck(self.L.ft_set_stage_timing(self.h, int(enable)))

    def stage_times(self):
        ms = np.zeros(len(STAGES), np.float32)
        self._ck(self.L.ft_get_stage_times(self.h, _ptr(ms), len(STAGES)))
        return {k: float(v) for k, v in zip(STAGES, ms) if v >= 0}

    def synchronize(self):
        self._ck(self.L.ft_synchronize(self.h))

    def stream(self):
        return self.L.ft_context_stream(self.h)

    def set_use_graph(self, enable):
        self._ck(self.L.ft_set_use_graph(self.h, int(enable)))

    def launch_counts(self):
        v = [C.c_int() for _ in range(3)]
        self._ck(self.L.ft_launch_counts(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # ---- diagnostics ----
    def level_dims(self, level):
        w, h = C.c_int(), C.c_int()
        self._ck(self.L.ft_debug_level_dims(self.h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level_image(self, eye, level, blurred=False):
        w, h = self.level_dims(level)
        out = np.zeros((h, w), np.uint8)
        self._ck(self.L.ft_debug_level_image(self.h, eye, level, int(blurred), _ptr(out)))
        return out

    def level_candidates(self, eye, level):
        n = C.c_int()
        self._ck(self.L.ft_debug_level_candidates(self.h, eye, level, 0, None, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.float32)
        self._ck(self.L.ft_debug_level_candidates(self.h, eye, level, n.value, _ptr(out), C.byref(n)))
        return out[: n.value]

    def track(self, M):
        ti = np.zeros((max(M, 1), 4), np.int32); tf = np.zeros((max(M, 1), 9), np.float32)
        self._ck(self.L.ft_debug_track(self.h, M, _ptr(ti), _ptr(tf)))
        return ti[:M], tf[:M]

    def grid(self, right=False):
        counts = np.zeros(64 * 48, np.int32)
        idx = np.zeros(2 * self.cap, np.int32)
        n = C.c_int()
        self._ck(self.L.ft_debug_grid(self.h, int(right), _ptr(counts), _ptr(idx), C.byref(n)))
        return counts, idx[: n.value].copy()

    def level_counts(self, eye):
        cand = np.zeros(self.nlevels, np.int32); kp = np.zeros(self.nlevels, np.int32)
        self._ck(self.L.ft_debug_level_counts(self.h, eye, _ptr(cand), _ptr(kp)))
        return cand, kp

    def stats(self):
        s = np.zeros(8, np.int64)
        self._ck(self.L.ft_debug_stats(self.h, _ptr(s), 8))
        names = ["cand_left", "cand_right", "kp_left", "kp_right", "stereo_tested", "stereo_refined", "sbp_candidates",
                 "sbp_rounds"]
        return dict(zip(names, [int(x) for x in s]))


def keypoints_as_array(kps):
    """structured ft_keypoint array -> float32 [n,6] (x,y,size,angle,response,octave), the oracle's layout"""
    out = np.zeros((len(kps), 6), np.float32)
    for i, k in enumerate(("x", "y", "size", "angle", "response")):
        out[:, i] = kps[k]
    out[:, 5] = kps["octave"]
    return out


def device_sort(words):
    """test hook: sort uint64 words by their high 32 bits with the device introsort"""
    a = np.ascontiguousarray(words, np.uint64).copy()
    st = load_library().ft_debug_sort(a.ctypes.data, len(a))
    if st != 0:
        raise FtError(st, "ft_debug_sort failed")
    return a
