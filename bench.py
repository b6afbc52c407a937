#!/usr/bin/env python
"""bench.py -- stereo front-end throughput on B200 (contract in the task brief).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = one synthetic EuRoC-shaped stereo frame (752x480, 1200 features, 8 levels, x1.2) through
extract(L,R) -> ComputeStereoMatches -> SearchLocalPoints(M map points). One independent sequence per GPU
(no collective on the data path; torch.distributed over gloo is only used for the barrier and the max-over-ranks
time: no NCCL anywhere). Prints ONE JSON line on rank 0. Besides the headline the line carries `configs` (every
BASELINE.json configuration: CPU-extractor image, pinhole stereo pair, fisheye pair at 1000 / 2000 features, projection
search at M = 5k / 10k / 20k x th = 1 / 2 / 6 / 10), `next_rows` (last-frame search, rectification / input resize /
undistortion front, bag of words) and `reference_gpu` (the reference's own CUDA kernels timed on the same GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fasttrack_b200 import replicas, synth  # noqa: E402

E = synth.EUROC
N_FRAMES = 104           # distinct pre-generated frames per rank, cycled: 104 x (2 images + map snapshot) = 132 MB > L2 (126 MB)
M_POINTS = 8000          # local map size per frame
TH = 3.0
STORE_UPSERTS = 400     # rows the mapping side changes per frame in the map-store e2e leg
WORKLOAD = ("euroc_752x480_stereo_sequence: extract(L,R; 1200 features, 8 levels, x1.2) + ComputeStereoMatches + "
            "SearchLocalPoints(M=%d, th=%g)" % (M_POINTS, TH))
METRIC = "frames/sec (ORB extract L+R + stereo match + projection search)"


def shared_config(n_seq):
    """`config` of the JSON line, identical for both arms (the driver compares the two dicts): the workload, the number of
    independent sequences, and how the GPU arm keeps its inputs out of L2 (the CPU arm has no such cache to defeat)."""
    return {"workload": WORKLOAD, "sequences": n_seq,
            "l2": "GPU arm: throughput leg cycles %d distinct frames (inputs %.0f MB > 126 MB L2), latency leg flushes L2 between steps "
                  "(256 MiB write); CPU arm: n/a" % (N_FRAMES, N_FRAMES * (2 * E["width"] * E["height"] + 68 * M_POINTS) / 1e6)}


_T0 = time.perf_counter()


def log(msg):
    """progress on stderr with FT_BENCH_LOG=1 (which leg takes how long)"""
    if os.environ.get("FT_BENCH_LOG"):
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg)); sys.stderr.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock + throttle-reason sampler running during the timed regions (profiling recipe's clocks line): NVML
    polled every 20 ms from a thread; falls back to an `nvidia-smi -lms 100` child when NVML is unusable."""

    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, uuid=None):
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.stop_flag = threading.Event()
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            self.h = h or pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            self.source = "nvml"
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.BITS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.th.join(timeout=1)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def make_frames(seed, n):
    sc = synth.StereoScene(seed=seed)
    out = []
    for t in range(n):
        out.append(sc.pair(pan=sc.sequence_pan(7 * t), noise_seed=seed * 100 + t))
    return out


def fast_mappoints(keys, desc, scale_factors, M, seed):
    """vectorised variant of synth.mappoints for bench setup (same distribution, no per-point Python loop)"""
    rng = np.random.default_rng(seed)
    N = len(keys); nl = len(scale_factors)
    fx, fy, cx, cy = E["fx"], E["fy"], E["cx"], E["cy"]
    inside = rng.random(M) < 0.7
    anchored = (rng.random(M) < 0.5) & inside & (N > 0)
    z = rng.uniform(0.4, 25.0, M)
    lvl = rng.integers(0, nl, M)
    k = rng.integers(0, max(N, 1), M)
    u = rng.uniform(0, E["width"], M); v = rng.uniform(0, E["height"], M)
    d = rng.integers(0, 256, (M, 32), dtype=np.uint8)
    if N:
        lvl = np.where(anchored, keys[k, 5].astype(np.int64), lvl)
        jit = 1.5 * scale_factors[lvl]
        u = np.where(anchored, keys[k, 0] + rng.uniform(-1, 1, M) * jit, u)
        v = np.where(anchored, keys[k, 1] + rng.uniform(-1, 1, M) * jit, v)
        bits = np.unpackbits(desc[k], axis=1)
        nflip = rng.integers(0, 49, M)
        flip = rng.random((M, 256)).argsort(axis=1) < nflip[:, None]
        d = np.where(anchored[:, None], np.packbits(bits ^ flip.astype(np.uint8), axis=1), d)
    P = np.stack([(u - cx) * z / fx, (v - cy) * z / fy, z], 1)
    dist = np.linalg.norm(P, axis=1)
    n = P / dist[:, None] + 0.26 * rng.standard_normal((M, 3))
    n /= np.linalg.norm(n, axis=1)[:, None]
    maxd = dist * scale_factors[lvl] * 0.95
    mode = rng.integers(0, 4, M)
    out_ = ~inside
    P[out_ & (mode == 0), 2] *= -1
    P[out_ & (mode == 1), 0] += (2.0 * E["width"] / fx) * z[out_ & (mode == 1)]
    maxd = np.where(out_ & (mode == 2), dist * 0.3, maxd)
    n[out_ & (mode == 3)] *= -1
    mind = maxd / scale_factors[nl - 1]
    flags = np.full(M, 2, np.int32)
    flags[rng.random(M) < 0.02] |= 1
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(pos=f32(P), normal=f32(n), minmax=f32(np.stack([mind, maxd], 1)), desc=np.ascontiguousarray(d), flags=flags)


def bench_bow(ctx, ft, torch, stream, frames, device_id, with_cpu):
    """SURVEY.md 8f row 4 (bag of words) measured beside the headline: Frame::ComputeBoW (device time, CUDA events on the
    context's stream) and ORBmatcher::SearchByBoW(KeyFrame, Frame) (wall clock through the C ABI with host buffers) on an
    ORBvoc-shaped vocabulary (k = 10, L = 6, 1,111,110 nodes), with the oracle's CPU time for the same calls."""
    import fasttrack_b200.synth as synth
    K, LV, LEVELSUP, REP = 10, 6, 4, 50
    parent, leaf, vdesc, weight = synth.make_vocabulary_bfs(K, LV, seed=11)
    ctx.extract_stereo(frames[0][0], frames[0][1]); ctx.stereo_match()
    kf = ctx.download(0)
    spread = np.nonzero(leaf)[0][::max(1, int(leaf.sum()) // kf["n"])][:kf["n"]]     # plant the KeyFrame's descriptors as words
    vdesc[spread] = kf["desc"][:len(spread)]
    voc = ft.Vocabulary.from_arrays(K, LV, 0, 0, parent, leaf, vdesc, weight, device_id=device_id)
    kf_node = voc.transform(kf["desc"], LEVELSUP)["node"]
    kf_has = np.ones(kf["n"], np.uint8)
    ctx.extract_stereo(frames[1][0], frames[1][1]); ctx.stereo_match()
    fr = ctx.download(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ctx.compute_bow(voc, LEVELSUP)
    ctx.synchronize()
    with torch.cuda.stream(stream):
        ev[0].record()
    for _ in range(REP):
        ctx.compute_bow(voc, LEVELSUP)
    with torch.cuda.stream(stream):
        ev[1].record()
    ctx.synchronize()
    bow_ms = ev[0].elapsed_time(ev[1]) / REP
    nm, _ = ctx.search_by_bow(kf["desc"], kf["kps"]["angle"], kf_node, kf_has, 0.7, True)
    t0 = time.perf_counter()
    for _ in range(REP):
        ctx.search_by_bow(kf["desc"], kf["kps"]["angle"], kf_node, kf_has, 0.7, True)
    search_ms = (time.perf_counter() - t0) * 1e3 / REP
    n = fr["n"]
    alg = n * (LV * K * 32 + 32 + 8)            # descent reads k child descriptors per level + the feature + its outputs
    pk, _ = peaks()
    out = {"vocabulary": "synthetic, k=%d L=%d (%d nodes, %.1f MB of descriptors), levelsup=%d" % (K, LV, len(parent) + 1,
                                                                                                len(parent) * 32 / 1e6, LEVELSUP),
           "features": int(n), "compute_bow_ms": bow_ms, "compute_bow_launches": 2,
           "compute_bow_algorithmic_bytes": int(alg), "compute_bow_hbm_frac": alg / (bow_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
           "search_by_bow_ms_e2e": search_ms, "search_by_bow_matches": int(nm), "keyframe_features": int(kf["n"]),
           "note": "compute_bow: CUDA events around %d asynchronous calls; search_by_bow: host wall clock per call incl. the "
                   "KeyFrame upload and the match download" % REP}
    if with_cpu:
        import oracle
        vo = oracle.Vocabulary.from_arrays(K, LV, 0, 0, parent, leaf, vdesc, weight)
        t0 = time.perf_counter()
        for _ in range(10):
            fo = vo.transform(fr["desc"], LEVELSUP)
        out["cpu_compute_bow_ms"] = (time.perf_counter() - t0) * 1e3 / 10
        ang = np.ascontiguousarray(fr["kps"]["angle"]); kang = np.ascontiguousarray(kf["kps"]["angle"])
        t0 = time.perf_counter()
        for _ in range(10):
            nm_o, _ = oracle.search_by_bow(kf["desc"], kang, kf_node, kf_has, fr["desc"], ang, fo["node"], -1, 0.7, True)
        out["cpu_search_by_bow_ms"] = (time.perf_counter() - t0) * 1e3 / 10
        if nm_o != nm:
            raise SystemExit("bench.py: SearchByBoW disagrees with the oracle (%d vs %d matches)" % (nm, nm_o))
    voc.close()
    return out


def _dev_ms(torch, stream, fn, reps, flush=None, warm=3):   # flush(stream): L2 flush enqueued on the SAME stream as fn
    """mean CUDA-event ms of fn() enqueued on `stream` (events recorded on that stream, synchronised per repetition)"""
    for _ in range(warm):
        fn()
    stream.synchronize()
    t = []
    for _ in range(reps):
        if flush:
            flush(stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream)
        stream.synchronize()
        t.append(e0.elapsed_time(e1))
    return float(np.mean(t))


def _extract_bytes(ctx, st):
    """SURVEY.md 8d algorithmic bytes of one extraction (both eyes) with the measured counts"""
    px = [w * h for w, h in (ctx.level_dims(l) for l in range(ctx.nlevels))]
    sumP = sum(px)
    C_ = st["cand_left"] + st["cand_right"]; K_ = st["kp_left"] + st["kp_right"]
    return (2 * (sum(px[:-1]) + sum(px[1:])) + 2 * 2 * sumP + 2 * sumP + 16 * C_ + 16 * C_ + 20 * K_ + (749 + 4 + 512 + 32) * K_)


def bench_configs(ft, torch, local, frames, with_cpu, flush_l2):
    """Every BASELINE.json configuration and every built 'next' row of SURVEY.md 8f, each with its device time (CUDA events on
    the context's stream), its algorithmic bytes / HBM fraction and the oracle's CPU time beside it. Side legs never break the
    headline: a failing leg reports {"unavailable": ...}."""
    import fasttrack_b200.synth as synth
    pk, _ = peaks()
    dev = torch.device("cuda", local)
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    oracle = None
    if with_cpu:
        import oracle as _o
        oracle = _o
    out = {}
    frac = lambda b, ms: b / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    def leg(name, fn):
        log("config " + name)
        try:
            out[name] = fn()
        except Exception as e:   # noqa: BLE001
            out[name] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}

    L0, R0 = frames[0]

    # ---- BASELINE config 0: one 752x480 image through the extractor alone (monocular Frame) ----
    def cfg_mono():
        c = ft.Context(E["width"], E["height"], nfeatures=E["nfeatures"], nlevels=E["nlevels"], scale_factor=E["scale"],
                       cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), device_id=local)
        c.set_sensor(1)     # FT_SENSOR_MONOCULAR
        st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        hp = pin(L0)
        ms = _dev_ms(torch, st_, lambda: c._ck(c.L.ft_extract_mono(c.h, hp.data_ptr(), E["width"])), 30, flush_l2)
        n = c.counts()["n_left"]
        r = {"workload": "ORBextractor::operator() on one 752x480 image (1200 features, 8 levels): upload + extraction",
             "gpu_ms": ms, "keypoints": int(n), "timing": "CUDA events on the context's stream around ft_extract_mono (pinned host image, H2D inside)"}
        if oracle:
            ex = oracle.Extractor()
            t0 = time.perf_counter()
            for _ in range(5):
                ex.extract(L0)
            r["cpu_ms"] = (time.perf_counter() - t0) * 1e3 / 5
            r["cpu"] = "oracle port, 1 thread"
        c.close()
        return r
    leg("cpu_extractor_image_752x480", cfg_mono)

    # ---- BASELINE config 1: EuRoC-shaped pair, extract + ComputeStereoMatches ----
    def cfg_pinhole():
        c = ft.Context(E["width"], E["height"], nfeatures=E["nfeatures"], nlevels=E["nlevels"], scale_factor=E["scale"],
                       cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), device_id=local)
        st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        dl, dr = torch.from_numpy(L0).to(dev), torch.from_numpy(R0).to(dev)
        ms = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dl.data_ptr(), E["width"], dr.data_ptr(), E["width"]), 30, flush_l2)
        s_ = c.stats()
        b = _extract_bytes(c, s_) + 44 * (s_["kp_left"] + s_["kp_right"]) + 44 * s_["stereo_tested"] + 352 * s_["stereo_refined"] + 12 * s_["kp_left"]
        r = {"workload": "752x480 pair: extract(L,R) + ComputeStereoMatches (one CUDA graph)", "gpu_ms": ms,
             "algorithmic_bytes": int(b), "hbm_frac": frac(b, ms), "stereo_matches": int(s_.get("stereo_refined", 0)),
             "timing": "CUDA events around ft_frame_enqueue_device, images resident in HBM, L2 flushed"}
        if oracle:
            exL, exR = oracle.Extractor(), oracle.Extractor()
            t = [oracle.time_stereo_frame(exL, exR, L0, R0, float(mbf), float(mb), two_threads=True)[0] for _ in range(4)]
            r["cpu_ms"] = float(np.mean(t[1:])); r["cpu"] = "oracle port, 2 extraction threads (Frame.cc:127-130)"
        c.close()
        return r
    leg("stereo_pinhole_euroc_752x480", cfg_pinhole)

    # ---- BASELINE config 2: TUM-VI-shaped fisheye pair, extract + ComputeStereoFishEyeMatches ----
    def cfg_fisheye(nf):
        T = synth.TUMVI
        Lf, Rf = synth.fisheye_pair(seed=3)
        Rlr, tlr, Rrl, trl = synth.tumvi_extrinsics()
        c = ft.Context(512, 512, nfeatures=nf, camera_type=1, cam1=T["cam1"], cam2=T["cam2"], lap_left=T["lap"], lap_right=T["lap"],
                       bf=T["bf"], Tlr=np.hstack([Rlr, tlr[:, None]]), device_id=local)
        st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        dl, dr = torch.from_numpy(Lf).to(dev), torch.from_numpy(Rf).to(dev)
        ms = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dl.data_ptr(), 512, dr.data_ptr(), 512), 30, flush_l2)
        c.set_stage_timing(True)
        acc = []
        for _ in range(10):
            flush_l2(st_); c.frame_enqueue_device(dl.data_ptr(), 512, dr.data_ptr(), 512); acc.append(c.stage_times())
        c.set_stage_timing(False)
        s_ = c.stats()
        nl, nr = s_["kp_left"], s_["kp_right"]
        b = _extract_bytes(c, s_) + 32 * (nl + nr) + 28 * nl
        r = {"workload": "512x512 KB8 pair, %d features: extract(L,R) + ComputeStereoFishEyeMatches" % nf, "gpu_ms": ms,
             "algorithmic_bytes": int(b), "hbm_frac": frac(b, ms), "keypoints": [int(nl), int(nr)], "hamming_pairs": int(nl) * int(nr),
             "fisheye_match_kernel_ms": float(np.mean([a.get("stereo_match", 0.0) for a in acc])),
             "timing": "CUDA events around ft_frame_enqueue_device, images resident in HBM, L2 flushed"}
        if oracle:
            exL, exR = oracle.Extractor(nf), oracle.Extractor(nf)
            t0 = time.perf_counter()
            mL, kL, dL_ = exL.extract(Lf, lap=T["lap"]); mR, kR, dR_ = exR.extract(Rf, lap=T["lap"])
            t1 = time.perf_counter()
            oracle.fisheye(T["cam1"], T["cam2"], Rlr, tlr, exL.sigma2, kL, dL_, mL, kR, dR_, mR)
            t2 = time.perf_counter()
            r["cpu_ms"] = (t2 - t0) * 1e3; r["cpu_extract_ms"] = (t1 - t0) * 1e3; r["cpu_match_ms"] = (t2 - t1) * 1e3
            r["cpu"] = "oracle port, 1 thread (both extractions back to back)"
        c.close()
        return r
    leg("stereo_fisheye_tumvi_512_1000", lambda: cfg_fisheye(1000))
    leg("stereo_fisheye_tumvi_512_2000", lambda: cfg_fisheye(2000))

    # ---- BASELINE config 3: SearchByProjection alone, 5k-20k local MapPoints against a 1200-keypoint frame ----
    def cfg_sbp():
        c = ft.Context(E["width"], E["height"], nfeatures=E["nfeatures"], nlevels=E["nlevels"], scale_factor=E["scale"],
                       cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), max_map_points=25000, device_id=local)
        st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        c.frame_construct(L0, R0)
        g = c.download(0, stereo=True)
        keys = ft.keypoints_as_array(g["kps"])
        scale = c.scale_tables()["scale"]
        c.set_pose(np.eye(3), np.zeros(3)); c.upload_holders(None, None)
        F = None
        if oracle:
            exL = oracle.Extractor()
            F = oracle.Frame(keys, g["desc"], exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                             mbf=float(mbf), u_right=g["u_right"])
        res = {}
        for M in (5000, 10000, 20000):
            mp = fast_mappoints(keys, g["desc"], scale, M, 4242 + M)
            dm = {k: torch.from_numpy(v).to(dev) for k, v in mp.items()}
            c.bind_map_points_device(M, dm["pos"].data_ptr(), dm["normal"].data_ptr(), dm["minmax"].data_ptr(), dm["desc"].data_ptr(),
                                     dm["flags"].data_ptr())
            for th in (1.0, 2.0, 6.0, 10.0):
                ms = _dev_ms(torch, st_, lambda: c.search_resident(th), 20, flush_l2)
                s_ = c.stats()
                b = 68 * M + 52 * s_["sbp_candidates"] + 4 * s_["sbp_candidates"] + 8 * M
                e = {"gpu_ms": ms, "algorithmic_bytes": int(b), "hbm_frac": frac(b, ms), "candidates": int(s_["sbp_candidates"]),
                     "resolve_rounds": int(s_.get("sbp_rounds", 0))}
                if F is not None:
                    t0 = time.perf_counter()
                    nm = F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], th,
                                               np.full(len(keys), -1, np.int32), np.zeros(len(keys), np.uint8))[0]
                    e["cpu_ms"] = (time.perf_counter() - t0) * 1e3
                    e["matches"] = int(nm)
                res["M%d_th%g" % (M, th)] = e
        c.close()
        return {"workload": "SearchLocalPoints (isInFrustum + SearchByProjection) against one 1200-keypoint EuRoC frame, map points resident "
                            "in HBM", "timing": "CUDA events around ft_search_resident (gather + resolve), L2 flushed", "cpu": "oracle port, 1 thread",
                "cases": res}
    leg("search_by_projection_5k_20k", cfg_sbp)

    nxt = {}

    def nleg(name, fn):
        log("next row " + name)
        try:
            nxt[name] = fn()
        except Exception as e:   # noqa: BLE001
            nxt[name] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---- next row 1: frame-to-last-frame SearchByProjection (TrackWithMotionModel) ----
    def row_last_frame():
        c = ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), device_id=local)
        c.frame_construct(L0, R0)
        g = c.download(0, stereo=True)
        keys = ft.keypoints_as_array(g["kps"])
        a = 0.02
        Rcw = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
        tcw = np.array([0.03, -0.02, -0.05], np.float32)
        lf = synth.last_frame_points(keys, g["desc"], 1000, seed=100, Rcw=Rcw, tcw=tcw, fx=E["fx"], fy=E["fy"], cx=E["cx"], cy=E["cy"])
        c.set_pose(Rcw, tcw)
        N = len(keys)
        h0 = np.full(N, -1, np.int32); ho0 = np.zeros(N, np.uint8)
        call = lambda: c.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], np.eye(3, dtype=np.float32),
                                           np.zeros(3, np.float32), 7.0, h0, ho0, b_mono=False, check_ori=True)
        for _ in range(3):
            nm = call()[0]
        t0 = time.perf_counter()
        for _ in range(50):
            call()
        ms = (time.perf_counter() - t0) * 1e3 / 50
        r = {"workload": "SearchByProjection(CurrentFrame, LastFrame, th=7) + rotation histogram, 1000 last-frame map points",
             "gpu_ms_e2e": ms, "matches": int(nm), "timing": "host wall clock per ft_search_last_frame call (H2D of the points, gather + resolve, D2H)"}
        if oracle:
            exL = oracle.Extractor()
            F = oracle.Frame(keys, g["desc"], exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                             mbf=float(mbf), u_right=g["u_right"], Rcw=Rcw, tcw=tcw)
            t0 = time.perf_counter()
            for _ in range(5):
                F.search_last_frame(lf["pos"], lf["desc"], lf["octave"], lf["angle"], lf["flags"], 7.0, 0, h0, ho0, True)
            r["cpu_ms"] = (time.perf_counter() - t0) * 1e3 / 5
        c.close()
        return r
    nleg("last_frame_search", row_last_frame)

    # ---- next row 2: rectification remap / input resize / keypoint undistortion in front of / behind the extractor ----
    def row_front():
        mk = lambda: ft.Context(E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), device_id=local)
        W, H = E["width"], E["height"]
        r = {}
        c = mk(); st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        dl, dr = torch.from_numpy(L0).to(dev), torch.from_numpy(R0).to(dev)
        base = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dl.data_ptr(), W, dr.data_ptr(), W), 20, flush_l2)
        r["plain_frame_gpu_ms"] = base
        # cv::remap with a mild radial map (System.cc:273-281)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        rr2 = ((xx - E["cx"]) / E["fx"]) ** 2 + ((yy - E["cy"]) / E["fy"]) ** 2
        mx = (xx + (xx - E["cx"]) * 0.05 * rr2).astype(np.float32); my = (yy + (yy - E["cy"]) * 0.05 * rr2).astype(np.float32)
        c.set_rectification(W, H, mx, my, mx, my)
        ms = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dl.data_ptr(), W, dr.data_ptr(), W), 20, flush_l2)
        r["rectify_remap"] = {"gpu_ms_frame": ms, "gpu_ms_added": ms - base, "algorithmic_bytes": int(2 * (W * H * 8 + 2 * W * H)),
                              "note": "k_remap writes level 0 from the raw pair (fixed-point map table 8 B / pixel)"}
        try:
            import cv2
            cv2.setNumThreads(1)
            t0 = time.perf_counter()
            for _ in range(10):
                cv2.remap(L0, mx, my, cv2.INTER_LINEAR); cv2.remap(R0, mx, my, cv2.INTER_LINEAR)
            r["rectify_remap"]["cpu_ms"] = (time.perf_counter() - t0) * 1e3 / 10
            r["rectify_remap"]["cpu"] = "cv2.remap x2, 1 thread"
        except Exception:   # noqa: BLE001
            pass
        c.close()
        # cv::resize of a 2x larger raw pair (System.cc:282-285)
        c = mk(); st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        raw = np.ascontiguousarray(np.kron(L0, np.ones((2, 2), np.uint8)))
        dlr = torch.from_numpy(raw).to(dev)
        c.set_input_resize(2 * W, 2 * H)
        ms = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dlr.data_ptr(), 2 * W, dlr.data_ptr(), 2 * W), 20, flush_l2)
        r["input_resize"] = {"gpu_ms_frame": ms, "gpu_ms_added": ms - base, "algorithmic_bytes": int(2 * (4 * W * H + W * H)),
                             "note": "1504x960 raw pair resized into level 0 by k_resize_input"}
        try:
            import cv2
            t0 = time.perf_counter()
            for _ in range(10):
                cv2.resize(raw, (W, H), interpolation=cv2.INTER_LINEAR); cv2.resize(raw, (W, H), interpolation=cv2.INTER_LINEAR)
            r["input_resize"]["cpu_ms"] = (time.perf_counter() - t0) * 1e3 / 10
        except Exception:   # noqa: BLE001
            pass
        c.close()
        # Frame::UndistortKeyPoints (Frame.cc:771-835) inside the frame-grid kernel
        c = mk(); st_ = torch.cuda.ExternalStream(c.stream(), device=dev)
        dist = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
        c.set_distortion(dist)
        ms = _dev_ms(torch, st_, lambda: c.frame_enqueue_device(dl.data_ptr(), W, dr.data_ptr(), W), 20, flush_l2)
        r["undistort_keypoints"] = {"gpu_ms_frame": ms, "gpu_ms_added": ms - base,
                                    "note": "EuRoC distortion coefficients; the undistortion runs per left keypoint inside k_grid_build (parallel branch)"}
        if oracle:
            g = c.download(0)
            xy = np.ascontiguousarray(ft.keypoints_as_array(g["kps"])[:, :2], np.float32)
            K = np.array([[E["fx"], 0, E["cx"]], [0, E["fy"], E["cy"]], [0, 0, 1]], np.float64)
            t0 = time.perf_counter()
            for _ in range(20):
                oracle.undistort_points(xy, K, np.asarray(dist, np.float64))
            r["undistort_keypoints"]["cpu_ms"] = (time.perf_counter() - t0) * 1e3 / 20
        c.close()
        return r
    nleg("front_end_preprocessing", row_front)
    return out, nxt


def _reference_gpu_worker():
    """child process of reference_gpu_legs: the reference's GPU code calls exit() on any CUDA error (CudaUtils.cu:17-22), so it
    never runs inside the bench process itself"""
    import oracle
    if oracle.ref_gpu_lib() is None:
        print(json.dumps({"unavailable": "oracle/_ref/libft_ref_orbextractor_gpu.so did not travel with the snapshot"}))
        return
    frames = make_frames(5, 6)
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    L, R = frames[0]
    st = oracle.ref_gpu_stage_times(L, E["nlevels"], E["scale"], 20, 7, reps=20)
    ms_img, nkp = oracle.ref_gpu_extract_ms([f[0] for f in frames] + [frames[0][0]], E["nfeatures"], E["scale"], E["nlevels"], 20, 7)
    exL, exR = oracle.Extractor(), oracle.Extractor()
    _, kL, dL = exL.extract(L); _, kR, dR = exR.extract(R)
    ms_st, kept = oracle.ref_gpu_stereo_ms(L, R, kL, dL, kR, dR, float(mbf), float(mb), E["nlevels"], E["scale"], reps=10)
    print(json.dumps({
        "per_image_launcher_ms": {k: v for k, v in st.items() if k != "corners"}, "corners_all_levels": st["corners"],
        "extract_operator_ms_per_image": ms_img, "extract_keypoints": nkp,
        "stereo_match_ms_per_call": ms_st, "stereo_matches": kept,
        "note": "one 752x480 image per launcher chain (the reference runs one ORBextractor per eye); its resize kernel computes every "
                "level straight from level 0 in one launch (not what its CPU ComputePyramid does: SURVEY.md 2.3), this repo's pyramid "
                "is the bit-exact chain; operator() = H2D + kernels + D2H of every corner + DistributeOctTreeGPU on the host, wall clock; "
                "stereo = Frame::ComputeStereoMatchesGPU (row table, StereoMatchKernel::launch with its per-call cudaMemcpy's, host sort "
                "+ median filter), wall clock; SearchLocalPointsKernel.cu / PoseEstimationKernel.cu need Eigen / Sophus headers: not "
                "buildable here"}))


def reference_gpu_legs(frames=None, mbf=None, mb=None):
    """FastTrack's OWN CUDA kernels on this box's GPU (oracle/_ref/libft_ref_orbextractor_gpu.so = the reference's
    src/{resize,gaussian_blur,fast,orientation,descriptor}.cu + src/Kernels/StereoMatchKernel.cu compiled from where they lie):
    per-launcher CUDA-event times, ORBextractor::operator() in GPU run mode end to end, and the GPU stereo matching call.
    The performance bar of SURVEY.md 2.1, never a correctness oracle. Runs in a child process (see _reference_gpu_worker)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-gpu-worker"], capture_output=True, text=True, timeout=240)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": ("exit code %d: " % r.returncode) + (r.stderr.strip().splitlines() or ["no output"])[-1][:240]}
        return json.loads(lines[-1])
    except Exception as e:   # noqa: BLE001 -- never let the extra leg break the arm
        return {"unavailable": str(e)[:300]}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, restated (oracle port), on the box's host cores with
    the reference's threading: two threads for L/R extraction (Frame.cc:127-130), everything else on one. Under torchrun
    rank 0 alone runs and prints; it drives `world` independent CPU sequences concurrently (one Python thread each; the
    oracle calls release the GIL), so the line is like-for-like with the N-GPU arm: N sequences, whole-job frames/s."""
    if rank != 0:
        return
    import oracle
    n_seq = max(1, world)
    frames = make_frames(5, min(N_FRAMES, 6))
    mbf = np.float32(E["fx"] * E["baseline"]); mb = np.float32(mbf / np.float32(E["fx"]))
    exs = [(oracle.Extractor(), oracle.Extractor()) for _ in range(n_seq)]
    exL, exR = exs[0]
    maps = []
    for (L, R) in frames:
        _, kL, dL = exL.extract(L)
        maps.append(fast_mappoints(kL, dL, exL.scale, M_POINTS, 99))
    def step(i, seq=0):
        k = i % len(frames)
        L, R = frames[k]
        a, b = exs[seq]
        # Frame ctor: two extractor threads, then ComputeStereoMatches (timed inside the oracle in C++)
        ms, nl, nr, ns = oracle.time_stereo_frame(a, b, L, R, float(mbf), float(mb), two_threads=True)
        # Tracking::SearchLocalPoints on that frame (frame model incl. AssignFeaturesToGrid is rebuilt per step)
        _, kL, dL = pre[k]
        t0 = time.perf_counter()
        F = oracle.Frame(kL, dL, exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                         mbf=float(mbf), u_right=pre_ur[k])
        mp = maps[k]
        F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], TH,
                              np.full(len(kL), -1, np.int32), np.zeros(len(kL), np.uint8))
        return ms + (time.perf_counter() - t0) * 1e3
    pre, pre_ur = [], []
    for (L, R) in frames:
        mL, kL, dL = exL.extract(L); mR, kR, dR = exR.extract(R)
        pre.append((mL, kL, dL))
        pre_ur.append(oracle.stereo(exL, exR, kL, dL, kR, dR, float(mbf), float(mb))["uRight"])
    for i in range(args.warmup):
        step(i)
    if n_seq == 1:
        t = [step(i) for i in range(args.steps)]
        ms = float(np.mean(t))
        val = 1000.0 / ms
        cores = 2
    else:
        # N independent sequences at once: whole-job rate = N * steps / wall time of the slowest
        def run_seq(q):
            for i in range(args.steps):
                step(i, q)
        th_ = [threading.Thread(target=run_seq, args=(q,)) for q in range(n_seq)]
        t0 = time.perf_counter()
        [x.start() for x in th_]; [x.join() for x in th_]
        wall = time.perf_counter() - t0
        ms = wall * 1e3 / args.steps
        val = n_seq * args.steps / wall
        cores = min(2 * n_seq, os.cpu_count() or 1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": shared_config(n_seq),
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d frames of the bench workload on each of %d concurrent sequences; L/R extraction on 2 threads "
                                       "per sequence as Frame.cc:127-130, stereo + SearchLocalPoints on 1; host has %d cores"
                                       % (args.steps, n_seq, os.cpu_count())},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # Beside it, when oracle/_ref travelled with the snapshot: the reference's OWN code (src/ORBextractor.cc compiled as a whole,
    # ComputeStereoMatches / isInFrustum / SearchByProjection compiled from their text) on the same frames. It runs on the
    # OpenCV stand-in (scalar primitives, extra copies), so it is slower than the port above; the port stays the baseline.
    try:
        if oracle.ref_frame_lib() is not None:
            rL, rR = oracle.RefExtractor(), oracle.RefExtractor()
            n_ref = max(3, min(args.steps, 8))
            t_ref = []
            for i in range(n_ref + 1):
                L, R = frames[i % len(frames)]
                out = {}
                t0 = time.perf_counter()
                th = [threading.Thread(target=lambda key, ex, im: out.__setitem__(key, ex.extract(im)), args=a)
                      for a in (("L", rL, L), ("R", rR, R))]
                [x.start() for x in th]; [x.join() for x in th]
                (_, kL, dL), (_, kR, dR) = out["L"], out["R"]
                st = oracle.ref_stereo(rL, rR, kL, dL, kR, dR, float(mbf), float(mb))
                F = oracle.RefFrame(kL, dL, exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                                    mbf=float(mbf), u_right=st["uRight"])
                mp = maps[i % len(frames)]
                F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], TH,
                                      np.full(len(kL), -1, np.int32), np.zeros(len(kL), np.uint8))
                if i:
                    t_ref.append((time.perf_counter() - t0) * 1e3)
            line["reference_compiled"] = {"ms_per_step": float(np.mean(t_ref)), "value": 1000.0 / float(np.mean(t_ref)), "frames": n_ref,
                                          "note": "oracle/_ref: the reference's own ORBextractor.cc + Frame / ORBmatcher function text on the "
                                                  "OpenCV stand-in (includes the pyramid hand-over through Python); informational"}
    except Exception as e:   # never let the extra leg break the arm
        line["reference_compiled"] = {"unavailable": str(e)[:200]}
    # What the real reference's OpenCV (SIMD / IPP) spends on the image primitives alone -- a lower bound for its extractor that
    # the scalar port above cannot show: pyramid resize chain, the 8 Gaussian blurs and cv::FAST at iniThFAST on whole levels,
    # for both eyes on one thread (no octree, orientation, descriptors, stereo or search). Informational.
    try:
        import cv2
        cv2.setNumThreads(1)
        sizes = [(exL.level_image(l).shape[1], exL.level_image(l).shape[0]) for l in range(E["nlevels"])]
        fast = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
        t_cv = []
        for i in range(6):
            L, R = frames[i % len(frames)]
            t0 = time.perf_counter()
            for im in (L, R):
                lv = im
                for l in range(E["nlevels"]):
                    if l:
                        lv = cv2.resize(lv, sizes[l], interpolation=cv2.INTER_LINEAR)
                    cv2.GaussianBlur(lv, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
                    fast.detect(lv, None)
            if i:
                t_cv.append((time.perf_counter() - t0) * 1e3)
        line["opencv_primitives"] = {"ms_per_step": float(np.mean(t_cv)), "cv2": cv2.__version__, "threads": 1,
                                     "note": "cv2 resize chain + 8 GaussianBlur + FAST(20) on whole levels, both eyes; lower bound of the "
                                             "reference's extractor with a SIMD OpenCV; informational"}
    except Exception as e:
        line["opencv_primitives"] = {"unavailable": str(e)[:200]}
    line["reference_gpu"] = reference_gpu_legs(frames, mbf, mb)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=0.3, help="the K-step timed region is repeated until this much device time is measured")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration / next-row / reference-GPU legs")
    ap.add_argument("--profile-steps", type=int, default=40, help="steps of the per-kernel CUDA-event pass")
    ap.add_argument("--e2e-depth", type=int, default=4, help="frames in flight in the end-to-end legs (<= pipeline depth); the headline loop collects frame t+1 while frame t is searched, so three is one too few: frame t+1 would still be extracting")
    ap.add_argument("--pipeline-depth", type=int, default=4, help="frames in flight in the throughput leg (contexts/streams)")
    ap.add_argument("--ref-gpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--watchdog", type=float, default=900.0, help="dump every thread's Python stack to stderr and exit if the run takes longer (s)")
    args = ap.parse_args()
    import faulthandler
    faulthandler.enable()
    if args.watchdog > 0:
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    if args.ref_gpu_worker:
        _reference_gpu_worker()
        return
    rank, local, world = replicas.env_rank()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import fasttrack_b200 as ft
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    pinned_cores = None
    if world > 1:
        # one rank per GPU on one node: give every rank its own slice of the host cores (the end-to-end legs are bound by the
        # tracker thread's memcpy + submit path; eight unpinned ranks migrate and collide on the same cores)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[(local % world) * per:(local % world) * per + per] or cores
            os.sched_setaffinity(0, mine)
            pinned_cores = mine
        except (AttributeError, OSError):
            pass
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # barrier + MAX of the per-rank times only: gloo on the host, no NCCL (the data path has no collective)
        dist.init_process_group("gloo", rank=rank, world_size=world)

    mbf = np.float32(E["fx"] * E["baseline"])
    def make_ctx():
        return ft.Context(E["width"], E["height"], nfeatures=E["nfeatures"], nlevels=E["nlevels"], scale_factor=E["scale"],
                          cam1=[E["fx"], E["fy"], E["cx"], E["cy"]], bf=float(mbf), max_map_points=25000, device_id=local)
    # two contexts = a 2-deep software pipeline over ONE sequence: frame t+1 is extracted while frame t is searched
    D = max(1, args.pipeline_depth)
    DE = max(2, min(args.e2e_depth, max(D, 2)))   # frames in flight in the end-to-end legs
    ctxs = [make_ctx() for _ in range(max(D, 2))]
    ctx = ctxs[0]
    streams = [torch.cuda.ExternalStream(c.stream(), device=torch.device("cuda", local)) for c in ctxs]
    stream = streams[0]
    scale = ctx.scale_tables()["scale"]

    log("contexts created")
    # ---- inputs: one independent synthetic sequence per GPU (seed = 5 + rank), prepared outside the timed region ----
    frames = make_frames(replicas.sequence_seed(rank), N_FRAMES)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hL = [pin(L) for L, R in frames]; hR = [pin(R) for L, R in frames]
    dL = [t.cuda(non_blocking=False) for t in hL]; dR = [t.cuda(non_blocking=False) for t in hR]
    maps, dmaps = [], []
    for i in range(N_FRAMES):
        ctx.extract_stereo(frames[i][0], frames[i][1])
        g = ctx.download(0)
        mp = fast_mappoints(ft.keypoints_as_array(g["kps"]), g["desc"], scale, M_POINTS, 1000 * rank + i)
        maps.append(mp)
        dmaps.append({k: torch.from_numpy(v).cuda() for k, v in mp.items()})   # snapshot resident in HBM
    cap = ctx.cap
    cap_dev = int(ctx.L.ft_max_keypoints(ctx.h))
    out_kps = [torch.empty(cap * 24, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_desc = [torch.empty(cap * 32, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_ur = torch.empty(cap, dtype=torch.float32).pin_memory(); out_dp = torch.empty(cap, dtype=torch.float32).pin_memory()
    holder = torch.empty(cap, dtype=torch.int32).pin_memory(); hobs = torch.empty(cap, dtype=torch.uint8).pin_memory()
    best = torch.empty(M_POINTS * 2, dtype=torch.int32).pin_memory()
    counts4 = torch.zeros(4, dtype=torch.int32).pin_memory()
    for c_ in ctxs:
        c_.set_pose(np.eye(3), np.zeros(3))
        c_.upload_holders(None, None)
    import ctypes as C
    L_ = ctx.L
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def flush_l2(on=None):
        with torch.cuda.stream(on or stream):
            flush_buf.add_(1)

    # ---- step variants ----
    def bind_map(c_, k):
        m = dmaps[k]
        c_.bind_map_points_device(M_POINTS, m["pos"].data_ptr(), m["normal"].data_ptr(), m["minmax"].data_ptr(),
                                  m["desc"].data_ptr(), m["flags"].data_ptr())

    def step_resident(i, c_=None):
        """inputs already in HBM: images (device), map-point snapshot (device); results stay on the device"""
        c_ = c_ or ctx
        k = i % N_FRAMES
        c_.frame_enqueue_device(dL[k].data_ptr(), E["width"], dR[k].data_ptr(), E["width"])
        bind_map(c_, k)
        c_.search_resident(TH)

    def step_e2e(i):
        """the user-facing call sequence with HOST buffers: Frame construction (upload, extract, stereo, download of
        every host vector) in one C-ABI call, then SearchLocalPoints (H2D of the map-point arrays, D2H of the result)"""
        k = i % N_FRAMES
        ctx._ck(L_.ft_frame_construct(ctx.h, hL[k].data_ptr(), E["width"], hR[k].data_ptr(), E["width"],
                                      out_kps[0].data_ptr(), out_desc[0].data_ptr(), out_kps[1].data_ptr(),
                                      out_desc[1].data_ptr(), counts4.data_ptr(), out_ur.data_ptr(), out_dp.data_ptr(),
                                      None, None, None))
        nl, nr = int(counts4[0]), int(counts4[2])
        # marshal the local map into the context's pinned staging (what the reference's CudaMapPoint loop does), then
        # one H2D + kernels + one D2H
        m = maps[k]
        for key in ("pos", "normal", "minmax", "desc", "flags"):
            np.copyto(stg[key], m[key])
        stg["holder"].fill(-1); stg["holder_obs"].fill(0)
        nm, h_out, ho_out, best_out = ctx.search_staged(M_POINTS, nl, TH)
        return nl, nr

    stgs = [c_.map_point_staging(M_POINTS, cap_dev) for c_ in ctxs]   # views over each context's pinned staging
    stg = stgs[0]

    def e2e_submit(c_, k):
        c_._ck(L_.ft_frame_submit(c_.h, hL[k].data_ptr(), E["width"], hR[k].data_ptr(), E["width"]))

    def e2e_collect_and_search(c_, sg, k):
        c_._ck(L_.ft_frame_collect(c_.h, out_kps[0].data_ptr(), out_desc[0].data_ptr(), out_kps[1].data_ptr(),
                                   out_desc[1].data_ptr(), counts4.data_ptr(), out_ur.data_ptr(), out_dp.data_ptr(),
                                   None, None, None))
        nl = int(counts4[0])
        m = maps[k]
        for key in ("pos", "normal", "minmax", "desc", "flags"):
            np.copyto(sg[key], m[key])
        sg["holder"].fill(-1); sg["holder_obs"].fill(0)
        return c_.search_staged(M_POINTS, nl, TH)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(max(args.warmup, 3)):
        for c_ in ctxs:
            step_resident(i, c_); c_.synchronize()
    for i in range(3):
        step_e2e(i)

    log("warm-up done")
    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(local), "uuid", None))
    # ---- timed region 1a: per-frame LATENCY, one frame at a time, CUDA events per step, L2 flushed between steps ----
    n_lat = min(args.steps, 100)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_lat)]
    barrier()
    for i in range(n_lat):
        flush_l2()
        ev[i][0].record(stream)
        step_resident(i)
        ev[i][1].record(stream)
    barrier()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    # the same loop without the flush: the distinct frames of the sequence are cycled (inputs larger than L2), kernel code and
    # the context's tables stay warm in L2, as they do for a tracker that runs this path back to back
    ev_w = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_lat)]
    for i in range(n_lat):
        ev_w[i][0].record(stream)
        step_resident(i)
        ev_w[i][1].record(stream)
    barrier()
    warm_ms = np.array([a.elapsed_time(b) for a, b in ev_w])

    # ---- timed region 1b: THROUGHPUT of the sequence, inputs resident in HBM and larger than L2 (no flush needed).
    # 2-deep pipeline over one sequence: frames alternate between two contexts (streams); frame t+1 is extracted
    # while frame t is searched, and the search of frame t+1 still waits for the search of frame t (tracking is
    # frame-sequential: pose(t+1) follows from the matches of frame t).
    # `--steps K` times EXACTLY K steps per region (barrier + synchronize on both sides). A region of a few milliseconds is
    # mostly pipeline fill and drain, so the region is repeated until about --min-seconds of device time have been
    # measured; ms_per_step = total device time / (regions x K).
    def throughput_region():
        done = [torch.cuda.Event(enable_timing=False) for _ in range(args.steps)]
        ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        ev_start.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_event(ev_start)
        for i in range(args.steps):
            c_, s_ = ctxs[i % D], streams[i % D]
            k = i % N_FRAMES
            c_.frame_enqueue_device(dL[k].data_ptr(), E["width"], dR[k].data_ptr(), E["width"])
            if i > 0:
                s_.wait_event(done[i - 1])
            bind_map(c_, k)
            c_.search_resident(TH)
            done[i].record(s_)
        for j in range(1, min(D, args.steps) + 1):
            streams[0].wait_event(done[args.steps - j])
        submit_s = time.perf_counter() - w0
        ev_end.record(streams[0])
        barrier()
        return float(ev_start.elapsed_time(ev_end)), submit_s, time.perf_counter() - w0

    log("latency leg done")
    dev_ms_total, host_submit_s, wall_total = throughput_region()
    n_regions = 1
    est = dev_ms_total / 1e3
    target_regions = int(min(64, max(1, np.ceil(args.min_seconds / max(est, 1e-6)))))
    if world > 1:   # every rank must run the same number of regions (they contain barriers)
        (tr,) = replicas.reduce_max(dist, world, [float(target_regions)])
        target_regions = int(tr)
    while n_regions < target_regions:
        a_, b_, c2_ = throughput_region()
        dev_ms_total += a_; host_submit_s += b_; wall_total += c2_
        n_regions += 1
    steps_timed = n_regions * args.steps
    log("throughput leg done (%d regions)" % n_regions)
    ext_l, st_l, se_l = ctx.launch_counts()
    launches_per_step = ext_l + st_l + se_l

    # ---- timed region 2: end to end through the C ABI with host buffers, one frame at a time (ft_frame_construct +
    # marshal + ft_search_staged; every step: H2D of both images and the map snapshot, D2H of every host vector) ----
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        nl, nr = step_e2e(i)
    barrier()
    e2e_serial_s = time.perf_counter() - t0
    # ---- timed region 2b: the same calls split as ft_frame_submit / ft_frame_collect over two contexts: the upload,
    # extraction and stereo of frame t+1 run while the host marshals and searches frame t. Same bytes per step. ----
    barrier()
    t0 = time.perf_counter()
    for j in range(min(DE - 1, args.steps)):
        e2e_submit(ctxs[j % DE], j % N_FRAMES)
    for i in range(args.steps):
        j = i + DE - 1
        if j < args.steps:
            e2e_submit(ctxs[j % DE], j % N_FRAMES)
        e2e_collect_and_search(ctxs[i % DE], stgs[i % DE], i % N_FRAMES)
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- timed region 2c: as 2b, with the local map named as rows of the persistent device-side map store
    # (SURVEY.md 8f row 3): per step the host upserts STORE_UPSERTS changed rows and sends 8 bytes per map point
    # (row, flags) instead of the 68-byte snapshot. The store is loaded before the timed region (it persists).
    ctxs[0].map_store_create(N_FRAMES * M_POINTS)
    for c_ in ctxs[1:]:
        c_.map_store_attach(ctxs[0])
    for k in range(N_FRAMES):
        m = maps[k]
        ctxs[0].map_store_update(np.arange(k * M_POINTS, (k + 1) * M_POINTS, dtype=np.int32), m["pos"], m["normal"], m["minmax"], m["desc"])
    rows = [np.arange(k * M_POINTS, (k + 1) * M_POINTS, dtype=np.int32) for k in range(N_FRAMES)]
    h_in = np.full(cap_dev, -1, np.int32); ho_in = np.zeros(cap_dev, np.uint8)
    h_io = np.empty(cap_dev, np.int32); ho_io = np.empty(cap_dev, np.uint8)
    best_io = np.empty((M_POINTS, 2), np.int32); nm_io = C.c_int()

    def e2e_collect_and_search_store(c_, k):
        c_._ck(L_.ft_frame_collect(c_.h, out_kps[0].data_ptr(), out_desc[0].data_ptr(), out_kps[1].data_ptr(),
                                   out_desc[1].data_ptr(), counts4.data_ptr(), out_ur.data_ptr(), out_dp.data_ptr(),
                                   None, None, None))
        m = maps[k]; r = rows[k]; u = STORE_UPSERTS
        c_._ck(L_.ft_map_store_update(c_.h, u, r.ctypes.data, m["pos"].ctypes.data, m["normal"].ctypes.data,
                                      m["minmax"].ctypes.data, m["desc"].ctypes.data))
        np.copyto(h_io, h_in); np.copyto(ho_io, ho_in)
        c_._ck(L_.ft_search_store(c_.h, M_POINTS, r.ctypes.data, m["flags"].ctypes.data, TH, 0, 50.0, 0.8, h_io.ctypes.data,
                                  ho_io.ctypes.data, best_io.ctypes.data, C.byref(nm_io)))
        return nm_io.value

    # the store path returns what the snapshot path returns (checked on a few frames outside the timed region)
    for k in range(3):
        e2e_submit(ctxs[k & 1], k)
        nm_snap = e2e_collect_and_search(ctxs[k & 1], stgs[k & 1], k)[0]
        e2e_submit(ctxs[k & 1], k)
        nm_store = e2e_collect_and_search_store(ctxs[k & 1], k)
        if nm_snap != nm_store:
            raise SystemExit("bench.py: map-store search differs from the snapshot search (%d vs %d)" % (nm_store, nm_snap))
    barrier()
    t0 = time.perf_counter()
    for j in range(min(DE - 1, args.steps)):
        e2e_submit(ctxs[j % DE], j % N_FRAMES)
    for i in range(args.steps):
        j = i + DE - 1
        if j < args.steps:
            e2e_submit(ctxs[j % DE], j % N_FRAMES)
        e2e_collect_and_search_store(ctxs[i % DE], i % N_FRAMES)
    barrier()
    e2e_store_s = time.perf_counter() - t0
    # ---- timed region 2d: the same three end-to-end loops driven from C++ (fasttrack_b200/host/ft_sequence_driver.cpp,
    # public C ABI only): what a C++ tracking thread pays, without the Python/ctypes overhead of 2/2b/2c ----
    log("python e2e legs done")
    drv = C.CDLL(os.path.join(os.path.dirname(ft.library_path()), "libft_sequence_driver.so"))

    class Seq(C.Structure):
        _fields_ = [("n_frames", C.c_int), ("width", C.c_int), ("height", C.c_int), ("M", C.c_int)] + \
                   [(k_, C.POINTER(C.c_void_p)) for k_ in ("imgL", "imgR", "pos", "normal", "minmax", "desc", "flags", "rows")]
    arr = lambda ptrs: (C.c_void_p * len(ptrs))(*ptrs)
    keep = dict(imgL=arr([t.data_ptr() for t in hL]), imgR=arr([t.data_ptr() for t in hR]),
                rows=arr([r.ctypes.data for r in rows]))
    for key in ("pos", "normal", "minmax", "desc", "flags"):
        keep[key] = arr([m[key].ctypes.data for m in maps])
    seq = Seq(N_FRAMES, E["width"], E["height"], M_POINTS, *[C.cast(keep[k_], C.POINTER(C.c_void_p)) for k_ in
                                                           ("imgL", "imgR", "pos", "normal", "minmax", "desc", "flags", "rows")])
    drv.ftd_run_serial.restype = C.c_double
    drv.ftd_run_serial.argtypes = [C.c_void_p, C.POINTER(Seq), C.c_int, C.c_float, C.POINTER(C.c_longlong)]
    drv.ftd_run_pipelined.restype = C.c_double
    drv.ftd_run_pipelined.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Seq), C.c_int, C.c_float, C.c_int, C.c_int,
                                      C.POINTER(C.c_longlong)]
    hctx = (C.c_void_p * DE)(*[c_.h for c_ in ctxs[:DE]])
    nmatch = [C.c_longlong(), C.c_longlong(), C.c_longlong()]

    def native_once(fn):
        barrier()
        t = fn()
        if t < 0:
            raise SystemExit("bench.py: native sequence driver failed: %s" % L_.ft_last_error().decode())
        (t,) = replicas.reduce_max(dist, world, [t])
        return t

    def native(fn):
        """seconds per K-step call (max over ranks), averaged over enough calls to cover --min-seconds (each call is exactly
        K steps, synchronised on both sides, so a 20-step run is not one 3 ms interval of pipeline fill and drain)"""
        t = native_once(fn)
        reps = int(min(64, max(1, np.ceil(args.min_seconds / max(t, 1e-6)))))
        if world > 1:
            (r_,) = replicas.reduce_max(dist, world, [float(reps)])
            reps = int(r_)
        tot = t
        for _ in range(reps - 1):
            tot += native_once(fn)
        return tot / reps
    native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq), 20, TH, 0, 0, C.byref(nmatch[1])))      # warm-up
    n_serial_s = native(lambda: drv.ftd_run_serial(ctxs[0].h, C.byref(seq), args.steps, TH, C.byref(nmatch[0])))
    n_pipe_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq), args.steps, TH, 0, 0, C.byref(nmatch[1])))
    # store loop with the synchronous ft_search_store (round 1 / early round 2 headline), then the same loop with the search
    # split into ft_search_store_submit / ft_search_collect and the next frame's ft_frame_submit issued between the halves
    nm_sync = C.c_longlong()
    n_store_sync_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq), args.steps, TH, 1, STORE_UPSERTS, C.byref(nm_sync)))
    drv.ftd_phase_seconds.argtypes = [C.POINTER(C.c_double)]
    drv.ftd_phase_seconds.restype = None
    PHASES = ("frame_submit", "frame_collect", "store_update", "search_submit", "search_collect", "search_sync", "marshal")

    def host_phases():
        """host wall clock per phase of the last ftd_run_pipelined call, us per frame (the tracking thread's time line)"""
        buf = (C.c_double * len(PHASES))()
        drv.ftd_phase_seconds(buf)
        return {k_: buf[i_] * 1e6 / args.steps for i_, k_ in enumerate(PHASES) if buf[i_] > 0}
    phases_sync = host_phases()
    nm_split = C.c_longlong()
    n_store_split_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq), args.steps, TH, 2, STORE_UPSERTS, C.byref(nm_split)))
    # ... and with frame t+1's host vectors collected and its upserts enqueued in the shadow of the search of frame t as well
    # (needs three frames in flight; falls back to the split loop otherwise)
    n_store_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq), args.steps, TH, 3, STORE_UPSERTS, C.byref(nmatch[2])))
    phases_e2e = host_phases()
    if not (nm_sync.value == nm_split.value == nmatch[2].value):
        raise SystemExit("bench.py: the split-search loops disagree with the synchronous one on the matches found")
    # the same store loop with PAGEABLE input images (what a caller holding a plain cv::Mat hands over): upload_images stages
    # them through the context's pinned buffer (ft_context.cu), one extra host copy of both images per frame
    pg = dict(imgL=arr([f[0].ctypes.data for f in frames]), imgR=arr([f[1].ctypes.data for f in frames]))
    seq_pg = Seq(N_FRAMES, E["width"], E["height"], M_POINTS, C.cast(pg["imgL"], C.POINTER(C.c_void_p)), C.cast(pg["imgR"], C.POINTER(C.c_void_p)),
                 *[C.cast(keep[k_], C.POINTER(C.c_void_p)) for k_ in ("pos", "normal", "minmax", "desc", "flags", "rows")])
    nmatch.append(C.c_longlong())
    n_store_pg_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq_pg), args.steps, TH, 3, STORE_UPSERTS, C.byref(nmatch[3])))
    if nmatch[3].value != nmatch[2].value:
        raise SystemExit("bench.py: pageable-input loop disagrees on the matches found")
    # ... and with the same pageable buffers registered once (ft_host_register): what INTEGRATION.md recommends for a caller
    # that keeps its camera buffers
    for f_ in frames:
        for im in f_:
            ctx._ck(L_.ft_host_register(im.ctypes.data, im.nbytes))
    nmatch.append(C.c_longlong())
    n_store_reg_s = native(lambda: drv.ftd_run_pipelined(hctx, DE, C.byref(seq_pg), args.steps, TH, 3, STORE_UPSERTS, C.byref(nmatch[4])))
    for f_ in frames:
        for im in f_:
            L_.ft_host_unregister(im.ctypes.data)
    if not (nmatch[0].value == nmatch[1].value == nmatch[2].value) or nmatch[0].value <= 0:
        raise SystemExit("bench.py: the end-to-end loops disagree on the matches found: %s" % [v.value for v in nmatch])
    log("native e2e legs done")
    clocks = sampler.stop()
    h2d = 2 * E["width"] * E["height"] + M_POINTS * (12 + 12 + 8 + 32 + 4) + 2 * cap_dev * 5
    d2h = 64 + 2 * cap_dev * (24 + 32) + cap_dev * 8 + 64 + 2 * cap_dev * 5 + M_POINTS * 8

    # ---- per-kernel pass: CUDA events around every kernel (direct launches), same inputs ----
    ctx.set_stage_timing(True)
    acc = {}
    for i in range(args.profile_steps):
        flush_l2(); step_resident(i)
        for k_, v in ctx.stage_times().items():
            acc.setdefault(k_, []).append(v)
    stage_ms = {k_: float(np.mean(v)) for k_, v in acc.items()}
    # the same pass with every launch on ONE stream: each stage alone on the GPU (what a per-kernel comparison needs)
    ctx.set_stage_timing(2)
    acc2 = {}
    for i in range(args.profile_steps):
        flush_l2(); step_resident(i)
        for k_, v in ctx.stage_times().items():
            acc2.setdefault(k_, []).append(v)
    ctx.set_stage_timing(False)
    stage_iso_ms = {k_: float(np.mean(v)) for k_, v in acc2.items()}
    st = ctx.stats()

    log("per-kernel pass done")
    # ---- reductions over ranks (max time) ----
    dev_ms_total, e2e_s, wall_total, e2e_serial_s, e2e_store_s = replicas.reduce_max(dist, world, [dev_ms_total, e2e_s, wall_total, e2e_serial_s, e2e_store_s])
    ms_per_step = dev_ms_total / steps_timed
    value = replicas.aggregate_throughput(world, steps_timed, dev_ms_total / 1e3)
    wall_s = wall_total
    e2e_value = replicas.aggregate_throughput(world, args.steps, e2e_s)

    # ---- roofline of the dominant kernel (SURVEY.md 8d byte formulas with the measured counts) ----
    pk, pk_kind = peaks()
    lv = [ctx.level_dims(l) for l in range(E["nlevels"])]
    px = [w * h for w, h in lv]
    sumP = sum(px)
    cL, kLv = ctx.level_counts(0); cR, kRv = ctx.level_counts(1)
    C0, K0 = int(cL[0] + cR[0]), int(kLv[0] + kRv[0])
    C_ = st["cand_left"] + st["cand_right"]; K_ = st["kp_left"] + st["kp_right"]
    # per launch (both eyes): bytes each kernel has to move at minimum (SURVEY.md 8d formulas, measured counts)
    alg_bytes = {
        "resize": 2 * (sum(px[:-1]) + sum(px[1:])),
        "blur_l0": 2 * 2 * px[0],
        "blur": 2 * 2 * (sumP - px[0]),
        "fast_cells_l0": 2 * px[0] + 16 * C0,
        "fast_cells": 2 * (sumP - px[0]) + 16 * (C_ - C0),
        "octree_l0": 16 * C0 + 20 * K0,
        "octree": 16 * (C_ - C0) + 20 * (K_ - K0),
        "orient_desc": (749 + 4 + 512 + 32) * K_,
        "grid": 24 * st["kp_left"] + 4 * st["kp_left"] + 4 * 3073,
        "stereo_match": 44 * K_ + 44 * st["stereo_tested"] + 352 * st["stereo_refined"] + 12 * st["kp_left"],
        "gather": 68 * M_POINTS + 52 * st["sbp_candidates"],
        "resolve": 4 * st["sbp_candidates"] + 8 * M_POINTS,
    }
    # dominant kernel = the single launch with the largest CUDA-event time ("resize" is a chain of 7 launches, reported in stages_ms)
    # (times of the isolated pass: a kernel's own duration, not what the kernels running beside it in the product topology add)
    kt = stage_iso_ms if stage_iso_ms else stage_ms
    top = max((k_ for k_ in kt if k_ in alg_bytes and k_ != "resize"), key=lambda k_: kt[k_])
    ach = alg_bytes[top] / (kt[top] * 1e-3) / 1e9
    traffic = None
    traffic_src = None
    for tp in (os.path.join(ROOT, "profiles", "r2_traffic.json"), os.path.join(ROOT, "profiles", "r1_traffic.json")):
        if os.path.exists(tp):   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this workload
            traffic = json.load(open(tp))["dram_bytes_per_launch"].get(top)
            traffic_src = os.path.basename(tp)
            break
    roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_kind,
                "algorithmic_bytes_per_launch": alg_bytes[top], "launch_ms": kt[top],
                "timing": "CUDA events around the launch on its stream, every launch alone on the GPU (stages_isolated_ms), L2 flushed"}
    # every kernel against the same roofline: algorithmic bytes of the launch(es) / CUDA-event time, as a fraction of the HBM peak
    stages_roofline = {k_: {"algorithmic_bytes": int(alg_bytes[k_]), "ms": kt[k_],
                            "achieved_gbs": alg_bytes[k_] / (kt[k_] * 1e-3) / 1e9,
                            "frac": alg_bytes[k_] / (kt[k_] * 1e-3) / 1e9 / pk["hbm_gbs"]}
                       for k_ in kt if k_ in alg_bytes and kt[k_] > 0}
    frame_bytes = sum(alg_bytes.values())
    lat_ms = float(step_ms.mean())

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": shared_config(world),
            "config_detail": {"frames_cycled": N_FRAMES,
                              "pipeline": "%d frames in flight over one sequence (extract t+1.. || search t); searches stay ordered" % D,
                              "pipeline_depth": D, "e2e_depth": DE,
                              "parallelism": "independent sequence per GPU, no collective (gloo barrier + MAX only)",
                              "pinned_cores": pinned_cores},
            # headline end-to-end path = the design's answer to the per-frame CudaMapPoint marshalling of the reference:
            # host images in, every host vector of the Frame out, the local map named as rows of the persistent device-side
            # store (8 B / point + the rows the mapping side changed). The 68 B / point snapshot variant is reported beside it.
            "e2e": {"value": replicas.aggregate_throughput(world, args.steps, n_store_s), "unit": "frames/s",
                    "h2d_bytes_per_step": int(2 * E["width"] * E["height"] + 8 * M_POINTS + 72 * STORE_UPSERTS + 2 * cap_dev * 5),
                    "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": n_store_s * 1e3 / args.steps, "frames_in_flight": DE,
                    "path": "ft_search_store_submit(t) + ft_frame_submit(t+D-1) + ft_frame_collect(t+1) + ft_map_store_update(%d rows, t+1) + "
                            "ft_search_collect(t) (SURVEY 8f row 3)" % STORE_UPSERTS,
                    "driver": "C++ loop over the C ABI (fasttrack_b200/host/ft_sequence_driver.cpp, use_store = 3): the search of "
                              "frame t is enqueued first; while it runs the next camera frame is handed over and frame t+1 delivers "
                              "its host vectors and takes its upserts; host wall clock, max over ranks; pinned host images",
                    "upserts_per_step": STORE_UPSERTS,
                    "host_phases_us_per_frame": phases_e2e,
                    "map_store": {"value": replicas.aggregate_throughput(world, args.steps, n_store_s), "ms_per_step": n_store_s * 1e3 / args.steps},
                    "split_search": {"value": replicas.aggregate_throughput(world, args.steps, n_store_split_s),
                                     "ms_per_step": n_store_split_s * 1e3 / args.steps,
                                     "note": "use_store = 2: only ft_frame_submit(next frame) between ft_search_store_submit and "
                                             "ft_search_collect; frame t+1 is collected after the search of frame t"},
                    "sync_search": {"value": replicas.aggregate_throughput(world, args.steps, n_store_sync_s),
                                    "ms_per_step": n_store_sync_s * 1e3 / args.steps, "host_phases_us_per_frame": phases_sync,
                                    "note": "the same loop with ft_frame_submit(next frame) in front of the synchronous "
                                            "ft_search_store (the headline loop until the search was split)"},
                    "pageable_images": {"value": replicas.aggregate_throughput(world, args.steps, n_store_pg_s),
                                        "ms_per_step": n_store_pg_s * 1e3 / args.steps,
                                        "note": "same loop, input images in pageable host memory (a plain cv::Mat): staged through the "
                                                "context's pinned buffer inside ft_frame_submit"},
                    "registered_images": {"value": replicas.aggregate_throughput(world, args.steps, n_store_reg_s),
                                          "ms_per_step": n_store_reg_s * 1e3 / args.steps,
                                          "note": "same pageable buffers after one ft_host_register each (camera ring buffers are reused)"},
                    "snapshot": {"value": replicas.aggregate_throughput(world, args.steps, n_pipe_s),
                                 "ms_per_step": n_pipe_s * 1e3 / args.steps, "h2d_bytes_per_step": int(h2d),
                                 "note": "local map re-marshalled and uploaded every frame (68 B / point), as the reference's "
                                         "CudaMapPoint loop does"},
                    "serial_ms_per_step": n_serial_s * 1e3 / args.steps,
                    "serial_value": replicas.aggregate_throughput(world, args.steps, n_serial_s),
                    "matches_per_frame": nmatch[0].value / args.steps,
                    "python_loop": {"note": "the same calls issued from Python/ctypes (bench.py step functions)",
                                    "ms_per_step": e2e_s * 1e3 / args.steps, "serial_ms_per_step": e2e_serial_s * 1e3 / args.steps,
                                    "map_store_ms_per_step": e2e_store_s * 1e3 / args.steps}},
            "gpu_launches": int(launches_per_step * steps_timed),
            "timed_regions": n_regions,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roofline,
            "stages_ms": stage_ms,
            "stages_isolated_ms": stage_iso_ms,
            "stages_roofline": stages_roofline,
            "frame_algorithmic_bytes": int(frame_bytes),
            "frame_hbm_roofline_frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"],
            "frame_hbm_roofline_frac_latency": frame_bytes / (lat_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "counts": st,
            "wall_s_resident_loop": wall_s, "host_submit_s_resident_loop": host_submit_s,
            "latency": {"ms_per_frame_mean": float(step_ms.mean()), "p50": float(np.percentile(step_ms, 50)),
                        "p95": float(np.percentile(step_ms, 95)), "frames": int(n_lat),
                        "warm_p50": float(np.percentile(warm_ms, 50)), "warm_p95": float(np.percentile(warm_ms, 95)),
                        "note": "one frame at a time on one context, CUDA events per frame; p50 / p95: L2 flushed between frames "
                                "(kernel code and tables come from HBM as well); warm_*: no flush, the distinct frames cycled"}}

    # ---- CPU baseline beside it (rank 0, N == 1): the oracle port on a bounded sample of the same workload ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        exL, exR = oracle.Extractor(), oracle.Extractor()
        mb = np.float32(mbf / np.float32(E["fx"]))
        nsample = 30
        t_cpu = []
        for i in range(nsample + 2):
            L, R = frames[i % N_FRAMES]
            ms, _, _, _ = oracle.time_stereo_frame(exL, exR, L, R, float(mbf), float(mb), two_threads=True)
            _, kL, dLd = exL.extract(L) if i < 2 else (0, None, None)
            if i >= 2:
                t_cpu.append(ms)
        # projection search of the oracle on 6 frames
        sbp = []
        for i in range(6):
            L, R = frames[i]
            _, kL, dLd = exL.extract(L); _, kR, dRd = exR.extract(R)
            so = oracle.stereo(exL, exR, kL, dLd, kR, dRd, float(mbf), float(mb))
            F = oracle.Frame(kL, dLd, exL.scale, E["width"], E["height"], cam1=[E["fx"], E["fy"], E["cx"], E["cy"], 0, 0, 0, 0],
                             mbf=float(mbf), u_right=so["uRight"])
            mp = maps[i]
            t0 = time.perf_counter()
            F.search_local_points(mp["pos"], mp["normal"], mp["minmax"], mp["desc"], mp["flags"], TH,
                                  np.full(len(kL), -1, np.int32), np.zeros(len(kL), np.uint8))
            sbp.append((time.perf_counter() - t0) * 1e3)
        cpu_ms = float(np.mean(t_cpu) + np.mean(sbp))
        line["cpu_baseline"] = {"value": 1000.0 / cpu_ms, "unit": "frames/s", "cores": 2, "kind": "port",
                                "ms_per_frame": cpu_ms,
                                "sample": "%d frames (extract L/R on 2 threads + stereo) + 6 frames SearchLocalPoints "
                                          "of the same workload; host has %d cores" % (nsample, os.cpu_count())}
    log("cpu baseline done")
    if rank == 0 and world == 1:
        line["next_rows"] = {"bow": bench_bow(ctx, ft, torch, stream, frames, local, not args.no_cpu_baseline)}
        if not args.no_configs:
            cfgs, nxt = bench_configs(ft, torch, local, frames, not args.no_cpu_baseline, flush_l2)
            cfgs["sequence_2000_frames_per_gpu"] = {"workload": WORKLOAD, "see": "value / ms_per_step / latency / e2e of this line (the headline)"}
            line["configs"] = cfgs
            line["next_rows"].update(nxt)
            log("reference gpu legs")
            rg = reference_gpu_legs(frames, mbf, np.float32(mbf / np.float32(E["fx"])))
            line["reference_gpu"] = rg
            if "per_image_launcher_ms" in rg:
                # FastTrack's own kernels process ONE image per launcher chain (one ORBextractor per eye, two streams); this repo's
                # launches process BOTH eyes. vs_ref_gpu = reference ms for one image / this repo's ms for both eyes (>= 1: faster even
                # if the reference overlapped its two eyes perfectly); vs_ref_gpu_two_images assumes they run back to back.
                rs = rg["per_image_launcher_ms"]
                # both sides timed the same way: every launch alone on the GPU, CUDA events around it
                mine = {"resize": stage_iso_ms.get("resize", 0.0),
                        "gaussian_blur": stage_iso_ms.get("blur", 0.0) + stage_iso_ms.get("blur_l0", 0.0),
                        "fast_extract": stage_iso_ms.get("fast_cells", 0.0) + stage_iso_ms.get("fast_cells_l0", 0.0),
                        "orientation_descriptor": stage_iso_ms.get("orient_desc", 0.0)}
                ref = {"resize": rs["resize"], "gaussian_blur": rs["gaussian_blur"], "fast_extract": rs["fast_extract"],
                       "orientation_descriptor": rs["compute_orientation"] + rs["compute_descriptor"]}
                line["stages_vs_ref_gpu"] = {k_: {"ref_ms_one_image": ref[k_], "ms_both_eyes": mine[k_],
                                                  "vs_ref_gpu": ref[k_] / mine[k_] if mine[k_] > 0 else None,
                                                  "vs_ref_gpu_two_images": 2 * ref[k_] / mine[k_] if mine[k_] > 0 else None}
                                             for k_ in mine}
                ext_ms = cfgs.get("stereo_pinhole_euroc_752x480", {}).get("gpu_ms")
                if ext_ms:
                    line["stages_vs_ref_gpu"]["extract_plus_stereo_pair"] = {
                        "ref_ms": 2 * rg["extract_operator_ms_per_image"] + rg["stereo_match_ms_per_call"], "ms": ext_ms,
                        "vs_ref_gpu": (2 * rg["extract_operator_ms_per_image"] + rg["stereo_match_ms_per_call"]) / ext_ms,
                        "note": "reference: two ORBextractor::operator() calls in GPU run mode (host octree included) + "
                                "ComputeStereoMatchesGPU, wall clock; this repo: extract + stereo graph, device time"}
    if rank == 0:
        print(json.dumps(line))
    for c_ in ctxs:
        c_.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
