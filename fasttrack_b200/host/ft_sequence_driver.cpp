// ft_sequence_driver.cpp -- host-side C++ loop over a stereo sequence, written against the public C ABI only
// (include/fasttrack_b200.h). It is what a C++ tracking thread does per frame with HOST buffers:
//   Frame constructor   : ft_frame_construct, or ft_frame_submit / ft_frame_collect with several frames in flight
//   SearchLocalPoints   : marshal the local map into the context's pinned staging (the reference's CudaMapPoint
//                         loop, src/Kernels/CudaWrappers/CudaMapPoint.cc:15-34) + ft_search_staged, or name the local
//                         map as rows of the persistent store (ft_map_store_update + ft_search_store)
// bench.py times these loops for its end-to-end legs (the Python loop around the same calls is reported beside them);
// every host<->device copy of a step happens inside the timed region. Built by fasttrack_b200/build.py with g++.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/fasttrack_b200.h"

extern "C" {

typedef struct {
  int n_frames, width, height, M;
  const uint8_t* const* imgL;      // [n_frames] host images, row pitch = width
  const uint8_t* const* imgR;
  const float* const* pos;         // [n_frames] local map of frame k: pos[M][3] ...
  const float* const* normal;
  const float* const* minmax;
  const uint8_t* const* desc;
  const int* const* flags;
  const int* const* rows;          // [n_frames] store rows of the local map (store variant), else NULL
} ftd_sequence;

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct HostFrame {   // the host vectors of ORB_SLAM3::Frame that the constructor fills
  std::vector<ft_keypoint> kL, kR;
  std::vector<uint8_t> dL, dR;
  std::vector<float> ur, dp;
  std::vector<int> holder; std::vector<uint8_t> hobs; std::vector<int> best;
  int counts[4];
  explicit HostFrame(int cap, int M) : kL(cap), kR(cap), dL((size_t)cap * 32), dR((size_t)cap * 32), ur(cap), dp(cap),
                                       holder(2 * (size_t)cap), hobs(2 * (size_t)cap), best(2 * (size_t)M + 2) {}
};

static ft_status search_snapshot(ft_context* c, const ftd_sequence* s, int k, int nl, float th, long long* matches) {
  float *p, *n, *mm; uint8_t* d; int *f, *h; uint8_t* ho;
  ft_status st = ft_map_point_staging(c, s->M, &p, &n, &mm, &d, &f, &h, &ho);
  if (st != FT_OK) return st;
  const size_t M = (size_t)s->M;
  memcpy(p, s->pos[k], 12 * M); memcpy(n, s->normal[k], 12 * M); memcpy(mm, s->minmax[k], 8 * M);
  memcpy(d, s->desc[k], 32 * M); memcpy(f, s->flags[k], 4 * M);
  memset(h, 0xFF, sizeof(int) * (size_t)nl); memset(ho, 0, (size_t)nl);   // F.mvpMapPoints all NULL after the constructor
  const int *hOut, *bOut; const uint8_t* oOut; int nm = 0;
  st = ft_search_staged(c, s->M, th, 0, 50.f, 0.8f, &hOut, &oOut, &bOut, &nm);
  *matches += nm;
  return st;
}

#define FTD(call) do { ft_status st_ = (call); if (st_ != FT_OK) return -(double)st_; } while (0)

// Host wall clock per phase of the pipelined loop, summed over the last ftd_run_pipelined call (two clock reads per
// phase, about 50 ns): what the tracking thread spends where. ftd_phase_seconds copies the sums out.
enum { PH_FRAME_SUBMIT, PH_FRAME_COLLECT, PH_STORE_UPDATE, PH_SEARCH_SUBMIT, PH_SEARCH_COLLECT, PH_SEARCH_SYNC, PH_MARSHAL, PH_COUNT };
static double g_phase[PH_COUNT];
struct PhaseTimer {
  int id; double t0;
  explicit PhaseTimer(int id_) : id(id_), t0(now_s()) {}
  ~PhaseTimer() { g_phase[id] += now_s() - t0; }
};
#define FTD_PH(id, call) do { PhaseTimer pt_(id); FTD(call); } while (0)
void ftd_phase_seconds(double* out7) { for (int i = 0; i < PH_COUNT; i++) out7[i] = g_phase[i]; }

// one frame at a time on one context: returns wall seconds for `steps` frames, < 0 on error (ft_last_error has the text)
double ftd_run_serial(ft_context* c, const ftd_sequence* s, int steps, float th, long long* matches) {
  HostFrame F(ft_max_keypoints(c), s->M);
  *matches = 0;
  FTD(ft_synchronize(c));
  const double t0 = now_s();
  for (int i = 0; i < steps; i++) {
    const int k = i % s->n_frames;
    FTD(ft_frame_construct(c, s->imgL[k], s->width, s->imgR[k], s->width, F.kL.data(), F.dL.data(), F.kR.data(), F.dR.data(),
                           F.counts, F.ur.data(), F.dp.data(), nullptr, nullptr, nullptr));
    FTD(search_snapshot(c, s, k, F.counts[0], th, matches));
  }
  FTD(ft_synchronize(c));
  return now_s() - t0;
}

// Frame i's host vectors + the rows the mapping side changed for it (store variants)
static ft_status collect_frame(ft_context* c, const ftd_sequence* s, int k, HostFrame& F, int use_store, int upserts) {
  ft_status st;
  {
    PhaseTimer pt(PH_FRAME_COLLECT);
    st = ft_frame_collect(c, F.kL.data(), F.dL.data(), F.kR.data(), F.dR.data(), F.counts, F.ur.data(), F.dp.data(), nullptr,
                          nullptr, nullptr);
  }
  if (st != FT_OK || !use_store) return st;
  PhaseTimer pt(PH_STORE_UPDATE);
  return ft_map_store_update(c, upserts, s->rows[k], s->pos[k], s->normal[k], s->minmax[k], s->desc[k]);
}

// D frames in flight over D contexts of one sequence: frame i+D-1 is submitted (upload + extraction + stereo + result
// download enqueued) before frame i is collected, marshalled and searched. use_store != 0: the local map is named as
// rows of the persistent store (created by the caller on ctxs[0], attached to the others), `upserts` rows are
// re-uploaded per frame.
// use_store == 2: the search is issued as ft_search_store_submit / ft_search_collect and frame i+D-1 is handed to
// ft_frame_submit BETWEEN the two halves: the search of frame i is first in the GPU's queues and the host's work for the
// next camera frame (two image uploads, one graph launch) overlaps the search instead of preceding it.
// use_store == 3 (D >= 3): additionally the host vectors of frame i+1 are collected and its upserts enqueued while the
// search of frame i runs (they do not depend on the tracking of frame i: the Frame constructor's outputs and the mapping
// side's changed rows); only the holders of frame i+1 -- F.mvpMapPoints after tracking -- wait for the search of frame i.
// Two host frames alternate, so frame i's vectors stay valid until its search has been collected.
double ftd_run_pipelined(ft_context** ctxs, int D, const ftd_sequence* s, int steps, float th, int use_store, int upserts,
                         long long* matches) {
  if (D < 1) return -1.0;
  HostFrame F0(ft_max_keypoints(ctxs[0]), s->M), F1(ft_max_keypoints(ctxs[0]), s->M);
  HostFrame* Fs[2] = {&F0, &F1};
  *matches = 0;
  for (int i = 0; i < PH_COUNT; i++) g_phase[i] = 0.0;
  for (int j = 0; j < D; j++) FTD(ft_synchronize(ctxs[j]));
  const double t0 = now_s();
  for (int j = 0; j < D - 1 && j < steps; j++) {
    const int k = j % s->n_frames;
    FTD_PH(PH_FRAME_SUBMIT, ft_frame_submit(ctxs[j % D], s->imgL[k], s->width, s->imgR[k], s->width));
  }
  const bool split = use_store >= 2 && D >= 2;   // with one context the next frame would overwrite the one being searched
  const bool early = use_store == 3 && D >= 3;   // frame i+1 must have been submitted in an earlier iteration
  if (early && steps > 0) FTD(collect_frame(ctxs[0], s, 0, *Fs[0], use_store, upserts));
  for (int i = 0; i < steps; i++) {
    const int j = i + D - 1;                     // the camera frame handed over in this iteration
    const int kj = j % s->n_frames;
    if (j < steps && !split) FTD_PH(PH_FRAME_SUBMIT, ft_frame_submit(ctxs[j % D], s->imgL[kj], s->width, s->imgR[kj], s->width));
    ft_context* c = ctxs[i % D];
    const int k = i % s->n_frames;
    HostFrame& F = *Fs[early ? (i & 1) : 0];
    if (!early) FTD(collect_frame(c, s, k, F, use_store, upserts));
    const int nl = F.counts[0];
    if (!use_store) {
      FTD_PH(PH_SEARCH_SYNC, search_snapshot(c, s, k, nl, th, matches));
      continue;
    }
    {
      PhaseTimer pt(PH_MARSHAL);
      std::fill(F.holder.begin(), F.holder.begin() + nl, -1);
      std::fill(F.hobs.begin(), F.hobs.begin() + nl, (uint8_t)0);
    }
    int nm = 0;
    if (!split) {
      FTD_PH(PH_SEARCH_SYNC, ft_search_store(c, s->M, s->rows[k], s->flags[k], th, 0, 50.f, 0.8f, F.holder.data(), F.hobs.data(), F.best.data(), &nm));
    } else {
      FTD_PH(PH_SEARCH_SUBMIT, ft_search_store_submit(c, s->M, s->rows[k], s->flags[k], th, 0, 50.f, 0.8f, F.holder.data(), F.hobs.data(), 1));
      // in the shadow of the search: the next camera frame goes to the context frame i-1 left ...
      if (j < steps) FTD_PH(PH_FRAME_SUBMIT, ft_frame_submit(ctxs[j % D], s->imgL[kj], s->width, s->imgR[kj], s->width));
      // ... and frame i+1, in flight since the previous iteration, delivers its vectors and takes its upserts
      if (early && i + 1 < steps)
        FTD(collect_frame(ctxs[(i + 1) % D], s, (i + 1) % s->n_frames, *Fs[(i + 1) & 1], use_store, upserts));
      FTD_PH(PH_SEARCH_COLLECT, ft_search_collect(c, F.holder.data(), F.hobs.data(), F.best.data(), &nm));
    }
    *matches += nm;
  }
  for (int j = 0; j < D; j++) FTD(ft_synchronize(ctxs[j]));
  return now_s() - t0;
}

}  // extern "C"
