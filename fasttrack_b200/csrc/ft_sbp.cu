// ft_sbp.cu -- Tracking::SearchLocalPoints on the device-resident frame (sm_100a).
//
//   k_grid_build   Frame::AssignFeaturesToGrid + PosInGrid          (reference src/Frame.cc:409-440,749-759)
//   k_frustum      Frame::isInFrustum / isInFrustumChecks + MapPoint::PredictScale
//                  (src/Frame.cc:536-598,1308-1382; src/MapPoint.cc:502-546)
//   k_gather       Frame::GetFeaturesInArea in its traversal order + Hamming distance per candidate
//                  (src/Frame.cc:681-747; src/ORBmatcher.cc:85-114)
//   k_resolve      the best / second-best scan and the loop-carried keypoint claims of
//                  ORBmatcher::SearchByProjection (src/ORBmatcher.cc:57-225), solved as a fix-point
//   k_resolve_seq  the same loop executed in order by one thread: only used for fisheye rigs when a searched
//                  map point has Observations()==0 (its mirrored write can un-block a keypoint)
//
// Claims. In the reference a keypoint taken by map point i (with Observations()>0) is invisible to every j>i.
// Every read/write gets a time stamp: left search of map point j = 2j, right search = 2j+1. A slot is blocked
// at time t iff it was blocked on entry or some blocking write has a stamp < t. Given the decisions of all
// map points the per-slot minimum stamp is an atomicMin; given the stamps every decision is independent. A
// decision only depends on decisions with smaller stamps, so iterating the two steps reaches the unique
// sequential result; each round finalises at least the next undecided map point and in practice the depth
// of the longest chain of displaced matches (reported as `rounds`).
#include <cstdlib>

#include "ft_device.cuh"
#include "ft_internal.h"
#include "ft_camera.cuh"

#define GRID_CELLS (FT_GRID_COLS * FT_GRID_ROWS)

// ---- frame grid -------------------------------------------------------------------------
#define GRID_SIDX 4096   // keypoints whose index lists are built in shared memory
__global__ void __launch_bounds__(1024) k_grid_build(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                     const __grid_constant__ FtGridBuffers g, int fisheye, float minX,
                                                     float minY, float gridWInv, float gridHInv,
                                                     const __grid_constant__ FtUndistort und) {
  __shared__ int sCnt[GRID_CELLS];
  __shared__ int sStart[GRID_CELLS + 1];
  __shared__ int sWarp[32];
  __shared__ int sIdx[GRID_SIDX];
  const int tid = threadIdx.x;
  const int nEyes = fisheye ? 2 : 1;
  for (int eye = 0; eye < nEyes; eye++) {
    const FtEye& E = b.eye[eye];
    const int n = E.counts[0];
    int* cellStart = g.cellStart + eye * FT_GRID_STRIDE;
    int* cellIdx = g.cellIdx + eye * p.maxKp;
    for (int c = tid; c < GRID_CELLS; c += 1024) sCnt[c] = 0;
    __syncthreads();
    float4* rec = g.rec + eye * p.maxKp;
    for (int i = tid; i < n; i += 1024) {
      const ft_keypoint kp = E.kps[i];
      float kx = kp.x, ky = kp.y;
      if (eye == 0) {   // mvKeysUn (Frame::UndistortKeyPoints): the grid and the search use the undistorted keypoint
        if (und.on) ft_undistort_point(und, kp.x, kp.y, kx, ky);
        g.kpUn[i] = make_float2(kx, ky);
      }
      rec[i] = make_float4(kx, ky, __int_as_float(kp.octave), 0.f);   // 16-byte search record read by k_gather
      const int px = (int)roundf(__fmul_rn(__fsub_rn(kx, minX), gridWInv));
      const int py = (int)roundf(__fmul_rn(__fsub_rn(ky, minY), gridHInv));
      if (px < 0 || px >= FT_GRID_COLS || py < 0 || py >= FT_GRID_ROWS) continue;
      atomicAdd(&sCnt[px * FT_GRID_ROWS + py], 1);
    }
    __syncthreads();
    // exclusive scan over 3072 cells: 3 per thread
    {
      const int c0 = tid * 3;
      const int a0 = sCnt[c0], a1 = sCnt[c0 + 1], a2 = sCnt[c0 + 2];
      const int sum = a0 + a1 + a2;
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += t;
      }
      if ((tid & 31) == 31) sWarp[tid >> 5] = incl;
      __syncthreads();
      if (tid < 32) {
        int w = sWarp[tid];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
          if (tid >= o) wi += t;
        }
        sWarp[tid] = wi - w;
      }
      __syncthreads();
      const int base = sWarp[tid >> 5] + incl - sum;
      sStart[c0] = base; sStart[c0 + 1] = base + a0; sStart[c0 + 2] = base + a0 + a1;
      if (tid == 1023) sStart[GRID_CELLS] = base + sum;
    }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += 1024) sCnt[c] = 0;
    __syncthreads();
    // the index lists are built and ordered in shared memory (one coalesced write at the end) whenever the frame's
    // keypoints fit; larger frames build them in place in global memory
    const bool inS = n <= GRID_SIDX;
    int* idxW = inS ? sIdx : cellIdx;
    for (int i = tid; i < n; i += 1024) {
      const float4 kr = rec[i];   // written above by this thread
      const int px = (int)roundf(__fmul_rn(__fsub_rn(kr.x, minX), gridWInv));
      const int py = (int)roundf(__fmul_rn(__fsub_rn(kr.y, minY), gridHInv));
      if (px < 0 || px >= FT_GRID_COLS || py < 0 || py >= FT_GRID_ROWS) continue;
      const int c = px * FT_GRID_ROWS + py;
      idxW[sStart[c] + atomicAdd(&sCnt[c], 1)] = i;
    }
    __syncthreads();
    // cells list keypoints in ascending index (insertion order of the reference's push_back loop)
    for (int c = tid; c < GRID_CELLS; c += 1024) {
      const int s0 = sStart[c], s1 = sStart[c + 1];
      for (int i = s0 + 1; i < s1; i++) {
        const int v = idxW[i];
        int j = i - 1;
        while (j >= s0 && idxW[j] > v) { idxW[j + 1] = idxW[j]; j--; }
        idxW[j + 1] = v;
      }
    }
    if (inS) {
      __syncthreads();
      const int total = sStart[GRID_CELLS];
      for (int i = tid; i < total; i += 1024) cellIdx[i] = sIdx[i];
    }
    for (int c = tid; c <= GRID_CELLS; c += 1024) cellStart[c] = sStart[c];
    __syncthreads();
  }
}

// ---- frustum ----------------------------------------------------------------------------

__device__ bool ft_frustum_checks(const FtFrustumArgs& a, const float* P, const float* Pn, float minDistRaw,
                                  float maxDistRaw, bool right, bool pinholeMode, float& u, float& v, float& xr,
                                  float& depth, float& viewCosOut, int& level) {
  float mR[9], mt[3], twc[3];
  if (right) {
    // mR = Rrl * Rcw; mt = Rrl * tcw + trl; twc = Rwc * tlr + Ow (Frame.cc:1314-1320)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        mR[3 * i + j] = __fadd_rn(__fadd_rn(__fmul_rn(a.pose.Rrl[3 * i], a.pose.Rcw[j]), __fmul_rn(a.pose.Rrl[3 * i + 1], a.pose.Rcw[3 + j])),
                                  __fmul_rn(a.pose.Rrl[3 * i + 2], a.pose.Rcw[6 + j]));
    float Rt[3], Rw[3];
    ft_mat3_vec(a.pose.Rrl, a.pose.tcw, Rt);
    ft_mat3_vec(a.pose.Rwc, a.pose.tlr, Rw);
    for (int i = 0; i < 3; i++) { mt[i] = __fadd_rn(Rt[i], a.pose.trl[i]); twc[i] = __fadd_rn(Rw[i], a.pose.Ow[i]); }
  } else {
    for (int i = 0; i < 9; i++) mR[i] = a.pose.Rcw[i];
    for (int i = 0; i < 3; i++) { mt[i] = a.pose.tcw[i]; twc[i] = a.pose.Ow[i]; }
  }
  float Pc[3];
  ft_mat3_vec(mR, P, Pc);
  for (int i = 0; i < 3; i++) Pc[i] = __fadd_rn(Pc[i], mt[i]);
  const float PcDist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(Pc[0], Pc[0]), __fmul_rn(Pc[1], Pc[1])), __fmul_rn(Pc[2], Pc[2])));
  const float PcZ = Pc[2];
  if (PcZ < 0.0f) return false;
  float uv[2];
  ft_cam_project(right ? a.cam2 : a.cam1, Pc, uv);
  if (uv[0] < a.minX || uv[0] > a.maxX) return false;
  if (uv[1] < a.minY || uv[1] > a.maxY) return false;
  if (pinholeMode) { u = uv[0]; v = uv[1]; }   // mTrackProjX/Y are written before the distance checks (Frame.cc:563-564)
  const float maxDistance = __fmul_rn(1.2f, maxDistRaw), minDistance = __fmul_rn(0.8f, minDistRaw);
  const float PO[3] = {__fsub_rn(P[0], twc[0]), __fsub_rn(P[1], twc[1]), __fsub_rn(P[2], twc[2])};
  const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(PO[0], PO[0]), __fmul_rn(PO[1], PO[1])), __fmul_rn(PO[2], PO[2])));
  if (dist < minDistance || dist > maxDistance) return false;
  const float viewCos = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(PO[0], Pn[0]), __fmul_rn(PO[1], Pn[1])), __fmul_rn(PO[2], Pn[2])), dist);
  if (viewCos < a.viewCosLimit) return false;
  // MapPoint::PredictScale (MapPoint.cc:531-546): ceil(log(maxDistRaw/dist)/logScale), float log via double
  const float ratio = __fdiv_rn(maxDistRaw, dist);
  // float log first; the double evaluation (which defines the result) only when ceil() could go either way
  float q = __fdiv_rn(logf(ratio), a.logScale);
  if (fabsf(q - rintf(q)) < 1e-3f || !(fabsf(q) < 1e6f)) q = __fdiv_rn((float)log((double)ratio), a.logScale);
  int nScale = (int)ceilf(q);
  if (nScale < 0) nScale = 0;
  else if (nScale >= a.nlevels) nScale = a.nlevels - 1;
  u = uv[0]; v = uv[1];
  xr = pinholeMode ? __fsub_rn(uv[0], __fmul_rn(a.mbf, __fdiv_rn(1.0f, PcZ))) : 0.f;
  depth = PcDist; viewCosOut = viewCos; level = nScale;
  return true;
}

// ---- bulk asynchronous copies (TMA engine, 1-D) with an mbarrier that counts the bytes landed --------------------
__device__ __forceinline__ uint32_t ft_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ft_saddr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ft_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ft_saddr(bar)), "r"(bytes) : "memory");
}
// size: multiple of 16 bytes; source and destination 16-byte aligned
__device__ __forceinline__ void ft_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(ft_saddr(dst)), "l"(src), "r"(bytes), "r"(ft_saddr(bar)) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(ft_saddr(bar)), "r"(parity) : "memory");
  }
}

// ---- frustum + candidate gathering, one warp per map point ---------------------------------------
// All lanes evaluate the (cheap) frustum test redundantly from broadcast loads, so the scratch values never make
// a round trip through memory; lane 0 stores them for the caller. In-view map points then walk their grid window
// once: passing keypoints are collected in traversal order into a small per-warp shared-memory buffer
// (GA_BUF entries; longer lists take a second walk straight into the pool), and the lanes finally split the
// candidates evenly for the Hamming distances.
#ifndef GA_WARPS
#define GA_WARPS 16
#endif
#define GA_CTAS_PER_SM (GA_WARPS >= 32 ? 1 : 2)
#define GA_BUF 96
#define GA_INLINE 16        // candidate slots every (map point, branch) owns in the pool; longer lists allocate from the overflow area
#define GA_ACTIVE_CAP 512   // map points one CTA can process: ceil(max_map_points / resident CTAs) + slack
__global__ void __launch_bounds__(GA_WARPS * 32) k_gather(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                          const __grid_constant__ FtGridBuffers g,
                                                          const __grid_constant__ FtStereoBuffers st,
                                                          const __grid_constant__ FtSbpBuffers s,
                                                          const __grid_constant__ FtFrustumArgs fa,
                                                          const __grid_constant__ FtGatherArgs a, int M, int stage) {
  extern __shared__ __align__(16) uint8_t gaSmem[];
  __shared__ unsigned short sBuf[GA_WARPS][GA_BUF];
  __shared__ int sCol[GA_WARPS][FT_GRID_COLS], sPre[GA_WARPS][FT_GRID_COLS];
  // per-CTA totals, folded into the global cursors once per CTA (thousands of same-address global atomics were the
  // kernel's critical resource): active map points of this CTA, candidates found, searched non-blocking map points
  __shared__ int sActive[GA_ACTIVE_CAP];
  __shared__ int sNActive, sNCand, sNNonBlocking, sActiveBase;
  __shared__ __align__(8) unsigned long long sStageBar;
  FT_PDL_TRIGGER();     // the claim-resolution cluster may be scheduled; it waits for this grid before it reads
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nEyes = a.fisheye ? 2 : 1;
  if (tid == 0) { sNActive = 0; sNCand = 0; sNNonBlocking = 0; }   // visible after the barrier below
  // The frame-side search structure (grid CSR + 16-byte keypoint records + uRight) is ~40 KB: every CTA stages it
  // in shared memory once, so the window walks below never leave the SM. One thread issues four to seven bulk copies
  // (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier: the TMA engine moves the arrays while
  // the warps already run the frustum tests of their first map points); a warp waits for the barrier in front of
  // its first window walk. Whole arrays are copied (capacity, not the frame's keypoint count), so nothing has to be read
  // before the copies can be issued. (stage == 0: structure too large for shared memory, e.g. 10k features; the same
  // code then reads it from L2.)
  const float4* recP[2] = {g.rec, g.rec + p.maxKp};
  const int* cellStartP[2] = {g.cellStart, g.cellStart + FT_GRID_STRIDE};
  const int* cellIdxP[2] = {g.cellIdx, g.cellIdx + p.maxKp};
  const float* uRightP = st.uRight;
  bool staged = !stage;
  if (stage) {
    uint8_t* q = gaSmem;
    const uint32_t recBytes = (uint32_t)sizeof(float4) * p.maxKp, csBytes = (uint32_t)sizeof(int) * FT_GRID_STRIDE,
                   ciBytes = (uint32_t)sizeof(int) * p.maxKp, urBytes = (uint32_t)sizeof(float) * p.maxKp;
    if (tid == 0) {
      ft_mbar_init(&sStageBar, 1);
      ft_mbar_expect_tx(&sStageBar, (uint32_t)nEyes * (recBytes + csBytes + ciBytes) + (a.fisheye ? 0u : urBytes));
    }
    for (int e = 0; e < nEyes; e++) {
      float4* r = reinterpret_cast<float4*>(q); q += recBytes;
      int* cs = reinterpret_cast<int*>(q); q += csBytes;
      int* ci = reinterpret_cast<int*>(q); q += ciBytes;
      if (tid == 0) {
        ft_bulk_g2s(r, recP[e], recBytes, &sStageBar);
        ft_bulk_g2s(cs, cellStartP[e], csBytes, &sStageBar);
        ft_bulk_g2s(ci, cellIdxP[e], ciBytes, &sStageBar);
      }
      recP[e] = r; cellStartP[e] = cs; cellIdxP[e] = ci;
    }
    if (!a.fisheye) {
      float* u = reinterpret_cast<float*>(q);
      if (tid == 0) ft_bulk_g2s(u, st.uRight, urBytes, &sStageBar);
      uRightP = u;
    }
  }
  __syncthreads();      // the counters above and the initialised mbarrier are visible to every warp
  // cursors [0] pool, [3] non-blocking count, [4] active count are accumulated here; they are zero on entry
  // (cleared at allocation and by the resolve kernel of the previous search)
  for (int mp = blockIdx.x * GA_WARPS + warp; mp < M; mp += gridDim.x * GA_WARPS) {
    const int flags = s.flags[mp];
    const int src = s.slot ? s.slot[mp] : mp;   // persistent map store: the local map is a list of rows
    int inView = 0, inViewR = 0, level = -1, levelR = -1;
    float f[9] = {-1.f, -1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float f2fRadius = 0.f;
    if (a.mode == 1) {
      // frame-to-last-frame search (ORBmatcher.cc:1797-1822): plain projection of the last frame's map point, window
      // th * scale[octave of the last-frame keypoint]; minmax[] carries {angle, octave} of that keypoint
      if (!(flags & 1)) {
        const float P[3] = {s.pos[3 * mp], s.pos[3 * mp + 1], s.pos[3 * mp + 2]};
        float Pc[3];
        ft_mat3_vec(fa.pose.Rcw, P, Pc);
        for (int i = 0; i < 3; i++) Pc[i] = __fadd_rn(Pc[i], fa.pose.tcw[i]);
        const float invzc = (float)(1.0 / (double)Pc[2]);
        if (!(invzc < 0)) {
          float uv[2];
          ft_cam_project(fa.cam1, Pc, uv);
          if (!(uv[0] < fa.minX || uv[0] > fa.maxX) && !(uv[1] < fa.minY || uv[1] > fa.maxY)) {
            const int oct = (int)s.minmax[2 * mp + 1];
            inView = 1; level = oct;
            f[0] = uv[0]; f[1] = uv[1];
            f[2] = __fsub_rn(uv[0], __fmul_rn(fa.mbf, invzc));   // ur = uv(0) - mbf*invzc (:1857)
            f2fRadius = __fmul_rn(a.th, p.scale[oct]);
            if (fa.fisheye) {
              // right image: Trl * x3Dc projected with the LEFT camera model, as the reference does (:1919-1921)
              float Pr[3], uvr[2];
              ft_mat3_vec(fa.pose.Rrl, Pc, Pr);
              for (int i = 0; i < 3; i++) Pr[i] = __fadd_rn(Pr[i], fa.pose.trl[i]);
              ft_cam_project(fa.cam1, Pr, uvr);
              inViewR = 1; levelR = oct; f[5] = uvr[0]; f[6] = uvr[1];
            }
          }
        }
      }
    } else if (!(flags & 1)) {
      const float P[3] = {s.pos[3 * src], s.pos[3 * src + 1], s.pos[3 * src + 2]};
      const float Pn[3] = {s.normal[3 * src], s.normal[3 * src + 1], s.normal[3 * src + 2]};
      const float mn = s.minmax[2 * src], mx = s.minmax[2 * src + 1];
      if (!fa.fisheye) {
        float u = -1, v = -1, xr = 0, d = 0, vc = 0; int lv = -1;
        const bool ok = ft_frustum_checks(fa, P, Pn, mn, mx, false, true, u, v, xr, d, vc, lv);
        f[0] = u; f[1] = v;
        if (ok) { inView = 1; f[2] = xr; f[3] = d; f[4] = vc; level = lv; }
      } else {
        float u = 0, v = 0, xr = 0, d = 0, vc = 0; int lv = -1;
        if (ft_frustum_checks(fa, P, Pn, mn, mx, false, false, u, v, xr, d, vc, lv)) {
          inView = 1; f[0] = u; f[1] = v; f[3] = d; f[4] = vc; level = lv;
        }
        if (ft_frustum_checks(fa, P, Pn, mn, mx, true, false, u, v, xr, d, vc, lv)) {
          inViewR = 1; f[5] = u; f[6] = v; f[7] = d; f[8] = vc; levelR = lv;
        }
      }
    }
    if (lane < 9) s.trF[9 * mp + lane] = f[lane];
    if (lane == 0) {
      *reinterpret_cast<int4*>(s.trI + 4 * mp) = make_int4(inView, inViewR, level, levelR);
      *reinterpret_cast<int2*>(s.sel + 2 * mp) = make_int2(-1, -1);
    }
    bool searched = (inView || inViewR) && !(flags & 1);
    if (a.bFar && f[3] > a.thFar) searched = false;   // mTrackDepth > thFarPoints (ORBmatcher.cc:66)
    if (lane == 0 && searched && !(flags & 2)) atomicAdd(&sNNonBlocking, 1);
    int2 lens = make_int2(0, 0), offs = make_int2(0, 0);
    if (searched) {
      const uint4* md = reinterpret_cast<const uint4*>(s.desc + (size_t)src * 32);
      const uint4 md0 = md[0], md1 = md[1];
      for (int br = 0; br < nEyes; br++) {
        const bool active = (br == 0 ? inView : inViewR) && (br == 0 || levelR != -1);
        if (!active) continue;
        const int lvl = br == 0 ? level : levelR;
        const float x = br == 0 ? f[0] : f[5];
        const float y = br == 0 ? f[1] : f[6];
        const float viewCos = br == 0 ? f[4] : f[8];
        float r = ((double)viewCos > 0.998) ? 2.5f : 4.0f;           // RadiusByViewingCos (ORBmatcher.cc:314-320)
        if (br == 0 && a.bFactor) r = __fmul_rn(r, a.th);
        float rr = __fmul_rn(r, p.scale[lvl]);
        int minLevel = lvl - 1, maxLevel = lvl;
        if (a.mode == 1) {   // (:1826-1834)
          rr = f2fRadius;
          if (a.direction > 0) { minLevel = lvl; maxLevel = -1; }
          else if (a.direction < 0) { minLevel = 0; maxLevel = lvl; }
          else { minLevel = lvl - 1; maxLevel = lvl + 1; }
        }
        // GetFeaturesInArea cell window (Frame.cc:689-707)
        const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.minX), rr), a.gridWInv)));
        const int cx1 = min(FT_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.minX), rr), a.gridWInv)));
        const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.minY), rr), a.gridHInv)));
        const int cy1 = min(FT_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.minY), rr), a.gridHInv)));
        // an empty LEFT window ends the last-frame search of this point before its right-eye half:
        // `if(vIndices2.empty()) continue;` (ORBmatcher.cc:1836-1837) sits in front of the Nleft != -1 block (:1915)
        if (!(cx0 < FT_GRID_COLS && cx1 >= 0 && cy0 < FT_GRID_ROWS && cy1 >= 0 && cx1 >= cx0 && cy1 >= cy0)) {
          if (a.mode == 1 && br == 0) break;
          continue;
        }
        if (!staged) { ft_mbar_wait(&sStageBar, 0); staged = true; }   // the bulk copies of the search structure have landed
        const int* cellStart = cellStartP[br];
        const int* cellIdx = cellIdxP[br];
        const float4* rec = recP[br];
        const FtEye& E = b.eye[br];
        const float projXR = f[2];
        auto passes = [&](int idx) -> bool {
          const float4 kp = rec[idx];
          const int oct = __float_as_int(kp.z);
          if (oct < minLevel) return false;
          if (maxLevel >= 0 && oct > maxLevel) return false;
          const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
          if (!(fabsf(dx) < rr && fabsf(dy) < rr)) return false;
          if (!a.fisheye) {
            const float ur = uRightP[idx];
            if (ur > 0) {
              const float er = fabsf(__fsub_rn(projXR, ur));
              if (er > rr) return false;                       // stereo consistency (ORBmatcher.cc:105-110)
            }
          }
          return true;
        };
        // Cells of one grid column are consecutive in the CSR (cell = ix*48 + iy), so the window is ncol contiguous
        // index ranges that, concatenated in ix order, are exactly the traversal order (ix outer, iy inner, insertion
        // order). Lanes first fetch the per-column ranges, then split the flattened element list evenly.
        const int ncol = cx1 - cx0 + 1;
        int T = 0;
        __syncwarp();   // the previous walk (which may have left through `continue`) has finished reading sCol / sPre
        for (int j0 = 0; j0 < ncol; j0 += 32) {
          const int j = j0 + lane;
          int cs = 0, ce = 0;
          if (j < ncol) { const int cb = (cx0 + j) * FT_GRID_ROWS; cs = cellStart[cb + cy0]; ce = cellStart[cb + cy1 + 1]; }
          const int len = ce - cs;
          int incl = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
          }
          if (j < ncol) { sCol[warp][j] = cs; sPre[warp][j] = T + incl - len; }
          T += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        __syncwarp();
        auto walk = [&](uint32_t* dst) -> int {
          int run = 0;
          for (int e0 = 0; e0 < T; e0 += 32) {
            const int e = e0 + lane;
            bool ok = false;
            int idx = 0;
            if (e < T) {
              int lo = 0, hi = ncol;                 // last column whose first element is <= e
              while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sPre[warp][mid] <= e) lo = mid; else hi = mid; }
              idx = cellIdx[sCol[warp][lo] + (e - sPre[warp][lo])];
              ok = passes(idx);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (ok) {
              const int w = run + __popc(m & ((1u << lane) - 1u));
              if (dst) dst[w] = (uint32_t)idx;
              else if (w < GA_BUF) sBuf[warp][w] = (unsigned short)idx;
            }
            run += __popc(m);
          }
          return run;
        };
        const int total = walk(nullptr);
        if (total == 0) {
          if (a.mode == 1 && br == 0) break;     // (:1836-1837), see above
          continue;
        }
        // short lists (the common case) live in the slots the (map point, branch) pair owns; only longer ones
        // allocate from the overflow area behind the M*2*GA_INLINE inline slots
        int base = (2 * mp + br) * GA_INLINE;
        if (total > GA_INLINE) {
          if (lane == 0) base = 2 * M * GA_INLINE + atomicAdd(&s.cursor[0], total);
          base = __shfl_sync(0xFFFFFFFFu, base, 0);
          if (base + total > s.poolCap) {
            if (lane == 0) atomicOr(b.status, FT_ST_SBP_POOL_OVERFLOW);
            continue;
          }
        }
        if (lane == 0) atomicAdd(&sNCand, total);
        if (total > GA_BUF) walk(s.pool + base);
        __syncwarp();
        // Hamming distance + octave per candidate
        for (int k = lane; k < total; k += 32) {
          const int idx = total > GA_BUF ? (int)s.pool[base + k] : (int)sBuf[warp][k];
          const uint4* dd = reinterpret_cast<const uint4*>(E.desc + (size_t)idx * 32);
          const int dist = ft_hamming256(md0, md1, dd[0], dd[1]);
          s.pool[base + k] = (uint32_t)idx | ((uint32_t)dist << 16) | ((uint32_t)__float_as_int(rec[idx].z) << 25);
        }
        __syncwarp();
        if (br == 0) { lens.x = total; offs.x = base; } else { lens.y = total; offs.y = base; }
      }
    }
    if (lane == 0) {
      *reinterpret_cast<int2*>(s.listOff + 2 * mp) = offs;
      *reinterpret_cast<int2*>(s.listLen + 2 * mp) = lens;
      if (lens.x | lens.y) {
        const int k = atomicAdd(&sNActive, 1);
        if (k < GA_ACTIVE_CAP) sActive[k] = mp; else s.active[atomicAdd(&s.cursor[4], 1)] = mp;   // (cannot happen with the launch geometry)
      }
    }
  }
  if (!staged) ft_mbar_wait(&sStageBar, 0);   // no CTA retires while bulk copies into its shared memory are in flight
  __syncthreads();
  const int nAct = min(sNActive, GA_ACTIVE_CAP);
  if (tid == 0) {
    sActiveBase = nAct ? atomicAdd(&s.cursor[4], nAct) : 0;
    if (sNCand) atomicAdd(&s.cursor[9], sNCand);
    if (sNNonBlocking) atomicAdd(&s.cursor[3], sNNonBlocking);
  }
  __syncthreads();
  for (int i = tid; i < nAct; i += GA_WARPS * 32) s.active[sActiveBase + i] = sActive[i];
}

// ---- claim resolution ---------------------------------------------------------------------

// best / second-best scan of one candidate list (ORBmatcher.cc:88-141). Returns the accepted keypoint or -1;
// *cont is set when the reference executes `continue` (ratio test failed on same-level neighbours).
template <typename BlockedFn>
__device__ __forceinline__ int ft_scan_list(const uint32_t* list, int len, float nnratio, BlockedFn blocked, bool* cont,
                                            bool bestOnly = false) {
  int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
  for (int k = 0; k < len; k++) {
    const uint32_t e = __ldg(list + k);
    const int idx = (int)(e & 0xFFFFu);
    if (blocked(idx)) continue;
    const int dist = (int)((e >> 16) & 0x1FFu), oct = (int)(e >> 25);
    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
    else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
  }
  *cont = false;
  if (bestDist <= 100) {   // TH_HIGH
    if (!bestOnly && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) { *cont = true; return -1; }
    return bestIdx;
  }
  return -1;
}

// One thread-block cluster (8 or 16 CTAs, co-scheduled on one GPC) so that the best/second-best scans of all
// candidate lists issue from several SMs at once: the rounds are latency-bound, not bandwidth-bound. The
// pre-blocked flags are replicated into every CTA's shared memory. The per-slot minimum stamps never leave the
// cluster: slot i is OWNED by CTA i / per, which keeps three rotating stamp buffers for its slots in its own shared
// memory; decisions are scattered with red.min through distributed shared memory into the owner's buffer, and after
// the round's single cluster barrier every CTA copies the whole table (a few KB) from the owners into its local
// mkS[] for the next round's scans. Against the round-1 scheme (atomicMin at L2 + read back with ld.global.cg) this
// takes two L2 round trips and the wait for the atomics' acknowledgements out of every round. The "something changed"
// flags, the final highest-stamp table and the rotation histogram live in shared memory the same way (flags and
// histogram in CTA 0). Only map points with a non-empty candidate list are visited (compact list written by k_gather).
#include <cooperative_groups.h>
#include <type_traits>
#include <algorithm>
namespace cg = cooperative_groups;

#define RS_THREADS 512
#define RS_LCAP 12   // candidate entries per thread cached in shared memory (longer lists continue from L2)

// best / second-best scan over a list whose first RS_LCAP entries sit in shared memory as sort keys (column-major, one
// column per thread) and the rest in global memory. Key of the entry at list position k: dist << 23 | k << 19 | octave << 15
// (0xFFFFFFFF beyond the end of the list): the reference's in-order loop with its two strict comparisons keeps exactly
// the two smallest (dist, position) pairs -- the old best becomes the second best, ties keep the earlier entry -- so the
// scan is a branch-free (min, second-min) reduction over keys whose blocked entries are replaced by 0xFFFFFFFF.
template <typename BlockedFn>
__device__ __forceinline__ int ft_scan_cached(const uint32_t* sKey, const uint16_t* sIdx, const uint32_t* gList, int len, float nnratio,
                                              BlockedFn blocked, bool* cont, bool bestOnly = false) {
  // The cached entries and their "blocked" look-ups are independent of each other: fetch them all first (waves of
  // shared-memory loads in flight instead of a chain of dependent ones), then reduce on registers.
  uint32_t key[RS_LCAP];
  int idx[RS_LCAP];
  bool blk[RS_LCAP];
#pragma unroll
  for (int k = 0; k < RS_LCAP; k++) { key[k] = sKey[k * RS_THREADS]; idx[k] = sIdx[k * RS_THREADS]; }
#pragma unroll
  for (int k = 0; k < RS_LCAP; k++) blk[k] = blocked(idx[k]);
  uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
#pragma unroll
  for (int k = 0; k < RS_LCAP; k++) {
    const uint32_t x = blk[k] ? 0xFFFFFFFFu : key[k];
    m2 = min(m2, max(m1, x));
    m1 = min(m1, x);
  }
  int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
  if ((m1 >> 23) < 256u) { bestDist = (int)(m1 >> 23); bestLevel = (int)((m1 >> 15) & 15u); bestIdx = sIdx[((m1 >> 19) & 15u) * RS_THREADS]; }
  if ((m2 >> 23) < 256u) { bestDist2 = (int)(m2 >> 23); bestLevel2 = (int)((m2 >> 15) & 15u); }
  for (int k = RS_LCAP; k < len; k++) {          // longer lists continue from L2 (rare)
    const uint32_t eg = __ldg(gList + k);
    const int id = (int)(eg & 0xFFFFu);
    if (blocked(id)) continue;
    const int dist = (int)((eg >> 16) & 0x1FFu), oct = (int)(eg >> 25);
    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = id; }
    else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
  }
  *cont = false;
  if (bestDist <= 100) {
    if (!bestOnly && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) { *cont = true; return -1; }
    return bestIdx;
  }
  return -1;
}

#ifdef FT_RS_CLOCK
#define RS_TICK(slot) do { if (threadIdx.x == 0 && b.eye[1].octClock) b.eye[1].octClock[blockIdx.x * 64 + (slot)] = clock64(); } while (0)
#else
#define RS_TICK(slot) do { } while (0)
#endif
__device__ __forceinline__ int rs_per(int nSlots, int nCta) { return max(8, (((nSlots + nCta - 1) / nCta) + 7) & ~7); }
// distributed shared memory through the shared::cluster window (no generic addressing, reductions without a return value)
__device__ __forceinline__ uint32_t rs_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rs_mapa(uint32_t saddr, int rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void rs_red_min(uint32_t a, int v) { asm volatile("red.relaxed.cluster.shared::cluster.min.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void rs_red_max(uint32_t a, int v) { asm volatile("red.relaxed.cluster.shared::cluster.max.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void rs_red_add(uint32_t a, int v) { asm volatile("red.relaxed.cluster.shared::cluster.add.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void rs_st(uint32_t a, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int rs_ld(uint32_t a) { int v; asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

__global__ void __launch_bounds__(RS_THREADS, 1) k_resolve(const __grid_constant__ FtBuffers b, const __grid_constant__ FtSbpBuffers s,
                                                           const __grid_constant__ FtStereoBuffers st,
                                                           const __grid_constant__ FtResolveArgs a0) {
  extern __shared__ int sMem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int nCta = gridDim.x;
  const int nThreads = nCta * RS_THREADS;
  const int gtid = blockIdx.x * RS_THREADS + tid;
  // map points are dealt round-robin over the CTAs of the cluster so that every SM gets an equal share of the
  // candidate-list scans (the active list is usually shorter than the cluster's thread count)
  const int vid = tid * nCta + blockIdx.x;
  RS_TICK(0);
  // Everything up to FT_PDL_WAIT reads only what was complete before k_gather started (keypoint counts, the caller's
  // holders) and writes only shared memory: it overlaps the tail of k_gather.
  FtResolveArgs a = a0;
  a.nLeft = b.eye[0].counts[0];
  a.nSlots = a.fisheye ? a.nLeft + b.eye[1].counts[0] : a.nLeft;
  const int nS = a.nSlots;
  const bool bestOnly = a.mode == 1;
  // shared memory: mkS[cap] int (this round's stamps, whole table; slots that were taken on entry hold INT_MIN),
  // own[3][perCap] (the stamp buffers of the slots this CTA owns), ownInit[perCap] (their value when nobody claims them),
  // last[perCap] (highest stamp per owned slot), flag[4] + hist[32] (used in CTA 0 only), keyL / keyR [RS_LCAP][RS_THREADS] u32,
  // idxL / idxR [RS_LCAP][RS_THREADS] u16 (the right eye's only for fisheye rigs). a0.nSlots = capacity bound (2*maxKp).
  const int perCap = rs_per(a0.nSlots, nCta);
  const int per = rs_per(nS, nCta);              // slots per owner for this frame
  const uint32_t perMagic = 0xFFFFFFFFu / (uint32_t)per + 1u;   // slot / per == __umulhi(slot, perMagic) for slot * per < 2^32
  int* mkS = sMem;
  int* own = mkS + a0.nSlots;
  int* ownInit = own + 3 * perCap;
  int* last = ownInit + perCap;
  int* flag = last + perCap;
  int* hist = flag + 4;
  uint32_t* keyL = reinterpret_cast<uint32_t*>(hist + 32);
  uint32_t* keyR = keyL + RS_LCAP * RS_THREADS;
  uint16_t* idxL = reinterpret_cast<uint16_t*>(keyL + (a0.fisheye ? 2 : 1) * RS_LCAP * RS_THREADS);
  uint16_t* idxR = idxL + RS_LCAP * RS_THREADS;
  const uint32_t flag0 = rs_mapa(rs_saddr(flag), 0), hist0 = rs_mapa(rs_saddr(hist), 0);
  // the first slot whose final holder this thread writes (the owner CTA writes its slots): caller's holder fetched early
  const int mySlot = (tid < per && blockIdx.x * per + tid < nS) ? blockIdx.x * per + tid : -1;
  int hInit = -1; uint8_t hObsInit = 0;
  if (mySlot >= 0) { hInit = s.holderInit[mySlot]; hObsInit = s.holderObsInit[mySlot]; }
  // round 0 sees no stamps except the slots that hold a map point with observations on entry: blocked for everybody
  for (int i = tid; i < nS; i += RS_THREADS) mkS[i] = (s.holderInit[i] != -1 && s.holderObsInit[i]) ? (int)0x80000000 : 0x7FFFFFFF;
  for (int i = tid; i < per; i += RS_THREADS) {
    const int slot = blockIdx.x * per + i;
    const int v = (slot < nS && s.holderInit[slot] != -1 && s.holderObsInit[slot]) ? (int)0x80000000 : 0x7FFFFFFF;
    ownInit[i] = v; own[perCap + i] = v; last[i] = -1;     // buffer 1: round 0 scatters into it
  }
  if (tid < 4) flag[tid] = 0;
  if (tid >= 32 && tid < 64) hist[tid - 32] = 0;
  FT_PDL_WAIT();        // launched as a programmatic dependent of k_gather
  const int nA = s.cursor[4];
  RS_TICK(1);
  // k_resolve_seq resolves fisheye local-map searches with non-blocking map points (uniform over the cluster); the
  // last-frame search has no mirrored writes, so the stamp rule is exact for it in every case
  const bool seqMode = a.fisheye && a.mode == 0 && s.cursor[3] > 0;
  if (seqMode) {
    // working holders start from the caller's (F.mvpMapPoints on entry); k_resolve_seq continues from them
    for (int i = gtid; i < nS; i += nThreads) { s.holder[i] = s.holderInit[i]; s.holderObs[i] = s.holderObsInit[i]; }
    return;
  }
  if (gtid == 3) s.cursor[1] = 0;
  // this thread's first map point: list heads cached in shared memory as sort keys, decision kept in registers
  int myMp = -1, myFlags = 0;
  int2 myOff = make_int2(0, 0), myLen = make_int2(0, 0);
  int selL = -1, selR = -1;
  if (vid < nA) {
    myMp = __ldg(&s.active[vid]);
    myFlags = __ldg(&s.flags[myMp]);
    myOff = __ldg(reinterpret_cast<const int2*>(s.listOff) + myMp);
    myLen = __ldg(reinterpret_cast<const int2*>(s.listLen) + myMp);
    uint32_t eL[RS_LCAP], eR[RS_LCAP];
#pragma unroll
    for (int k = 0; k < RS_LCAP; k++) {
      eL[k] = k < myLen.x ? __ldg(s.pool + myOff.x + k) : 0xFFFFFFFFu;
      eR[k] = k < myLen.y ? __ldg(s.pool + myOff.y + k) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < RS_LCAP; k++) {
      keyL[k * RS_THREADS + tid] = k < myLen.x ? (((eL[k] >> 16) & 0x1FFu) << 23) | ((uint32_t)k << 19) | ((eL[k] >> 25) << 15) : 0xFFFFFFFFu;
      idxL[k * RS_THREADS + tid] = k < myLen.x ? (uint16_t)(eL[k] & 0xFFFFu) : (uint16_t)0;
      if (a.fisheye) {
        keyR[k * RS_THREADS + tid] = k < myLen.y ? (((eR[k] >> 16) & 0x1FFu) << 23) | ((uint32_t)k << 19) | ((eR[k] >> 25) << 15) : 0xFFFFFFFFu;
        idxR[k * RS_THREADS + tid] = k < myLen.y ? (uint16_t)(eR[k] & 0xFFFFu) : (uint16_t)0;
      }
    }
  }
  RS_TICK(2);
  cluster.sync();       // every CTA's tables are initialised before anybody scatters into them
  RS_TICK(3);
  // Round r scans against mkS (the stamps scattered in round r - 1), scatters the decisions it makes into the owners'
  // buffer (r+1)%3 and resets its own buffer (r+2)%3: one cluster barrier per round.
  auto stampMin = [&](uint32_t buf, int slot, int key) {
    const int o = (int)__umulhi((uint32_t)slot, perMagic);
    rs_red_min(rs_mapa(buf, o) + 4u * (uint32_t)(slot - o * per), key);
  };
  int rounds = 0;
  for (;;) {
    int* ownNext = own + ((rounds + 1) % 3) * perCap;
    int* ownClear = own + ((rounds + 2) % 3) * perCap;
    const uint32_t ownNextS = rs_saddr(ownNext);
    __syncthreads();    // mkS is complete (copied at the end of the previous round)
    for (int i = tid; i < per; i += RS_THREADS) ownClear[i] = ownInit[i];
    // The "changed" flag of round r lives in CTA 0's flag[r % 3]. Clear the flag of round r + 1 here: it was last read at
    // the end of round r - 2, and the cluster barrier of round r - 1 lies in between. (Clearing the flag of round r + 2 =
    // r - 1 would race with CTAs that have not read it yet at the end of round r - 1: a CTA reading 0 leaves the loop
    // while the others continue, and the cluster deadlocks in its next barrier.)
    if (gtid == 0) flag[(rounds + 1) % 3] = 0;
    if (rounds < 8) RS_TICK(4 + 4 * rounds);
    int changed = 0;
    // one map point: scans, comparison with the previous decision, stamps for the next round
    auto decide = [&](int mp, int2 off, int2 len, bool blocking, auto mineTag) {
      constexpr bool mine = decltype(mineTag)::value;
      int newL = -1, newR = -1;
      bool cont = false;
      const int t = 2 * mp;
      if (len.x > 0) {
        auto blk = [&](int idx) { return mkS[idx] < t; };
        if constexpr (mine) newL = ft_scan_cached(keyL + tid, idxL + tid, s.pool + off.x, len.x, a.nnratio, blk, &cont, bestOnly);
        else newL = ft_scan_list(s.pool + off.x, len.x, a.nnratio, blk, &cont, bestOnly);
      }
      if (len.y > 0 && !cont) {
        // own left writes (stamp 2mp) are handled explicitly, older stamps through the table
        const int ownMirror = (a.mode == 0 && blocking && newL >= 0 && st.l2r[newL] != -1) ? st.l2r[newL] : -1;
        bool contR = false;
        auto blk = [&](int idx) { return (mkS[idx + a.nLeft] < t) | (idx == ownMirror); };
        if constexpr (mine) newR = ft_scan_cached(keyR + tid, idxR + tid, s.pool + off.y, len.y, a.nnratio, blk, &contR, bestOnly);
        else newR = ft_scan_list(s.pool + off.y, len.y, a.nnratio, blk, &contR, bestOnly);
      }
      if constexpr (mine) {
        if (newL != selL || newR != selR) { changed = 1; selL = newL; selR = newR; }
      } else {
        const int2 old = make_int2(__ldcg(&s.sel[2 * mp]), __ldcg(&s.sel[2 * mp + 1]));
        if (newL != old.x || newR != old.y) { changed = 1; *reinterpret_cast<int2*>(s.sel + 2 * mp) = make_int2(newL, newR); }
      }
      if (blocking) {   // stamps the next round reads (writes of map points without observations never block)
        const bool mirror = a.fisheye && a.mode == 0;   // mirrored assignments exist only in the local-map search
        if (newL >= 0) {
          stampMin(ownNextS, newL, 2 * mp);
          if (mirror && st.l2r[newL] != -1) stampMin(ownNextS, st.l2r[newL] + a.nLeft, 2 * mp);
        }
        if (newR >= 0) {
          stampMin(ownNextS, newR + a.nLeft, 2 * mp + 1);
          if (mirror && st.r2l[newR] != -1) stampMin(ownNextS, st.r2l[newR], 2 * mp + 1);
        }
      }
    };
    if (vid < nA) decide(myMp, myOff, myLen, (myFlags & 2) != 0, std::true_type());
    for (int k = vid + nThreads; k < nA; k += nThreads) {      // more active map points than threads in the cluster (rare)
      const int mp = __ldg(&s.active[k]);
      decide(mp, __ldg(reinterpret_cast<const int2*>(s.listOff) + mp), __ldg(reinterpret_cast<const int2*>(s.listLen) + mp),
             (__ldg(&s.flags[mp]) & 2) != 0, std::false_type());
    }
    if (__any_sync(0xFFFFFFFFu, changed) && (tid & 31) == 0) rs_st(flag0 + 4u * (uint32_t)(rounds % 3), 1);
    if (rounds < 8) RS_TICK(5 + 4 * rounds);
    cluster.sync();
    if (rounds < 8) RS_TICK(6 + 4 * rounds);
    // the "anything changed" flag and the stamps the next round reads are fetched together (one trip through distributed
    // shared memory); every thread of the CTA is past its scans, so the local copy may be overwritten. Every CTA starts
    // with the slots it owns and walks the owners in its own rotation, so the 16 readers do not queue up at one owner.
    const int ch = rs_ld(flag0 + 4u * (uint32_t)(rounds % 3));
    for (int w = tid >> 5; w < nCta; w += RS_THREADS / 32) {     // one warp per owner: coalesced rows, no address arithmetic
      int o = w + blockIdx.x; if (o >= nCta) o -= nCta;
      const uint32_t src = rs_mapa(ownNextS, o);
      const int base = o * per;
      for (int j = tid & 31; j < per && base + j < nS; j += 32) mkS[base + j] = rs_ld(src + 4u * (uint32_t)j);
    }
    if (rounds < 8) RS_TICK(7 + 4 * rounds);
    rounds++;
    if (!ch) break;
    if (rounds > a.M + 2) { if (gtid == 0) atomicOr(b.status, FT_ST_RESOLVE_NOCONV); break; }
  }
  // final holders: the write with the highest stamp wins each slot; count matches (ORBmatcher.cc:142-155,207-222)
  const bool mirror = a.fisheye && a.mode == 0;
  const bool rotCheck = a.mode == 1 && a.checkOri;
  auto rotBin = [&](int mp, int idx, bool right) -> int {   // (ORBmatcher.cc:1890-1899): factor = 1/HISTO_LENGTH
    const float angLF = s.minmax[2 * mp];
    const float angCF = right ? b.eye[1].kps[idx].angle : b.eye[0].kps[idx].angle;
    float rot = __fsub_rn(angLF, angCF);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
    if (bin == 30) bin = 0;
    return bin;
  };
  const uint32_t lastS = rs_saddr(last);
  auto stampMax = [&](int slot, int key) {
    const int o = (int)__umulhi((uint32_t)slot, perMagic);
    rs_red_max(rs_mapa(lastS, o) + 4u * (uint32_t)(slot - o * per), key);
  };
  int nm = 0;
  for (int k = vid; k < nA; k += nThreads) {
    const bool mine = (k == vid);
    const int mp = mine ? myMp : __ldg(&s.active[k]);
    const int sl = mine ? selL : __ldcg(&s.sel[2 * mp]), sr = mine ? selR : __ldcg(&s.sel[2 * mp + 1]);
    if (mine) *reinterpret_cast<int2*>(s.sel + 2 * mp) = make_int2(sl, sr);
    if (sl >= 0) {
      stampMax(sl, 2 * mp); nm++;
      if (mirror && st.l2r[sl] != -1) { stampMax(st.l2r[sl] + a.nLeft, 2 * mp); nm++; }
      if (rotCheck) rs_red_add(hist0 + 4u * (uint32_t)rotBin(mp, sl, false), 1);
    }
    if (sr >= 0) {
      stampMax(sr + a.nLeft, 2 * mp + 1); nm++;
      if (mirror && st.r2l[sr] != -1) { stampMax(st.r2l[sr], 2 * mp + 1); nm++; }
      if (rotCheck) rs_red_add(hist0 + 4u * (uint32_t)rotBin(mp, sr, true), 1);
    }
  }
  if (nm) atomicAdd(&s.cursor[1], nm);
  RS_TICK(40);
  cluster.sync();
  RS_TICK(41);
  for (int i = tid; i < per; i += RS_THREADS) {    // the owner writes each slot's holder exactly once
    const int slot = blockIdx.x * per + i;
    if (slot >= nS) break;
    const int k = last[i];
    if (k >= 0) { s.holder[slot] = k >> 1; s.holderObs[slot] = (uint8_t)((s.flags[k >> 1] >> 1) & 1); }
    else if (i == tid) { s.holder[slot] = hInit; s.holderObs[slot] = hObsInit; }
    else { s.holder[slot] = s.holderInit[slot]; s.holderObs[slot] = s.holderObsInit[slot]; }
  }
  if (rotCheck) {
    // ComputeThreeMaxima (ORBmatcher.cc:2210-2254) on the bin counts, then every match outside the three strongest
    // bins is withdrawn (:2057-2079)
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < 30; i++) {
      const int c = rs_ld(hist0 + 4u * (uint32_t)i);
      if (c > max1) { max3 = max2; max2 = max1; max1 = c; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (c > max2) { max3 = max2; max2 = c; ind3 = ind2; ind2 = i; }
      else if (c > max3) { max3 = c; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
    cluster.sync();     // holders are written and CTA 0's histogram has been read by everybody
    int removed = 0;
    for (int k = vid; k < nA; k += nThreads) {
      const int mp = __ldg(&s.active[k]);
      const int sl = __ldcg(&s.sel[2 * mp]), sr = __ldcg(&s.sel[2 * mp + 1]);
      if (sl >= 0) {
        const int bn = rotBin(mp, sl, false);
        if (bn != ind1 && bn != ind2 && bn != ind3) { s.holder[sl] = -1; s.holderObs[sl] = 0; removed++; }
      }
      if (sr >= 0) {
        const int bn = rotBin(mp, sr, true);
        if (bn != ind1 && bn != ind2 && bn != ind3) { s.holder[sr + a.nLeft] = -1; s.holderObs[sr + a.nLeft] = 0; removed++; }
      }
    }
    if (removed) atomicSub(&s.cursor[1], removed);
    cluster.sync();
  }
  RS_TICK(42);
  if (gtid == 0) { s.cursor[2] = rounds; s.cursor[7] = s.cursor[9]; s.cursor[9] = 0; s.cursor[0] = 0; s.cursor[3] = 0; s.cursor[4] = 0; s.cursor[8] = *b.status; }
}

// In-order execution by one thread (fisheye rigs with non-blocking map points only).
__global__ void k_resolve_seq(const __grid_constant__ FtBuffers b, const __grid_constant__ FtSbpBuffers s,
                              const __grid_constant__ FtStereoBuffers st, const __grid_constant__ FtResolveArgs a0) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  FtResolveArgs a = a0;
  a.nLeft = b.eye[0].counts[0];
  a.nSlots = a.fisheye ? a.nLeft + b.eye[1].counts[0] : a.nLeft;
  if (!(a.fisheye && a.mode == 0 && s.cursor[3] > 0)) return;   // the fix-point kernel handles this frame
  int nm = 0;
  for (int mp = 0; mp < a.M; mp++) {
    const int lenL = s.listLen[2 * mp], lenR = s.listLen[2 * mp + 1];
    s.sel[2 * mp] = -1; s.sel[2 * mp + 1] = -1;
    if (lenL == 0 && lenR == 0) continue;
    const uint8_t obs = (uint8_t)((s.flags[mp] >> 1) & 1);
    bool cont = false;
    if (lenL > 0) {
      const int sl = ft_scan_list(s.pool + s.listOff[2 * mp], lenL, a.nnratio,
                                  [&](int idx) { return s.holder[idx] != -1 && s.holderObs[idx]; }, &cont);
      if (sl >= 0) {
        s.sel[2 * mp] = sl;
        s.holder[sl] = mp; s.holderObs[sl] = obs; nm++;
        if (a.fisheye && st.l2r[sl] != -1) { s.holder[st.l2r[sl] + a.nLeft] = mp; s.holderObs[st.l2r[sl] + a.nLeft] = obs; nm++; }
      }
    }
    if (lenR > 0 && !cont) {
      bool contR = false;
      const int sr = ft_scan_list(s.pool + s.listOff[2 * mp + 1], lenR, a.nnratio,
                                  [&](int idx) { return s.holder[idx + a.nLeft] != -1 && s.holderObs[idx + a.nLeft]; }, &contR);
      if (sr >= 0) {
        s.sel[2 * mp + 1] = sr;
        if (st.r2l[sr] != -1) { s.holder[st.r2l[sr]] = mp; s.holderObs[st.r2l[sr]] = obs; nm++; }
        s.holder[sr + a.nLeft] = mp; s.holderObs[sr + a.nLeft] = obs; nm++;
      }
    }
  }
  s.cursor[1] = nm; s.cursor[2] = 0; s.cursor[7] = s.cursor[9]; s.cursor[9] = 0; s.cursor[0] = 0; s.cursor[3] = 0; s.cursor[4] = 0; s.cursor[8] = *b.status;
}

// ---- host launchers -----------------------------------------------------------------------
static int g_resolveCluster = 8;
static size_t ft_gather_smem(const FtParams& p, int fisheye) {
  const size_t perEye = sizeof(float4) * p.maxKp + sizeof(int) * (GRID_CELLS + 4) + sizeof(int) * p.maxKp;
  return (fisheye ? 2 : 1) * perEye + (fisheye ? 0 : sizeof(float) * p.maxKp) + 16;
}
static int ft_resolve_per(int slots, int nCta) { return std::max(8, (((slots + nCta - 1) / nCta) + 7) & ~7); }   // = rs_per on the device
static size_t ft_resolve_smem(int slotCap, int nCta, int fisheye) {
  return (size_t)slotCap * 4 + (size_t)5 * ft_resolve_per(slotCap, nCta) * 4 + (4 + 32) * 4 +
         (size_t)(fisheye ? 2 : 1) * RS_LCAP * RS_THREADS * (4 + 2) + 16;
}
cudaError_t ft_launch_sbp_setup(const FtParams& p) {
  // per function and per device: the opt-in maximum once, never lowered by a later, smaller context
  cudaError_t e = ft_set_max_dynamic_smem((const void*)k_resolve);
  if (e != cudaSuccess) return e;
  e = ft_set_max_dynamic_smem((const void*)k_gather);
  if (e != cudaSuccess) return e;
  // 16-CTA clusters are a non-portable size: opt in, and fall back to the portable 8 when the device cannot place one
  g_resolveCluster = 8;
  if (cudaFuncSetAttribute(k_resolve, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(RS_THREADS); cfg.dynamicSmemBytes = ft_resolve_smem(2 * p.maxKp, 16, 1);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nClusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nClusters, k_resolve, &cfg) == cudaSuccess && nClusters >= 1) g_resolveCluster = 16;
  }
  if (const char* e = getenv("FT_RESOLVE_CLUSTER")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) g_resolveCluster = v; }
  cudaGetLastError();
  return cudaSuccess;
}
// Persistent map store: upsert n packed records (staging layout: slots | pos | normal | minmax | desc) into the rows
// they name. One thread per 32-bit word of a record (16 words = 64 bytes; the slot id is the 17th word of the staging record).
__global__ void k_store_scatter(int n, const int* __restrict__ slots, const uint32_t* __restrict__ pos,
                                const uint32_t* __restrict__ normal, const uint32_t* __restrict__ minmax,
                                const uint32_t* __restrict__ desc, uint32_t* dPos, uint32_t* dNormal, uint32_t* dMinmax,
                                uint32_t* dDesc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t >> 4, w = t & 15;
  if (r >= n) return;
  const size_t row = (size_t)slots[r];
  if (w < 3) dPos[row * 3 + w] = pos[(size_t)r * 3 + w];
  else if (w < 6) dNormal[row * 3 + (w - 3)] = normal[(size_t)r * 3 + (w - 3)];
  else if (w < 8) dMinmax[row * 2 + (w - 6)] = minmax[(size_t)r * 2 + (w - 6)];
  else dDesc[row * 8 + (w - 8)] = desc[(size_t)r * 8 + (w - 8)];
}
void ft_launch_store_scatter(int n, const uint8_t* staged, float* pos, float* normal, float* minmax, uint8_t* desc,
                             cudaStream_t st) {
  if (n <= 0) return;
  const int* slots = reinterpret_cast<const int*>(staged);
  const uint32_t* p = reinterpret_cast<const uint32_t*>(staged + (size_t)4 * n);
  const uint32_t* nm = reinterpret_cast<const uint32_t*>(staged + (size_t)16 * n);
  const uint32_t* mm = reinterpret_cast<const uint32_t*>(staged + (size_t)28 * n);
  const uint32_t* d = reinterpret_cast<const uint32_t*>(staged + (size_t)36 * n);
  const long long threads = 16LL * n;
  k_store_scatter<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, slots, p, nm, mm, d, reinterpret_cast<uint32_t*>(pos),
                                                                   reinterpret_cast<uint32_t*>(normal),
                                                                   reinterpret_cast<uint32_t*>(minmax),
                                                                   reinterpret_cast<uint32_t*>(desc));
}
void ft_launch_grid(const FtParams& p, const FtBuffers& b, const FtGridBuffers& g, int fisheye, float minX, float minY,
                    float gridWInv, float gridHInv, const FtUndistort& und, cudaStream_t st) {
  k_grid_build<<<1, 1024, 0, st>>>(p, b, g, fisheye, minX, minY, gridWInv, gridHInv, und);
}
void ft_launch_gather(const FtParams& p, const FtBuffers& b, const FtGridBuffers& g, const FtStereoBuffers& stb,
                      const FtSbpBuffers& s, const FtFrustumArgs& fa, const FtGatherArgs& ga, int M, cudaStream_t st) {
  const size_t smem = ft_gather_smem(p, ga.fisheye);
  const int stage = smem <= 200 * 1024;
  const int ctas = min((M + GA_WARPS - 1) / GA_WARPS, GA_CTAS_PER_SM * 148);   // resident CTAs only, grid-stride over map points
  k_gather<<<ctas, GA_WARPS * 32, stage ? smem : 0, st>>>(p, b, g, stb, s, fa, ga, M, stage);
}
void ft_launch_resolve(const FtBuffers& b, const FtSbpBuffers& s, const FtStereoBuffers& stb, const FtResolveArgs& ra,
                       cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g_resolveCluster); cfg.blockDim = dim3(RS_THREADS); cfg.dynamicSmemBytes = ft_resolve_smem(ra.nSlots, g_resolveCluster, ra.fisheye); cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = g_resolveCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // predecessor in the stream: k_gather
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = ft_pdl_enabled() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, k_resolve, b, s, stb, ra);
  if (ra.fisheye) k_resolve_seq<<<1, 32, 0, st>>>(b, s, stb, ra);   // exits immediately unless it is needed
}

// ---- the search chain as a CUDA graph ------------------------------------------------------------------------------
// gather -> resolve (-> resolve_seq on fisheye rigs) captured once per context; the pose, M, th and the map-point
// pointers change with every call, so the kernel nodes get their parameters (and the gather grid) through
// cudaGraphExecKernelNodeSetParams: the arguments stay in the constant bank, one cudaGraphLaunch per search.
cudaError_t ft_search_graph_run(FtSearchGraph* G, const FtParams& p, const FtBuffers& b, const FtGridBuffers& g,
                                const FtStereoBuffers& stb, const FtSbpBuffers& s, const FtFrustumArgs& fa,
                                const FtGatherArgs& ga, const FtResolveArgs& ra, int M, cudaStream_t st) {
  cudaError_t e;
  if (!G->exec) {
    cudaGraph_t graph = nullptr;
    e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return e;
    ft_launch_gather(p, b, g, stb, s, fa, ga, M, st);
    ft_launch_resolve(b, s, stb, ra, st);
    e = cudaStreamEndCapture(st, &graph);
    if (e != cudaSuccess) return e;
    e = cudaGraphInstantiate(&G->exec, graph, 0);
    if (e == cudaSuccess) {
      cudaGraphNode_t nodes[8];
      size_t n = 8;
      e = cudaGraphGetNodes(graph, nodes, &n);
      for (size_t i = 0; e == cudaSuccess && i < n; i++) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
        if (kp.func == (void*)k_gather) G->gather = nodes[i];
        else if (kp.func == (void*)k_resolve) G->resolve = nodes[i];
        else if (kp.func == (void*)k_resolve_seq) G->seq = nodes[i];
      }
      if (e == cudaSuccess && (!G->gather || !G->resolve || (ra.fisheye && !G->seq))) e = cudaErrorUnknown;
    }
    // the exec graph keeps the node handles of the graph it was instantiated from: the template stays alive with it
    G->graph = graph;
    if (e != cudaSuccess) return e;
  } else {
    const size_t smem = ft_gather_smem(p, ga.fisheye);
    int stage = smem <= 200 * 1024;
    int Marg = M;
    const int ctas = min((M + GA_WARPS - 1) / GA_WARPS, GA_CTAS_PER_SM * 148);
    void* ka[9] = {(void*)&p, (void*)&b, (void*)&g, (void*)&stb, (void*)&s, (void*)&fa, (void*)&ga, (void*)&Marg, (void*)&stage};
    cudaKernelNodeParams kp = {};
    kp.func = (void*)k_gather; kp.gridDim = dim3(ctas); kp.blockDim = dim3(GA_WARPS * 32);
    kp.sharedMemBytes = (unsigned)(stage ? smem : 0); kp.kernelParams = ka;
    e = cudaGraphExecKernelNodeSetParams(G->exec, G->gather, &kp);
    if (e != cudaSuccess) return e;
    void* kr[4] = {(void*)&b, (void*)&s, (void*)&stb, (void*)&ra};
    cudaKernelNodeParams kq = {};
    kq.func = (void*)k_resolve; kq.gridDim = dim3(g_resolveCluster); kq.blockDim = dim3(RS_THREADS);
    kq.sharedMemBytes = (unsigned)ft_resolve_smem(ra.nSlots, g_resolveCluster, ra.fisheye); kq.kernelParams = kr;
    e = cudaGraphExecKernelNodeSetParams(G->exec, G->resolve, &kq);
    if (e != cudaSuccess) return e;
    if (G->seq) {
      cudaKernelNodeParams ks = {};
      ks.func = (void*)k_resolve_seq; ks.gridDim = dim3(1); ks.blockDim = dim3(32); ks.sharedMemBytes = 0; ks.kernelParams = kr;
      e = cudaGraphExecKernelNodeSetParams(G->exec, G->seq, &ks);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaGraphLaunch(G->exec, st);
}
void ft_search_graph_destroy(FtSearchGraph* G) {
  if (G->exec) cudaGraphExecDestroy(G->exec);
  if (G->graph) cudaGraphDestroy(G->graph);
  *G = FtSearchGraph();
}
