// ft_extract.cu -- ORB extraction kernels for sm_100a (both eyes per launch).
//
// Stage map (reference = CPU branch of src/ORBextractor.cc; results are bit-exact):
//   k_resize        ComputePyramid               (:1495-1520)  cv::resize INTER_LINEAR, 11-bit fixed point
//   k_blur          GaussianBlur 7x7 sigma 2     (:1456-1457)  8.8 / 16.16 fixed point, reflect-101
//   k_fast_cells    per-cell cv::FAST + fallback (:1131-1203)  one CTA per cell, smem tile, ordered compaction
//   k_octree        DistributeOctTree            (:660-884)    one CTA per (eye, level), node list in smem
//   k_orient_desc   IC_Angle + rBRIEF + tail     (:39-108, :1392-1492) one warp per keypoint
#include <cstdio>
#include <cstdlib>

#include "ft_device.cuh"
#include "ft_internal.h"
#include "ft_sort.h"
#include "ft_camera.cuh"

__constant__ __align__(16) int8_t c_pattern[1024] = {
#include "../../include/ft_orb_pattern.inc"
};

// ------------------------------------------------------------------------------------
// Pyramid: level l from level l-1. 4 destination pixels per thread, uchar4 store.
// Coefficients come from tables built once on the host (they depend on sizes only).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ft_resize_rows(const FtBuffers& b, const uint8_t* src, int srcW, int srcPitch, uint8_t* dst,
                                               const FtLevel& L) {
  const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int dy = blockIdx.y * 8 + threadIdx.y;
  if (dx0 >= L.w || dy >= L.h) return;
  const int2 yt = b.yTab[L.yTab + dy];
  const int sy0 = yt.x & 0xFFFF, sy1 = yt.x >> 16;
  const int b0 = (short)(yt.y & 0xFFFF), b1 = (short)(yt.y >> 16);
  const uint8_t* r0 = src + (size_t)sy0 * srcPitch;
  const uint8_t* r1 = src + (size_t)sy1 * srcPitch;
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int dx = dx0 + k;
    if (dx < L.w) {
      const int2 xt = b.xTab[L.xTab + dx];
      const int sx = xt.x;
      const int sx1 = min(sx + 1, srcW - 1);
      const int a0 = (short)(xt.y & 0xFFFF), a1 = (short)(xt.y >> 16);
      const int h0 = r0[sx] * a0 + r0[sx1] * a1;
      const int h1 = r1[sx] * a0 + r1[sx1] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      out |= (uint32_t)(v & 0xFF) << (8 * k);
    }
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)dy * L.pitch + dx0) = out;
}

__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                int level) {
  const FtLevel& L = p.lv[level];
  const FtLevel& S = p.lv[level - 1];
  const int eye = blockIdx.z;
  FT_PDL_TRIGGER();     // the next level's resize may be scheduled; it waits for this grid before it reads
  FT_PDL_WAIT();        // levels >= 2 are launched as programmatic dependents of the previous level's resize
  ft_resize_rows(b, b.eye[eye].pyr + S.offset, S.w, S.pitch, b.eye[eye].pyr + L.offset, L);
}

// cv::resize(im, imToFeed, newImSize) of System::TrackStereo (reference src/System.cc:282-285, Settings::needToResize):
// the raw image is resized straight into level 0 with the level-0 slots of the coefficient tables.
__global__ void __launch_bounds__(256) k_resize_input(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                      const uint8_t* rawL, const uint8_t* rawR, int rawW) {
  const FtLevel& L = p.lv[0];
  const int eye = blockIdx.z;
  ft_resize_rows(b, eye ? rawR : rawL, rawW, rawW, b.eye[eye].pyr + L.offset, L);
}

// ------------------------------------------------------------------------------------
// Stereo rectification ("next" row 2, reference System::TrackStereo, src/System.cc:273-281):
// cv::remap(raw, M1, M2, INTER_LINEAR) straight into level 0 of the pyramid. The float maps are converted once on
// the host to OpenCV's fixed-point form (integer source pixel + 5-bit fractions); the 2x2 weights are
// (32-fy)(32-fx)*32 ... scaled by 2^15 with BilinearTab_i's single saturated entry {32767,0,0,1} at fx = fy = 0.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_remap(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                               const uint8_t* rawL, const uint8_t* rawR, const int2* tab, int rawW, int rawH) {
  const FtLevel& L = p.lv[0];
  const int eye = blockIdx.z;
  const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
  if (x >= L.w || y >= L.h) return;
  const uint8_t* src = eye ? rawR : rawL;
  const int2 t = tab[(size_t)eye * L.w * L.h + (size_t)y * L.w + x];
  const int ix = (short)(t.x & 0xFFFF), iy = (short)(t.x >> 16);
  const int fx = t.y & 0xFF, fy = (t.y >> 8) & 0xFF;
  int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  if ((fx | fy) == 0) { w00 = 32767; w11 = 1; }
  auto px = [&](int yy, int xx) -> int { return (xx >= 0 && xx < rawW && yy >= 0 && yy < rawH) ? (int)src[(size_t)yy * rawW + xx] : 0; };
  const int v = w00 * px(iy, ix) + w01 * px(iy, ix + 1) + w10 * px(iy + 1, ix) + w11 * px(iy + 1, ix + 1);
  b.eye[eye].pyr[L.offset + (size_t)y * L.pitch + x] = (uint8_t)min(max((v + (1 << 14)) >> 15, 0), 255);
}

// ------------------------------------------------------------------------------------
// Gaussian blur, all levels of both eyes in one launch. Tile = 64 x 32 outputs per CTA.
// ------------------------------------------------------------------------------------
#define BLUR_TW 64
#define BLUR_TH 32
__device__ __forceinline__ int ft_reflect101(int q, int n) {
  if (n == 1) return 0;
  while (q < 0 || q >= n) q = q < 0 ? -q : 2 * (n - 1) - q;
  return q;
}

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                              int levelBegin, int levelEnd) {
  // Word-oriented: the tile (+3 px halo, padded to whole words) is staged with aligned 32-bit loads, a thread of the
  // horizontal pass produces 4 neighbouring sums from 3 words, a thread of the vertical pass slides a 7-row window
  // down 4 rows of a column pair. Same integer arithmetic as before (8.8 horizontal, 16.16 vertical, round-half-up).
  constexpr int WPR = BLUR_TW / 4 + 2;                       // staged words per row: columns [tx-4, tx+TW+4)
  __shared__ __align__(16) uint32_t sIn[BLUR_TH + 6][WPR];
  __shared__ __align__(16) uint16_t sH[BLUR_TH + 6][BLUR_TW];
  const int eye = blockIdx.y;
  int level = levelBegin;
  const int tile = blockIdx.x + p.lv[levelBegin].blurTileBase;
  while (level + 1 < levelEnd && tile >= p.lv[level + 1].blurTileBase) level++;
  const FtLevel& L = p.lv[level];
  const int t = tile - L.blurTileBase;
  const int tx = (t % L.blurTilesX) * BLUR_TW, ty = (t / L.blurTilesX) * BLUR_TH;
  const uint8_t* src = b.eye[eye].pyr + L.offset;
  uint8_t* dst = b.eye[eye].blur + L.offset;
  const int tid = threadIdx.x;
  for (int i = tid; i < (BLUR_TH + 6) * WPR; i += 256) {
    const int ry = i / WPR, wx = i - ry * WPR;
    const int gy = ft_reflect101(ty + ry - 3, L.h);
    const int gx0 = tx - 4 + 4 * wx;
    const uint8_t* row = src + (size_t)gy * L.pitch;
    uint32_t w;
    if (gx0 >= 0 && gx0 + 3 < L.w) {
      w = *reinterpret_cast<const uint32_t*>(row + gx0);       // rows are 64-byte aligned, gx0 is a multiple of 4
    } else {                                                   // border word: reflect-101 per byte
      w = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) w |= (uint32_t)row[ft_reflect101(gx0 + k, L.w)] << (8 * k);
    }
    sIn[ry][wx] = w;
  }
  __syncthreads();
  // horizontal: item = (row, group of 4 columns); output column x reads staged bytes x+1 .. x+7
  for (int i = tid; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += 256) {
    const int ry = i / (BLUR_TW / 4), g = i % (BLUR_TW / 4);
    const uint32_t w0 = sIn[ry][g], w1 = sIn[ry][g + 1], w2 = sIn[ry][g + 2];
    unsigned by[10];
    by[0] = (w0 >> 8) & 0xFF; by[1] = (w0 >> 16) & 0xFF; by[2] = w0 >> 24;
    by[3] = w1 & 0xFF; by[4] = (w1 >> 8) & 0xFF; by[5] = (w1 >> 16) & 0xFF; by[6] = w1 >> 24;
    by[7] = w2 & 0xFF; by[8] = (w2 >> 8) & 0xFF; by[9] = (w2 >> 16) & 0xFF;
    unsigned h[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      h[k] = 18u * (by[k] + by[k + 6]) + 34u * (by[k + 1] + by[k + 5]) + 48u * (by[k + 2] + by[k + 4]) + 56u * by[k + 3];
    *reinterpret_cast<uint2*>(&sH[ry][4 * g]) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  }
  __syncthreads();
  {
    const int cp = tid & 31;            // column pair
    const int yg = (tid >> 5) * 4;      // 8 groups of 4 rows
    const int gx = tx + 2 * cp;
    if (gx < L.w) {
      unsigned lo[10], hi[10];
#pragma unroll
      for (int r = 0; r < 10; r++) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(&sH[yg + r][2 * cp]);
        lo[r] = w & 0xFFFFu; hi[r] = w >> 16;
      }
      const bool two = gx + 1 < L.w;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int gy = ty + yg + k;
        if (gy < L.h) {
          const unsigned v0 = 18u * (lo[k] + lo[k + 6]) + 34u * (lo[k + 1] + lo[k + 5]) + 48u * (lo[k + 2] + lo[k + 4]) + 56u * lo[k + 3];
          const unsigned v1 = 18u * (hi[k] + hi[k + 6]) + 34u * (hi[k + 1] + hi[k + 5]) + 48u * (hi[k + 2] + hi[k + 4]) + 56u * hi[k + 3];
          const unsigned o0 = (v0 + 32768u) >> 16, o1 = (v1 + 32768u) >> 16;
          uint8_t* q = dst + (size_t)gy * L.pitch + gx;          // gx is even
          if (two) *reinterpret_cast<uint16_t*>(q) = (uint16_t)(o0 | (o1 << 8));
          else *q = (uint8_t)o0;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// FAST-9/16 per cell. One CTA (FAST_THREADS threads) per cell of one eye.
// score(p) = max over the 16 arcs of 9 contiguous ring pixels of min(+-diff) - 1; corner(th) <=> score >= th.
// The cell first tries iniThFAST; when no pixel survives NMS it falls back to minThFAST
// (ORBextractor.cc:1157-1177). NMS only sees scores inside the cell's interior, as cv::FAST on a ROI does.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int ft_fast_score(const uint8_t* c, int stride) {
  // ring order (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
  const int v = c[0];
  int d[16];
  d[0] = v - c[3 * stride];      d[1] = v - c[3 * stride + 1];  d[2] = v - c[2 * stride + 2];  d[3] = v - c[stride + 3];
  d[4] = v - c[3];               d[5] = v - c[-stride + 3];     d[6] = v - c[-2 * stride + 2]; d[7] = v - c[-3 * stride + 1];
  d[8] = v - c[-3 * stride];     d[9] = v - c[-3 * stride - 1]; d[10] = v - c[-2 * stride - 2]; d[11] = v - c[-stride - 3];
  d[12] = v - c[-3];             d[13] = v - c[stride - 3];     d[14] = v - c[2 * stride - 2]; d[15] = v - c[3 * stride - 1];
  // two signed 16-bit lanes per word: low = d (bright-centre arcs), high = -d (dark-centre arcs)
  unsigned x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = ((unsigned)d[i] & 0xFFFFu) | ((unsigned)(-d[i]) << 16);
  unsigned m2[16], m4[16], m8[16];
#pragma unroll
  for (int i = 0; i < 16; i++) m2[i] = __vmins2(x[i], x[(i + 1) & 15]);
#pragma unroll
  for (int i = 0; i < 16; i++) m4[i] = __vmins2(m2[i], m2[(i + 2) & 15]);
#pragma unroll
  for (int i = 0; i < 16; i++) m8[i] = __vmins2(m4[i], m4[(i + 4) & 15]);
  unsigned best = __vmins2(m8[0], x[8]);
#pragma unroll
  for (int i = 1; i < 16; i++) best = __vmaxs2(best, __vmins2(m8[i], x[(i + 8) & 15]));
  const int pos = (short)(best & 0xFFFFu), neg = (short)(best >> 16);
  return max(pos, neg) - 1;
}

// floor(i / n) for the small run-time divisors of the cell geometry (interior width, words per row, cell columns):
// c_recip[n] = ceil(2^32 / n), exact for i * n < 2^32; n = 1 has no 32-bit reciprocal and is passed through.
#define FT_RECIP_N 256
__constant__ unsigned c_recip[FT_RECIP_N];
__device__ __forceinline__ int ft_div_small(int i, int n, unsigned m) { return n == 1 ? i : (int)__umulhi((unsigned)i, m); }

#ifndef FAST_THREADS
#define FAST_THREADS 256
#endif
__global__ void __launch_bounds__(FAST_THREADS, 1536 / FAST_THREADS) k_fast_cells(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                    int levelBegin, int levelEnd) {
  extern __shared__ uint8_t smem[];
  __shared__ int sWarp[FAST_THREADS / 32];
  __shared__ int sAny, sPass;
  FT_PDL_TRIGGER();     // the level's octree kernel may be scheduled; it waits for this grid before it reads
  const int eye = blockIdx.y;
  const int cell = blockIdx.x + p.lv[levelBegin].cellBase;
  int level = levelBegin;
  while (level + 1 < levelEnd && cell >= p.lv[level + 1].cellBase) level++;
  const FtLevel& L = p.lv[level];
  const int ci = ft_div_small(cell - L.cellBase, L.nCols, c_recip[L.nCols]), cj = (cell - L.cellBase) - ci * L.nCols;
  const FtEye& E = b.eye[eye];
  const int tid = threadIdx.x;
  // cell window (ORBextractor.cc:1136-1153); all quantities are integers held in floats there
  const int iniY = FT_MIN_BORDER + ci * L.hCell, iniX = FT_MIN_BORDER + cj * L.wCell;
  int maxY = iniY + L.hCell + 6, maxX = iniX + L.wCell + 6;
  const bool skip = (iniY >= L.maxBorderY - 3) || (iniX >= L.maxBorderX - 6);
  if (maxY > L.maxBorderY) maxY = L.maxBorderY;
  if (maxX > L.maxBorderX) maxX = L.maxBorderX;
  const int rw = maxX - iniX, rh = maxY - iniY;
  if (skip || rw < 7 || rh < 7) {
    if (tid == 0) E.cellCount[cell] = 0;
    return;
  }
  const int iw = rw - 6, ih = rh - 6;
  // row / column of a flat interior index without an integer division per pixel: floor(i / iw) = umulhi(i, ceil(2^32 / iw))
  // (exact for i < 2^26 with iw <= 64)
  const unsigned mIw = c_recip[iw];
  const int ax = iniX & 3;                      // the tile is staged with aligned 32-bit loads: ax bytes of slack in front
  const int rwPad = (rw + 6) & ~3;              // >= ax + rw, multiple of 4
  const int S = rwPad >> 2;                     // words per staged row
  // score planes: 1-px zero border so that the 3x3 NMS needs no edge cases, rows a whole number of words;
  // interior pixel (y, x) lives at [(y + 1) * sS + x + 4]
  const int sS = (((iw + 3) >> 2) + 3) * 4;
  const int planeBytes = ((ih + 2) * sS + 15) & ~15;
  uint8_t* sImg = smem;                         // [rh][rwPad], cell pixel (y, x) at sImg[y*rwPad + ax + x]
  uint8_t* sSc = smem + ((rh * rwPad + 15) & ~15);   // score (0 below minTh)
  uint8_t* sMx = sSc + planeBytes;                   // NMS survivors
  uint16_t* sList = reinterpret_cast<uint16_t*>(sMx + planeBytes);   // pixels that pass the quick test (y << 8 | x)
  {
    const uint8_t* srcA = E.pyr + L.offset + (size_t)iniY * L.pitch + (iniX - ax);   // 4-byte aligned
    const int wpr = (ax + rw + 3) >> 2;                                               // words per row
    uint32_t* sW = reinterpret_cast<uint32_t*>(sImg);
    const unsigned mWpr = c_recip[wpr];
    for (int i = tid; i < rh * wpr; i += FAST_THREADS) {
      const int y = ft_div_small(i, wpr, mWpr), wx = i - y * wpr;
      sW[y * S + wx] = *reinterpret_cast<const uint32_t*>(srcA + (size_t)y * L.pitch + 4 * wx);
    }
    uint32_t* z = reinterpret_cast<uint32_t*>(sSc);
    for (int i = tid; i < planeBytes / 4; i += FAST_THREADS) z[i] = 0u;
  }
  if (tid == 0) { sAny = 0; sPass = 0; }
  __syncthreads();
  const int total = iw * ih;
  // The reference runs cv::FAST(iniTh) and, only when that leaves the cell empty AFTER non-maximum suppression, cv::FAST(minTh)
  // (:1157-1177). Same order here: on textured images nearly every cell is done after the first attempt, whose quick test
  // passes a third fewer pixels to the expensive score than a single pass at minTh would.
#pragma unroll 1
  for (int attempt = 0; attempt < 2; attempt++) {
  const int thr = attempt ? p.minTh : p.iniTh;
  // pass 1, four pixels per thread (one staged word): quick reject at thr -- every 9-arc holds one pixel of each
  // opposite ring pair, so a corner needs (p0 | p8) & (p4 | p12) darker, or brighter, than the centre by more than
  // thr. Byte-SIMD compares; saturating +-thr gives the same predicate as the integer test. Survivors are
  // compacted into a list so that the expensive score runs on dense warps.
  {
    const uint32_t* W = reinterpret_cast<const uint32_t*>(sImg);
    const int k0 = (ax + 3) >> 2, k1 = (ax + 3 + iw - 1) >> 2, nW = k1 - k0 + 1;
    const unsigned mNW = c_recip[nW];
    const unsigned th4 = (unsigned)thr * 0x01010101u;
    const int items = ih * nW;
    for (int i0 = 0; i0 < items; i0 += FAST_THREADS) {
      const int i = i0 + tid;
      unsigned pass = 0;
      int y = 0, xb = 0;
      if (i < items) {
        y = ft_div_small(i, nW, mNW);
        const int k = k0 + (i - y * nW);
        const uint32_t* r = W + (y + 3) * S + k;
        const unsigned c = r[0], prev = r[k > 0 ? -1 : 0], next = r[k + 1 < S ? 1 : 0];
        const unsigned t0 = r[3 * S], t8 = r[-3 * S];
        const unsigned t12 = __funnelshift_r(prev, c, 8), t4 = __funnelshift_r(c, next, 24);   // columns -3 / +3
        const unsigned lo = __vsubus4(c, th4), hi = __vaddus4(c, th4);
        const unsigned dark = (__vcmpltu4(t0, lo) | __vcmpltu4(t8, lo)) & (__vcmpltu4(t4, lo) | __vcmpltu4(t12, lo));
        const unsigned bright = (__vcmpgtu4(t0, hi) | __vcmpgtu4(t8, hi)) & (__vcmpgtu4(t4, hi) | __vcmpgtu4(t12, hi));
        pass = dark | bright;
        xb = 4 * k - ax - 3;                       // interior x of byte 0 of this word
        // bytes outside the interior columns do not count
#pragma unroll
        for (int j = 0; j < 4; j++) if (xb + j < 0 || xb + j >= iw) pass &= ~(0xFFu << (8 * j));
      }
      const int cnt = __popc(pass) >> 3;
      // warp-aggregated append: exclusive prefix of cnt over the lanes
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += n;
      }
      const int wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (wtot) {
        int base = 0;
        if ((tid & 31) == 31) base = atomicAdd(&sPass, wtot);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        int w = base + incl - cnt;
#pragma unroll
        for (int j = 0; j < 4; j++) if (pass & (0xFFu << (8 * j))) sList[w++] = (uint16_t)((y << 8) | (xb + j));
      }
    }
  }
  __syncthreads();
  // pass 2, survivors only: the exact score (order of the list is irrelevant)
  const int nPass = sPass;
  for (int j = tid; j < nPass; j += FAST_THREADS) {
    const int e = sList[j];
    const int y = e >> 8, x = e & 0xFF;
    int sc = ft_fast_score(&sImg[(y + 3) * rwPad + ax + (x + 3)], rwPad);
    if (sc < thr) sc = 0;
    sSc[(y + 1) * sS + x + 4] = (uint8_t)sc;
  }
  __syncthreads();
  // 3x3 strict-maximum NMS inside the interior, four pixels per thread; the zero border stands for "no corner"
  bool any20 = false;
  {
    const uint32_t* P = reinterpret_cast<const uint32_t*>(sSc);
    uint32_t* Q = reinterpret_cast<uint32_t*>(sMx);
    const int sW4 = sS >> 2;
    const int nW = ((3 + iw) >> 2);               // words 1 .. nW hold the interior columns 4 .. 3 + iw
    const unsigned mNW = c_recip[nW];
    const unsigned ini4 = (unsigned)p.iniTh * 0x01010101u;
    const int items = ih * nW;
    for (int i = tid; i < items; i += FAST_THREADS) {
      const int y = ft_div_small(i, nW, mNW);
      const int k = 1 + (i - y * nW);
      const uint32_t* r = P + (y + 1) * sW4 + k;
      const unsigned c = r[0];
      unsigned keep = 0;
      if (c) {
        unsigned m = __vmaxu4(__funnelshift_r(r[-1], c, 24), __funnelshift_r(c, r[1], 8));
        const uint32_t* u = r - sW4;
        const uint32_t* d = r + sW4;
        const unsigned uc = u[0], dc = d[0];
        m = __vmaxu4(m, __vmaxu4(uc, dc));
        m = __vmaxu4(m, __vmaxu4(__funnelshift_r(u[-1], uc, 24), __funnelshift_r(uc, u[1], 8)));
        m = __vmaxu4(m, __vmaxu4(__funnelshift_r(d[-1], dc, 24), __funnelshift_r(dc, d[1], 8)));
        keep = c & __vcmpgtu4(c, m);
        any20 |= __vcmpgeu4(keep, ini4) != 0;
      }
      Q[(y + 1) * sW4 + k] = keep;
    }
  }
  if (any20) sAny = 1;
  __syncthreads();
  if (sAny || attempt) break;                 // block-uniform
  // nothing survived at iniTh: again at minTh. Scores written above belong to pixels that pass at minTh too and are
  // rewritten with the same values; the NMS plane is rewritten completely.
  if (tid == 0) sPass = 0;
  __syncthreads();
  }
  const int th = sAny ? p.iniTh : p.minTh;
  // ordered compaction, four pixels at a time: the items are the words of the NMS plane in row-major order, thread t owns
  // the contiguous item range [t*per, (t+1)*per); a byte-SIMD compare marks the survivors of a word
  const int nWc = (3 + iw) >> 2;                  // words 1 .. nWc of a plane row hold the interior columns
  const unsigned mNWc = c_recip[nWc];
  const int itemsC = ih * nWc;
  const int per = (itemsC + FAST_THREADS - 1) / FAST_THREADS;
  const int ibeg = min(tid * per, itemsC), iend = min(ibeg + per, itemsC);
  const unsigned th4c = (unsigned)th * 0x01010101u;     // th >= 1: a survivor is a byte >= th
  const uint32_t* Pm = reinterpret_cast<const uint32_t*>(sMx);
  const int sW4c = sS >> 2;
  int cnt = 0;
  for (int i = ibeg; i < iend; i++) {
    const int y = ft_div_small(i, nWc, mNWc), k = 1 + (i - y * nWc);
    cnt += __popc(__vcmpgeu4(Pm[(y + 1) * sW4c + k], th4c)) >> 3;
  }
  // block exclusive scan of cnt
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if ((tid & 31) >= o) incl += n;
  }
  if ((tid & 31) == 31) sWarp[tid >> 5] = incl;
  __syncthreads();
  int base = 0, totalKp = 0;
#pragma unroll
  for (int w = 0; w < FAST_THREADS / 32; w++) { const int v = sWarp[w]; if (w < (tid >> 5)) base += v; totalKp += v; }
  int pos = base + incl - cnt;
  uint32_t* out = E.cellKp + L.cellKpBase + (size_t)(cell - L.cellBase) * L.cellCap;
  if (cnt) {
    for (int i = ibeg; i < iend; i++) {
      const int y = ft_div_small(i, nWc, mNWc), k = 1 + (i - y * nWc);
      const unsigned w = Pm[(y + 1) * sW4c + k];
      unsigned m = __vcmpgeu4(w, th4c);
      while (m) {
        const int j = (__ffs(m) - 1) >> 3;
        m &= ~(0xFFu << (8 * j));
        if (pos < L.cellCap) out[pos] = ft_pack_xys(4 * (k - 1) + j + 3 + cj * L.wCell, y + 3 + ci * L.hCell, (w >> (8 * j)) & 0xFFu);
        pos++;
      }
    }
  }
  if (L.octDenseDepth > 0) {
    // the octree's dense path (ft_octree.cu): count and best (response, canonical position) of the quadtree cell of depth
    // octDenseDepth that holds each corner -- one corner per lane, read back from the slab this CTA just wrote
    __syncthreads();
    const int nOut = min(totalKp, L.cellCap);
    for (int j = tid; j < nOut; j += FAST_THREADS) {
      const uint32_t w = out[j];
      const uint32_t q = (b.octTabX[L.octTabX + ft_px(w)] | b.octTabY[L.octTabY + ft_py(w)]) ^ FT_OCT_EVEN_MASK;
      const int di = L.octDenseBase + (int)(q >> (2 * (FT_OCT_D - L.octDenseDepth)));
      atomicAdd(&E.octCnt[di], 1);
      atomicMax(&E.octBest[di], ((unsigned)ft_ps(w) << 20) | (0xFFFFFu - (((unsigned)(cell - L.cellBase) << 9) | (unsigned)j)));
    }
  }
  if (tid == 0) {
    E.cellCount[cell] = min(totalKp, L.cellCap);
    if (totalKp > L.cellCap) atomicOr(b.status, FT_ST_CELL_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------
// Orientation + descriptor + final ordering. One warp per kept keypoint.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float ft_fast_atan2(float y, float x) {
  // cv::fastAtan2 scalar path; every operation individually rounded (no FMA contraction)
  const float sc = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc, p5 = 0.1555786518463281f * sc,
              p7 = -0.04432655554792128f * sc;
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = (float)2.2204460492503131e-16;
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

#define OCT_THREADS 512      // k_debug_sort (the octree kernel itself lives in ft_octree.cu)
#define OD_WARPS 8
static_assert(OD_WARPS * 32 == 256, "k_orient_desc stages the 256-word rBRIEF pattern with one word per thread");
__global__ void __launch_bounds__(OD_WARPS * 32) k_orient_desc(const __grid_constant__ FtParams p,
                                                               const __grid_constant__ FtBuffers b) {
  __shared__ int sLvlOff[FT_MAX_LEVELS + 1], sLvlCnt[FT_MAX_LEVELS];
  __shared__ int sScan[OD_WARPS];
  __shared__ int sMonoBefore;   // mono (non-lapping) keypoints before this block's first keypoint
  __shared__ int sMonoTotal;
  __shared__ int sPat[256];     // rBRIEF pattern staged from constant memory (lanes read different rows)
  const int eye = blockIdx.y;
  const FtEye& E = b.eye[eye];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  FT_PDL_TRIGGER();     // the stereo kernel may be scheduled; it waits for this grid before it reads
  sPat[tid] = reinterpret_cast<const int*>(c_pattern)[tid];
  // per-level keypoint counts: one load per thread (all in flight at once), then the prefix by one thread
  if (tid < FT_MAX_LEVELS) sLvlCnt[tid] = tid < p.nlevels ? E.lvlKpCount[tid] : 0;
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int l = 0; l < p.nlevels; l++) { sLvlOff[l] = o; o += sLvlCnt[l]; }
    sLvlOff[p.nlevels] = o;
    sMonoBefore = 0; sMonoTotal = 0;
  }
  __syncthreads();
  const int total = sLvlOff[p.nlevels];
  const int lap0 = p.lap[eye][0], lap1 = p.lap[eye][1];
  const int first = blockIdx.x * OD_WARPS;   // first global keypoint order index of this block
  if (blockIdx.x == 0 && eye == 0 && tid < 2 && b.stereoStats) b.stereoStats[tid] = 0ull;   // counters of the stereo kernel that follows
  if (blockIdx.x == 0 && tid == 0) {
    E.counts[0] = min(total, p.maxKp);
    if (total > p.maxKp) atomicOr(b.status, FT_ST_KP_OVERFLOW);
  }
  if (first >= total || total > p.maxKp) {
    // block 0 still has to publish monoIndex when there are no keypoints
    if (blockIdx.x == 0 && tid == 0) E.counts[1] = 0;
    return;
  }
  // count non-lapping keypoints (a) before `first`, (b) in total. Keypoints sit 19 px inside every level, so a
  // lapping area left of x = 19 (pinhole rigs pass {0,0}) contains none of them and one spanning the image all.
  if (lap1 < FT_EDGE_THRESHOLD) {
    if (tid == 0) { sMonoBefore = first; sMonoTotal = total; }
  } else if (lap0 <= FT_EDGE_THRESHOLD && lap1 >= p.width) {
    // all lapping: nothing to count
  } else {
    int before = 0, all = 0;
    for (int g = tid; g < total; g += OD_WARPS * 32) {
      int l = 0;
      while (g >= sLvlOff[l + 1]) l++;
      const uint32_t pk = E.lvlKp[p.lv[l].lvlKpBase + (g - sLvlOff[l])];
      float x = (float)ft_px(pk);
      if (l != 0) x = __fmul_rn(x, p.scale[l]);
      const bool inLap = (x >= (float)lap0) && (x <= (float)lap1);
      if (!inLap) { all++; if (g < first) before++; }
    }
    atomicAdd(&sMonoBefore, before);
    atomicAdd(&sMonoTotal, all);
  }
  __syncthreads();
  const int g = first + warp;
  // in-block order: lapping flags of the (up to) OD_WARPS keypoints of this block
  int l = 0;
  bool valid = g < total;
  uint32_t pk = 0;
  if (valid) {
    while (g >= sLvlOff[l + 1]) l++;
    pk = E.lvlKp[p.lv[l].lvlKpBase + (g - sLvlOff[l])];
  }
  const int kx = ft_px(pk), ky = ft_py(pk);
  float fxs = (float)kx, fys = (float)ky;
  if (l != 0) { fxs = __fmul_rn(fxs, p.scale[l]); fys = __fmul_rn(fys, p.scale[l]); }
  const bool inLap = valid && (fxs >= (float)lap0) && (fxs <= (float)lap1);
  if (lane == 0) sScan[warp] = (valid && !inLap) ? 1 : 0;
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) E.counts[1] = sMonoTotal;
  if (!valid) return;
  int monoBefore = sMonoBefore;
  for (int w = 0; w < warp; w++) monoBefore += sScan[w];
  // destination index (ORBextractor.cc:1408,1476-1486): lapping points fill from the end backwards
  const int dstIdx = inLap ? (total - 1 - (g - monoBefore)) : monoBefore;

  const FtLevel& L = p.lv[l];
  const uint8_t* img = E.pyr + L.offset;
  const uint8_t* center = img + (size_t)ky * L.pitch + kx;
  // IC_Angle (ORBextractor.cc:39-66): integer moments over the 749-px disc; lanes span u = -15..15
  int m01 = 0, m10 = 0;
  {
    const int u = lane - 15;
    if (lane < 31) {
#pragma unroll
      for (int v = -FT_HALF_PATCH; v <= FT_HALF_PATCH; v++) {
        const int d = p.umax[v < 0 ? -v : v];
        if (u >= -d && u <= d) {
          const int val = center[v * L.pitch + u];
          m10 += u * val;
          m01 += v * val;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
      m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
    }
  }
  const float angle = ft_fast_atan2((float)m01, (float)m10);
  // computeOrbDescriptor (ORBextractor.cc:68-108)
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  const float ang = __fmul_rn(angle, factorPI);
  float a, bb;
  ft_glibc_sincosf(ang, bb, a);
  const uint8_t* bimg = E.blur + L.offset;
  const uint8_t* bc = bimg + (size_t)ky * L.pitch + kx;
  int val = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int pw = sPat[lane * 8 + k];
    const int x0 = (int8_t)(pw & 0xFF), y0 = (int8_t)((pw >> 8) & 0xFF), x1 = (int8_t)((pw >> 16) & 0xFF), y1 = (int8_t)(pw >> 24);
    const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn((float)x0, bb), __fmul_rn((float)y0, a)));
    const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn((float)x0, a), __fmul_rn((float)y0, bb)));
    const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn((float)x1, bb), __fmul_rn((float)y1, a)));
    const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn((float)x1, a), __fmul_rn((float)y1, bb)));
    const int t0 = bc[ry0 * L.pitch + rx0], t1 = bc[ry1 * L.pitch + rx1];
    val |= (t0 < t1) << k;
  }
  E.desc[(size_t)dstIdx * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    ft_keypoint kp;
    kp.x = fxs; kp.y = fys;
    kp.size = (float)(int)__fmul_rn((float)FT_PATCH, p.scale[l]);
    kp.angle = angle;
    kp.response = (float)ft_ps(pk);
    kp.octave = l;
    E.kps[dstIdx] = kp;
  }
}

// Test hook: the device sort on its own (tests/test_gpu_sort.py fuzzes it against std::sort).
__global__ void __launch_bounds__(OCT_THREADS) k_debug_sort(unsigned long long* data, int n) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  unsigned long long* a = (unsigned long long*)smemRaw;
  unsigned long long* o = a + n;
  int* posA = (int*)(o + n);
  int* posB = posA + n;
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += OCT_THREADS) a[i] = data[i];
  __syncthreads();
  const long long t0 = clock64();
  ftsort::cta_introsort_loop(a, n, posA, posB, posB + n);
  __syncthreads();
  const long long t1 = clock64();
  ftsort::stable_rank(a, o, n, tid, OCT_THREADS);
  __syncthreads();
  const long long t2 = clock64();
  for (int i = tid; i < n; i += OCT_THREADS) data[i] = o[i];
  if (tid == 0) { data[n] = (unsigned long long)(t1 - t0); data[n + 1] = (unsigned long long)(t2 - t1); }
}

extern "C" int ft_debug_sort(unsigned long long* keys_inout, int n) {
  if (!keys_inout || n < 0 || n > 4096) return FT_ERR_INVALID;
  if (n == 0) return FT_OK;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, sizeof(unsigned long long) * (n + 2)) != cudaSuccess) return FT_ERR_CUDA;
  cudaMemcpy(d, keys_inout, sizeof(unsigned long long) * n, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)n * 24 + 6 * (n / 16 + 2) * 4;
  cudaFuncSetAttribute(k_debug_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_debug_sort<<<1, OCT_THREADS, smem>>>(d, n);
  cudaError_t e = cudaMemcpy(keys_inout, d, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost);
  if (getenv("FT_SORT_CLOCK")) { unsigned long long t[2]; cudaMemcpy(t, d + n, 16, cudaMemcpyDeviceToHost); fprintf(stderr, "ft_debug_sort n=%d loop=%llu rank=%llu cycles\n", n, t[0], t[1]); }
  cudaFree(d);
  return e == cudaSuccess ? FT_OK : FT_ERR_CUDA;
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
size_t ft_fast_smem_bytes(const FtParams& p) {
  size_t mx = 0;
  for (int l = 0; l < p.nlevels; l++) {
    const FtLevel& L = p.lv[l];
    const int rw = L.wCell + 6, rh = L.hCell + 6;
    const int rwPad = (rw + 6) & ~3;
    const int sS = (((L.wCell + 3) >> 2) + 3) * 4;
    const size_t plane = ((size_t)(L.hCell + 2) * sS + 15) & ~(size_t)15;
    size_t s = ((rh * rwPad + 15) & ~15) + 2 * plane + 2 * ((L.wCell * L.hCell + 15) & ~15);
    if (s > mx) mx = s;
  }
  return mx;
}
cudaError_t ft_launch_extract_setup(const FtParams& p) {
  {
    unsigned recip[FT_RECIP_N];
    recip[0] = 0; recip[1] = 0;
    for (unsigned n = 2; n < FT_RECIP_N; n++) recip[n] = 0xFFFFFFFFu / n + 1u;
    cudaError_t e0 = cudaMemcpyToSymbol(c_recip, recip, sizeof(recip));
    if (e0 != cudaSuccess) return e0;
  }
  cudaError_t e = ft_launch_octree_setup();
  if (e != cudaSuccess) return e;
  return ft_set_max_dynamic_smem((const void*)k_fast_cells);
}

void ft_launch_remap(const FtParams& p, const FtBuffers& b, const uint8_t* rawL, const uint8_t* rawR, const int2* tab, int rawW,
                     int rawH, cudaStream_t st) {
  dim3 blk(64, 4), grd((p.lv[0].w + 63) / 64, (p.lv[0].h + 3) / 4, p.nEyes);
  k_remap<<<grd, blk, 0, st>>>(p, b, rawL, rawR, tab, rawW, rawH);
}
void ft_launch_resize_input(const FtParams& p, const FtBuffers& b, const uint8_t* rawL, const uint8_t* rawR, int rawW,
                            cudaStream_t st) {
  dim3 blk(32, 8), grd((p.lv[0].w + 127) / 128, (p.lv[0].h + 7) / 8, p.nEyes);
  k_resize_input<<<grd, blk, 0, st>>>(p, b, rawL, rawR, rawW);
}
void ft_launch_resize(const FtParams& p, const FtBuffers& b, int level, cudaStream_t st) {
  dim3 blk(32, 8), grd((p.lv[level].w + 127) / 128, (p.lv[level].h + 7) / 8, p.nEyes);
  if (level >= 2) ft_launch_pdl(k_resize, grd, blk, 0, st, p, b, level);   // predecessor in the stream: resize of level - 1
  else k_resize<<<grd, blk, 0, st>>>(p, b, level);
}
void ft_launch_blur(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  const int tiles = (l1 < p.nlevels ? p.lv[l1].blurTileBase : p.totalBlurTiles) - p.lv[l0].blurTileBase;
  k_blur<<<dim3(tiles, p.nEyes), 256, 0, st>>>(p, b, l0, l1);
}
void ft_launch_fast(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  const int cells = (l1 < p.nlevels ? p.lv[l1].cellBase : p.totalCells) - p.lv[l0].cellBase;
  k_fast_cells<<<dim3(cells, p.nEyes), FAST_THREADS, ft_fast_smem_bytes(p), st>>>(p, b, l0, l1);
}
void ft_launch_orient_desc(const FtParams& p, const FtBuffers& b, cudaStream_t st) {
  k_orient_desc<<<dim3((p.maxKp + OD_WARPS - 1) / OD_WARPS, p.nEyes), OD_WARPS * 32, 0, st>>>(p, b);
}
