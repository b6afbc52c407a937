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

#define FAST_THREADS 256
__global__ void __launch_bounds__(FAST_THREADS, 6) k_fast_cells(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                    int levelBegin, int levelEnd) {
  extern __shared__ uint8_t smem[];
  __shared__ int sWarp[FAST_THREADS / 32];
  __shared__ int sAny, sPass;
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
  // ordered compaction: thread t owns the contiguous pixel range [t*seg, (t+1)*seg) in row-major order
  const int seg = (total + FAST_THREADS - 1) / FAST_THREADS;
  const int beg = min(tid * seg, total), end = min(beg + seg, total);
  int cnt = 0;
  {
    int y = ft_div_small(beg, iw, mIw), x = beg - y * iw;
    for (int i = beg; i < end; i++) {
      const int v = sMx[(y + 1) * sS + x + 4];
      cnt += v >= th && v > 0;
      if (++x == iw) { x = 0; y++; }
    }
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
  int base = 0;
  for (int w = 0; w < (tid >> 5); w++) base += sWarp[w];
  int totalKp = 0;
  for (int w = 0; w < FAST_THREADS / 32; w++) totalKp += sWarp[w];
  int pos = base + incl - cnt;
  uint32_t* out = E.cellKp + L.cellKpBase + (size_t)(cell - L.cellBase) * L.cellCap;
  if (cnt) {
    int y = ft_div_small(beg, iw, mIw), x = beg - y * iw;
    for (int i = beg; i < end; i++) {
      const int v = sMx[(y + 1) * sS + x + 4];
      if (v >= th && v > 0) {
        if (pos < L.cellCap) out[pos] = ft_pack_xys(x + 3 + cj * L.wCell, y + 3 + ci * L.hCell, v);
        pos++;
      }
      if (++x == iw) { x = 0; y++; }
    }
  }
  if (tid == 0) {
    E.cellCount[cell] = min(totalKp, L.cellCap);
    if (totalKp > L.cellCap) atomicOr(b.status, FT_ST_CELL_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------
// Octree distribution. One CTA per (eye, level). The std::list of nodes becomes an array in
// list order that is rebuilt every pass; a node's keypoints are not moved: every candidate
// carries a 16-bit code (node slot * 4 + quadrant) that is remapped through a table.
// ------------------------------------------------------------------------------------
#define OCT_THREADS 512
#ifdef FT_OCT_CLOCK
#define OCT_TICK(slot) do { if (tid == 0 && E.octClock && (slot) < 64) E.octClock[(eye * FT_MAX_LEVELS + level) * 64 + (slot)] = clock64(); } while (0)
#else
#define OCT_TICK(slot) do {} while (0)
#endif
#define OCT_SMEM_CANDS 16384   // candidates of one level held in shared memory (6 B each); larger levels use HBM scratch

struct OctSmem {
  // carved from dynamic shared memory, all arrays sized nodeCap
  short4* bnd[2];      // ULx, ULy, BRx, BRy  (ping-pong)
  int* cnt[2];         // keypoints per node
  int* ccnt;           // [cap][4] child counts of the pass
  int* posA;           // scan scratch
  int* posB;
  int* posC;
  int* map;            // [cap*4] code -> node index in the current list
  unsigned long long* vec[2];  // expandable nodes (key<<32 | node index), ping-pong
  int* vecPos;         // node index -> position in processing order, -1 when not in vec
};

// exclusive scans of three per-node quantities by warp 0; returns totals through smem
__device__ void oct_scan3(int n, const int* a, const int* b3, const int* c3, int* oa, int* ob, int* oc, int* totals,
                          bool reverseA) {
  // reverseA: oa[i] = sum of a[j] for j > i (children of later nodes go in front of earlier ones)
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int ra = 0, rb = 0, rc = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const int ia = reverseA ? (n - 1 - i) : i;
      int va = (i < n) ? a[ia] : 0, vb = (i < n) ? b3[i] : 0, vc = (i < n) ? c3[i] : 0;
      int sa = va, sb = vb, sc = vc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ta = __shfl_up_sync(0xFFFFFFFFu, sa, o), tb = __shfl_up_sync(0xFFFFFFFFu, sb, o),
                  tc = __shfl_up_sync(0xFFFFFFFFu, sc, o);
        if (lane >= o) { sa += ta; sb += tb; sc += tc; }
      }
      if (i < n) { oa[ia] = ra + sa - va; ob[i] = rb + sb - vb; oc[i] = rc + sc - vc; }
      ra += __shfl_sync(0xFFFFFFFFu, sa, 31); rb += __shfl_sync(0xFFFFFFFFu, sb, 31); rc += __shfl_sync(0xFFFFFFFFu, sc, 31);
    }
    if (lane == 0) { totals[0] = ra; totals[1] = rb; totals[2] = rc; }
  }
}

// Counter increments of the candidate loops. While a pass has at most OCT_PRIV counters (the first passes funnel
// thousands of candidates into a handful of counters and same-address shared atomics serialise) every warp
// counts into its own private copy; oct_count_flush folds the copies into the real counters.
#define OCT_PRIV 64
__device__ __forceinline__ void oct_count(int* counters, int* priv, bool usePriv, int key) {
  if (usePriv) atomicAdd(&priv[(threadIdx.x >> 5) * OCT_PRIV + key], 1);
  else atomicAdd(&counters[key], 1);
}
__device__ __forceinline__ void oct_priv_clear(int* priv) {
  for (int i = threadIdx.x; i < (OCT_THREADS / 32) * OCT_PRIV; i += OCT_THREADS) priv[i] = 0;
}
__device__ __forceinline__ void oct_count_flush(int* counters, const int* priv, int nCounters) {
  for (int k = threadIdx.x; k < nCounters; k += OCT_THREADS) {
    int sum = 0;
#pragma unroll
    for (int w = 0; w < OCT_THREADS / 32; w++) sum += priv[w * OCT_PRIV + k];
    counters[k] += sum;
  }
}

__device__ __forceinline__ void oct_child_bounds(short4 bd, int q, short4& out) {
  // DivideNode (ORBextractor.cc:510-538): halfX = ceil((UR.x-UL.x)/2.f), halfY = ceil((BR.y-UL.y)/2.f)
  const int halfX = (bd.z - bd.x + 1) >> 1, halfY = (bd.w - bd.y + 1) >> 1;
  const int mx = bd.x + halfX, my = bd.y + halfY;
  out.x = (q & 1) ? mx : bd.x;
  out.z = (q & 1) ? bd.z : mx;
  out.y = (q & 2) ? my : bd.y;
  out.w = (q & 2) ? bd.w : my;
}

__global__ void __launch_bounds__(OCT_THREADS) k_octree(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                        int levelBegin) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  __shared__ int sTot[4];
  __shared__ int sN, sMode, sVecN, sP, sCellTot;
  __shared__ int sScan[OCT_THREADS / 32];
  __shared__ int sChildP, sGrowP, sVecP;
  __shared__ int sPriv[(OCT_THREADS / 32) * OCT_PRIV];
  const int level = levelBegin + blockIdx.x;
  const int eye = blockIdx.y;
  const FtLevel& L = p.lv[level];
  const FtEye& E = b.eye[eye];
  const int tid = threadIdx.x;
  const int cap = L.nodeCap;
  const int N = L.quota;
  const int nCells = L.nCols * L.nRows;

  OctSmem S;
  int* cellOff;          // [nCells + 1] exclusive prefix of the per-cell candidate counts
  uint32_t* candS;       // [OCT_SMEM_CANDS]
  uint16_t* codeS;       // [OCT_SMEM_CANDS]
  {
    uint8_t* q = smemRaw;
    S.vec[0] = (unsigned long long*)q; q += sizeof(unsigned long long) * cap;
    S.vec[1] = (unsigned long long*)q; q += sizeof(unsigned long long) * cap;
    S.bnd[0] = (short4*)q; q += sizeof(short4) * cap;
    S.bnd[1] = (short4*)q; q += sizeof(short4) * cap;
    S.cnt[0] = (int*)q; q += 4 * cap;
    S.cnt[1] = (int*)q; q += 4 * cap;
    S.ccnt = (int*)q; q += 16 * cap;
    S.posA = (int*)q; q += 4 * cap;
    S.posB = (int*)q; q += 4 * cap;
    S.posC = (int*)q; q += 4 * cap;
    S.map = (int*)q; q += 16 * cap;
    S.vecPos = (int*)q; q += 4 * cap;
    cellOff = (int*)q; q += 4 * (nCells + 1);
    q = (uint8_t*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    candS = (uint32_t*)q; q += 4 * OCT_SMEM_CANDS;
    codeS = (uint16_t*)q;
  }

  OCT_TICK(0);
  // ---- flat canonical order (cell row-major, then row-major inside the cell) ----
  // exclusive scan of the per-cell counts, then every candidate finds its cell by binary search: the copy out of
  // the per-cell slabs is one coalesced pass instead of a per-cell serial loop.
  {
    int running = 0;
    for (int c0 = 0; c0 < nCells; c0 += OCT_THREADS) {
      const int c = c0 + tid;
      const int cnt = c < nCells ? E.cellCount[L.cellBase + c] : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += n;
      }
      if ((tid & 31) == 31) sScan[tid >> 5] = incl;
      __syncthreads();
      int off = running;
      for (int w = 0; w < (tid >> 5); w++) off += sScan[w];
      if (c < nCells) cellOff[c] = off + incl - cnt;
      int tot = 0;
      for (int w = 0; w < OCT_THREADS / 32; w++) tot += sScan[w];
      running += tot;
      __syncthreads();
    }
    if (tid == 0) {
      int C = running;
      cellOff[nCells] = C;
      if (C > L.candCap) { C = L.candCap; atomicOr(b.status, FT_ST_CAND_OVERFLOW); }
      sCellTot = C;
      E.lvlCandCount[level] = C;
    }
    __syncthreads();
  }
  const int C = sCellTot;
  OCT_TICK(1);
  uint32_t* outKp = E.lvlKp + L.lvlKpBase;
  if (C == 0) {
    if (tid == 0) E.lvlKpCount[level] = 0;
    return;
  }
  uint32_t* candG = E.cand + L.candBase;          // global copy: read back by ft_debug_level_candidates
  const bool inSmem = C <= OCT_SMEM_CANDS;
  uint32_t* cand = inSmem ? candS : candG;
  uint16_t* code = inSmem ? codeS : (E.candNode + L.candBase);
  {
    // one warp per cell, four cells in flight per warp so the slab reads overlap
    const int lane = tid & 31, warp = tid >> 5, nWarps = OCT_THREADS / 32;
    for (int cell0 = warp * 4; cell0 < nCells; cell0 += nWarps * 4) {
      uint32_t v[4]; int off[4], cnt[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int cell = cell0 + u;
        off[u] = 0; cnt[u] = 0; v[u] = 0;
        if (cell < nCells) {
          off[u] = cellOff[cell]; cnt[u] = min(cellOff[cell + 1], C) - off[u];
          if (lane < cnt[u]) v[u] = E.cellKp[L.cellKpBase + (size_t)cell * L.cellCap + lane];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (lane < cnt[u]) { candG[off[u] + lane] = v[u]; if (inSmem) candS[off[u] + lane] = v[u]; }
        for (int k = 32 + lane; k < cnt[u]; k += 32) {   // rare: more than 32 survivors in one cell
          const uint32_t w = E.cellKp[L.cellKpBase + (size_t)(cell0 + u) * L.cellCap + k];
          candG[off[u] + k] = w; if (inSmem) candS[off[u] + k] = w;
        }
      }
    }
  }

  OCT_TICK(2);
  // ---- roots (ORBextractor.cc:664-706) ----
  int cur = 0;  // ping-pong index of the current list
  const int nIni = L.nIni;
  const float hX = L.hX;
  const int H = L.maxBorderY - FT_MIN_BORDER;
  for (int i = tid; i < nIni; i += OCT_THREADS) {
    short4 bd;
    bd.x = (short)(int)__fmul_rn(hX, (float)i);
    bd.z = (short)(int)__fmul_rn(hX, (float)(i + 1));
    bd.y = 0; bd.w = (short)H;
    S.bnd[0][i] = bd;
    S.cnt[0][i] = 0;
  }
  const bool privRoots = nIni <= OCT_PRIV;
  if (privRoots) oct_priv_clear(sPriv);
  __syncthreads();
  for (int c = tid; c < C; c += OCT_THREADS) {
    const uint32_t pk = cand[c];
    int r = (int)__fdiv_rn((float)ft_px(pk), hX);
    if (r >= nIni) r = nIni - 1;
    code[c] = (uint16_t)(r * 4);
    oct_count(S.cnt[0], sPriv, privRoots, r);
  }
  __syncthreads();
  if (privRoots) { oct_count_flush(S.cnt[0], sPriv, nIni); __syncthreads(); }
  // drop empty roots
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < nIni; i++) {
      if (S.cnt[0][i] > 0) {
        S.bnd[1][n] = S.bnd[0][i]; S.cnt[1][n] = S.cnt[0][i];
        for (int q = 0; q < 4; q++) S.map[i * 4 + q] = n;
        n++;
      } else {
        for (int q = 0; q < 4; q++) S.map[i * 4 + q] = -1;
      }
    }
    sN = n; sMode = 0; sVecN = 0;
  }
  cur = 1;
  __syncthreads();

  OCT_TICK(3);
  int tick = 4;
  // ---- main loop ----
  // sMode: 0 = normal pass, 1 = careful pass (largest-first with early break), 2 = finished
  int vcur = 0;  // ping-pong of the expandable-node vector
  for (int iter = 0; iter < 64; iter++) {
    const int n = sN;
    const int mode = sMode;
    if (mode == 2) break;
    OCT_TICK(tick); tick++;
    if (tid == 0 && E.octClock && tick < 60) E.octClock[(eye * FT_MAX_LEVELS + level) * 64 + 40 + (tick - 5)] = mode * 100000 + n;
    short4* bnd = S.bnd[cur]; int* cnt = S.cnt[cur];
    short4* bnd2 = S.bnd[cur ^ 1]; int* cnt2 = S.cnt[cur ^ 1];
    const int m = sVecN;                 // size of the vector entering a careful pass
    unsigned long long* vecPrev = S.vec[vcur];
    unsigned long long* vecNew = S.vec[vcur ^ 1];

    const bool usePriv = 4 * n <= OCT_PRIV;
    if (usePriv) oct_priv_clear(sPriv);
    // which nodes are split (speculatively, in careful mode) this pass
    for (int i = tid; i < n; i += OCT_THREADS) {
      S.vecPos[i] = -1;
      S.ccnt[4 * i] = 0; S.ccnt[4 * i + 1] = 0; S.ccnt[4 * i + 2] = 0; S.ccnt[4 * i + 3] = 0;
    }
    if (mode == 1) {
      // std::sort(..., compareNodes) (ORBextractor.cc:805): warp 0 replays libstdc++'s introsort loop, then the
      // final insertion sort (= stable sort of what the loop leaves) is a parallel rank computation
      OCT_TICK(20);
      __syncthreads();   // the loop above initialises vecPos/ccnt; posC doubles as the range queue of the sort
      ftsort::cta_introsort_loop(vecPrev, m, S.posA, S.posB, S.posC);
      __syncthreads();
      OCT_TICK(21);
      ftsort::stable_rank(vecPrev, vecNew, m, tid, OCT_THREADS);
      __syncthreads();
      // processing order r = 0..m-1 walks the sorted vector from the back (:806)
      for (int r = tid; r < m; r += OCT_THREADS) {
        const unsigned long long e = vecNew[m - 1 - r];
        vecPrev[m - 1 - r] = e;
        S.vecPos[(int)(e & 0xFFFFFFFFu)] = r;
      }
    }
    __syncthreads();
    if (mode == 1) OCT_TICK(22); else OCT_TICK(27);
    // candidates: remap code -> node, count children of nodes being split
    for (int c = tid; c < C; c += OCT_THREADS) {
      const int node = S.map[code[c]];
      const bool split = (mode == 0) ? (cnt[node] > 1) : (S.vecPos[node] >= 0);
      int q = 0;
      if (split) {
        const uint32_t pk = cand[c];
        const short4 bd = bnd[node];
        const int halfX = (bd.z - bd.x + 1) >> 1, halfY = (bd.w - bd.y + 1) >> 1;
        q = (ft_px(pk) < bd.x + halfX ? 0 : 1) | (ft_py(pk) < bd.y + halfY ? 0 : 2);
        oct_count(S.ccnt, sPriv, usePriv, 4 * node + q);
      }
      code[c] = (uint16_t)(node * 4 + q);
    }
    __syncthreads();
    if (usePriv) { oct_count_flush(S.ccnt, sPriv, 4 * n); __syncthreads(); }

    if (mode == 1) OCT_TICK(23); else OCT_TICK(28);
    if (mode == 0) {
      // per node: k = non-empty children, e = children with more than one keypoint
      for (int i = tid; i < n; i += OCT_THREADS) {
        int k = 0, e = 0, nm = 0;
        if (cnt[i] > 1) {
          for (int q = 0; q < 4; q++) { k += S.ccnt[4 * i + q] > 0; e += S.ccnt[4 * i + q] > 1; }
        } else nm = 1;
        S.posA[i] = k; S.posB[i] = nm; S.posC[i] = e;
      }
      __syncthreads();
      OCT_TICK(29);
      // posA <- children of later nodes (they end up in front), posB <- noMore nodes before i, posC <- vec offset
      oct_scan3(n, S.posA, S.posB, S.posC, S.posA, S.posB, S.posC, sTot, true);
      __syncthreads();
      OCT_TICK(30);
      const int totalChildren = sTot[0], nNew = sTot[0] + sTot[1], nToExpand = sTot[2];
      if (nNew > cap) {  // cannot happen (list <= N+3, roots*4); guarded so a logic slip cannot corrupt memory
        if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlKpCount[level] = 0; }
        return;
      }
      for (int i = tid; i < n; i += OCT_THREADS) {
        if (cnt[i] > 1) {
          const short4 bd = bnd[i];
          int after = 0;   // non-empty children with a higher quadrant index are pushed later => sit in front
          int vpos = S.posC[i];
          int pos[4];
          for (int q = 3; q >= 0; q--) { pos[q] = S.posA[i] + after; after += S.ccnt[4 * i + q] > 0; }
          for (int q = 0; q < 4; q++) {
            const int cc = S.ccnt[4 * i + q];
            if (cc > 0) {
              short4 cb; oct_child_bounds(bd, q, cb);
              bnd2[pos[q]] = cb; cnt2[pos[q]] = cc;
              S.map[4 * i + q] = pos[q];
              if (cc > 1) vecNew[vpos++] = ((unsigned long long)(((unsigned)cc << 12) | (unsigned)cb.x) << 32) | (unsigned)pos[q];
            } else S.map[4 * i + q] = -1;
          }
        } else {
          const int np = totalChildren + S.posB[i];
          bnd2[np] = bnd[i]; cnt2[np] = cnt[i];
          S.map[4 * i] = np; S.map[4 * i + 1] = np; S.map[4 * i + 2] = np; S.map[4 * i + 3] = np;
        }
      }
      __syncthreads();
      OCT_TICK(31);
      if (tid == 0) {
        sN = nNew; sVecN = nToExpand;
        if (nNew >= N || nNew == n) sMode = 2;                 // (:790)
        else if (nNew + nToExpand * 3 > N) sMode = 1;          // (:794)
      }
      cur ^= 1; vcur ^= 1;
      __syncthreads();
    } else {
      // careful pass: nodes of the sorted vector are split from the back until the list reaches N (:806-853)
      // posA[r] = non-empty children of the r-th processed node, posB[r] = growth (children - 1), posC[r] = expandable
      for (int r = tid; r < m; r += OCT_THREADS) {
        const int node = (int)(vecPrev[m - 1 - r] & 0xFFFFFFFFu);
        int k = 0, e = 0;
        for (int q = 0; q < 4; q++) { k += S.ccnt[4 * node + q] > 0; e += S.ccnt[4 * node + q] > 1; }
        S.posA[r] = k; S.posB[r] = k - 1; S.posC[r] = e;
      }
      if (tid == 0) sP = m;
      __syncthreads();
      // exclusive prefix over processing order: posA -> children before r, posB -> growth before r, posC -> vec offset
      oct_scan3(m, S.posA, S.posB, S.posC, S.posA, S.posB, S.posC, sTot, false);
      __syncthreads();
      // number processed P: first r with n + growthBefore(r) + growth(r) >= N, else m
      for (int r = tid; r < m; r += OCT_THREADS) {
        const int node = (int)(vecPrev[m - 1 - r] & 0xFFFFFFFFu);
        int k = 0;
        for (int q = 0; q < 4; q++) k += S.ccnt[4 * node + q] > 0;
        if (n + S.posB[r] + (k - 1) >= N) atomicMin(&sP, r + 1);
      }
      __syncthreads();
      const int P = sP;
      OCT_TICK(24);
      // totals restricted to the processed prefix
      if (tid == 0) {
        if (P == m) { sChildP = sTot[0]; sGrowP = sTot[1]; sVecP = sTot[2]; }
        else { sChildP = S.posA[P]; sGrowP = S.posB[P]; sVecP = S.posC[P]; }
      }
      __syncthreads();
      const int childP = sChildP, nNew = n + sGrowP, vecP = sVecP;
      if (nNew > cap) {
        if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlKpCount[level] = 0; }
        return;
      }
      // processed nodes: children go to the front, later-processed first
      for (int r = tid; r < P; r += OCT_THREADS) {
        const int node = (int)(vecPrev[m - 1 - r] & 0xFFFFFFFFu);
        const short4 bd = bnd[node];
        int k = 0;
        for (int q = 0; q < 4; q++) k += S.ccnt[4 * node + q] > 0;
        const int start = childP - S.posA[r] - k;   // children of nodes processed after r sit in front
        int after = 0, vpos = S.posC[r];
        int pos[4];
        for (int q = 3; q >= 0; q--) { pos[q] = start + after; after += S.ccnt[4 * node + q] > 0; }
        for (int q = 0; q < 4; q++) {
          const int cc = S.ccnt[4 * node + q];
          if (cc > 0) {
            short4 cb; oct_child_bounds(bd, q, cb);
            bnd2[pos[q]] = cb; cnt2[pos[q]] = cc;
            S.map[4 * node + q] = pos[q];
            if (cc > 1) vecNew[vpos++] = ((unsigned long long)(((unsigned)cc << 12) | (unsigned)cb.x) << 32) | (unsigned)pos[q];
          } else S.map[4 * node + q] = -1;
        }
      }
      __syncthreads();   // posA/posB of the processed prefix are consumed; they are reused below
      OCT_TICK(25);
      // surviving old nodes keep their relative order behind the new children
      for (int i = tid; i < n; i += OCT_THREADS) {
        const int r = S.vecPos[i];
        S.posB[i] = (r >= 0 && r < P) ? 0 : 1;
      }
      __syncthreads();
      if (tid < 32) {
        // exclusive scan of the survive flags into posA
        const int lane = tid;
        int run = 0;
        for (int base = 0; base < n; base += 32) {
          const int i = base + lane;
          const int v = i < n ? S.posB[i] : 0;
          int s = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
          if (i < n) S.posA[i] = run + s - v;
          run += __shfl_sync(0xFFFFFFFFu, s, 31);
        }
      }
      __syncthreads();
      for (int i = tid; i < n; i += OCT_THREADS) {
        if (S.posB[i]) {
          const int np = childP + S.posA[i];
          bnd2[np] = bnd[i]; cnt2[np] = cnt[i];
          S.map[4 * i] = np; S.map[4 * i + 1] = np; S.map[4 * i + 2] = np; S.map[4 * i + 3] = np;
        }
      }
      __syncthreads();
      OCT_TICK(26);
      if (tid == 0) {
        sN = nNew; sVecN = vecP; E.octClock ? (void)(E.octClock[(eye * FT_MAX_LEVELS + level) * 64 + 63] = m) : (void)0;
        if (nNew >= N || nNew == n) sMode = 2;   // (:855)
      }
      cur ^= 1; vcur ^= 1;
      __syncthreads();
    }
  }

  OCT_TICK(tick); tick++;
  // ---- best keypoint per node: highest response, first in input order wins ties (:862-881) ----
  const int n = sN;
  unsigned* best = (unsigned*)S.posA;
  for (int i = tid; i < n; i += OCT_THREADS) best[i] = 0;
  __syncthreads();
  for (int c = tid; c < C; c += OCT_THREADS) {
    const int node = S.map[code[c]];
    const unsigned key = ((unsigned)ft_ps(cand[c]) << 20) | (unsigned)(0xFFFFF - c);
    atomicMax(&best[node], key);
  }
  __syncthreads();
  if (n > L.lvlKpCap) {
    if (tid == 0) { atomicOr(b.status, FT_ST_KP_OVERFLOW); E.lvlKpCount[level] = 0; }
    return;
  }
  for (int i = tid; i < n; i += OCT_THREADS) {
    const int c = 0xFFFFF - (int)(best[i] & 0xFFFFFu);
    const uint32_t pk = cand[c];
    outKp[i] = ft_pack_xys(ft_px(pk) + FT_MIN_BORDER, ft_py(pk) + FT_MIN_BORDER, ft_ps(pk));
  }
  if (tid == 0) E.lvlKpCount[level] = n;
  OCT_TICK(tick);
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

#define OD_WARPS 8
__global__ void __launch_bounds__(OD_WARPS * 32) k_orient_desc(const __grid_constant__ FtParams p,
                                                               const __grid_constant__ FtBuffers b) {
  __shared__ int sLvlOff[FT_MAX_LEVELS + 1];
  __shared__ int sScan[OD_WARPS];
  __shared__ int sMonoBefore;   // mono (non-lapping) keypoints before this block's first keypoint
  __shared__ int sMonoTotal;
  __shared__ int sPat[256];     // rBRIEF pattern staged from constant memory (lanes read different rows)
  const int eye = blockIdx.y;
  const FtEye& E = b.eye[eye];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  sPat[tid] = reinterpret_cast<const int*>(c_pattern)[tid];
  if (tid == 0) {
    int o = 0;
    for (int l = 0; l < p.nlevels; l++) { sLvlOff[l] = o; o += E.lvlKpCount[l]; }
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
size_t ft_octree_smem_bytes(const FtParams& p, int level) {
  const FtLevel& L = p.lv[level];
  return (size_t)L.nodeCap * (8 + 8 + 8 + 8 + 4 + 4 + 16 + 4 + 4 + 4 + 16 + 4) + 4 * (size_t)(L.nCols * L.nRows + 1) + 16 +
         (size_t)OCT_SMEM_CANDS * 6 + 64;
}

cudaError_t ft_launch_extract_setup(const FtParams& p) {
  {
    unsigned recip[FT_RECIP_N];
    recip[0] = 0; recip[1] = 0;
    for (unsigned n = 2; n < FT_RECIP_N; n++) recip[n] = 0xFFFFFFFFu / n + 1u;
    cudaError_t e0 = cudaMemcpyToSymbol(c_recip, recip, sizeof(recip));
    if (e0 != cudaSuccess) return e0;
  }
  size_t mx = 0;
  for (int l = 0; l < p.nlevels; l++) mx = ft_octree_smem_bytes(p, l) > mx ? ft_octree_smem_bytes(p, l) : mx;
  cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mx);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_fast_smem_bytes(p));
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
  k_resize<<<grd, blk, 0, st>>>(p, b, level);
}
void ft_launch_blur(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  const int tiles = (l1 < p.nlevels ? p.lv[l1].blurTileBase : p.totalBlurTiles) - p.lv[l0].blurTileBase;
  k_blur<<<dim3(tiles, p.nEyes), 256, 0, st>>>(p, b, l0, l1);
}
void ft_launch_fast(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  const int cells = (l1 < p.nlevels ? p.lv[l1].cellBase : p.totalCells) - p.lv[l0].cellBase;
  k_fast_cells<<<dim3(cells, p.nEyes), FAST_THREADS, ft_fast_smem_bytes(p), st>>>(p, b, l0, l1);
}
void ft_launch_octree(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  size_t mx = 0;
  for (int l = l0; l < l1; l++) mx = ft_octree_smem_bytes(p, l) > mx ? ft_octree_smem_bytes(p, l) : mx;
  k_octree<<<dim3(l1 - l0, p.nEyes), OCT_THREADS, mx, st>>>(p, b, l0);
}
void ft_launch_orient_desc(const FtParams& p, const FtBuffers& b, cudaStream_t st) {
  k_orient_desc<<<dim3((p.maxKp + OD_WARPS - 1) / OD_WARPS, p.nEyes), OD_WARPS * 32, 0, st>>>(p, b);
}
