// ft_octree.cu -- ORBextractor::DistributeOctTree (reference src/ORBextractor.cc:660-884, DivideNode :510-566,
// compareNodes :626-641) as a sort-once "trie" kernel for sm_100a. One CTA per (eye, pyramid level).
//
// The reference grows a std::list of nodes pass by pass, moving keypoints from parent to child vectors. DivideNode is a
// fixed spatial quadtree (child bounds are ceil-halved parent bounds), so nothing a pass decides changes WHERE a keypoint
// goes, only how deep its node is split. That gives this formulation (tools/octree_trie_proto.py is the executable
// derivation, checked against the oracle's list-based restatement):
//
//  1. path key   every candidate gets key = (root, q1 .. q12), qd = quadrant at depth d (n1..n4 = 0..3). The x and y
//                halves of the descent are independent, so they come from two per-level tables built on the host
//                (already spread to even / odd bit positions): key = (tabX[x] | tabY[y]) ^ 0x333333.
//  2. list order push_front of n1..n4 while the list is walked front to back reverses the order at every depth:
//                order_k = (order_{k-1} of the parent descending, quadrant descending). With the digits of even depths
//                complemented (the xor above; call the result Q) the depth-k nodes sit in the list in ascending Q for
//                even k and in descending Q for odd k.
//  3. one sort   candidates are sorted by Q once (dense histogram over the leading digits = bins of bounded pixel area,
//                then an exact rank inside each bin). Every node of every depth is a contiguous range of that array.
//  4. normal passes in closed form: a pass splits EVERY node with more than one keypoint, so after pass k the list holds
//                the non-empty depth-k nodes created by that pass plus the single-keypoint nodes that settled earlier.
//                With cl(i) = number of leading digits sorted neighbours i, i+1 share, size_k = 1 + #{cl <= k} and
//                nToExpand_k = #{i : cl(i-1) <= k < cl(i)}: two 14-bin histograms replace all candidate sweeps, and the
//                pass K at which the reference stops (:790) or switches to its careful phase (:794) is scalar logic.
//  5. the list after pass K is written directly: depth-K nodes first, then the singles that settled at depth K-1, K-2,
//                .. 0, each group in its own direction (rank inside a class = ballot prefix over the sorted array).
//  6. careful phase (:794-858) on node records (range, depth): the (size, UL.x) sort with libstdc++'s tie order
//                (ft_sort.h), children = sub-ranges found by binary search on the next digit, early break by prefix
//                sums over the processing order, the vector of the next careful pass in creation order.
//  7. best keypoint per node (:862-881): highest response, first in input order on ties.
//
// Dense fast path. Steps 1-5 and 7 do not need the candidates at all when the decisions stay shallow, which is the
// normal case (thousands of candidates, a quota of a few hundred: the list reaches N around depth 4-5). k_fast_cells
// already holds every surviving corner in a register when it writes it out; it also adds it to two dense arrays of the
// level, indexed by the leading digits of Q down to depth Dd (cells of a few pixels): a count and the best
// (response, canonical position) key. The octree kernel then builds the count / best pyramid of depths Dd-1 .. 0 by
// summing four children at a time, reads size_k and nToExpand_k off it, compacts the flagged cells of depth K, K-1, .. 0
// into the list, runs the careful phase with child counts that are plain lookups, and reads each node's keypoint from
// the best pyramid: no per-candidate pass on its single SM. Whenever a decision would need a depth below Dd (few
// candidates, so the passes run deep; or a third careful pass) the kernel falls back to the general sort-based path,
// which reads the per-cell slabs and computes the same list.
#include <cstdio>
#include <cstdlib>

#include "ft_device.cuh"
#include "ft_sort.h"

#define OCT_D FT_OCT_D                 // quadrant digits in a path key (enough for 4096-px roots)
#define OCT_EVEN_MASK FT_OCT_EVEN_MASK  // digits of even depth (2, 4, .. 12), two bits each, depth 12 in bits 0-1
#define OCT_MAX_THREADS 1024
#define OCT_MAX_WARPS (OCT_MAX_THREADS / 32)
#define OCT_NCLS 16              // classes of the list construction: depth-K heads, depth-K ends, singles of depth 0..13

#ifdef FT_OCT_CLOCK
#define OCT_CTICK(slot) do { if (threadIdx.x == 0 && clk) clk[slot] = clock64(); } while (0)
#else
#define OCT_CTICK(slot) do { (void)clk; } while (0)
#endif

struct OctShared {
  int scanW[3][OCT_MAX_WARPS];
  int scanTot[3];
  int hist[OCT_MAX_WARPS][16];   // per-warp copies: neighbour pairs by common-digit count
  int diff[OCT_MAX_WARPS][16];   // per-warp copies: difference array of nToExpand over depth
  int clsTot[OCT_NCLS], clsBase[OCT_NCLS];
  int sizeD[16], nexpD[16];      // dense path: non-empty cells / cells with more than one keypoint per depth
  int C, K, mode, n, m, P, fallback;
};

// Exclusive scan in place of up to three int arrays (b, c may be null) by the whole block; totals returned in
// registers (uniform over the block). Contains __syncthreads; callers synchronise before (inputs written) and may
// read the outputs right after it returns.
__device__ __forceinline__ void oct_block_scan3(int n, int* a, int* b, int* c, OctShared& sh, int& ta, int& tb, int& tc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nW = T >> 5;
  int ca = 0, cb = 0, cc = 0;
  for (int base = 0; base < n; base += T) {
    const int i = base + tid;
    const int va = i < n ? a[i] : 0, vb = (b && i < n) ? b[i] : 0, vc = (c && i < n) ? c[i] : 0;
    int sa = va, sb = vb, sc = vc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int xa = __shfl_up_sync(0xFFFFFFFFu, sa, o), xb = __shfl_up_sync(0xFFFFFFFFu, sb, o), xc = __shfl_up_sync(0xFFFFFFFFu, sc, o);
      if (lane >= o) { sa += xa; sb += xb; sc += xc; }
    }
    if (lane == 31) { sh.scanW[0][warp] = sa; sh.scanW[1][warp] = sb; sh.scanW[2][warp] = sc; }
    __syncthreads();
    if (warp < 3) {
      const int v = lane < nW ? sh.scanW[warp][lane] : 0;
      int s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += x; }
      if (lane < nW) sh.scanW[warp][lane] = s - v;
      if (lane == 31) sh.scanTot[warp] = s;
    }
    __syncthreads();
    if (i < n) {
      a[i] = ca + sh.scanW[0][warp] + sa - va;
      if (b) b[i] = cb + sh.scanW[1][warp] + sb - vb;
      if (c) c[i] = cc + sh.scanW[2][warp] + sc - vc;
    }
    ca += sh.scanTot[0]; cb += sh.scanTot[1]; cc += sh.scanTot[2];
    if (base + T < n) __syncthreads();   // scanW / scanTot are rewritten by the next round
  }
  ta = ca; tb = cb; tc = cc;
}

// number of leading digits (root digit included) two path keys share: 0 .. OCT_D (equal keys do not occur: one pixel
// is one candidate; they would count as OCT_D)
__device__ __forceinline__ int oct_common(uint32_t a, uint32_t b) {
  const uint32_t x = a ^ b;
  if (x == 0) return OCT_D;
  const int hb = 31 - __clz(x);
  if (hb >= 2 * OCT_D) return 0;
  return OCT_D - (hb >> 1);
}

// quadrant digit of depth d (1..OCT_D) as stored in Q
__device__ __forceinline__ int oct_digit(uint32_t key, int d) { return (key >> (2 * (OCT_D - d))) & 3; }

// UL.x of the node that holds `key` at depth d (replays the x half of DivideNode from the root)
__device__ __forceinline__ int oct_node_ulx(uint32_t key, int d, float hX) {
  const int r = key >> (2 * OCT_D);
  int x0 = (int)__fmul_rn(hX, (float)r), x1 = (int)__fmul_rn(hX, (float)(r + 1));
  const uint32_t plain = key ^ OCT_EVEN_MASK;   // undo the complement: real quadrants
  for (int j = 1; j <= d; j++) {
    const int mx = x0 + ((x1 - x0 + 1) >> 1);
    if ((plain >> (2 * (OCT_D - j))) & 1) x0 = mx; else x1 = mx;
  }
  return x0;
}

// first index in [lo, hi) of the sorted keys whose digit of depth d is >= v (digits ascend inside a node's range)
__device__ __forceinline__ int oct_lower(const uint32_t* skey, int lo, int hi, int d, int v) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (oct_digit(skey[mid], d) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}


#define OCT_LO(w) ((int)((w) & 0xFFFFFu))
#define OCT_DEPTH(w) ((int)((w) >> 20))

// Arrays of the node list and of the careful phase, carved from dynamic shared memory (60 bytes per node slot).
struct OctNodes {
  unsigned long long* vec[2];    // expandable nodes: (size << 12 | UL.x) << 32 | list position
  uint32_t* lo[2];               // general path: range start | depth << 20;  dense path: cell index | depth << 20
  uint32_t* hi[2];               // general path: range end;                   dense path: keypoints in the cell
  int *posA, *posB, *posC, *vecPos, *split1, *split2, *split3;
};

// What distinguishes the two paths inside the careful phase: how a node's size, UL.x and child counts are found.
struct OctGeneralView {          // nodes are ranges of the Q-sorted candidate array
  const uint32_t* skey;
  __device__ __forceinline__ int size(uint32_t lo, uint32_t hi) const { return (int)hi - OCT_LO(lo); }
  __device__ __forceinline__ uint32_t key(uint32_t lo) const { return skey[OCT_LO(lo)]; }
  __device__ __forceinline__ int first(uint32_t lo) const { return OCT_LO(lo); }
  __device__ __forceinline__ int last(uint32_t hi) const { return (int)hi; }
  // bnd[dq] .. bnd[dq + 1] = the child whose stored digit is dq
  __device__ __forceinline__ bool children(uint32_t lo, uint32_t hi, int bnd[5]) const {
    const int a = OCT_LO(lo), d = OCT_DEPTH(lo) + 1, e = (int)hi;
    bnd[0] = a; bnd[4] = e;
    bnd[1] = oct_lower(skey, a, e, d, 1); bnd[2] = oct_lower(skey, bnd[1], e, d, 2); bnd[3] = oct_lower(skey, bnd[2], e, d, 3);
    return true;
  }
  __device__ __forceinline__ void child(uint32_t plo, int d, int dq, const int bnd[5], uint32_t& clo, uint32_t& chi) const {
    (void)plo;
    clo = (uint32_t)bnd[dq] | ((uint32_t)d << 20); chi = (uint32_t)bnd[dq + 1];
  }
};
struct OctDenseView {            // nodes are cells of the count pyramid
  const int* cnt;                // all depths back to back
  int nIni, maxDepth;
  // deepest level first (every level starts on a multiple of four entries: children are read as one int4)
  __device__ __forceinline__ int off(int d) const { return nIni * (((1 << (2 * (maxDepth + 1))) - (1 << (2 * (d + 1)))) / 3); }
  __device__ __forceinline__ int size(uint32_t lo, uint32_t hi) const { (void)lo; return (int)hi; }
  __device__ __forceinline__ uint32_t key(uint32_t lo) const { return (uint32_t)OCT_LO(lo) << (2 * (OCT_D - OCT_DEPTH(lo))); }
  __device__ __forceinline__ int first(uint32_t lo) const { (void)lo; return 0; }
  __device__ __forceinline__ int last(uint32_t hi) const { return (int)hi; }
  __device__ __forceinline__ bool children(uint32_t lo, uint32_t hi, int bnd[5]) const {
    (void)hi;
    const int d = OCT_DEPTH(lo) + 1;
    if (d > maxDepth) return false;
    const int4 c = *reinterpret_cast<const int4*>(cnt + off(d) + 4 * OCT_LO(lo));
    bnd[0] = 0; bnd[1] = c.x; bnd[2] = c.x + c.y; bnd[3] = c.x + c.y + c.z; bnd[4] = c.x + c.y + c.z + c.w;
    return true;
  }
  __device__ __forceinline__ void child(uint32_t plo, int d, int dq, const int bnd[5], uint32_t& clo, uint32_t& chi) const {
    clo = (uint32_t)(4 * OCT_LO(plo) + dq) | ((uint32_t)d << 20); chi = (uint32_t)(bnd[dq + 1] - bnd[dq]);
  }
};

// The careful phase (ORBextractor.cc:794-858) on the list nd.lo/hi[cur][0..n) whose first nG entries are the depth-K
// nodes of the last normal pass. Returns the final list size (cur is updated), -1 on node overflow, -2 when the view
// cannot supply a child count (dense path: deeper than its pyramid).
template <class View>
__device__ int oct_careful(const View& view, const OctNodes& nd, OctShared& sh, int n, int nG, int N, int cap, float hX,
                           int& cur, long long* clk) {
  const int tid = threadIdx.x, T = blockDim.x;
  int *posA = nd.posA, *posB = nd.posB, *posC = nd.posC, *vecPos = nd.vecPos;
  // vector of expandable nodes in creation order = the depth-K nodes with more than one keypoint, back to front
  for (int r = tid; r < nG; r += T) {
    const int pos = nG - 1 - r;
    posA[r] = view.size(nd.lo[cur][pos], nd.hi[cur][pos]) > 1;
  }
  __syncthreads();
  int m, t1, t2;
  oct_block_scan3(nG, posA, nullptr, nullptr, sh, m, t1, t2);
  for (int r = tid; r < nG; r += T) {
    const int pos = nG - 1 - r;
    const uint32_t lo = nd.lo[cur][pos];
    const int sz = view.size(lo, nd.hi[cur][pos]);
    if (sz > 1) {
      const unsigned ulx = (unsigned)oct_node_ulx(view.key(lo), OCT_DEPTH(lo), hX);
      nd.vec[0][posA[r]] = ((unsigned long long)(((unsigned)sz << 12) | ulx) << 32) | (unsigned)pos;
    }
  }
  if (tid == 0) sh.fallback = 0;
  __syncthreads();
  int vcur = 0;
  for (int iter = 0; iter < 64; iter++) {
    if (iter < 8) OCT_CTICK(7 + iter);
    unsigned long long* vecPrev = nd.vec[vcur];
    unsigned long long* vecNew = nd.vec[vcur ^ 1];
    uint32_t *lo1 = nd.lo[cur], *hi1 = nd.hi[cur], *lo2 = nd.lo[cur ^ 1], *hi2 = nd.hi[cur ^ 1];
    for (int i = tid; i < n; i += T) vecPos[i] = -1;
    // std::sort(..., compareNodes) (:805): warp-parallel replay of libstdc++'s introsort loop, then the final
    // insertion sort (= stable sort of what the loop leaves) as a parallel rank computation
    __syncthreads();
    if (iter == 0) OCT_CTICK(30);
    ftsort::cta_introsort_loop(vecPrev, m, posA, posB, posC);
    __syncthreads();
    if (iter == 0) OCT_CTICK(31);
    ftsort::stable_rank(vecPrev, vecNew, m, tid, T);
    __syncthreads();
    if (iter == 0) OCT_CTICK(32);
    // processing order r = 0..m-1 walks the sorted vector from the back (:806)
    for (int r = tid; r < m; r += T) {
      const unsigned long long e = vecNew[m - 1 - r];
      vecPrev[r] = e;
      const int node = (int)(e & 0xFFFFFFFFu);
      vecPos[node] = r;
      int bnd[5];
      if (!view.children(lo1[node], hi1[node], bnd)) { sh.fallback = 1; bnd[0] = bnd[1] = bnd[2] = bnd[3] = bnd[4] = 0; }
      nd.split1[r] = bnd[1]; nd.split2[r] = bnd[2]; nd.split3[r] = bnd[3];
      const int c0 = bnd[1] - bnd[0], c1 = bnd[2] - bnd[1], c2 = bnd[3] - bnd[2], c3 = bnd[4] - bnd[3];
      const int k = (c0 > 0) + (c1 > 0) + (c2 > 0) + (c3 > 0);
      posA[r] = k; posB[r] = k - 1; posC[r] = (c0 > 1) + (c1 > 1) + (c2 > 1) + (c3 > 1);
    }
    if (tid == 0) sh.P = m;
    __syncthreads();
    if (sh.fallback) return -2;
    if (iter == 0) OCT_CTICK(33);
    // exclusive prefixes over the processing order: children, growth, new vector entries
    int totA, totB, totC;
    oct_block_scan3(m, posA, posB, posC, sh, totA, totB, totC);
    __syncthreads();
    // number processed P: first r whose split brings the list to N (:851), else all
    for (int r = tid; r < m; r += T) {
      const int next = r + 1 < m ? posB[r + 1] : totB;      // growth up to and including r
      if (n + next >= N) atomicMin(&sh.P, r + 1);
    }
    __syncthreads();
    if (iter == 0) OCT_CTICK(34);
    const int P = sh.P;
    const int childP = P == m ? totA : posA[P], growP = P == m ? totB : posB[P], vecP = P == m ? totC : posC[P];
    const int nNew = n + growP;
    if (nNew > cap) return -1;
    // processed nodes: children to the front, later-processed first, n4 .. n1 inside a group (push_front order)
    for (int r = tid; r < P; r += T) {
      const int node = (int)(vecPrev[r] & 0xFFFFFFFFu);
      const uint32_t plo = lo1[node];
      const int d = OCT_DEPTH(plo) + 1;
      int bnd[5];
      bnd[1] = nd.split1[r]; bnd[2] = nd.split2[r]; bnd[3] = nd.split3[r];
      bnd[0] = view.first(plo); bnd[4] = view.last(hi1[node]);
      const int kAll = (r + 1 < m ? posA[r + 1] : totA) - posA[r];
      const int start = childP - posA[r] - kAll;
      int vpos = posC[r];
      const bool flip = (d & 1) == 0;            // digits of even depth are stored complemented
      int after = 0;
      int posq[4];
      for (int q = 3; q >= 0; q--) {
        const int dq = flip ? 3 - q : q;
        posq[q] = start + after;
        after += bnd[dq + 1] > bnd[dq];
      }
      for (int q = 0; q < 4; q++) {
        const int dq = flip ? 3 - q : q;
        if (bnd[dq + 1] > bnd[dq]) {
          uint32_t clo, chi;
          view.child(plo, d, dq, bnd, clo, chi);
          lo2[posq[q]] = clo; hi2[posq[q]] = chi;
          const int sz = view.size(clo, chi);
          if (sz > 1) {
            const unsigned ulx = (unsigned)oct_node_ulx(view.key(clo), d, hX);
            vecNew[vpos++] = ((unsigned long long)(((unsigned)sz << 12) | ulx) << 32) | (unsigned)posq[q];
          }
        }
      }
    }
    __syncthreads();   // posA / posB / posC of the processing order are consumed (every thread has read its prefix values and
                       // the entries at P): posB is reused below
    if (iter == 0) OCT_CTICK(35);
    // surviving old nodes keep their relative order behind the new children
    for (int i = tid; i < n; i += T) {
      const int r = vecPos[i];
      posB[i] = (r >= 0 && r < P) ? 0 : 1;
    }
    __syncthreads();
    int nSurv, u1, u2;
    oct_block_scan3(n, posB, nullptr, nullptr, sh, nSurv, u1, u2);
    for (int i = tid; i < n; i += T) {
      const int r = vecPos[i];
      if (!(r >= 0 && r < P)) { lo2[childP + posB[i]] = lo1[i]; hi2[childP + posB[i]] = hi1[i]; }
    }
    __syncthreads();
    const bool done = nNew >= N || nNew == n;   // (:855)
    n = nNew; m = vecP;
    cur ^= 1; vcur ^= 1;
    if (done) break;
  }
  return n;
}

// ------------------------------------------------------------------------------------
// General path: sort the level's candidates by Q, everything else on ranges of that array.
// ------------------------------------------------------------------------------------
__device__ void oct_general_path(const FtParams& p, const FtBuffers& b, const FtLevel& L, const FtEye& E, int level, int eye,
                                 const OctNodes& nd, uint8_t* area, OctShared& sh, long long* clk) {
  (void)p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nWarps = T >> 5;
  const int cap = L.nodeCap;
  const int N = L.quota;
  const int nCells = L.nCols * L.nRows;
  const int tabW = L.maxBorderX - FT_MIN_BORDER + 1, tabH = L.maxBorderY - FT_MIN_BORDER + 1;
  const int kb = L.octBinDepth;
  const int nBins = L.nIni << (2 * kb);
  const int binShift = 2 * (OCT_D - kb);
  uint32_t *tabX, *tabY;
  int *cellOff, *bin, *binCur;
  uint8_t* candArea;
  {
    uint8_t* q = area;
    tabX = (uint32_t*)q; q += 4 * (size_t)tabW;
    tabY = (uint32_t*)q; q += 4 * (size_t)tabH;
    cellOff = (int*)q; q += 4 * (size_t)(nCells + 1);
    bin = (int*)q; q += 4 * (size_t)(nBins + 1);
    binCur = (int*)q; q += 4 * (size_t)nBins;
    candArea = (uint8_t*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
  }
  OCT_CTICK(0);
  // stage the path tables (they do not depend on the frame) while the cell counts are in flight
  for (int i = tid; i < tabW; i += T) tabX[i] = b.octTabX[L.octTabX + i];
  for (int i = tid; i < tabH; i += T) tabY[i] = b.octTabY[L.octTabY + i];
  for (int i = tid; i <= nBins; i += T) { bin[i] = 0; if (i < nBins) binCur[i] = 0; }
  for (int i = tid; i < OCT_MAX_WARPS * 16; i += T) { (&sh.hist[0][0])[i] = 0; (&sh.diff[0][0])[i] = 0; }
  // ---- flat canonical order (cell row-major, then row-major inside the cell): exclusive scan of the cell counts ----
  for (int c = tid; c < nCells; c += T) cellOff[c] = E.cellCount[L.cellBase + c];
  if (tid == 0) cellOff[nCells] = 0;
  __syncthreads();
  int C;
  {
    int t0, t1, t2;
    oct_block_scan3(nCells + 1, cellOff, nullptr, nullptr, sh, t0, t1, t2);
    C = t0;
  }
  __syncthreads();
  if (C > L.candCap) { C = L.candCap; if (tid == 0) atomicOr(b.status, FT_ST_CAND_OVERFLOW); }
  if (tid == 0) E.lvlCandCount[level] = C;
  uint32_t* outKp = E.lvlKp + L.lvlKpBase;
  if (C == 0) {
    if (tid == 0) E.lvlKpCount[level] = 0;
    return;
  }
  OCT_CTICK(1);
  // ---- per-candidate arrays ----
  uint32_t* candG = E.cand + L.candBase;          // canonical-order list of a level too large for shared memory
  const bool inSmem = C <= L.octCandSmem;
  uint32_t *cand, *key, *tmpKey, *tmpIdx, *skey, *sidx;
  {
    uint8_t* q = inSmem ? candArea : E.octScratch + (size_t)L.candBase * 20;
    const size_t stride = 4 * (size_t)(inSmem ? L.octCandSmem : L.candCap);
    if (inSmem) { cand = (uint32_t*)q; q += stride; } else cand = candG;
    key = (uint32_t*)q; q += stride;
    tmpKey = (uint32_t*)q; q += stride;
    tmpIdx = (uint32_t*)q; q += stride;
    skey = (uint32_t*)q; q += stride;
    sidx = (uint32_t*)q;
  }
  uint8_t* cl = (uint8_t*)key;                    // key[] is dead once the candidates are scattered into their bins
  int* chunkCnt = (int*)tmpKey;                   // tmpKey/tmpIdx are dead once the sorted arrays exist: 2 ints per candidate
  // ---- copy out of the per-cell slabs, path keys, bin histogram. One warp per cell, four cells in flight ----
  for (int cell0 = warp * 4; cell0 < nCells; cell0 += nWarps * 4) {
    uint32_t v[4]; int off[4], cnt[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int cell = cell0 + u;
      off[u] = 0; cnt[u] = 0; v[u] = 0;
      if (cell < nCells) {
        off[u] = min(cellOff[cell], C); cnt[u] = min(cellOff[cell + 1], C) - off[u];
        if (lane < cnt[u]) v[u] = E.cellKp[L.cellKpBase + (size_t)cell * L.cellCap + lane];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      for (int k = lane; k < cnt[u]; k += 32) {
        const uint32_t w = k < 32 ? v[u] : E.cellKp[L.cellKpBase + (size_t)(cell0 + u) * L.cellCap + k];   // > 32 survivors in one cell: rare
        const int c = off[u] + k;
        cand[c] = w;
        const uint32_t kq = (tabX[min(ft_px(w), tabW - 1)] | tabY[min(ft_py(w), tabH - 1)]) ^ OCT_EVEN_MASK;
        key[c] = kq;
        atomicAdd(&bin[kq >> binShift], 1);
      }
    }
  }
  __syncthreads();
  OCT_CTICK(2);
  // ---- bin offsets, scatter into bins (order inside a bin is arbitrary here), exact rank inside the bin ----
  {
    int t0, t1, t2;
    oct_block_scan3(nBins + 1, bin, nullptr, nullptr, sh, t0, t1, t2);
  }
  __syncthreads();
  for (int c = tid; c < C; c += T) {
    const uint32_t kq = key[c];
    const int bi = kq >> binShift;
    const int slot = bin[bi] + atomicAdd(&binCur[bi], 1);
    tmpKey[slot] = kq; tmpIdx[slot] = (uint32_t)c;
  }
  __syncthreads();
  OCT_CTICK(3);
  for (int s = tid; s < C; s += T) {
    const uint32_t kq = tmpKey[s];
    const uint32_t ci = tmpIdx[s];
    const int bi = kq >> binShift;
    const int b0 = bin[bi], b1 = bin[bi + 1];
    int r = 0;
    for (int j = b0; j < b1; j++) {
      const uint32_t kj = tmpKey[j];
      r += kj < kq;
      if (kj == kq && tmpIdx[j] < ci) r++;     // equal keys do not occur (one pixel, one candidate); kept total anyway
    }
    skey[b0 + r] = kq;
    sidx[b0 + r] = ci;
  }
  __syncthreads();
  OCT_CTICK(4);
  // ---- neighbour statistics: cl[i], histogram of cl, difference array of nToExpand ----
  for (int i = tid; i < C; i += T) {
    const uint32_t k = skey[i];
    const int Lc = i > 0 ? oct_common(skey[i - 1], k) : 0;
    const int Rc = i + 1 < C ? oct_common(k, skey[i + 1]) : 0;
    cl[i] = (uint8_t)Rc;
    if (i + 1 < C) atomicAdd(&sh.hist[warp][Rc], 1);
    if (Rc > Lc) { atomicAdd(&sh.diff[warp][Lc], 1); atomicSub(&sh.diff[warp][Rc], 1); }
  }
  __syncthreads();
  if (tid < 32) {
    int h = 0, d = 0;
    if (tid < 16) for (int w = 0; w < nWarps; w++) { h += sh.hist[w][tid]; d += sh.diff[w][tid]; }
    // inclusive prefix over depth: size_k = 1 + sum_{j<=k} hist[j], nexp_k = sum_{j<=k} diff[j]
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int xh = __shfl_up_sync(0xFFFFFFFFu, h, o), xd = __shfl_up_sync(0xFFFFFFFFu, d, o);
      if (lane >= o) { h += xh; d += xd; }
    }
    const int sizeK = 1 + h, nexpK = d;       // valid for lanes 0 .. OCT_D (deeper: size stays, nothing to expand)
    // the pass at which the reference stops (:790) or enters the careful phase (:794); pass 1 always runs
    int K = 1, mode = 2;
    for (;; K++) {
      const int kc = min(K, OCT_D), kp = min(K - 1, OCT_D);
      const int cur = __shfl_sync(0xFFFFFFFFu, sizeK, kc), prev = __shfl_sync(0xFFFFFFFFu, sizeK, kp);
      const int ne = K <= OCT_D ? __shfl_sync(0xFFFFFFFFu, nexpK, kc) : 0;
      if (cur >= N || cur == prev) { mode = 2; break; }
      if (cur + 3 * ne > N) { mode = 1; break; }
    }
    const int nK = __shfl_sync(0xFFFFFFFFu, sizeK, min(K, OCT_D));
    if (tid == 0) { sh.K = K; sh.mode = mode; sh.n = nK; }
  }
  __syncthreads();
  OCT_CTICK(5);
  const int K = sh.K;
  int n = sh.n;
  if (n > cap) {   // cannot happen (a pass only runs while size + 3 * nToExpand <= N; the first yields <= 4 * roots)
    if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlKpCount[level] = 0; }
    return;
  }
  // ---- the list after pass K. Classes: 0 = first element of a depth-K node, 1 = last element of one,
  //      2 + j = single keypoint that settled at depth j < K. Rank inside a class by ballots over 32-element chunks. ----
  const int nChunks = (C + 31) >> 5;
  const int nCls = K + 2;
  auto classify = [&](int i, bool& head, bool& end, int& single) {
    head = false; end = false; single = -1;
    if (i < C) {
      const int Lc = i > 0 ? cl[i - 1] : 0, Rc = cl[i];
      const int settle = max(Lc, Rc);
      if (settle >= K) { head = Lc <= K; end = Rc <= K; }
      else single = settle;
    }
  };
  for (int ch = warp; ch < nChunks; ch += nWarps) {
    bool head, end; int single;
    classify(ch * 32 + lane, head, end, single);
    const unsigned bh = __ballot_sync(0xFFFFFFFFu, head), be = __ballot_sync(0xFFFFFFFFu, end);
    int mine = lane == 0 ? __popc(bh) : __popc(be);
    for (int j = 0; j < K; j++) {
      const unsigned bs = __ballot_sync(0xFFFFFFFFu, single == j);
      if (lane == j + 2) mine = __popc(bs);
    }
    if (lane < nCls) chunkCnt[lane * nChunks + ch] = mine;
  }
  __syncthreads();
  for (int cls = warp; cls < nCls; cls += nWarps) {   // exclusive scan over the chunks, one warp per class
    int run = 0;
    int* a = chunkCnt + cls * nChunks;
    for (int base = 0; base < nChunks; base += 32) {
      const int i = base + lane;
      const int v = i < nChunks ? a[i] : 0;
      int s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += x; }
      if (i < nChunks) a[i] = run + s - v;
      run += __shfl_sync(0xFFFFFFFFu, s, 31);
    }
    if (lane == 0) sh.clsTot[cls] = run;
  }
  __syncthreads();
  if (tid == 0) {
    // depth-K nodes in front, then the singles of depth K-1, K-2, .. 0
    sh.clsBase[0] = 0; sh.clsBase[1] = 0;
    int o = sh.clsTot[0];
    for (int j = K - 1; j >= 0; j--) { sh.clsBase[2 + j] = o; o += sh.clsTot[2 + j]; }
  }
  __syncthreads();
  const int nG = sh.clsTot[0];
  int cur = 0;
  for (int ch = warp; ch < nChunks; ch += nWarps) {
    const int i = ch * 32 + lane;
    bool head, end; int single;
    classify(i, head, end, single);
    const unsigned lt = (1u << lane) - 1u;
    const unsigned bh = __ballot_sync(0xFFFFFFFFu, head), be = __ballot_sync(0xFFFFFFFFu, end);
    if (head) {
      int r = chunkCnt[ch] + __popc(bh & lt);
      if (K & 1) r = nG - 1 - r;
      nd.lo[0][r] = (uint32_t)i | ((uint32_t)K << 20);
    }
    if (end) {
      int r = chunkCnt[nChunks + ch] + __popc(be & lt);
      if (K & 1) r = nG - 1 - r;
      nd.hi[0][r] = (uint32_t)(i + 1);
    }
    for (int j = 0; j < K; j++) {
      const unsigned bs = __ballot_sync(0xFFFFFFFFu, single == j);
      if (single == j) {
        int r = chunkCnt[(2 + j) * nChunks + ch] + __popc(bs & lt);
        if (j & 1) r = sh.clsTot[2 + j] - 1 - r;
        const int pos = sh.clsBase[2 + j] + r;
        nd.lo[0][pos] = (uint32_t)i | ((uint32_t)j << 20);
        nd.hi[0][pos] = (uint32_t)(i + 1);
      }
    }
  }
  __syncthreads();
  OCT_CTICK(6);
  // ---- careful phase ----
  if (sh.mode == 1) {
    OctGeneralView view; view.skey = skey;
    const int r = oct_careful(view, nd, sh, n, nG, N, cap, L.hX, cur, clk);
    if (r < 0) {
      if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlKpCount[level] = 0; }
      return;
    }
    n = r;
  }
  OCT_CTICK(20);
  // ---- best keypoint per node: highest response, first in input order wins ties (:862-881) ----
  if (n > L.lvlKpCap) {
    if (tid == 0) { atomicOr(b.status, FT_ST_KP_OVERFLOW); E.lvlKpCount[level] = 0; }
    return;
  }
  {
    const uint32_t* lo1 = nd.lo[cur];
    const uint32_t* hi1 = nd.hi[cur];
    for (int i = warp; i < n; i += nWarps) {
      const int nlo = OCT_LO(lo1[i]), nhi = (int)hi1[i];
      unsigned best = 0;
      for (int j = nlo + lane; j < nhi; j += 32) {
        const uint32_t c = sidx[j];
        best = max(best, ((unsigned)ft_ps(cand[c]) << 20) | (0xFFFFFu - c));
      }
      best = __reduce_max_sync(0xFFFFFFFFu, best);
      if (lane == 0) {
        const uint32_t pk = cand[0xFFFFFu - (best & 0xFFFFFu)];
        outKp[i] = ft_pack_xys(ft_px(pk) + FT_MIN_BORDER, ft_py(pk) + FT_MIN_BORDER, ft_ps(pk));
      }
    }
  }
  if (tid == 0) E.lvlKpCount[level] = n;
  OCT_CTICK(21);
}



// ------------------------------------------------------------------------------------
// Dense path: the count / best pyramid over the cells k_fast_cells filled. Returns false when a decision needs a
// depth below the pyramid (the caller then runs the general path); the global arrays are zeroed for the next frame
// either way.
// ------------------------------------------------------------------------------------
__device__ bool oct_dense_path(const FtBuffers& b, const FtLevel& L, const FtEye& E, int level, const OctNodes& nd,
                               uint8_t* area, OctShared& sh, long long* clk) {
  const int tid = threadIdx.x, lane = tid & 31, T = blockDim.x;
  const int cap = L.nodeCap, N = L.quota, Dd = L.octDenseDepth, nIni = L.nIni;
  const int nDense = nIni << (2 * Dd);
  OctDenseView view;
  view.nIni = nIni; view.maxDepth = Dd;
  const int totalCells = view.off(0) + nIni;
  int* cnt = (int*)area;
  unsigned* best = (unsigned*)(cnt + ((totalCells + 3) & ~3));
  int* tcount = (int*)(best + ((totalCells + 3) & ~3));     // [2][T] per-thread flag counts of the list construction
  view.cnt = cnt;
  OCT_CTICK(0);
  if (tid < 16) { sh.sizeD[tid] = 0; sh.nexpD[tid] = 0; }
  {
    int4* gC = reinterpret_cast<int4*>(E.octCnt + L.octDenseBase);
    uint4* gB = reinterpret_cast<uint4*>(E.octBest + L.octDenseBase);
    const int4 z = make_int4(0, 0, 0, 0);
    for (int i = tid; i < nDense / 4; i += T) {
      const int4 c = gC[i];
      const uint4 bb = gB[i];
      reinterpret_cast<int4*>(cnt)[i] = c;
      reinterpret_cast<uint4*>(best)[i] = bb;
      gC[i] = z;
      gB[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  OCT_CTICK(1);
  // pyramid: a cell of depth d = its four children of depth d + 1 (consecutive entries: the index is the Q prefix)
  for (int d = Dd - 1; d >= 0; d--) {
    const int cells = nIni << (2 * d);
    const int oc = view.off(d + 1), op = view.off(d);
    int ne = 0, mu = 0;
    for (int i = tid; i < cells; i += T) {
      const int4 c = *reinterpret_cast<const int4*>(cnt + oc + 4 * i);
      const uint4 bb = *reinterpret_cast<const uint4*>(best + oc + 4 * i);
      cnt[op + i] = c.x + c.y + c.z + c.w;
      best[op + i] = max(max(bb.x, bb.y), max(bb.z, bb.w));
      ne += (c.x > 0) + (c.y > 0) + (c.z > 0) + (c.w > 0);
      mu += (c.x > 1) + (c.y > 1) + (c.z > 1) + (c.w > 1);
    }
    if (cells >= 32 || tid < 32) {
      ne = __reduce_add_sync(0xFFFFFFFFu, ne); mu = __reduce_add_sync(0xFFFFFFFFu, mu);
      if (lane == 0 && (ne | mu)) { atomicAdd(&sh.sizeD[d + 1], ne); atomicAdd(&sh.nexpD[d + 1], mu); }
    }
    __syncthreads();
  }
  OCT_CTICK(2);
  if (tid == 0) {
    int C = 0, ne = 0, mu = 0;
    for (int i = 0; i < nIni; i++) { const int c = cnt[view.off(0) + i]; C += c; ne += c > 0; mu += c > 1; }
    sh.sizeD[0] = ne; sh.nexpD[0] = mu;
    sh.C = C;
    // the pass at which the reference stops (:790) or enters the careful phase (:794); pass 1 always runs
    int K = 1, mode = 2, fb = 0;
    for (;; K++) {
      if (K > Dd) { fb = 1; break; }
      const int cur = sh.sizeD[K], prev = sh.sizeD[K - 1];
      if (cur >= N || cur == prev) { mode = 2; break; }
      if (cur + 3 * sh.nexpD[K] > N) { mode = 1; if (K >= Dd) fb = 1; break; }   // the careful phase reads depth K + 1
    }
    sh.K = K; sh.mode = mode; sh.fallback = fb;
    sh.n = fb ? 0 : sh.sizeD[K];
    if (C > L.candCap) atomicOr(b.status, FT_ST_CAND_OVERFLOW);
  }
  __syncthreads();
  const int C = sh.C;
  if (C == 0) {
    if (tid == 0) { E.lvlCandCount[level] = 0; E.lvlKpCount[level] = 0; }
    return true;
  }
  if (sh.fallback) return false;
  const int K = sh.K;
  int n = sh.n;
  if (n > cap) {
    if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlCandCount[level] = C; E.lvlKpCount[level] = 0; }
    return true;
  }
  // ---- the list after pass K: flagged cells of depth K (all of them in front), then the single-keypoint cells that
  //      settled at depth K-1, .. 0; ascending cell index at even depth, descending at odd depth. The groups are laid
  //      out back to back as one virtual array and compacted with one scan. ----
  const int total = nIni * (((1 << (2 * (K + 1))) - 1) / 3);
  const int items = (total + T - 1) / T;
  const int v0 = min(tid * items, total), v1 = min(v0 + items, total);
  auto decode = [&](int v, int& j, int& i, int& c, bool& flag) {
    j = K;
    int g = nIni << (2 * K);
    while (v >= g) { v -= g; j--; g >>= 2; }
    i = (j & 1) ? g - 1 - v : v;
    c = cnt[view.off(j) + i];
    const bool parentMulti = j == 0 || cnt[view.off(j - 1) + (i >> 2)] > 1;
    flag = parentMulti && (j == K ? c >= 1 : c == 1);
  };
  {
    int f = 0, fg = 0;
    for (int v = v0; v < v1; v++) {
      int j, i, c; bool flag;
      decode(v, j, i, c, flag);
      f += flag; fg += flag && j == K;
    }
    tcount[tid] = f; tcount[T + tid] = fg;
  }
  __syncthreads();
  int nList, nG, u2;
  oct_block_scan3(T, tcount, tcount + T, nullptr, sh, nList, nG, u2);
  {
    int pos = tcount[tid];
    for (int v = v0; v < v1; v++) {
      int j, i, c; bool flag;
      decode(v, j, i, c, flag);
      if (flag) { nd.lo[0][pos] = (uint32_t)i | ((uint32_t)j << 20); nd.hi[0][pos] = (uint32_t)c; pos++; }
    }
  }
  __syncthreads();
  OCT_CTICK(6);
  int cur = 0;
  if (sh.mode == 1) {
    const int r = oct_careful(view, nd, sh, n, nG, N, cap, L.hX, cur, clk);
    if (r == -2) return false;
    if (r < 0) {
      if (tid == 0) { atomicOr(b.status, FT_ST_NODE_OVERFLOW); E.lvlCandCount[level] = C; E.lvlKpCount[level] = 0; }
      return true;
    }
    n = r;
  }
  OCT_CTICK(20);
  if (n > L.lvlKpCap) {
    if (tid == 0) { atomicOr(b.status, FT_ST_KP_OVERFLOW); E.lvlCandCount[level] = C; E.lvlKpCount[level] = 0; }
    return true;
  }
  // ---- best keypoint per node (:862-881): the pyramid holds (response, first canonical position) per cell ----
  uint32_t* outKp = E.lvlKp + L.lvlKpBase;
  for (int i = tid; i < n; i += T) {
    const uint32_t lo = nd.lo[cur][i];
    const unsigned bk = best[view.off(OCT_DEPTH(lo)) + OCT_LO(lo)];
    const unsigned where = 0xFFFFFu - (bk & 0xFFFFFu);         // cell << 9 | position in the cell's slab
    const uint32_t pk = E.cellKp[L.cellKpBase + (size_t)(where >> 9) * L.cellCap + (where & 511u)];
    outKp[i] = ft_pack_xys(ft_px(pk) + FT_MIN_BORDER, ft_py(pk) + FT_MIN_BORDER, ft_ps(pk));
  }
  if (tid == 0) { E.lvlCandCount[level] = C; E.lvlKpCount[level] = n; }
  OCT_CTICK(21);
  return true;
}

__global__ void __launch_bounds__(OCT_MAX_THREADS) k_octree(const __grid_constant__ FtParams p, const __grid_constant__ FtBuffers b,
                                                            int levelBegin) {
  extern __shared__ __align__(16) uint8_t smemRaw[];
  __shared__ OctShared sh;
  const int level = levelBegin + blockIdx.x;
  const int eye = blockIdx.y;
  const FtLevel& L = p.lv[level];
  const FtEye& E = b.eye[eye];
  const size_t cap = (size_t)L.nodeCap;
  // ---- carve shared memory (host side: ft_octree_smem_bytes): node arrays, then the path-specific area ----
  OctNodes nd;
  uint8_t* q = smemRaw;
  nd.vec[0] = (unsigned long long*)q; q += 8 * cap;
  nd.vec[1] = (unsigned long long*)q; q += 8 * cap;
  nd.lo[0] = (uint32_t*)q; q += 4 * cap;
  nd.lo[1] = (uint32_t*)q; q += 4 * cap;
  nd.hi[0] = (uint32_t*)q; q += 4 * cap;
  nd.hi[1] = (uint32_t*)q; q += 4 * cap;
  nd.posA = (int*)q; q += 4 * cap;
  nd.posB = (int*)q; q += 4 * cap;
  nd.posC = (int*)q; q += 4 * cap;
  nd.vecPos = (int*)q; q += 4 * cap;
  nd.split1 = (int*)q; q += 4 * cap;
  nd.split2 = (int*)q; q += 4 * cap;
  nd.split3 = (int*)q; q += 4 * cap;
  uint8_t* area = (uint8_t*)(((uintptr_t)q + 15) & ~(uintptr_t)15);
#ifdef FT_OCT_CLOCK
  long long* clk = E.octClock ? E.octClock + (eye * FT_MAX_LEVELS + level) * 64 : nullptr;
#else
  long long* clk = nullptr;
#endif
  FT_PDL_WAIT();        // launched as a programmatic dependent of the level's k_fast_cells
  if (L.octDenseDepth > 0) {
    if (oct_dense_path(b, L, E, level, nd, area, sh, clk)) return;
    __syncthreads();
#ifdef FT_OCT_CLOCK
    if (threadIdx.x == 0 && clk) clk[40] = 1;     // fell back
#endif
  }
  oct_general_path(p, b, L, E, level, eye, nd, area, sh, clk);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
#include "ft_internal.h"

// bytes of dynamic shared memory in front of the per-candidate arrays (mirrors the carve in k_octree)
static size_t oct_node_bytes(const FtLevel& L) { return (((size_t)L.nodeCap * 60) + 15) & ~(size_t)15; }
// general path: tables, cell offsets, bins in front of the per-candidate arrays
static size_t oct_general_fixed_bytes(const FtLevel& L) {
  const int tabW = L.maxBorderX - FT_MIN_BORDER + 1, tabH = L.maxBorderY - FT_MIN_BORDER + 1;
  const int nBins = L.nIni << (2 * L.octBinDepth);
  const size_t s = 4 * (size_t)(tabW + tabH + L.nCols * L.nRows + 1 + nBins + 1 + nBins);
  return (s + 15) & ~(size_t)15;
}
// dense path: count + best pyramids and the per-thread flag counts
static size_t oct_dense_bytes(const FtLevel& L, int depth) {
  const size_t cells = (size_t)L.nIni * (((size_t)1 << (2 * (depth + 1))) - 1) / 3;
  return 2 * 4 * ((cells + 3) & ~(size_t)3) + 2 * 4 * OCT_MAX_THREADS + 16;
}
size_t ft_octree_smem_bytes(const FtParams& p, int level) {
  const FtLevel& L = p.lv[level];
  const size_t g = oct_general_fixed_bytes(L) + 24 * (size_t)L.octCandSmem + 16;
  const size_t d = L.octDenseDepth > 0 ? oct_dense_bytes(L, L.octDenseDepth) : 0;
  return oct_node_bytes(L) + (g > d ? g : d) + 16;
}

// Per-level constants of the octree kernel: bin depth of the sort (bins of about 24 x 24 px, at most 4096 of them) and
// how many candidates the level keeps in shared memory. Returns false when even the node list does not fit.
bool ft_octree_plan(FtLevel& L, size_t smemBudget) {
  const float rootW = L.hX, rootH = (float)(L.maxBorderY - FT_MIN_BORDER);
  int kb = 0;
  while (kb < 5 && (L.nIni << (2 * (kb + 1))) <= 4096 && fmaxf(rootW, rootH) / (float)(1 << kb) > 24.f) kb++;
  L.octBinDepth = kb;
  const size_t fixed = oct_node_bytes(L) + oct_general_fixed_bytes(L) + 48;
  if (fixed + 24 * 64 > smemBudget) return false;
  const size_t room = (smemBudget - fixed) / 24;
  L.octCandSmem = (int)(room < (size_t)L.candCap ? room : (size_t)L.candCap);
  // dense path: cells of depth Dd no smaller than about 1.5 px, at most 8192 of them; the canonical position of a
  // candidate must fit the 20-bit tie-break field (cell < 2048, slot < 512)
  L.octDenseDepth = 0;
  if (L.nCols * L.nRows <= 2048 && L.cellCap <= 512) {
    int d = 0;
    while (d < 6 && (L.nIni << (2 * (d + 1))) <= 8192 && fminf(rootW, rootH) / (float)(1 << (d + 1)) >= 1.5f) d++;
    if (d >= 2 && oct_node_bytes(L) + oct_dense_bytes(L, d) + 48 <= smemBudget) L.octDenseDepth = d;
  }
  // With the dense path on, the general path is the rare fallback: give it only as much shared memory as the dense path
  // needs anyway (its candidates spill to the HBM scratch beyond that), so that an SM hosting an octree CTA keeps about
  // 100 KB for the FAST / blur CTAs of the frames in flight. FT_OCT_SMEM=full keeps the large reservation.
  const char* e = getenv("FT_OCT_SMEM");
  if (L.octDenseDepth > 0 && !(e && e[0] == 'f')) {
    const size_t dense = oct_dense_bytes(L, L.octDenseDepth), gfix = oct_general_fixed_bytes(L) + 16;
    const size_t fit = dense > gfix + 24 * 64 ? (dense - gfix) / 24 : 64;
    if ((size_t)L.octCandSmem > fit) L.octCandSmem = (int)fit;
  }
  return true;
}

// Path tables of one level: the x (y) half of DivideNode's descent for every column (row) of the level's keypoint area,
// 12 quadrant bits spread to the even (odd) bit positions, the root index above them (ORBextractor.cc:510-538, 664-706).
void ft_octree_tables(const FtLevel& L, uint32_t* tabX, uint32_t* tabY) {
  const int tabW = L.maxBorderX - FT_MIN_BORDER + 1, tabH = L.maxBorderY - FT_MIN_BORDER + 1;
  for (int x = 0; x < tabW; x++) {
    int r = (int)((float)x / L.hX);
    if (r >= L.nIni) r = L.nIni - 1;
    int x0 = (int)(L.hX * (float)r), x1 = (int)(L.hX * (float)(r + 1));
    uint32_t w = (uint32_t)r << (2 * OCT_D);
    for (int d = 1; d <= OCT_D; d++) {
      const int mx = x0 + ((x1 - x0 + 1) >> 1);
      if (x >= mx) { w |= 1u << (2 * (OCT_D - d)); x0 = mx; } else x1 = mx;
    }
    tabX[x] = w;
  }
  for (int y = 0; y < tabH; y++) {
    int y0 = 0, y1 = L.maxBorderY - FT_MIN_BORDER;
    uint32_t w = 0;
    for (int d = 1; d <= OCT_D; d++) {
      const int my = y0 + ((y1 - y0 + 1) >> 1);
      if (y >= my) { w |= 2u << (2 * (OCT_D - d)); y0 = my; } else y1 = my;
    }
    tabY[y] = w;
  }
}

cudaError_t ft_launch_octree_setup() {
  // per function and per device: the largest opt-in size, set once, never lowered by a later (smaller) context
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, k_octree);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

size_t ft_octree_smem_budget() {
  int dev = 0, optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
    return 0;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, k_octree) != cudaSuccess) return 0;
  const size_t room = (size_t)optin - fa.sharedSizeBytes;
  return room < 200 * 1024 ? room : 200 * 1024;
}

void ft_launch_octree(const FtParams& p, const FtBuffers& b, int l0, int l1, cudaStream_t st) {
  size_t mx = 0;
  int threads = 256;
  for (int l = l0; l < l1; l++) {
    mx = ft_octree_smem_bytes(p, l) > mx ? ft_octree_smem_bytes(p, l) : mx;
    // enough threads for a handful of candidates each on a typical frame (about a sixth of the NMS capacity)
    const int want = p.lv[l].candCap / 48;
    while (threads < OCT_MAX_THREADS && threads < want) threads *= 2;
  }
  ft_launch_pdl(k_octree, dim3(l1 - l0, p.nEyes), dim3(threads), mx, st, p, b, l0);   // predecessor in the stream: k_fast_cells
}
