// ft_bow.cu -- bag-of-words side of the front-end: Frame::ComputeBoW and
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) on the device-resident frame.
//
// Reference (paths relative to the reference repository):
//   Frame::ComputeBoW                               src/Frame.cc:762-769 -> ORBVocabulary::transform(desc, BowVec, FeatVec, 4)
//   TemplatedVocabulary::loadFromTextFile           Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1423
//   TemplatedVocabulary::transform (tree descent)   :1218-1260, (features -> BowVector / FeatureVector) :1127-1194
//   BowVector::addWeight / normalize                Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-86
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ..)  src/ORBmatcher.cc:322-523, ComputeThreeMaxima :2210-2254
//
// Device layout. The vocabulary tree lives in HBM in CHILD ORDER: the children of a node are contiguous
// (firstChild[node], nChild[node]), each child a 32-byte descriptor (two uint4) plus its node id, so one descent step of
// one feature is one coalesced read of k*32 bytes by the k lanes of a warp; ORBvoc (k = 10, L = 6, 1.1 M nodes) is
// 35.5 MB of descriptors and stays L2-resident between frames. Kernels:
//   k_bow_transform  one warp per feature; lane j scores child j (__popc over two uint4), warp arg-min with the
//                    first-minimum-wins rule of the reference's strict `<`; L dependent steps.
//   k_bow_vector     two CTAs side by side. CTA 0: bitonic sort of (word, feature) keys in shared memory, run heads ->
//                    distinct words, the reference's summation order for the weights and for the norm (sequential in
//                    ascending word id). CTA 1: the FeatureVector as a CSR (features sorted by (node, index), group heads).
//   k_bow_group      the same CSR for the uploaded KeyFrame side of a search.
//   k_bow_search     one warp per KeyFrame node group. Frame features of different nodes are disjoint, so the
//                    reference's loop-carried "already matched" test only couples KeyFrame features of one node: the
//                    warp walks them in order and the lanes split the frame features of that node.
//   k_bow_finish     ComputeThreeMaxima + withdrawal of the matches outside the three dominant rotation bins.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ft_internal.h"

#define FT_BOW_MAX_FEATURES 16384   // one shared-memory sort: 16384 x 8 bytes
#define FT_BOW_HISTO 30             // HISTO_LENGTH, src/ORBmatcher.cc:37
#define FT_BOW_TH_LOW 50            // ORBmatcher::TH_LOW, src/ORBmatcher.cc:36

struct ft_vocabulary {
  int device = 0;
  int k = 0, L = 0, scoring = 0, weighting = 0, nNodes = 0, nWords = 0;
  FtVocDevice D = {};
  std::vector<void*> allocs;
  // scratch of ft_vocabulary_transform (host descriptors), guarded by mu
  std::mutex mu;
  cudaStream_t stream = nullptr;
  FtBowFrame T = {};
  uint8_t* tDesc = nullptr;
  int tCap = 0;
  std::vector<void*> tAllocs;
};

// ----------------------------------------------------------------------------------------------------------------
// device code
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint4* bow_frame_desc(const FtBowSource& s, int i, int nLeft) {
  return reinterpret_cast<const uint4*>(i < nLeft ? s.desc0 + 32 * (size_t)i : s.desc1 + 32 * (size_t)(i - nLeft));
}

__device__ __forceinline__ void bow_counts(const FtBowSource& s, int& nLeft, int& n) {
  nLeft = s.cnt0 ? s.cnt0[0] : s.nFixed;
  n = nLeft + (s.cnt1 ? s.cnt1[0] : 0);
}

// transform(feature, word_id, weight, nid, levelsup): TemplatedVocabulary.h:1218-1260
__global__ void __launch_bounds__(256) k_bow_transform(FtVocDevice V, FtBowSource S, FtBowFrame F, int levelsup) {
  int nLeft, n;
  bow_counts(S, nLeft, n);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i == 0 && lane == 0) { F.meta[0] = n; F.meta[1] = nLeft; }
  if (i >= n) return;
  const uint4* f = bow_frame_desc(S, i, nLeft);
  const uint4 a0 = __ldg(f), a1 = __ldg(f + 1);
  const int nidLevel = V.L - levelsup;
  int cur = 0, level = 0, nid = 0;
  if (V.nNodes > 1) {
    for (;;) {
      const int nc = V.nChild[cur];
      if (nc == 0) break;
      const int fc = V.firstChild[cur];
      unsigned best = 0xFFFFFFFFu;
      for (int j0 = 0; j0 < nc; j0 += 32) {
        const int j = j0 + lane;
        if (j < nc) {
          const uint4* d = V.childDesc + 2 * (size_t)(fc + j);
          const unsigned key = ((unsigned)ft_hamming256(a0, a1, __ldg(d), __ldg(d + 1)) << 20) | (unsigned)j;   // first minimum wins
          best = min(best, key);
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
      cur = V.childNode[fc + (int)(best & 0xFFFFFu)];
      if (++level == nidLevel) nid = cur;
    }
  }
  if (lane == 0) {
    const bool live = V.nNodes > 1 && V.weight[cur] > 0;   // "not stopped"
    F.word[i] = V.nNodes > 1 ? V.wordId[cur] : -1;
    F.node[i] = live ? nid : -1;
  }
}

// in-place bitonic sort of P (power of two) 64-bit keys in shared memory by the whole CTA
__device__ void bow_block_sort(unsigned long long* key, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const unsigned long long x = key[lo], y = key[hi];
        const bool up = (lo & k) == 0;
        if ((x > y) == up) { key[lo] = y; key[hi] = x; }
      }
      __syncthreads();
    }
  }
}

// number of keys below the sentinel after the sort (keys are unique, sentinel = all ones)
__device__ int bow_count_valid(const unsigned long long* key, int P, int* sh) {
  if (threadIdx.x == 0) *sh = 0;
  __syncthreads();
  int c = 0;
  for (int t = threadIdx.x; t < P; t += blockDim.x) c += key[t] != ~0ull;
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(sh, c);
  __syncthreads();
  return *sh;
}

// Ranks the run heads of the sorted keys (a head = the high word differs from its predecessor): start[r] = position of
// the r-th head, start[count] = nv. Returns the number of runs. Ordered compaction by a block-wide scan.
__device__ int bow_run_heads(const unsigned long long* key, int nv, int* start, int* shScan /* [33] */) {
  const int per = (nv + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per, e = min(nv, b + per);
  int c = 0;
  for (int p = b; p < e; p++) c += (p == 0) || ((key[p] >> 32) != (key[p - 1] >> 32));
  // exclusive scan of c over the CTA
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = c;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) shScan[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? shScan[lane] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += u; }
    shScan[lane] = s - v;
    if (lane == 31) shScan[32] = s;
  }
  __syncthreads();
  int r = shScan[w] + inc - c;
  const int total = shScan[32];
  for (int p = b; p < e; p++)
    if ((p == 0) || ((key[p] >> 32) != (key[p - 1] >> 32))) start[r++] = p;
  if (threadIdx.x == 0) start[total] = nv;
  __syncthreads();
  return total;
}

// FeatureVector as a CSR: features with a node, sorted by (node, feature index); gStart = group heads
__device__ void bow_group(unsigned long long* key, int* shCnt, int* shScan, const int* node, int n, int* idxOut, int* nodeOut,
                          int* gStart, int* meta /* [0] nv, [1] groups */, int P) {
  for (int t = threadIdx.x; t < P; t += blockDim.x)
    key[t] = (t < n && node[t] >= 0) ? (((unsigned long long)(unsigned)node[t] << 32) | (unsigned)t) : ~0ull;
  __syncthreads();
  bow_block_sort(key, P);
  const int nv = bow_count_valid(key, P, shCnt);
  const int ng = bow_run_heads(key, nv, gStart, shScan);
  for (int p = threadIdx.x; p < nv; p += blockDim.x) { idxOut[p] = (int)(unsigned)key[p]; nodeOut[p] = (int)(key[p] >> 32); }
  if (threadIdx.x == 0) { meta[0] = nv; meta[1] = ng; }
}

// CTA 0: BowVector of the frame, transform(features, v, fv, levelsup) :1127-1194 + BowVector.cpp:34-86.
// CTA 1: its FeatureVector (the two sorts are independent and run side by side).
__global__ void __launch_bounds__(1024) k_bow_vector(FtVocDevice V, FtBowFrame F, int P) {
  extern __shared__ unsigned long long key[];
  __shared__ int shCnt, shScan[33];
  __shared__ double shNorm;
  const int n = F.meta[0];
  if (blockIdx.x == 1) { bow_group(key, &shCnt, shScan, F.node, n, F.fvIdx, F.fvNode, F.fvGStart, F.fvMeta, P); return; }
  for (int t = threadIdx.x; t < P; t += blockDim.x)
    key[t] = (t < n && F.node[t] >= 0) ? (((unsigned long long)(unsigned)F.word[t] << 32) | (unsigned)t) : ~0ull;
  __syncthreads();
  bow_block_sort(key, P);
  const int nv = bow_count_valid(key, P, &shCnt);
  const int nd = bow_run_heads(key, nv, F.bowStart, shScan);
  const bool tf = V.weighting == 0 || V.weighting == 1;     // TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist
  const bool must = V.scoring != 5;                          // every scoring but DOT_PRODUCT normalises (ScoringObject.h:74-89)
  for (int r = threadIdx.x; r < nd; r += blockDim.x) {
    const int p = F.bowStart[r], cnt = F.bowStart[r + 1] - p;
    const unsigned word = (unsigned)(key[p] >> 32);
    const double w = V.wordWeight[word];
    double v = w;
    if (tf) for (int c = 1; c < cnt; c++) v += w;            // `vit->second += v`, once per further feature of the word
    if (tf && !must) v /= (double)nd;
    F.bowIds[r] = word;
    F.bowVals[r] = v;
  }
  __syncthreads();
  if (must) {
    // BowVector::normalize sums in map order: one thread, strictly sequential, so the doubles equal the reference's.
    // The addends are staged in shared memory (the sort keys are dead by now) to keep the chain at add latency.
    double* val = reinterpret_cast<double*>(key);
    for (int r = threadIdx.x; r < nd; r += blockDim.x) {
      const double v = F.bowVals[r];
      val[r] = V.scoring != 1 ? fabs(v) : v * v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double norm = 0.0;
      int r = 0;
      for (; r + 4 <= nd; r += 4) {
        const double a = val[r], b = val[r + 1], c = val[r + 2], d = val[r + 3];
        norm += a; norm += b; norm += c; norm += d;
      }
      for (; r < nd; r++) norm += val[r];
      if (V.scoring == 1) norm = sqrt(norm);
      shNorm = norm;
    }
    __syncthreads();
    const double norm = shNorm;
    if (norm > 0.0) for (int r = threadIdx.x; r < nd; r += blockDim.x) F.bowVals[r] = F.bowVals[r] / norm;
  }
  if (threadIdx.x == 0) F.meta[2] = nd;
}

// the KeyFrame's FeatureVector, from one node per feature
__global__ void __launch_bounds__(1024) k_bow_group(const int* node, int n, int* idxOut, int* nodeOut, int* gStart, int* meta, int P) {
  extern __shared__ unsigned long long key[];
  __shared__ int shCnt, shScan[33];
  bow_group(key, &shCnt, shScan, node, n, idxOut, nodeOut, gStart, meta, P);
}

__global__ void k_bow_search_init(FtBowFrame F, FtBowSearch Q, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { Q.match[i] = -1; Q.matchBin[i] = -1; }
  if (i < 32) Q.hist[i] = 0;
  if (i == 0) Q.result[0] = 0;
}

// merge of two (best key, second distance) pairs; key = dist << 16 | position, second = distance only
__device__ __forceinline__ void bow_merge(unsigned& k1, int& d2, unsigned ok1, int od2) {
  if (ok1 < k1) { d2 = min(od2, (int)(k1 >> 16)); k1 = ok1; }
  else { d2 = min(d2, (int)(ok1 >> 16)); }
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), src/ORBmatcher.cc:322-523
// The frame features of the node are loaded once: each lane keeps up to BOW_CACHE of them (index, descriptor, claimed bit)
// in registers, so the in-order walk over the node's KeyFrame features has no dependent global load on its critical path
// (the next KeyFrame feature is prefetched while the current one is scored). Nodes with more than 32 * BOW_CACHE frame
// features continue from global memory with the claim table.
#define BOW_CACHE 4
__global__ void __launch_bounds__(128) k_bow_search(FtBowSource S, FtBowFrame F, FtBowSearch Q, float nnratio, int checkOri) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int ngKF = Q.kfMeta[1];
  if (g >= ngKF) return;
  int nLeft, n;
  bow_counts(S, nLeft, n);
  const int fNleft = S.cnt1 ? nLeft : -1;                     // F.Nleft: -1 unless the frame has a second camera
  const int kb = Q.kfGStart[g], ke = Q.kfGStart[g + 1];
  const int node = Q.kfNodeSorted[kb];
  // Fit = F.mFeatVec.lower_bound(node): binary search over the frame's group heads
  int lo = 0, hi = F.fvMeta[1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (F.fvNode[F.fvGStart[mid]] < node) lo = mid + 1; else hi = mid;
  }
  if (lo >= F.fvMeta[1] || F.fvNode[F.fvGStart[lo]] != node) return;
  const int fb = F.fvGStart[lo], fe = F.fvGStart[lo + 1];
  const float factor = 1.0f / FT_BOW_HISTO;
  // this lane's share of the node's frame features
  int cIdx[BOW_CACHE];
  uint4 c0[BOW_CACHE], c1[BOW_CACHE];
  unsigned claimed = 0;
#pragma unroll
  for (int c = 0; c < BOW_CACHE; c++) {
    const int p = fb + lane + 32 * c;
    cIdx[c] = p < fe ? F.fvIdx[p] : -1;
    if (cIdx[c] >= 0) { const uint4* df = bow_frame_desc(S, cIdx[c], nLeft); c0[c] = __ldg(df); c1[c] = __ldg(df + 1); }
    else { c0[c] = make_uint4(0, 0, 0, 0); c1[c] = c0[c]; }
  }
  const int cachedEnd = min(fe, fb + 32 * BOW_CACHE);
  // software pipeline over the KeyFrame features of the node
  int realIdxKF = Q.kfIdxSorted[kb];
  int has = Q.kfHasMp[realIdxKF];
  const uint4* dk = reinterpret_cast<const uint4*>(Q.kfDesc + 32 * (size_t)realIdxKF);
  uint4 a0 = __ldg(dk), a1 = __ldg(dk + 1);
  float kfAngle = Q.kfAngle[realIdxKF];
  for (int q = kb; q < ke; q++) {
    const int curKF = realIdxKF, curHas = has;
    const uint4 b0 = a0, b1 = a1;
    const float curAngle = kfAngle;
    if (q + 1 < ke) {
      realIdxKF = Q.kfIdxSorted[q + 1];
      has = Q.kfHasMp[realIdxKF];
      dk = reinterpret_cast<const uint4*>(Q.kfDesc + 32 * (size_t)realIdxKF);
      a0 = __ldg(dk); a1 = __ldg(dk + 1);
      kfAngle = Q.kfAngle[realIdxKF];
    }
    if (!curHas) continue;                                    // !pMP || pMP->isBad()
    unsigned k1 = 256u << 16, k1R = 256u << 16;               // bestDist1 = 256, bestIdxF = -1
    int d2 = 256, d2R = 256;
#pragma unroll
    for (int c = 0; c < BOW_CACHE; c++) {
      if (cIdx[c] < 0 || ((claimed >> c) & 1)) continue;      // vpMapPointMatches[realIdxF] already set
      const int dist = ft_hamming256(b0, b1, c0[c], c1[c]);
      const unsigned key = ((unsigned)dist << 16) | (unsigned)(lane + 32 * c);
      if (fNleft == -1 || cIdx[c] < fNleft) bow_merge(k1, d2, key, 256);
      else bow_merge(k1R, d2R, key, 256);
    }
    for (int p = cachedEnd + lane; p < fe; p += 32) {         // overflow of very large nodes
      const int realIdxF = F.fvIdx[p];
      if (((volatile int*)Q.match)[realIdxF] >= 0) continue;
      const uint4* df = bow_frame_desc(S, realIdxF, nLeft);
      const int dist = ft_hamming256(b0, b1, __ldg(df), __ldg(df + 1));
      const unsigned key = ((unsigned)dist << 16) | (unsigned)(p - fb);
      if (fNleft == -1 || realIdxF < fNleft) bow_merge(k1, d2, key, 256);
      else bow_merge(k1R, d2R, key, 256);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned ok1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o); const int od2 = __shfl_xor_sync(0xFFFFFFFFu, d2, o);
      const unsigned ok1R = __shfl_xor_sync(0xFFFFFFFFu, k1R, o); const int od2R = __shfl_xor_sync(0xFFFFFFFFu, d2R, o);
      bow_merge(k1, d2, ok1, od2);
      bow_merge(k1R, d2R, ok1R, od2R);
    }
    // every lane holds the same (best, second) pairs: the decisions are uniform, lane 0 publishes them
    const int bestDist1 = (int)(k1 >> 16), bestDist1R = (int)(k1R >> 16);
    int posL = -1, posR = -1;
    if (bestDist1 <= FT_BOW_TH_LOW) {
      if ((float)bestDist1 < __fmul_rn(nnratio, (float)d2)) posL = (int)(k1 & 0xFFFFu);
      if (bestDist1R <= FT_BOW_TH_LOW) posR = (int)(k1R & 0xFFFFu);                       // `ratio test || true` (:451)
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int pos = side ? posR : posL;
      if (pos < 0) continue;
      if (pos < 32 * BOW_CACHE && (pos & 31) == lane) claimed |= 1u << (pos >> 5);
      if (lane == 0) {
        const int bestIdxF = F.fvIdx[fb + pos];
        Q.match[bestIdxF] = curKF;
        if (checkOri) {
          const float fa = bestIdxF < nLeft ? S.kps0[bestIdxF].angle : S.kps1[bestIdxF - nLeft].angle;
          float rot = __fsub_rn(curAngle, fa);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          int bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == FT_BOW_HISTO) bin = 0;
          Q.matchBin[bestIdxF] = bin;
          atomicAdd(&Q.hist[bin], 1);
        }
        atomicAdd(&Q.result[0], 1);
      }
    }
    if (fe > cachedEnd) __syncwarp();   // claims of overflow features travel through the global table
  }
}

// ComputeThreeMaxima (src/ORBmatcher.cc:2210-2254) + withdrawal of the other bins (:498-513)
__global__ void __launch_bounds__(1024) k_bow_finish(FtBowFrame F, FtBowSearch Q) {
  __shared__ int ind[3];
  if (threadIdx.x == 0) {
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < FT_BOW_HISTO; i++) {
      const int s = Q.hist[i];
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
      else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
    ind[0] = ind1; ind[1] = ind2; ind[2] = ind3;
  }
  __syncthreads();
  const int n = F.meta[0];
  int removed = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = Q.matchBin[i];
    if (Q.match[i] >= 0 && b >= 0 && b != ind[0] && b != ind[1] && b != ind[2]) { Q.match[i] = -1; removed++; }
  }
  for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xFFFFFFFFu, removed, o);
  if ((threadIdx.x & 31) == 0 && removed) atomicSub(&Q.result[0], removed);
}

// ----------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------
#define BCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      ft_internal_set_err((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str());           \
      return FT_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

static int pow2_at_least(int n) { int p = 32; while (p < n) p <<= 1; return p; }

// function attributes are per device: opt the two sort kernels into large dynamic shared memory once on every device used
static bool g_bowAttr[64] = {};
static cudaError_t bow_kernel_attributes() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && g_bowAttr[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(k_bow_vector, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_BOW_MAX_FEATURES * 8);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_bow_group, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_BOW_MAX_FEATURES * 8);
  if (e == cudaSuccess && dev >= 0 && dev < 64) g_bowAttr[dev] = true;
  return e;
}

template <typename T>
static cudaError_t balloc(std::vector<void*>& owner, T** p, size_t n) {
  void* v = nullptr;
  cudaError_t e = cudaMalloc(&v, (n ? n : 1) * sizeof(T) + 256);
  if (e == cudaSuccess) { owner.push_back(v); *p = (T*)v; e = cudaMemset(v, 0, (n ? n : 1) * sizeof(T) + 256);
    // the memset runs on the legacy default stream, which the context's non-blocking stream does not wait for
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy); }
  return e;
}

cudaError_t ft_bow_frame_alloc(FtBowFrame* F, int cap, std::vector<void*>& owner) {
  cudaError_t e = bow_kernel_attributes();
  if (e == cudaSuccess) e = balloc(owner, &F->word, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->node, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->bowIds, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->bowVals, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->bowStart, cap + 1);
  if (e == cudaSuccess) e = balloc(owner, &F->fvIdx, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->fvNode, cap);
  if (e == cudaSuccess) e = balloc(owner, &F->fvGStart, cap + 1);
  if (e == cudaSuccess) e = balloc(owner, &F->fvMeta, 4);
  if (e == cudaSuccess) e = balloc(owner, &F->meta, 4);
  return e;
}

// Device side of a search. `result` and `match` are one allocation ([0] nmatches, 3 pad, match[capF]) so that one D2H
// brings the answer back; the uploaded KeyFrame side is one blob (desc | angle | node | hasMp, laid out for the call's nKF
// by ft_bow_search_bind) so that one H2D carries it.
cudaError_t ft_bow_search_alloc(FtBowSearch* Q, int capF, int capKF, std::vector<void*>& owner) {
  cudaError_t e = cudaSuccess;
  if (!Q->result) {
    e = balloc(owner, &Q->result, (size_t)capF + 4);
    if (e == cudaSuccess) Q->match = Q->result + 4;
    if (e == cudaSuccess) e = balloc(owner, &Q->matchBin, capF);
    if (e == cudaSuccess) e = balloc(owner, &Q->hist, 32);
    if (e == cudaSuccess) e = balloc(owner, &Q->kfMeta, 4);
  }
  if (e == cudaSuccess && capKF > Q->kfCap) {   // KeyFrame-side arrays grow with the largest KeyFrame seen
    void** old[] = {(void**)&Q->kfBlob, (void**)&Q->kfIdxSorted, (void**)&Q->kfNodeSorted, (void**)&Q->kfGStart};
    for (void** p : old) {
      if (!*p) continue;
      for (size_t i = 0; i < owner.size(); i++) if (owner[i] == *p) { owner.erase(owner.begin() + i); break; }
      cudaFree(*p); *p = nullptr;
    }
    e = balloc(owner, &Q->kfBlob, (size_t)capKF * 41 + 64);
    if (e == cudaSuccess) e = balloc(owner, &Q->kfIdxSorted, capKF);
    if (e == cudaSuccess) e = balloc(owner, &Q->kfNodeSorted, capKF);
    if (e == cudaSuccess) e = balloc(owner, &Q->kfGStart, capKF + 1);
    if (e == cudaSuccess) Q->kfCap = capKF;
  }
  return e;
}

// blob layout for nKF features: desc[nKF][32] | angle[nKF] | node[nKF] | hasMp[nKF]; returns its size in bytes
size_t ft_bow_search_bind(FtBowSearch* Q, int nKF) {
  Q->kfDesc = Q->kfBlob;
  Q->kfAngle = reinterpret_cast<float*>(Q->kfBlob + (size_t)32 * nKF);
  Q->kfNode = reinterpret_cast<int*>(Q->kfBlob + (size_t)36 * nKF);
  Q->kfHasMp = Q->kfBlob + (size_t)40 * nKF;
  return (size_t)41 * nKF;
}

// transform of up to maxN features (the actual count is read on the device): 2 launches
int ft_launch_bow_transform(const ft_vocabulary* voc, const FtBowSource& S, const FtBowFrame& F, int maxN, int levelsup,
                            cudaStream_t st) {
  const int P = pow2_at_least(maxN);
  k_bow_transform<<<(maxN + 7) / 8, 256, 0, st>>>(voc->D, S, F, levelsup);
  k_bow_vector<<<2, 1024, (size_t)P * 8, st>>>(voc->D, F, P);
  return 2;
}

// SearchByBoW against nKF uploaded KeyFrame features: 4 (5 with the orientation check) launches
int ft_launch_bow_search(const FtBowSource& S, const FtBowFrame& F, const FtBowSearch& Q, int nKF, int capF, float nnratio,
                         int checkOri, cudaStream_t st) {
  k_bow_search_init<<<(capF + 255) / 256, 256, 0, st>>>(F, Q, capF);
  k_bow_group<<<1, 1024, (size_t)pow2_at_least(nKF) * 8, st>>>(Q.kfNode, nKF, Q.kfIdxSorted, Q.kfNodeSorted, Q.kfGStart, Q.kfMeta,
                                                            pow2_at_least(nKF));
  k_bow_search<<<(nKF + 3) / 4, 128, 0, st>>>(S, F, Q, nnratio, checkOri);
  if (checkOri) k_bow_finish<<<1, 1024, 0, st>>>(F, Q);
  return checkOri ? 4 : 3;
}

int ft_vocabulary_device(const ft_vocabulary* v) { return v->device; }

// ---- vocabulary construction ----
static void voc_free(ft_vocabulary* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  for (void* p : v->allocs) cudaFree(p);
  for (void* p : v->tAllocs) cudaFree(p);
  if (v->stream) cudaStreamDestroy(v->stream);
  delete v;
}

static ft_status voc_build(int device, int k, int L, int scoring, int weighting, int n, const int* parent, const uint8_t* isLeaf,
                           const uint8_t* desc, const double* weight, ft_vocabulary** out) {
  if (!out || n < 0 || (n > 0 && (!parent || !isLeaf || !desc || !weight)) || scoring < 0 || scoring > 5 || weighting < 0 ||
      weighting > 3 || L < 0) {
    ft_internal_set_err("ft_vocabulary_create: bad argument"); return FT_ERR_INVALID;
  }
  const int N = n + 1;   // with the root
  for (int i = 0; i < n; i++)
    if (parent[i] < 0 || parent[i] > i) {   // a node's parent must already exist (the reference indexes m_nodes[pid])
      ft_internal_set_err("ft_vocabulary_create: parent of a node must precede it"); return FT_ERR_INVALID;
    }
  std::vector<int> nChild(N, 0), firstChild(N, 0), fill(N, 0), childNode(n, 0), wordId(N, 0);
  for (int i = 0; i < n; i++) nChild[parent[i]]++;
  for (int i = 1; i < N; i++) firstChild[i] = firstChild[i - 1] + nChild[i - 1];
  std::vector<uint8_t> childDesc((size_t)n * 32 + 32, 0);
  std::vector<double> w(N, 0.0), wordWeight;
  int nWords = 0;
  for (int i = 0; i < n; i++) {
    const int nid = i + 1, pid = parent[i];
    const int slot = firstChild[pid] + fill[pid]++;             // children in order of appearance (push_back)
    childNode[slot] = nid;
    memcpy(&childDesc[(size_t)slot * 32], desc + (size_t)i * 32, 32);
    w[nid] = weight[i];
    if (isLeaf[i]) { wordId[nid] = nWords++; wordWeight.push_back(weight[i]); }
  }
  // A childless node that was not flagged as a leaf keeps Node()'s word_id = 0 in the reference and is still where a
  // descent ends. With weight 0 (the phantom node of loadFromTextFile) its features are stopped and never reach the
  // BowVector; with a positive weight the reference would credit word 0 with a foreign weight -- rejected.
  for (int i = 0; i < n; i++)
    if (!isLeaf[i] && nChild[i + 1] == 0 && weight[i] > 0) {
      ft_internal_set_err("ft_vocabulary_create: childless node with a weight but without the leaf flag"); return FT_ERR_INVALID;
    }
  if (wordWeight.empty()) wordWeight.push_back(0.0);
  BCK(cudaSetDevice(device));
  ft_vocabulary* v = new ft_vocabulary();
  v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nNodes = N; v->nWords = nWords;
  int *dFirst = nullptr, *dN = nullptr, *dChild = nullptr, *dWord = nullptr;
  uint4* dDesc = nullptr; double *dW = nullptr, *dWW = nullptr;
  cudaError_t e = balloc(v->allocs, &dFirst, N);
  if (e == cudaSuccess) e = balloc(v->allocs, &dN, N);
  if (e == cudaSuccess) e = balloc(v->allocs, &dChild, n);
  if (e == cudaSuccess) e = balloc(v->allocs, &dWord, N);
  if (e == cudaSuccess) e = balloc(v->allocs, &dDesc, (size_t)2 * n + 2);
  if (e == cudaSuccess) e = balloc(v->allocs, &dW, N);
  if (e == cudaSuccess) e = balloc(v->allocs, &dWW, wordWeight.size());
  if (e == cudaSuccess) e = cudaMemcpy(dFirst, firstChild.data(), sizeof(int) * N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dN, nChild.data(), sizeof(int) * N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n) e = cudaMemcpy(dChild, childNode.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dWord, wordId.data(), sizeof(int) * N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n) e = cudaMemcpy(dDesc, childDesc.data(), (size_t)n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dW, w.data(), sizeof(double) * N, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dWW, wordWeight.data(), sizeof(double) * wordWeight.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    ft_internal_set_err((std::string("ft_vocabulary_create: ") + cudaGetErrorString(e)).c_str());
    voc_free(v);
    return FT_ERR_CUDA;
  }
  v->D.L = L; v->D.nNodes = N; v->D.scoring = scoring; v->D.weighting = weighting;
  v->D.firstChild = dFirst; v->D.nChild = dN; v->D.childNode = dChild; v->D.childDesc = dDesc; v->D.weight = dW; v->D.wordId = dWord;
  v->D.wordWeight = dWW;
  *out = v;
  return FT_OK;
}

extern "C" ft_status ft_vocabulary_create(int device_id, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                                          const uint8_t* is_leaf, const uint8_t* desc, const double* weight,
                                          ft_vocabulary** out) {
  return voc_build(device_id, k, L, scoring, weighting, n_nodes, parent, is_leaf, desc, weight, out);
}

// loadFromTextFile (TemplatedVocabulary.h:1338-1423) without iostreams: header `k L scoring weighting`, then one node
// per line `parent isLeaf d0 .. d31 weight`. The reference's `while(!f.eof())` loop also makes a node out of the empty
// line that follows a trailing newline: every extraction fails, so the node keeps the previous line's parent and leaf
// flag (the reference's uninitialised locals, as gcc lays them out), an unwritten descriptor (zero here) and weight 0,
// i.e. a stopped extra word under the last inner node. Reproduced, so that word ids and descents agree with a
// vocabulary loaded by the reference.
extern "C" ft_status ft_vocabulary_load_text(int device_id, const char* path, ft_vocabulary** out) {
  if (!path || !out) { ft_internal_set_err("ft_vocabulary_load_text: null argument"); return FT_ERR_INVALID; }
  FILE* f = fopen(path, "rb");
  if (!f) { ft_internal_set_err((std::string("ft_vocabulary_load_text: cannot open ") + path).c_str()); return FT_ERR_INVALID; }
  std::string buf;
  {
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.append(chunk, got);
    fclose(f);
  }
  size_t pos = buf.find('\n');
  const std::string header = buf.substr(0, pos == std::string::npos ? buf.size() : pos);
  int k = -1, L = -1, n1 = -1, n2 = -1;
  if (sscanf(header.c_str(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 ||
      n2 < 0 || n2 > 3) {
    ft_internal_set_err("ft_vocabulary_load_text: not a DBoW2 text vocabulary"); return FT_ERR_INVALID;
  }
  std::vector<int> parent; std::vector<uint8_t> leaf, desc; std::vector<double> weight;
  int pid = 0, nIsLeaf = 0;
  bool more = pos != std::string::npos;   // getline hit eof on the header: the node loop does not run
  size_t p = more ? pos + 1 : buf.size();
  while (more) {
    size_t e = buf.find('\n', p);
    more = e != std::string::npos;        // no newline: this getline reaches eof, the loop ends after this node
    const char* s = buf.c_str() + p;
    const char* end = buf.c_str() + (more ? e : buf.size());
    p = more ? e + 1 : buf.size();
    uint8_t d[32] = {0};
    double w = 0.0;
    bool ok = true;
    auto next_long = [&](long& v) {
      if (!ok) return;
      while (s < end && isspace((unsigned char)*s)) s++;
      char* q = nullptr;
      const long t = strtol(s, &q, 10);
      if (q == s || q > end) { ok = false; return; }
      v = t; s = q;
    };
    long v = 0;
    next_long(v); if (ok) pid = (int)v;
    next_long(v); if (ok) nIsLeaf = (int)v;
    for (int i = 0; i < 32 && ok; i++) { next_long(v); if (ok) d[i] = (uint8_t)v; }
    if (ok) {
      while (s < end && isspace((unsigned char)*s)) s++;
      char* q = nullptr;
      const double t = strtod(s, &q);
      if (q != s && q <= end) w = t;
    }
    if (pid < 0 || pid > (int)parent.size()) {
      ft_internal_set_err("ft_vocabulary_load_text: node refers to a parent that does not exist yet"); return FT_ERR_INVALID;
    }
    parent.push_back(pid); leaf.push_back(nIsLeaf > 0 ? 1 : 0); weight.push_back(w);
    desc.insert(desc.end(), d, d + 32);
  }
  const int n = (int)parent.size();
  return voc_build(device_id, k, L, n1, n2, n, parent.data(), leaf.data(), desc.data(), weight.data(), out);
}

extern "C" ft_status ft_vocabulary_destroy(ft_vocabulary* v) { voc_free(v); return FT_OK; }

extern "C" ft_status ft_vocabulary_info(ft_vocabulary* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
  if (!v) { ft_internal_set_err("ft_vocabulary_info: null vocabulary"); return FT_ERR_INVALID; }
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  if (n_nodes) *n_nodes = v->nNodes;
  if (n_words) *n_words = v->nWords;
  return FT_OK;
}

// transform of HOST descriptors (a KeyFrame's mDescriptors, KeyFrame::ComputeBoW, src/KeyFrame.cc:145-155)
extern "C" ft_status ft_vocabulary_transform(ft_vocabulary* v, const uint8_t* desc, int n, int levelsup, int* word_id,
                                             int* node_id, uint32_t* bow_ids, double* bow_vals, int bow_cap, int* n_bow) {
  if (!v || n < 0 || (n > 0 && !desc) || levelsup < 0) { ft_internal_set_err("ft_vocabulary_transform: bad argument"); return FT_ERR_INVALID; }
  if (n > FT_BOW_MAX_FEATURES) { ft_internal_set_err("ft_vocabulary_transform: more than 16384 features"); return FT_ERR_CAPACITY; }
  if (n_bow) *n_bow = 0;
  if (n == 0) return FT_OK;
  std::lock_guard<std::mutex> lk(v->mu);
  BCK(cudaSetDevice(v->device));
  if (n > v->tCap) {
    for (void* p : v->tAllocs) cudaFree(p);
    v->tAllocs.clear(); v->T = FtBowFrame(); v->tDesc = nullptr; v->tCap = 0;
    const int cap = pow2_at_least(n);
    BCK(ft_bow_frame_alloc(&v->T, cap, v->tAllocs));
    BCK(balloc(v->tAllocs, &v->tDesc, (size_t)cap * 32));
    v->tCap = cap;
  }
  BCK(cudaMemcpyAsync(v->tDesc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
  FtBowSource S = {};
  S.desc0 = v->tDesc; S.nFixed = n;
  ft_launch_bow_transform(v, S, v->T, n, levelsup, v->stream);
  BCK(cudaGetLastError());
  int meta[4] = {0, 0, 0, 0};
  BCK(cudaMemcpyAsync(meta, v->T.meta, sizeof(meta), cudaMemcpyDeviceToHost, v->stream));
  if (word_id) BCK(cudaMemcpyAsync(word_id, v->T.word, sizeof(int) * n, cudaMemcpyDeviceToHost, v->stream));
  if (node_id) BCK(cudaMemcpyAsync(node_id, v->T.node, sizeof(int) * n, cudaMemcpyDeviceToHost, v->stream));
  BCK(cudaStreamSynchronize(v->stream));
  const int nd = meta[2];
  if (n_bow) *n_bow = nd;
  if ((bow_ids || bow_vals) && bow_cap < nd) { ft_internal_set_err("ft_vocabulary_transform: bow capacity too small"); return FT_ERR_INVALID; }
  if (bow_ids && nd) BCK(cudaMemcpy(bow_ids, v->T.bowIds, sizeof(uint32_t) * nd, cudaMemcpyDeviceToHost));
  if (bow_vals && nd) BCK(cudaMemcpy(bow_vals, v->T.bowVals, sizeof(double) * nd, cudaMemcpyDeviceToHost));
  return FT_OK;
}
