// ft_sort.h -- introsort with the exact element movement of libstdc++'s std::sort.
//
// ORBextractor::DistributeOctTree (reference src/ORBextractor.cc:805) sorts
// (size, node*) pairs with a comparator that leaves ties (equal size and UL.x), and then
// splits nodes from the back with an early break (:851), so the order std::sort happens
// to leave equivalent elements in is observable in the final keypoint list. To stay
// bit-exact on the device the sort has to move elements the way libstdc++ does:
// introsort (median-of-3 pivot taken from first+1 / mid / last-1, unguarded Hoare
// partition, threshold 16, heapsort after 2*floor(log2 n) levels) followed by the final
// insertion sort. This header restates that published algorithm (bits/stl_algo.h,
// bits/stl_heap.h) for a plain array, callable from host and device.
// tests/test_sort.py checks it against std::sort on tie-heavy inputs.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FT_HD __host__ __device__ __forceinline__
#else
#define FT_HD inline
#endif

namespace ftsort {

// Element = 64-bit word; only the high 32 bits (the key) are compared: key = size<<12 | UL.x.
typedef unsigned long long elem_t;
FT_HD bool less(elem_t a, elem_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
FT_HD void swp(elem_t* a, elem_t* b) { elem_t t = *a; *a = *b; *b = t; }

FT_HD void push_heap(elem_t* first, int holeIndex, int topIndex, elem_t value) {
  int parent = (holeIndex - 1) / 2;
  while (holeIndex > topIndex && less(first[parent], value)) {
    first[holeIndex] = first[parent];
    holeIndex = parent;
    parent = (holeIndex - 1) / 2;
  }
  first[holeIndex] = value;
}

FT_HD void adjust_heap(elem_t* first, int holeIndex, int len, elem_t value) {
  const int topIndex = holeIndex;
  int secondChild = holeIndex;
  while (secondChild < (len - 1) / 2) {
    secondChild = 2 * (secondChild + 1);
    if (less(first[secondChild], first[secondChild - 1])) secondChild--;
    first[holeIndex] = first[secondChild];
    holeIndex = secondChild;
  }
  if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
    secondChild = 2 * (secondChild + 1);
    first[holeIndex] = first[secondChild - 1];
    holeIndex = secondChild - 1;
  }
  push_heap(first, holeIndex, topIndex, value);
}

FT_HD void heap_sort(elem_t* first, int len) {  // std::__partial_sort(first, last, last)
  if (len >= 2) {
    int parent = (len - 2) / 2;
    while (true) {
      elem_t value = first[parent];
      adjust_heap(first, parent, len, value);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) {
    --last;
    elem_t value = first[last];
    first[last] = first[0];
    adjust_heap(first, 0, last, value);
  }
}

FT_HD void move_median_to_first(elem_t* result, elem_t* a, elem_t* b, elem_t* c) {
  if (less(*a, *b)) {
    if (less(*b, *c)) swp(result, b);
    else if (less(*a, *c)) swp(result, c);
    else swp(result, a);
  } else if (less(*a, *c)) swp(result, a);
  else if (less(*b, *c)) swp(result, c);
  else swp(result, b);
}

FT_HD int unguarded_partition(elem_t* base, int first, int last, int pivot) {
  while (true) {
    while (less(base[first], base[pivot])) ++first;
    --last;
    while (less(base[pivot], base[last])) --last;
    if (!(first < last)) return first;
    swp(&base[first], &base[last]);
    ++first;
  }
}

FT_HD void unguarded_linear_insert(elem_t* base, int last) {
  elem_t val = base[last];
  int next = last - 1;
  while (less(val, base[next])) {
    base[last] = base[next];
    last = next;
    --next;
  }
  base[last] = val;
}

FT_HD void insertion_sort(elem_t* base, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (less(base[i], base[first])) {
      elem_t val = base[i];
      for (int k = i; k > first; --k) base[k] = base[k - 1];
      base[first] = val;
    } else {
      unguarded_linear_insert(base, i);
    }
  }
}

FT_HD int floor_log2(int n) {
  int l = 0;
  while (n > 1) { n >>= 1; ++l; }
  return l;
}

// std::sort(base, base + n, less). The recursion of __introsort_loop (recurse right, loop
// left) is unrolled with an explicit stack of (first, last, depth) ranges.
FT_HD void sort(elem_t* base, int n) {
  if (n <= 1) return;
  const int kThreshold = 16;
  int stackFirst[64], stackLast[64], stackDepth[64];
  int sp = 0;
  stackFirst[0] = 0; stackLast[0] = n; stackDepth[0] = 2 * floor_log2(n);
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = stackFirst[sp], last = stackLast[sp], depth = stackDepth[sp];
    // one activation of __introsort_loop(first, last, depth). The right-hand recursive calls
    // complete before the loop continues on the left part, but the two ranges are disjoint,
    // so deferring the right part does not change any element movement.
    while (last - first > kThreshold) {
      if (depth == 0) {
        heap_sort(base + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first(&base[first], &base[first + 1], &base[mid], &base[last - 1]);
      const int cut = unguarded_partition(base, first + 1, last, first);
      stackFirst[sp] = cut; stackLast[sp] = last; stackDepth[sp] = depth;
      ++sp;
      last = cut;
    }
  }
  if (n > kThreshold) {
    insertion_sort(base, 0, kThreshold);
    for (int i = kThreshold; i != n; ++i) unguarded_linear_insert(base, i);
  } else {
    insertion_sort(base, 0, n);
  }
}

#ifdef __CUDACC__
// ---- warp-cooperative version (device only) -----------------------------------------------------------
// Same element movement as sort() above, executed by one full warp:
//  * __unguarded_partition is data-parallel. Let A be the ascending positions in (first, last) whose element is
//    not less than the pivot (where the left scan stops) and B the descending positions in [first, last) whose
//    element is not greater (where the right scan stops; the pivot at `first` is the sentinel). Positions strictly
//    between two stops are never touched, so iteration k of the sequential loop stops at
//        i_k = (k == 0) ? A[0] : min(A[k], B[k-1]),   j_k = (k == 0) ? B[0] : max(B[k], A[k-1])
//    (the min/max account for the elements the previous swap put at B[k-1] / A[k-1]); it swaps while i_k < j_k,
//    and when the loop continues i_k = A[k], j_k = B[k]. The number of swaps K is the first k with !(i_k < j_k),
//    the returned cut is i_K, and the K swaps A[k] <-> B[k] touch disjoint positions.
//  * __final_insertion_sort is an insertion sort of the whole array, i.e. a stable sort of the arrangement the
//    introsort loop leaves; stable_rank() computes that permutation in parallel.
__device__ __forceinline__ int warp_partition(elem_t* base, int first, int last, int* posA, int* posB, int lane) {
  const uint32_t p = (uint32_t)(base[first] >> 32);
  const unsigned lt = (1u << lane) - 1u;
  int nA = 0, nB = 0;
  for (int c0 = first; c0 < last; c0 += 32) {
    const int i = c0 + lane;
    const bool in = i < last;
    const uint32_t k = in ? (uint32_t)(base[i] >> 32) : 0u;
    const bool fa = in && i > first && k >= p;
    const bool fb = in && k <= p;
    const unsigned ma = __ballot_sync(0xFFFFFFFFu, fa), mb = __ballot_sync(0xFFFFFFFFu, fb);
    if (fa) posA[nA + __popc(ma & lt)] = i;
    if (fb) posB[nB + __popc(mb & lt)] = i;
    nA += __popc(ma); nB += __popc(mb);
  }
  __syncwarp();
  int K = -1, cut = first + 1;
  for (int k0 = 0; K < 0; k0 += 32) {
    const int k = k0 + lane;
    const int Ak = k < nA ? posA[k] : 0x7FFFFFFF;
    const int Bk = k < nB ? posB[nB - 1 - k] : -1;
    int ik = Ak, jk = Bk;
    if (k > 0) {
      const int Bp = (k - 1) < nB ? posB[nB - k] : -1;
      const int Ap = (k - 1) < nA ? posA[k - 1] : 0x7FFFFFFF;
      ik = min(Ak, Bp); jk = max(Bk, Ap);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, !(ik < jk));
    if (m) { const int l = __ffs(m) - 1; K = k0 + l; cut = __shfl_sync(0xFFFFFFFFu, ik, l); }
  }
  for (int k = lane; k < K; k += 32) swp(&base[posA[k]], &base[posB[nB - 1 - k]]);
  __syncwarp();
  return cut;
}

// introsort loop only (no final insertion sort); posA/posB: scratch of n ints each. Call with a full warp.
// The explicit stack of pending right-hand ranges lives in registers, one entry per lane (its depth is bounded by
// the depth limit 2*floor(log2 n) < 32), so the loop touches no local memory.
__device__ __forceinline__ void warp_introsort_loop(elem_t* base, int n, int* posA, int* posB) {
  if (n <= 16) return;
  const int lane = threadIdx.x & 31;
  int myFirst = 0, myLast = 0, myDepth = 0;     // stack entry held by this lane
  int sp = 0;
  int first = 0, last = n, depth = 2 * floor_log2(n);
  bool have = true;
  while (have) {
    while (last - first > 16) {
      if (depth == 0) {
        if (lane == 0) heap_sort(base + first, last - first);
        __syncwarp();
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      if (lane == 0) move_median_to_first(&base[first], &base[first + 1], &base[mid], &base[last - 1]);
      __syncwarp();
      const int cut = warp_partition(base, first, last, posA, posB, lane);
      if (lane == sp) { myFirst = cut; myLast = last; myDepth = depth; }   // push the right-hand range
      ++sp;
      last = cut;
    }
    if (sp > 0) {
      --sp;
      first = __shfl_sync(0xFFFFFFFFu, myFirst, sp);
      last = __shfl_sync(0xFFFFFFFFu, myLast, sp);
      depth = __shfl_sync(0xFFFFFFFFu, myDepth, sp);
    } else {
      have = false;
    }
  }
}

// Block-cooperative introsort loop: the two halves a partition leaves are disjoint, so all pending ranges of one
// recursion level are processed concurrently, one warp per range (same element movement as the sequential loop,
// which only fixes the order of operations inside a range). The critical path is the recursion depth instead of
// the number of ranges. `queue` needs 6*(n/16+2) ints; posA/posB n ints each (a range uses its own slice).
// Call with every thread of the block (contains __syncthreads).
__device__ __forceinline__ void cta_introsort_loop(elem_t* base, int n, int* posA, int* posB, int* queue) {
  if (n <= 16) return;                       // uniform over the block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nWarps = blockDim.x >> 5;
  const int Q = n / 16 + 2;
  int* qf[2] = {queue, queue + 3 * Q};
  __shared__ int sCnt[2];
  if (tid == 0) { qf[0][0] = 0; qf[0][Q] = n; qf[0][2 * Q] = 2 * floor_log2(n); sCnt[0] = 1; sCnt[1] = 0; }
  __syncthreads();
  int cur = 0;
  for (;;) {
    const int cnt = sCnt[cur];
    if (cnt == 0) break;
    int* qc = qf[cur];
    int* qn = qf[cur ^ 1];
    for (int r = warp; r < cnt; r += nWarps) {
      const int first = qc[r], last = qc[Q + r];
      int depth = qc[2 * Q + r];
      if (depth == 0) {
        if (lane == 0) heap_sort(base + first, last - first);
        __syncwarp();
        continue;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      if (lane == 0) move_median_to_first(&base[first], &base[first + 1], &base[mid], &base[last - 1]);
      __syncwarp();
      const int cut = warp_partition(base, first, last, posA + first, posB + first, lane);
      if (lane == 0) {
        if (last - cut > 16) { const int k = atomicAdd(&sCnt[cur ^ 1], 1); qn[k] = cut; qn[Q + k] = last; qn[2 * Q + k] = depth; }
        if (cut - first > 16) { const int k = atomicAdd(&sCnt[cur ^ 1], 1); qn[k] = first; qn[Q + k] = cut; qn[2 * Q + k] = depth; }
      }
    }
    __syncthreads();
    if (tid == 0) sCnt[cur] = 0;
    cur ^= 1;
    __syncthreads();
  }
}

// out[rank] = in[i] with rank = #(smaller keys) + #(equal keys before i): the stable sort an insertion sort yields.
// Call with `nthreads` cooperating threads (tid in [0, nthreads)); caller synchronises before and after.
__device__ __forceinline__ void stable_rank(const elem_t* in, elem_t* out, int n, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) {
    const elem_t e = in[i];
    const uint32_t k = (uint32_t)(e >> 32);
    int r = 0;
    for (int j = 0; j < n; j++) {
      const uint32_t kj = (uint32_t)(in[j] >> 32);
      r += (kj < k) || (kj == k && j < i);
    }
    out[r] = e;
  }
}
#endif

}  // namespace ftsort
