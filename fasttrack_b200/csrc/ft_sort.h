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

}  // namespace ftsort
