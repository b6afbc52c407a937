// Host harness: ftsort::sort (the device introsort replica) vs libstdc++ std::sort.
#include <algorithm>
#include <cstdint>
#include "../../fasttrack_b200/csrc/ft_sort.h"
extern "C" {
void harness_ftsort(uint64_t* a, int n) { ftsort::sort((ftsort::elem_t*)a, n); }
void harness_stdsort(uint64_t* a, int n) {
  std::sort(a, a + n, [](const uint64_t& x, const uint64_t& y) { return (uint32_t)(x >> 32) < (uint32_t)(y >> 32); });
}
}
