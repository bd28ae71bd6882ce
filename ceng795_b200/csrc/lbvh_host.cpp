// lbvh_host.cpp — host-side entry to the index arithmetic of csrc/lbvh.cuh, so that the hierarchy rule the device
// kernels apply can be checked without a GPU (tests/test_abi.py).
#include "lbvh.cuh"
#include "api_guard.h"

extern "C" {
// ranges[3*i .. 3*i+2] = {first, last, split} of internal node i, for sorted keys (duplicates allowed)
int b200rt_lbvh_ranges(const uint64_t* keys, int n, int32_t* ranges) try {
  if (!keys || !ranges || n < 2) return B200RT_E_INVALID;
  for (int i = 0; i < n - 1; i++) {
    int f, l, s;
    lbvh_range(keys, n, i, &f, &l, &s);
    ranges[3 * i] = f; ranges[3 * i + 1] = l; ranges[3 * i + 2] = s;
  }
  return B200RT_OK;
} B200RT_CATCH
uint64_t b200rt_lbvh_morton63(float x, float y, float z) { return lbvh_morton63(x, y, z); }
}
