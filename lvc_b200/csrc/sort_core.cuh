// In-CTA helpers: bitonic sort of 64-bit keys in shared memory, descending.
#pragma once
#include "common.cuh"

namespace lvcb200 {

// sorts keys[0..n_pow2) descending; n_pow2 power of two; all threads of the CTA must call.
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// number of entries in a descending-sorted array with key > q  (strictly greater)
__device__ __forceinline__ int count_greater_desc(const float* a, int n, float q) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] > q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace lvcb200
