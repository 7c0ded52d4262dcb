// Shared helpers for liblvcb200 (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/lvcb200.h"

namespace lvcb200 {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;

inline int set_error(int code, const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
  return code;
}

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define LVC_REQUIRE(cond, msg) \
  do { if (!(cond)) return ::lvcb200::set_error(LVCB200_EINVAL, msg); } while (0)

#define LVC_CUDA(call) \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(::lvcb200::g_last_error, sizeof(::lvcb200::g_last_error), "%s: %s", #call, cudaGetErrorString(e_)); \
    return (int)e_; } } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// order-preserving float -> uint32 (larger float => larger key); NaN with sign bit 0 sorts above +inf
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// IoU > thr decision with the exact operation sequence of torchvision's CPU nms kernel
// (separately rounded mul / add / sub / div; no FMA contraction).
__device__ __forceinline__ bool iou_gt(float ax1, float ay1, float ax2, float ay2, float aarea,
                                       float bx1, float by1, float bx2, float by2, float barea, float thr) {
  float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
  float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
  float w = __fsub_rn(xx2, xx1); w = (w > 0.f) ? w : 0.f;
  float h = __fsub_rn(yy2, yy1); h = (h > 0.f) ? h : 0.f;
  float inter = __fmul_rn(w, h);
  float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
  return ovr > thr;
}
__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2) {
  return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}

}  // namespace lvcb200
