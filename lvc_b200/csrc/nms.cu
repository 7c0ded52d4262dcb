// Generic nms / batched_nms entry point (arbitrary n, arbitrary class ids): the detectron2.layers drop-in.
// Pipeline (all on `stream`, no host sync):
//   1. min/max of the coordinates (trick offset needs max; negative coordinates disable class segmentation
//      in trick mode because offset classes may then overlap, exactly as in the reference)
//   2. stable sort by descending score  -> global rank            (cub::DeviceRadixSort, 32-bit keys)
//   3. stable sort by (class, rank)     -> class segments         (cub::DeviceRadixSort, 64-bit keys)
//   4. segment starts by flag + exclusive scan                    (cub::DeviceScan)
//   5. one CTA per segment: greedy NMS (nms_core.cuh), kept flags scattered to rank order
//   6. exclusive scan of kept flags in rank order -> keep[] (int64 original indices), num_keep
// CUB (the CUDA toolkit's header-only primitives) is used for the two radix sorts and two scans of this
// generic op only; the engine's RPN / box-head paths (rpn.cu, detections.cu) sort inside their own CTAs.
#include <cub/cub.cuh>

#include "nms_core.cuh"

namespace lvcb200 {

struct NmsWs {
  size_t off_minmax, off_keys1, off_keys1b, off_vals1, off_vals1b, off_rank, off_keys2, off_keys2b, off_vals2, off_vals2b,
      off_segflag, off_segscan, off_segstart, off_kept, off_flagrank, off_flagscan, off_byteflags, off_cub, cub_bytes, total;
};

static NmsWs nms_layout(int64_t n) {
  NmsWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  w.off_minmax = take(16);
  w.off_keys1 = take(4 * n); w.off_keys1b = take(4 * n);
  w.off_vals1 = take(4 * n); w.off_vals1b = take(4 * n);
  w.off_rank = take(4 * n);
  w.off_keys2 = take(8 * n); w.off_keys2b = take(8 * n);
  w.off_vals2 = take(4 * n); w.off_vals2b = take(4 * n);
  w.off_segflag = take(4 * (n + 1)); w.off_segscan = take(4 * (n + 1)); w.off_segstart = take(4 * (n + 2));
  w.off_kept = take(5 * 4 * n);
  w.off_flagrank = take(4 * (n + 1)); w.off_flagscan = take(4 * (n + 1));
  w.off_byteflags = take(n);
  size_t b1 = 0, b2 = 0, b3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (uint32_t*)nullptr, (uint32_t*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, b3, (int*)nullptr, (int*)nullptr, (int)n + 1);
  w.cub_bytes = b1 > b2 ? b1 : b2;
  if (b3 > w.cub_bytes) w.cub_bytes = b3;
  w.off_cub = take(w.cub_bytes);
  w.total = o;
  return w;
}

__global__ void nms_prep_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int n,
                                uint32_t* __restrict__ keys1, int* __restrict__ vals1, uint32_t* __restrict__ minmax) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t mx = 0u, mn = 0xffffffffu;
  if (i < n) {
    float4 b = reinterpret_cast<const float4*>(boxes)[i];
    uint32_t k0 = float_to_ordered(b.x), k1 = float_to_ordered(b.y), k2 = float_to_ordered(b.z), k3 = float_to_ordered(b.w);
    mx = max(max(k0, k1), max(k2, k3));
    mn = min(min(k0, k1), min(k2, k3));
    keys1[i] = ~float_to_ordered(scores[i]);  // ascending sort of ~key == descending score, stable on ties
    vals1[i] = i;
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(&minmax[0], mx); atomicMin(&minmax[1], mn); }
}

__global__ void nms_init_kernel(uint32_t* minmax) { minmax[0] = 0u; minmax[1] = 0xffffffffu; }

// mode resolved on device; one_segment = plain nms, or trick with negative coordinates
__device__ __forceinline__ bool nms_single_segment(const int64_t* idxs, int mode, const uint32_t* minmax) {
  if (idxs == nullptr) return true;
  if (mode == 0 && ordered_to_float(minmax[1]) < 0.f) return true;
  return false;
}

__global__ void nms_keys2_kernel(const int* __restrict__ order1, const int64_t* __restrict__ idxs, int n, int mode,
                                 const uint32_t* __restrict__ minmax, int* __restrict__ rank,
                                 uint64_t* __restrict__ keys2, int* __restrict__ vals2) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int i = order1[r];
  rank[i] = r;
  bool single = nms_single_segment(idxs, mode, minmax);
  uint64_t cls = single ? 0ull : (uint64_t)(uint32_t)idxs[i];
  keys2[r] = (cls << 32) | (uint32_t)r;  // written in rank order; sort is by full key anyway
  vals2[r] = i;
}

__global__ void nms_segflag_kernel(const uint64_t* __restrict__ keys2s, int n, int* __restrict__ segflag) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  if (p == n) { segflag[p] = 1; return; }  // sentinel: closes the last segment
  segflag[p] = (p == 0 || (keys2s[p] >> 32) != (keys2s[p - 1] >> 32)) ? 1 : 0;
}

__global__ void nms_segstart_kernel(const int* __restrict__ segflag, const int* __restrict__ segscan, int n,
                                    int* __restrict__ segstart) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  if (segflag[p]) segstart[segscan[p]] = p;   // segstart[s] = first sorted position of segment s; segstart[nseg] = n
  if (p == n) segstart[n + 1] = segscan[p];   // number of segments
}

__global__ void __launch_bounds__(256)
nms_segments_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ idxs, const int* __restrict__ order2,
                    const int* __restrict__ rank, const int* __restrict__ segstart, int n, float thr, int mode,
                    const uint32_t* __restrict__ minmax, float* __restrict__ kept_ws, unsigned char* __restrict__ byteflags,
                    int* __restrict__ flag_by_rank) {
  __shared__ NmsShared sh;
  const int nseg = segstart[n + 1];
  const bool trick = (mode == 0) && idxs != nullptr;
  const float mul = __fadd_rn(ordered_to_float(minmax[0]), 1.0f);
  for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
    const int beg = segstart[s], end = segstart[s + 1], len = end - beg;
    float* kx1 = kept_ws + (size_t)beg * 5;
    float* ky1 = kx1 + len; float* kx2 = ky1 + len; float* ky2 = kx2 + len; float* kar = ky2 + len;
    auto get = [&](int j, float& x1, float& y1, float& x2, float& y2) {
      int i = order2[beg + j];
      float4 b = reinterpret_cast<const float4*>(boxes)[i];
      float off = trick ? __fmul_rn((float)idxs[i], mul) : 0.f;
      x1 = __fadd_rn(b.x, off); y1 = __fadd_rn(b.y, off); x2 = __fadd_rn(b.z, off); y2 = __fadd_rn(b.w, off);
      return true;
    };
    unsigned char* fl = byteflags + beg;
    segment_nms(sh, get, len, thr, kx1, ky1, kx2, ky2, kar, fl);
    __syncthreads();
    for (int j = threadIdx.x; j < len; j += blockDim.x) flag_by_rank[rank[order2[beg + j]]] = fl[j];
    __syncthreads();
  }
}

__global__ void nms_emit_kernel(const int* __restrict__ order1, const int* __restrict__ flag_by_rank,
                                const int* __restrict__ flagscan, int n, int64_t* __restrict__ keep,
                                int64_t* __restrict__ num_keep) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  if (r == n) { *num_keep = flagscan[n]; return; }
  if (flag_by_rank[r]) keep[flagscan[r]] = order1[r];
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" size_t lvcb200_batched_nms_workspace(int64_t n) {
  if (n <= 0) return 256;
  NmsWs w = nms_layout(n);
  return w.total;
}

extern "C" int lvcb200_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                                   float iou_threshold, int mode, int64_t* keep, int64_t* num_keep, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  LVC_REQUIRE(n >= 0 && n < (1ll << 30), "batched_nms: n out of range");
  LVC_REQUIRE(num_keep != nullptr, "batched_nms: num_keep is NULL");
  if (n == 0) { LVC_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int64_t), s)); return 0; }
  LVC_REQUIRE(boxes && scores && keep && workspace, "batched_nms: NULL pointer");
  LVC_REQUIRE(((uintptr_t)boxes % 16) == 0, "batched_nms: boxes must be 16-byte aligned");
  LVC_REQUIRE(mode >= -1 && mode <= 1, "batched_nms: mode must be -1, 0 or 1");
  if (workspace_bytes < lvcb200_batched_nms_workspace(n)) return set_error(LVCB200_EWORKSPACE, "batched_nms: workspace too small");
  if (mode < 0) mode = reference_cuda_nms_mode(n);
  NmsWs w = nms_layout(n);
  char* ws = (char*)workspace;
  uint32_t* minmax = (uint32_t*)(ws + w.off_minmax);
  uint32_t *keys1 = (uint32_t*)(ws + w.off_keys1), *keys1b = (uint32_t*)(ws + w.off_keys1b);
  int *vals1 = (int*)(ws + w.off_vals1), *vals1b = (int*)(ws + w.off_vals1b);
  int* rank = (int*)(ws + w.off_rank);
  uint64_t *keys2 = (uint64_t*)(ws + w.off_keys2), *keys2b = (uint64_t*)(ws + w.off_keys2b);
  int *vals2 = (int*)(ws + w.off_vals2), *vals2b = (int*)(ws + w.off_vals2b);
  int *segflag = (int*)(ws + w.off_segflag), *segscan = (int*)(ws + w.off_segscan), *segstart = (int*)(ws + w.off_segstart);
  float* kept = (float*)(ws + w.off_kept);
  int *flagrank = (int*)(ws + w.off_flagrank), *flagscan = (int*)(ws + w.off_flagscan);
  void* cubtmp = ws + w.off_cub;
  size_t cb = w.cub_bytes;
  const int T = 256;
  const unsigned B = (unsigned)ceil_div64(n + 1, T);
  int nn = (int)n;
  unsigned char* byteflags = (unsigned char*)(ws + w.off_byteflags);
  nms_init_kernel<<<1, 1, 0, s>>>(minmax);
  nms_prep_kernel<<<B, T, 0, s>>>(boxes, scores, nn, keys1, vals1, minmax);
  LVC_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp, cb, keys1, keys1b, vals1, vals1b, nn, 0, 32, s));
  nms_keys2_kernel<<<B, T, 0, s>>>(vals1b, idxs, nn, mode, minmax, rank, keys2, vals2);
  cb = w.cub_bytes;
  LVC_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp, cb, keys2, keys2b, vals2, vals2b, nn, 0, 64, s));
  nms_segflag_kernel<<<B, T, 0, s>>>(keys2b, nn, segflag);
  cb = w.cub_bytes;
  LVC_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp, cb, segflag, segscan, nn + 1, s));
  nms_segstart_kernel<<<B, T, 0, s>>>(segflag, segscan, nn, segstart);
  LVC_CUDA(cudaMemsetAsync(flagrank, 0, sizeof(int) * (n + 1), s));
  nms_segments_kernel<<<kNumSMs * 4, 256, 0, s>>>(boxes, idxs, vals2b, rank, segstart, nn, iou_threshold, mode, minmax, kept, byteflags, flagrank);
  cb = w.cub_bytes;
  LVC_CUDA(cub::DeviceScan::ExclusiveSum(cubtmp, cb, flagrank, flagscan, nn + 1, s));
  nms_emit_kernel<<<B, T, 0, s>>>(vals1b, flagrank, flagscan, nn, keep, num_keep);
  g_launch_count.fetch_add(11, std::memory_order_relaxed);
  return check_launch("batched_nms");
}
