// Greedy IoU-NMS over one score-sorted segment, executed by one CTA (256 threads).
// Decision arithmetic == torchvision's nms kernels (see iou_gt in common.cuh); the greedy order and the
// strict ">" make the result identical to the bitmask + host-sweep formulation (nms_rotated_cuda.cu:21-137
// documents that classic structure) without materialising an N x N/64 mask: candidates are visited in
// chunks of 64; a chunk is tested against the list of boxes kept so far (parallel over kept x candidate),
// then resolved internally with a 64x64 bit matrix and a 64-step serial sweep by one thread.
#pragma once
#include "common.cuh"

namespace lvcb200 {

struct NmsShared {
  float cx1[64], cy1[64], cx2[64], cy2[64], carea[64];
  unsigned long long row[64];
  unsigned long long kept_mask;
  unsigned int sup[64];
  int valid[64];
};

// Box accessor: get(j, x1,y1,x2,y2) -> valid.  kept_* : storage for kept boxes (>= n entries; smem or global).
// flags[j] (global or smem, n entries) receives 1 if sorted candidate j is kept.
// Returns number kept (uniform across the CTA).  blockDim.x must be 256.
template <typename Get>
__device__ int segment_nms(NmsShared& sh, Get get, int n, float thr, float* kx1, float* ky1, float* kx2, float* ky2,
                           float* karea, unsigned char* flags) {
  const int t = threadIdx.x;
  int K = 0;
  const bool skip_zero_inter = thr >= 0.f;  // inter == 0 -> ovr is 0 or NaN -> never > thr
  for (int c0 = 0; c0 < n; c0 += 64) {
    if (t < 64) {
      int j = c0 + t;
      float x1 = 0, y1 = 0, x2 = 0, y2 = 0;
      bool v = false;
      if (j < n) v = get(j, x1, y1, x2, y2);
      sh.cx1[t] = x1; sh.cy1[t] = y1; sh.cx2[t] = x2; sh.cy2[t] = y2;
      sh.carea[t] = box_area(x1, y1, x2, y2);
      sh.valid[t] = v ? 1 : 0;
      sh.sup[t] = 0;
      sh.row[t] = 0ull;
    }
    __syncthreads();
    {  // phase A: chunk vs kept list
      const int c = t & 63;
      if (sh.valid[c]) {
        float x1 = sh.cx1[c], y1 = sh.cy1[c], x2 = sh.cx2[c], y2 = sh.cy2[c], ar = sh.carea[c];
        bool s = false;
        for (int k = t >> 6; k < K && !s; k += 4) {
          float bx1 = kx1[k], by1 = ky1[k], bx2 = kx2[k], by2 = ky2[k];
          if (skip_zero_inter && (fminf(bx2, x2) <= fmaxf(bx1, x1) || fminf(by2, y2) <= fmaxf(by1, y1))) continue;
          s = iou_gt(bx1, by1, bx2, by2, karea[k], x1, y1, x2, y2, ar, thr);
        }
        if (s) sh.sup[c] = 1;
      }
    }
    {  // phase B: 64x64 intra-chunk matrix, 16 pairs per thread
      const int i = t >> 2;
      const int j0 = (t & 3) * 16;
      if (sh.valid[i]) {
        float x1 = sh.cx1[i], y1 = sh.cy1[i], x2 = sh.cx2[i], y2 = sh.cy2[i], ar = sh.carea[i];
        unsigned long long bits = 0ull;
#pragma unroll 4
        for (int jj = 0; jj < 16; jj++) {
          int j = j0 + jj;
          if (j <= i || !sh.valid[j]) continue;
          float bx1 = sh.cx1[j], by1 = sh.cy1[j], bx2 = sh.cx2[j], by2 = sh.cy2[j];
          if (skip_zero_inter && (fminf(bx2, x2) <= fmaxf(bx1, x1) || fminf(by2, y2) <= fmaxf(by1, y1))) continue;
          if (iou_gt(x1, y1, x2, y2, ar, bx1, by1, bx2, by2, sh.carea[j], thr)) bits |= (1ull << j);
        }
        if (bits) atomicOr(&sh.row[i], bits);
      }
    }
    __syncthreads();
    if (t == 0) {  // phase C: serial sweep over the chunk
      unsigned long long alive = 0ull;
      for (int i = 0; i < 64; i++)
        if (sh.valid[i] && !sh.sup[i]) alive |= (1ull << i);
      unsigned long long kept = 0ull;
      for (int i = 0; i < 64; i++) {
        if ((alive >> i) & 1ull) { kept |= (1ull << i); alive &= ~sh.row[i]; }
      }
      sh.kept_mask = kept;
    }
    __syncthreads();
    const unsigned long long kept = sh.kept_mask;
    if (t < 64) {
      bool k = (kept >> t) & 1ull;
      if (c0 + t < n) flags[c0 + t] = k ? 1 : 0;
      if (k) {
        int pos = K + __popcll(kept & ((1ull << t) - 1ull));
        kx1[pos] = sh.cx1[t]; ky1[pos] = sh.cy1[t]; kx2[pos] = sh.cx2[t]; ky2[pos] = sh.cy2[t];
        karea[pos] = sh.carea[t];
      }
    }
    K += __popcll(kept);
    __syncthreads();
  }
  return K;
}

// which branch would the reference take on a CUDA device (detectron2/layers/nms.py:19-29 over
// torchvision 0.26 batched_nms: trick iff numel <= 100000)
__host__ __device__ inline int reference_cuda_nms_mode(long long n) {
  if (n >= 40000) return 1;
  return (n * 4 > 100000) ? 1 : 0;
}

}  // namespace lvcb200
