// Device-side candidate filter between detection ("Label") and verification ("Verify"): get_ret_anns in score mode
// (tools/create_coco_dataset_from_dets_all.py:129-193) applied to the detector's per-image output block right behind the
// NMS, before anything leaves the GPU.  Per detection d of image i with class c (contiguous id):
//   valid = novel[c] && !excluded[i][c] && 0 < area < 1e10 && ar < area / image_area < 1        (:39-43, :133-137)
//           area = w * h of the XYWH box in double (pycocotools loadRes), w = fp32(x2 - x1), h = fp32(y2 - y1)
//   keep  = valid && k_min < score <= k_max            (left searchsorted on -scores, :169-174)      -> flags 1 (ignore_qe = 0)
//   --full: valid && !keep && the image holds a kept detection of class c                          -> flags 2 (ignore_qe = 1)
// One CTA per image (<= a few hundred detections): flags first, then the same-image / same-class scan out of shared memory.
#include "common.cuh"

namespace lvcb200 {

constexpr int kCandMax = 1024;   // detections per image held in shared memory

__global__ void __launch_bounds__(128)
candidate_filter_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ classes,
                        const int32_t* __restrict__ counts, const double* __restrict__ image_area, int topk, int num_classes,
                        const uint8_t* __restrict__ novel, const uint8_t* __restrict__ excluded, double k_min, double k_max, double ar,
                        int full, int8_t* __restrict__ flags, int32_t* __restrict__ n_keep) {
  __shared__ int s_cls[kCandMax];
  __shared__ int8_t s_flag[kCandMax];     // 0 dropped, 1 keep, 3 valid but not kept
  __shared__ int s_keep;
  const int img = blockIdx.x;
  const int cnt = min(counts[img], topk);
  const double ia = image_area[img];
  if (threadIdx.x == 0) s_keep = 0;
  __syncthreads();
  for (int d = threadIdx.x; d < topk; d += blockDim.x) {
    int8_t f = 0;
    int c = -1;
    if (d < cnt) {
      const float4 b = boxes[(size_t)img * topk + d];
      c = (int)classes[(size_t)img * topk + d];
      const double area = (double)__fsub_rn(b.z, b.x) * (double)__fsub_rn(b.w, b.y);
      const double ratio = area / ia;
      bool valid = c >= 0 && c < num_classes && novel[c] != 0 && (excluded == nullptr || excluded[(size_t)img * num_classes + c] == 0);
      valid = valid && area > 0.0 && area < 1e10 && ratio > ar && ratio < 1.0;
      const double s = (double)scores[(size_t)img * topk + d];
      if (valid) f = (s > k_min && s <= k_max) ? 1 : 3;
      if (f == 1) atomicAdd(&s_keep, 1);
    }
    s_cls[d] = c;
    s_flag[d] = f;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < topk; d += blockDim.x) {
    int8_t f = s_flag[d];
    if (f == 3) {
      f = 0;
      if (full) {
        const int c = s_cls[d];
        for (int e = 0; e < cnt; e++)
          if (s_flag[e] == 1 && s_cls[e] == c) { f = 2; break; }
      }
    }
    flags[(size_t)img * topk + d] = f;
  }
  if (n_keep != nullptr && threadIdx.x == 0) n_keep[img] = s_keep;
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_candidate_filter(const float* det_boxes, const float* det_scores, const int64_t* det_classes,
                                        const int32_t* det_counts, const double* image_area, int n_images, int topk, int num_classes,
                                        const uint8_t* novel, const uint8_t* excluded, double k_min, double k_max, double ar, int full,
                                        int8_t* flags, int32_t* n_keep, void* stream) {
  if (n_images == 0 || topk == 0) return 0;
  LVC_REQUIRE(det_boxes && det_scores && det_classes && det_counts && image_area && novel && flags, "candidate_filter: NULL pointer");
  LVC_REQUIRE(topk <= kCandMax && num_classes >= 1, "candidate_filter: at most 1024 detections per image");
  LVC_REQUIRE(((uintptr_t)det_boxes % 16) == 0, "candidate_filter: boxes must be 16-byte aligned");
  candidate_filter_kernel<<<n_images, 128, 0, (cudaStream_t)stream>>>((const float4*)det_boxes, det_scores, det_classes, det_counts, image_area,
                                                                      topk, num_classes, novel, excluded, k_min, k_max, ar, full, flags, n_keep);
  return check_launch("candidate_filter_kernel");
}
