// "Next" row f4 of the scope table: the training-side users of the path's operators.
//   roi_align_backward : autograd backward of detectron2.layers.roi_align (ROIAlign_cuda.cu:141-306), NCHW fp32, scatter by RED.ADD
//   pairwise_iou       : detectron2/structures/boxes.py:315-347, separately rounded fp32 ops (bit-exact with the library's)
//   match_boxes        : Matcher.__call__ + set_low_quality_matches_ (detectron2/modeling/matcher.py:61-126) on a given quality matrix
//                        or fused with the IoU computation (the [G, P] matrix of label_anchors / label_and_sample_proposals --
//                        rpn.py:277-326, lvc/modeling/roi_heads/roi_heads.py:173-278 -- is never materialised)
#include "common.cuh"

namespace lvcb200 {

__global__ void __launch_bounds__(256)
roi_align_backward_nchw_kernel(const float* __restrict__ gout, int C, int H, int W, const float* __restrict__ rois, int64_t total,
                               int ph_n, int pw_n, float scale, int sampling_ratio, bool aligned, float* __restrict__ gin) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % pw_n);
    const int ph = (int)((idx / pw_n) % ph_n);
    const int c = (int)((idx / ((int64_t)pw_n * ph_n)) % C);
    const int64_t n = idx / ((int64_t)pw_n * ph_n * C);
    const float* roi = rois + n * 5;
    const int b = (int)roi[0];
    const float off = aligned ? 0.5f : 0.0f;
    const float start_w = roi[1] * scale - off, start_h = roi[2] * scale - off;
    float rw = (roi[3] * scale - off) - start_w, rh = (roi[4] * scale - off) - start_h;
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bin_h = rh / (float)ph_n, bin_w = rw / (float)pw_n;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph_n);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw_n);
    const float count = (float)(gh * gw);
    const float g = gout[idx];
    float* p = gin + ((int64_t)b * C + c) * H * W;
    for (int iy = 0; iy < gh; iy++) {
      float yy = start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
      for (int ix = 0; ix < gw; ix++) {
        float x = start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw, y = yy;
        if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
        if (y <= 0) y = 0;
        if (x <= 0) x = 0;
        int yl = (int)y, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
        atomicAdd(p + yl * W + xl, g * (hy * hx) / count);
        atomicAdd(p + yl * W + xh, g * (hy * lx) / count);
        atomicAdd(p + yh * W + xl, g * (ly * hx) / count);
        atomicAdd(p + yh * W + xh, g * (ly * lx) / count);
      }
    }
  }
}

// boxes.py:328-346 with the library's separately rounded operations
__device__ __forceinline__ float iou_pair(const float4 a, const float4 b) {
  const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)), area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  w = w < 0.f ? 0.f : w;
  h = h < 0.f ? 0.f : h;
  const float inter = __fmul_rn(w, h);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area1, area2), inter)) : 0.f;
}

__global__ void __launch_bounds__(256)
pairwise_iou_kernel(const float4* __restrict__ b1, int64_t G, const float4* __restrict__ b2, int64_t P, float* __restrict__ iou) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 bp = b2[p];
  for (int64_t g = blockIdx.y; g < G; g += gridDim.y) iou[g * P + p] = iou_pair(__ldg(b1 + g), bp);
}

constexpr int kMatchMaxThr = 4;
struct MatchParams {
  float thr[kMatchMaxThr]; int8_t lab[kMatchMaxThr + 1]; int n_thr; int allow_low;
};

// thread per prediction: column max / first argmax over the G ground-truth rows, threshold labels; per-gt row maxima reduced over the
// warp and merged with one atomicMax per (warp, gt) on the float bits (IoU >= 0: the unsigned order is the float order)
__global__ void __launch_bounds__(256)
match_cols_kernel(const float4* __restrict__ gt, int64_t G, const float4* __restrict__ boxes, const float* __restrict__ quality, int64_t P,
                  MatchParams mp, int64_t* __restrict__ matches, int8_t* __restrict__ labels, float* __restrict__ vals,
                  unsigned int* __restrict__ row_max) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = p < P;
  float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in && boxes) bp = boxes[p];
  float best = -1.f;
  int64_t arg = 0;
  for (int64_t g = 0; g < G; g++) {
    float v = 0.f;
    if (in) v = quality ? quality[g * P + p] : iou_pair(__ldg(gt + g), bp);
    if (in && v > best) { best = v; arg = g; }
    if (mp.allow_low) {
      float m = in ? v : 0.f;
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0) atomicMax(row_max + g, __float_as_uint(m));
    }
  }
  if (!in) return;
  matches[p] = arg;
  if (vals) vals[p] = G > 0 ? best : 0.f;
  int8_t l = 1;
  for (int i = 0; i <= mp.n_thr; i++) {
    const float low = i == 0 ? -INFINITY : mp.thr[i - 1], high = i == mp.n_thr ? INFINITY : mp.thr[i];
    if (best >= low && best < high) l = mp.lab[i];
  }
  labels[p] = l;
}

// set_low_quality_matches_: every prediction attaining a gt's maximum quality (ties included) is labelled 1
__global__ void __launch_bounds__(256)
match_low_quality_kernel(const float4* __restrict__ gt, int64_t G, const float4* __restrict__ boxes, const float* __restrict__ quality,
                         int64_t P, const unsigned int* __restrict__ row_max, int8_t* __restrict__ labels) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (boxes) bp = boxes[p];
  bool hit = false;
  for (int64_t g = 0; g < G; g++) {
    const float v = quality ? quality[g * P + p] : iou_pair(__ldg(gt + g), bp);
    hit |= (__float_as_uint(v) == row_max[g]);
  }
  if (hit) labels[p] = 1;
}

// RPN.losses (rpn.py:328-400) before normalisation: objectness BCE-with-logits over anchors with label >= 0, smooth-L1 (beta < 1e-5: L1)
// between the predicted deltas and Box2BoxTransform.get_deltas(anchor, matched gt) (box_regression.py:38-71) over label == 1.
// Thread per (image, anchor), coalesced 16-byte loads of deltas / gt boxes; fp64 block reduction, one atomicAdd(double) pair per CTA.
__global__ void __launch_bounds__(256)
rpn_losses_kernel(const float4* __restrict__ anchors, const float* __restrict__ logits, const float4* __restrict__ deltas,
                  const int8_t* __restrict__ labels, const float4* __restrict__ gt, int64_t N, int64_t A, float4 w, float beta,
                  double* __restrict__ out2) {
  double cls = 0.0, loc = 0.0;
  const int64_t total = N * A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int8_t l = labels[i];
    if (l >= 0) {
      const float x = logits[i], y = (float)l;
      cls += (double)(fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))));
    }
    if (l == 1) {
      const float4 s = __ldg(anchors + (i % A)), t = gt[i], d = deltas[i];
      const float sw = s.z - s.x, sh = s.w - s.y, sx = s.x + 0.5f * sw, sy = s.y + 0.5f * sh;
      const float tw = t.z - t.x, th = t.w - t.y, tx = t.x + 0.5f * tw, ty = t.y + 0.5f * th;
      const float e[4] = {fabsf(d.x - w.x * (tx - sx) / sw), fabsf(d.y - w.y * (ty - sy) / sh), fabsf(d.z - w.z * logf(tw / sw)),
                          fabsf(d.w - w.w * logf(th / sh))};
#pragma unroll
      for (int k = 0; k < 4; k++) loc += (double)(beta < 1e-5f ? e[k] : (e[k] < beta ? 0.5f * e[k] * e[k] / beta : e[k] - 0.5f * beta));
    }
  }
  __shared__ double red[2][8];
  for (int o = 16; o; o >>= 1) { cls += __shfl_xor_sync(0xffffffffu, cls, o); loc += __shfl_xor_sync(0xffffffffu, loc, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = cls; red[1][threadIdx.x >> 5] = loc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; i++) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(out2, a);
    atomicAdd(out2 + 1, b);
  }
}

// FastRCNNOutputs.losses (lvc/modeling/roi_heads/fast_rcnn.py:267-358, 424-438) before the division by R: warp per RoI row --
// log-sum-exp cross entropy over the K + 1 logits (lanes stride the columns), and for foreground rows the smooth-L1 / L1 distance
// between the gt class's four predicted deltas and get_deltas(proposal, gt box).  fp64 reduction, one atomicAdd pair per CTA.
__global__ void __launch_bounds__(256)
fast_rcnn_losses_kernel(const float* __restrict__ logits, const float* __restrict__ deltas, int n_delta_cols, const int64_t* __restrict__ gt_classes,
                        const float4* __restrict__ proposals, const float4* __restrict__ gt_boxes, int64_t R, int K, float4 w, float beta,
                        double* __restrict__ out2) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double cls = 0.0, box = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * 8 + wib; r < R; r += (int64_t)gridDim.x * 8) {
    const float* x = logits + r * (K + 1);
    float mx = -INFINITY;
    for (int k = lane; k <= K; k += 32) mx = fmaxf(mx, x[k]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int k = lane; k <= K; k += 32) se += expf(x[k] - mx);
    for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) {
      const int64_t g = gt_classes[r];
      cls += (double)(logf(se) + mx - x[g]);
      if (g >= 0 && g < K) {
        const float4 s = proposals[r], t = gt_boxes[r];
        const float sw = s.z - s.x, sh = s.w - s.y, sx = s.x + 0.5f * sw, sy = s.y + 0.5f * sh;
        const float tw = t.z - t.x, th = t.w - t.y, tx = t.x + 0.5f * tw, ty = t.y + 0.5f * th;
        const float* pd = deltas + r * n_delta_cols + (n_delta_cols == 4 ? 0 : 4 * g);
        const float e[4] = {fabsf(pd[0] - w.x * (tx - sx) / sw), fabsf(pd[1] - w.y * (ty - sy) / sh), fabsf(pd[2] - w.z * logf(tw / sw)),
                            fabsf(pd[3] - w.w * logf(th / sh))};
#pragma unroll
        for (int k = 0; k < 4; k++) box += (double)(beta < 1e-5f ? e[k] : (e[k] < beta ? 0.5f * e[k] * e[k] / beta : e[k] - 0.5f * beta));
      }
    }
  }
  __shared__ double red[2][8];
  if (lane == 0) { red[0][wib] = cls; red[1][wib] = box; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; i++) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(out2, a);
    atomicAdd(out2 + 1, b);
  }
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_roi_align_backward_nchw_f32(const float* grad_output, const float* rois, int R, int N, int C, int H, int W,
                                                   int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                                                   float* grad_input, void* stream) {
  LVC_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, "roi_align_backward: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if ((int64_t)N * C * H * W == 0) return 0;
  LVC_REQUIRE(grad_input, "roi_align_backward: NULL grad_input");
  LVC_CUDA(cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)N * C * H * W, s));
  const int64_t total = (int64_t)R * C * pooled_h * pooled_w;
  if (total == 0) return 0;
  LVC_REQUIRE(grad_output && rois, "roi_align_backward: NULL pointer");
  // (a CTA-per-(RoI, 8 channels) variant that accumulates the footprint in shared memory first -- 5x fewer global atomics -- was
  //  measured slower, 2.0 vs 1.5 ms on 4096 RoIs x 256 channels: shared fp32 atomics and 131k tiny CTAs cost more than the RED.ADDs)
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 64) blocks = kNumSMs * 64;
  roi_align_backward_nchw_kernel<<<(unsigned)blocks, 256, 0, s>>>(grad_output, C, H, W, rois, total, pooled_h, pooled_w, spatial_scale,
                                                                   sampling_ratio, aligned != 0, grad_input);
  return check_launch("roi_align_backward_nchw_kernel");
}

extern "C" int lvcb200_pairwise_iou(const float* boxes1, int64_t G, const float* boxes2, int64_t P, float* iou, void* stream) {
  LVC_REQUIRE(G >= 0 && P >= 0, "pairwise_iou: bad sizes");
  if (G == 0 || P == 0) return 0;
  LVC_REQUIRE(boxes1 && boxes2 && iou, "pairwise_iou: NULL pointer");
  LVC_REQUIRE(((uintptr_t)boxes1 % 16) == 0 && ((uintptr_t)boxes2 % 16) == 0, "pairwise_iou: boxes must be 16-byte aligned");
  dim3 grid((unsigned)ceil_div64(P, 256), (unsigned)(G < 64 ? G : 64));
  pairwise_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes1, G, (const float4*)boxes2, P, iou);
  return check_launch("pairwise_iou_kernel");
}

extern "C" size_t lvcb200_match_boxes_workspace(int64_t G) { return align_up((size_t)(G > 0 ? G : 1) * 4, 256); }

extern "C" int lvcb200_match_boxes(const float* gt_boxes, int64_t G, const float* boxes, const float* quality, int64_t P,
                                   const float* thresholds, int n_thresholds, const int8_t* labels, int allow_low_quality_matches,
                                   int64_t* matches, int8_t* match_labels, float* matched_vals, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  LVC_REQUIRE(G >= 0 && P >= 0 && n_thresholds >= 1 && n_thresholds <= kMatchMaxThr && thresholds && labels, "match_boxes: bad arguments");
  for (int i = 0; i + 1 < n_thresholds; i++) LVC_REQUIRE(thresholds[i] <= thresholds[i + 1], "match_boxes: thresholds must be sorted");
  for (int i = 0; i <= n_thresholds; i++) LVC_REQUIRE(labels[i] >= -1 && labels[i] <= 1, "match_boxes: labels must be in {-1, 0, 1}");
  if (P == 0) return 0;
  LVC_REQUIRE(matches && match_labels, "match_boxes: NULL output");
  LVC_REQUIRE(quality || (boxes && (G == 0 || gt_boxes)), "match_boxes: need a quality matrix or both box lists");
  LVC_REQUIRE(((uintptr_t)gt_boxes % 16) == 0 && ((uintptr_t)boxes % 16) == 0, "match_boxes: boxes must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  MatchParams mp;
  for (int i = 0; i < kMatchMaxThr; i++) mp.thr[i] = i < n_thresholds ? thresholds[i] : 0.f;
  for (int i = 0; i <= kMatchMaxThr; i++) mp.lab[i] = i <= n_thresholds ? labels[i] : 0;
  mp.n_thr = n_thresholds;
  mp.allow_low = (allow_low_quality_matches && G > 0) ? 1 : 0;
  if (mp.allow_low) {
    if (!workspace || workspace_bytes < lvcb200_match_boxes_workspace(G)) return set_error(LVCB200_EWORKSPACE, "match_boxes: workspace too small");
    LVC_CUDA(cudaMemsetAsync(workspace, 0, (size_t)G * 4, s));
  }
  const unsigned blocks = (unsigned)ceil_div64(P, 256);
  // G == 0: the loop is empty, best stays -1 -> matches 0 and labels[0] (matcher.py:76-87: -1 lies in the first band [-inf, thr0))
  match_cols_kernel<<<blocks, 256, 0, s>>>((const float4*)gt_boxes, G, (const float4*)boxes, quality, P, mp, matches, match_labels,
                                           matched_vals, (unsigned int*)workspace);
  int rc = check_launch("match_cols_kernel");
  if (rc || !mp.allow_low) return rc;
  match_low_quality_kernel<<<blocks, 256, 0, s>>>((const float4*)gt_boxes, G, (const float4*)boxes, quality, P,
                                                  (const unsigned int*)workspace, match_labels);
  return check_launch("match_low_quality_kernel");
}

extern "C" int lvcb200_rpn_losses(const float* anchors, const float* logits, const float* deltas, const int8_t* labels,
                                  const float* gt_boxes, int64_t N, int64_t A, const float* weights, float smooth_l1_beta,
                                  double* out2, void* stream) {
  LVC_REQUIRE(N >= 0 && A >= 0 && weights && out2, "rpn_losses: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  LVC_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(double), s));
  if (N * A == 0) return 0;
  LVC_REQUIRE(anchors && logits && deltas && labels && gt_boxes, "rpn_losses: NULL pointer");
  LVC_REQUIRE(((uintptr_t)anchors % 16) == 0 && ((uintptr_t)deltas % 16) == 0 && ((uintptr_t)gt_boxes % 16) == 0, "rpn_losses: boxes / deltas must be 16-byte aligned");
  int64_t blocks = ceil_div64(N * A, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  rpn_losses_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float4*)anchors, logits, (const float4*)deltas, labels, (const float4*)gt_boxes, N, A,
                                                     make_float4(weights[0], weights[1], weights[2], weights[3]), smooth_l1_beta, out2);
  return check_launch("rpn_losses_kernel");
}

extern "C" int lvcb200_fast_rcnn_losses(const float* cls_logits, const float* box_deltas, int n_delta_cols, const int64_t* gt_classes,
                                        const float* proposals, const float* gt_boxes, int64_t R, int num_classes, const float* weights,
                                        float smooth_l1_beta, double* out2, void* stream) {
  LVC_REQUIRE(R >= 0 && num_classes >= 1 && weights && out2 && (n_delta_cols == 4 || n_delta_cols == 4 * num_classes), "fast_rcnn_losses: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  LVC_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(double), s));
  if (R == 0) return 0;
  LVC_REQUIRE(cls_logits && box_deltas && gt_classes && proposals && gt_boxes, "fast_rcnn_losses: NULL pointer");
  LVC_REQUIRE(((uintptr_t)proposals % 16) == 0 && ((uintptr_t)gt_boxes % 16) == 0, "fast_rcnn_losses: boxes must be 16-byte aligned");
  int64_t blocks = ceil_div64(R, 8);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  fast_rcnn_losses_kernel<<<(unsigned)blocks, 256, 0, s>>>(cls_logits, box_deltas, n_delta_cols, gt_classes, (const float4*)proposals,
                                                           (const float4*)gt_boxes, R, num_classes,
                                                           make_float4(weights[0], weights[1], weights[2], weights[3]), smooth_l1_beta, out2);
  return check_launch("fast_rcnn_losses_kernel");
}
