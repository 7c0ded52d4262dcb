/*
 * lvc_oracle.c -- CPU restatement (plain C, scalar, single thread) of the arithmetic on the LVC
 * pseudo-label mining hot path.  TEST INFRASTRUCTURE ONLY: this file is the *checker* for the CUDA
 * path in lvc_b200/csrc; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never routes through it.
 *
 * Every function cites the reference file:line (relative to /root/reference, prannaykaul/lvc @ 3b5e5fa)
 * it follows.  Where the reference delegates to a third-party library that is not vendored in the
 * reference tree (torchvision.ops.nms / batched_nms / roi_align -- unpinned in the reference's
 * requirements.txt; the version installed in this image and used to pin this oracle is
 * torchvision 0.26.0+cu128, torch 2.11.0+cu128), the published algorithm of that library is
 * restated and the call site in the reference is cited.
 *
 * Pinning: the reference ships no tests / golden vectors.  This oracle is pinned
 * against outputs of the reference's own Python executed in the build container through
 * oracle/ref_shim.py (fixtures in tests/golden/, generator oracle/make_golden.py) and against the
 * known-answer cases of SURVEY.md Appendix A (tests/test_oracle_*.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: NMS/IoU decisions must
 * reproduce the library's separate multiply / add roundings).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * RoIAlign forward, NCHW fp32.
 * Follows detectron2/layers/csrc/ROIAlign/ROIAlign_cpu.cpp:21-114 (pre_calc_for_bilinear_interpolate)
 * and :116-218 (ROIAlignForward), which is the in-tree spec of torchvision.ops.roi_align reached from
 * detectron2/layers/roi_align.py:15,106-108.  One deliberate deviation, following torchvision (the
 * code that actually runs): zero-size / inverted RoIs with aligned=True do not assert
 * (ROIAlign_cpu.cpp:149-152 would) but produce zeros (grid = ceil(<=0) = 0, count = max(0,1)).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_roi_align_forward(const float* input, int N, int C, int H, int W,
                                   const float* rois, int R, int pooled_h, int pooled_w,
                                   float spatial_scale, int sampling_ratio, int aligned,
                                   float* output) {
  (void)N;
  for (int n = 0; n < R; n++) {
    const float* roi = rois + 5 * n;
    int batch = (int)roi[0];
    float offset = aligned ? 0.5f : 0.0f;
    float roi_start_w = roi[1] * spatial_scale - offset;
    float roi_start_h = roi[2] * spatial_scale - offset;
    float roi_end_w = roi[3] * spatial_scale - offset;
    float roi_end_h = roi[4] * spatial_scale - offset;
    float roi_width = roi_end_w - roi_start_w;
    float roi_height = roi_end_h - roi_start_h;
    if (!aligned) {
      roi_width = roi_width > 1.f ? roi_width : 1.f;
      roi_height = roi_height > 1.f ? roi_height : 1.f;
    }
    float bin_size_h = roi_height / (float)pooled_h;
    float bin_size_w = roi_width / (float)pooled_w;
    int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)pooled_h);
    int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)pooled_w);
    int cnt = grid_h * grid_w;
    float count = (float)(cnt > 1 ? cnt : 1);
    for (int c = 0; c < C; c++) {
      const float* in = input + ((size_t)batch * C + c) * H * W;
      for (int ph = 0; ph < pooled_h; ph++) {
        for (int pw = 0; pw < pooled_w; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < grid_h; iy++) {
            float yy = roi_start_h + ph * bin_size_h + ((float)iy + .5f) * bin_size_h / (float)grid_h;
            for (int ix = 0; ix < grid_w; ix++) {
              float xx = roi_start_w + pw * bin_size_w + ((float)ix + .5f) * bin_size_w / (float)grid_w;
              float x = xx, y = yy;
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue; /* weights 0 */
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int y_low = (int)y, x_low = (int)x, y_high, x_high;
              if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
              if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
              float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * in[y_low * W + x_low] + w2 * in[y_low * W + x_high] +
                     w3 * in[y_high * W + x_low] + w4 * in[y_high * W + x_high];
            }
          }
          output[(((size_t)n * C + c) * pooled_h + ph) * pooled_w + pw] = acc / count;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * FPN level assignment.  Follows detectron2/modeling/poolers.py:51-59 (assign_boxes_to_levels) with
 * Boxes.area() detectron2/structures/boxes.py:172-181.  Returns level - min_level as int64.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_assign_boxes_to_levels(const float* boxes, int R, int min_level, int max_level,
                                        int canonical_box_size, int canonical_level, int64_t* out) {
  for (int i = 0; i < R; i++) {
    const float* b = boxes + 4 * i;
    float area = (b[2] - b[0]) * (b[3] - b[1]);
    float size = sqrtf(area);
    float lvl = floorf((float)canonical_level + log2f(size / (float)canonical_box_size + 1e-8f));
    if (lvl < (float)min_level) lvl = (float)min_level; /* NaN compares false: stays NaN like torch.clamp */
    if (lvl > (float)max_level) lvl = (float)max_level;
    out[i] = (int64_t)lvl - min_level;
  }
}

/* ------------------------------------------------------------------------------------------
 * Box2BoxTransform.apply_deltas.  Follows detectron2/modeling/box_regression.py:73-110.
 * deltas: [R, K*4], boxes: [R,4] -> out [R, K*4].  scale_clamp = log(1000/16) (box_regression.py:14).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_apply_deltas(const float* deltas, const float* boxes, int R, int K, float wx, float wy,
                              float ww, float wh, float scale_clamp, float* out) {
  for (int i = 0; i < R; i++) {
    const float* b = boxes + 4 * i;
    float widths = b[2] - b[0], heights = b[3] - b[1];
    float ctr_x = b[0] + 0.5f * widths, ctr_y = b[1] + 0.5f * heights;
    for (int k = 0; k < K; k++) {
      const float* d = deltas + ((size_t)i * K + k) * 4;
      float dx = d[0] / wx, dy = d[1] / wy, dw = d[2] / ww, dh = d[3] / wh;
      if (dw > scale_clamp) dw = scale_clamp;
      if (dh > scale_clamp) dh = scale_clamp;
      float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
      float pw = expf(dw) * widths, ph = expf(dh) * heights;
      float* o = out + ((size_t)i * K + k) * 4;
      o[0] = pcx - 0.5f * pw; o[1] = pcy - 0.5f * ph; o[2] = pcx + 0.5f * pw; o[3] = pcy + 0.5f * ph;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Greedy NMS.  torchvision.ops.nms CPU algorithm (reached from detectron2/layers/nms.py:7,25):
 * stable descending sort by score; box j suppressed by a kept higher-scoring box i iff
 * inter / (area_i + area_j - inter) > thr (strict), inter = max(0,xx2-xx1) * max(0,yy2-yy1).
 * Returns number kept; keep[] = original indices in descending-score order.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float s; int64_t i; } orc_si;
static int orc_cmp_desc(const void* a, const void* b) {
  const orc_si* x = (const orc_si*)a; const orc_si* y = (const orc_si*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i); /* stable: ties keep the lower index first */
}

ORC_API int64_t orc_nms(const float* boxes, const float* scores, int64_t n, float thr, int64_t* keep) {
  if (n == 0) return 0;
  orc_si* order = (orc_si*)malloc(sizeof(orc_si) * n);
  float* areas = (float*)malloc(sizeof(float) * n);
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  for (int64_t i = 0; i < n; i++) {
    order[i].s = scores[i]; order[i].i = i;
    areas[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  qsort(order, n, sizeof(orc_si), orc_cmp_desc);
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; _i++) {
    int64_t i = order[_i].i;
    if (sup[i]) continue;
    keep[nk++] = i;
    float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; _j++) {
      int64_t j = order[_j].i;
      if (sup[j]) continue;
      float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
      float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
      float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
      float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
      float w = xx2 - xx1; if (!(w > 0.f)) w = 0.f;
      float h = yy2 - yy1; if (!(h > 0.f)) h = 0.f;
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      if (ovr > thr) sup[j] = 1;
    }
  }
  free(order); free(areas); free(sup);
  return nk;
}

/* ------------------------------------------------------------------------------------------
 * batched_nms.  Follows detectron2/layers/nms.py:10-29 and torchvision.ops.boxes.batched_nms:
 *   mode 0 "coordinate trick": boxes + idx * (max(boxes) + 1) in fp32, then one nms over everything
 *   mode 1 "vanilla" (also the >= 40000-box loop of nms.py:22-29): nms per class, union, then
 *           order kept indices by descending score.
 * ------------------------------------------------------------------------------------------ */
ORC_API int64_t orc_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n,
                                float thr, int mode, int64_t* keep) {
  if (n == 0) return 0;
  if (mode == 0) {
    float mx = boxes[0];
    for (int64_t i = 1; i < 4 * n; i++) if (boxes[i] > mx) mx = boxes[i];
    float mul = mx + 1.0f;
    float* b2 = (float*)malloc(sizeof(float) * 4 * n);
    for (int64_t i = 0; i < n; i++) {
      float off = (float)idxs[i] * mul;
      for (int k = 0; k < 4; k++) b2[4 * i + k] = boxes[4 * i + k] + off;
    }
    int64_t nk = orc_nms(b2, scores, n, thr, keep);
    free(b2);
    return nk;
  }
  uint8_t* kept = (uint8_t*)calloc(n, 1);
  uint8_t* done = (uint8_t*)calloc(n, 1);
  float* cb = (float*)malloc(sizeof(float) * 4 * n);
  float* cs = (float*)malloc(sizeof(float) * n);
  int64_t* ci = (int64_t*)malloc(sizeof(int64_t) * n);
  int64_t* ck = (int64_t*)malloc(sizeof(int64_t) * n);
  for (int64_t s = 0; s < n; s++) {
    if (done[s]) continue;
    int64_t cls = idxs[s], m = 0;
    for (int64_t j = s; j < n; j++)
      if (idxs[j] == cls) { done[j] = 1; memcpy(cb + 4 * m, boxes + 4 * j, 16); cs[m] = scores[j]; ci[m] = j; m++; }
    int64_t nk = orc_nms(cb, cs, m, thr, ck);
    for (int64_t k = 0; k < nk; k++) kept[ci[ck[k]]] = 1;
  }
  int64_t nk = 0;
  orc_si* order = (orc_si*)malloc(sizeof(orc_si) * n);
  for (int64_t i = 0; i < n; i++) if (kept[i]) { order[nk].s = scores[i]; order[nk].i = i; nk++; }
  qsort(order, nk, sizeof(orc_si), orc_cmp_desc);
  for (int64_t i = 0; i < nk; i++) keep[i] = order[i].i;
  free(order); free(kept); free(done); free(cb); free(cs); free(ci); free(ck);
  return nk;
}

/* ------------------------------------------------------------------------------------------
 * Anchor grid for one level.  Follows detectron2/modeling/anchor_generator.py:37-49
 * (_create_grid_offsets), :157-171 (_grid_anchors), :173-208 (generate_cell_anchors); offset 0.0.
 * anchors: [H*W*A, 4] in (y, x, a) order.  cell anchors computed in double then cast, like
 * torch.tensor(python floats) does.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_cell_anchors(const double* sizes, int ns, const double* ratios, int nr, float* out) {
  int a = 0;
  for (int s = 0; s < ns; s++) {
    double area = sizes[s] * sizes[s];
    for (int r = 0; r < nr; r++, a++) {
      double w = sqrt(area / ratios[r]);
      double h = ratios[r] * w;
      out[4 * a] = (float)(-w / 2.0); out[4 * a + 1] = (float)(-h / 2.0);
      out[4 * a + 2] = (float)(w / 2.0); out[4 * a + 3] = (float)(h / 2.0);
    }
  }
}

ORC_API void orc_grid_anchors(const float* cell, int A, int H, int W, int stride, float* out) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int a = 0; a < A; a++) {
        float sx = (float)(x * stride), sy = (float)(y * stride);
        float* o = out + (((size_t)y * W + x) * A + a) * 4;
        o[0] = sx + cell[4 * a]; o[1] = sy + cell[4 * a + 1];
        o[2] = sx + cell[4 * a + 2]; o[3] = sy + cell[4 * a + 3];
      }
}

/* ------------------------------------------------------------------------------------------
 * Boxes.clip (detectron2/structures/boxes.py:183-196) and nonempty (:198-212), in place.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_clip_boxes(float* boxes, int64_t n, int h, int w) {
  for (int64_t i = 0; i < n; i++) {
    float* b = boxes + 4 * i;
    b[0] = fminf(fmaxf(b[0], 0.f), (float)w); b[1] = fminf(fmaxf(b[1], 0.f), (float)h);
    b[2] = fminf(fmaxf(b[2], 0.f), (float)w); b[3] = fminf(fmaxf(b[3], 0.f), (float)h);
  }
}

/* ------------------------------------------------------------------------------------------
 * Centred cosine kNN + class-mode vote.
 * Follows tools/run_nearest_neighbours.py:142-162 (run_nearest_neighbours: crop_mean = bank mean,
 * F.cosine_similarity(bank - mean, q - mean), topk(10)) and :214-227 (get_nn_class_confirmatory:
 * torch.mode(votes[:, :k]) == detector class -> keep).  F.cosine_similarity clamps each norm at
 * eps = 1e-8.  torch.mode returns the smallest most-frequent value.
 * topk ties: lower bank index first (the reference's order on ties is implementation-defined).
 * Outputs: top_idx [Q,topk] int64, top_sim [Q,topk] fp32, votes [Q,topk] int64, nn_class [Q] int64,
 * keep [Q] uint8.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_knn_verify(const float* bank, const int64_t* bank_cls, int S, int D, const float* q,
                            const int64_t* q_cls, int64_t Q, int topk, int knn, int64_t* top_idx,
                            float* top_sim, int64_t* votes, int64_t* nn_class, uint8_t* keep) {
  float* mean = (float*)calloc(D, sizeof(float));
  float* bc = (float*)malloc(sizeof(float) * (size_t)S * D);
  float* bn = (float*)malloc(sizeof(float) * S);
  float* qc = (float*)malloc(sizeof(float) * D);
  float* sim = (float*)malloc(sizeof(float) * S);
  /* mean over dim 0 (double accumulation: torch's CPU sum is pairwise/vectorised, fp32 result
   * correctly rounded to ~1 ulp; double accumulate + round reproduces that to 1 ulp) */
  for (int d = 0; d < D; d++) {
    double a = 0; for (int s = 0; s < S; s++) a += bank[(size_t)s * D + d];
    mean[d] = (float)(a / S);
  }
  for (int s = 0; s < S; s++) {
    double nn = 0;
    for (int d = 0; d < D; d++) { float v = bank[(size_t)s * D + d] - mean[d]; bc[(size_t)s * D + d] = v; nn += (double)v * v; }
    float nrm = (float)sqrt(nn); bn[s] = nrm > 1e-8f ? nrm : 1e-8f;
  }
  for (int64_t i = 0; i < Q; i++) {
    double nn = 0;
    for (int d = 0; d < D; d++) { float v = q[(size_t)i * D + d] - mean[d]; qc[d] = v; nn += (double)v * v; }
    float qn = (float)sqrt(nn); if (!(qn > 1e-8f)) qn = 1e-8f;
    for (int s = 0; s < S; s++) {
      double dot = 0; const float* b = bc + (size_t)s * D;
      for (int d = 0; d < D; d++) dot += (double)qc[d] * b[d];
      sim[s] = (float)(dot / ((double)qn * bn[s]));
    }
    /* top-k by repeated selection (S is small) */
    for (int k = 0; k < topk; k++) {
      int best = -1;
      for (int s = 0; s < S; s++) {
        int taken = 0;
        for (int p = 0; p < k; p++) if (top_idx[i * topk + p] == s) { taken = 1; break; }
        if (taken) continue;
        if (best < 0 || sim[s] > sim[best]) best = s;
      }
      top_idx[i * topk + k] = best; top_sim[i * topk + k] = sim[best]; votes[i * topk + k] = bank_cls[best];
    }
    /* mode of first knn votes, smallest value on ties */
    int64_t best_v = 0; int best_c = 0;
    for (int a = 0; a < knn; a++) {
      int64_t v = votes[i * topk + a]; int c = 0;
      for (int b2 = 0; b2 < knn; b2++) if (votes[i * topk + b2] == v) c++;
      if (c > best_c || (c == best_c && v < best_v)) { best_c = c; best_v = v; }
    }
    nn_class[i] = best_v;
    keep[i] = (q_cls[i] == best_v) ? 1 : 0;
  }
  free(mean); free(bc); free(bn); free(qc); free(sim);
}

/* ------------------------------------------------------------------------------------------
 * Row softmax (F.softmax, lvc/modeling/roi_heads/fast_rcnn.py:460-468).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_softmax_rows(const float* x, int64_t R, int K, float* out) {
  for (int64_t r = 0; r < R; r++) {
    const float* xr = x + r * K; float* o = out + r * K;
    float mx = xr[0]; for (int k = 1; k < K; k++) if (xr[k] > mx) mx = xr[k];
    double s = 0; for (int k = 0; k < K; k++) { o[k] = expf(xr[k] - mx); s += o[k]; }
    float fs = (float)s; for (int k = 0; k < K; k++) o[k] = o[k] / fs;
  }
}

/* ------------------------------------------------------------------------------------------
 * "Next" row f4: training-side users of the same operators.
 *
 * RoIAlign backward, NCHW fp32.  Follows detectron2/layers/csrc/ROIAlign/ROIAlign_cuda.cu:141-306
 * (bilinear_interpolate_gradient + RoIAlignBackwardFeature; the CPU twin is ROIAlign_cpu.cpp:220-330), the in-tree
 * spec of the autograd backward of torchvision.ops.roi_align reached from detectron2/layers/roi_align.py:15.  grad_in
 * is accumulated in roi / channel / bin / sample order (the reference's CUDA kernel uses atomicAdd: order unspecified).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_roi_align_backward(const float* grad_out, int N, int C, int H, int W, const float* rois, int R,
                                    int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                                    float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)N * C * H * W);
  for (int n = 0; n < R; n++) {
    const float* roi = rois + 5 * n;
    int batch = (int)roi[0];
    float offset = aligned ? 0.5f : 0.0f;
    float roi_start_w = roi[1] * spatial_scale - offset, roi_start_h = roi[2] * spatial_scale - offset;
    float roi_end_w = roi[3] * spatial_scale - offset, roi_end_h = roi[4] * spatial_scale - offset;
    float roi_width = roi_end_w - roi_start_w, roi_height = roi_end_h - roi_start_h;
    if (!aligned) {
      roi_width = roi_width > 1.f ? roi_width : 1.f;
      roi_height = roi_height > 1.f ? roi_height : 1.f;
    }
    float bin_size_h = roi_height / (float)pooled_h, bin_size_w = roi_width / (float)pooled_w;
    int grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)pooled_h);
    int grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)pooled_w);
    float count = (float)(grid_h * grid_w);
    for (int c = 0; c < C; c++) {
      float* gin = grad_in + ((size_t)batch * C + c) * H * W;
      for (int ph = 0; ph < pooled_h; ph++)
        for (int pw = 0; pw < pooled_w; pw++) {
          float g = grad_out[(((size_t)n * C + c) * pooled_h + ph) * pooled_w + pw];
          for (int iy = 0; iy < grid_h; iy++) {
            float yy = roi_start_h + ph * bin_size_h + ((float)iy + .5f) * bin_size_h / (float)grid_h;
            for (int ix = 0; ix < grid_w; ix++) {
              float xx = roi_start_w + pw * bin_size_w + ((float)ix + .5f) * bin_size_w / (float)grid_w;
              float x = xx, y = yy;
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;   /* weights 0, indices -1 */
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int y_low = (int)y, x_low = (int)x, y_high, x_high;
              if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
              if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
              float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              gin[y_low * W + x_low] += g * w1 / count;
              gin[y_low * W + x_high] += g * w2 / count;
              gin[y_high * W + x_low] += g * w3 / count;
              gin[y_high * W + x_high] += g * w4 / count;
            }
          }
        }
    }
  }
}

/* pairwise_iou (detectron2/structures/boxes.py:315-347; Boxes.area :172-181): separately rounded fp32 ops, iou = 0 where the
 * intersection is empty.  iou is [G, P] row-major (G = boxes1 = ground truth in the matcher's use). */
static float orc_iou_pair(const float* a, const float* b) {
  float area1 = (a[2] - a[0]) * (a[3] - a[1]), area2 = (b[2] - b[0]) * (b[3] - b[1]);
  float w = (a[2] < b[2] ? a[2] : b[2]) - (a[0] > b[0] ? a[0] : b[0]);
  float h = (a[3] < b[3] ? a[3] : b[3]) - (a[1] > b[1] ? a[1] : b[1]);
  if (w < 0.f) w = 0.f;
  if (h < 0.f) h = 0.f;
  float inter = w * h;
  return inter > 0.f ? inter / ((area1 + area2) - inter) : 0.f;
}
ORC_API void orc_pairwise_iou(const float* boxes1, int64_t G, const float* boxes2, int64_t P, float* iou) {
  for (int64_t g = 0; g < G; g++)
    for (int64_t p = 0; p < P; p++) iou[g * P + p] = orc_iou_pair(boxes1 + 4 * g, boxes2 + 4 * p);
}

/* Matcher.__call__ + set_low_quality_matches_ (detectron2/modeling/matcher.py:61-126) on the pairwise IoU of gt [G,4] and
 * proposal / anchor boxes [P,4]: matches[p] = argmax_g iou (first maximum), labels by thresholds (n_thr interior thresholds,
 * n_thr + 1 labels; the reference's list has -inf / +inf added at the ends, :52-54), optional low-quality promotion (every
 * prediction that attains a gt's maximum IoU, ties included, is labelled 1).  G == 0: matches 0, labels[0] (:76-87). */
ORC_API void orc_match_boxes(const float* gt, int64_t G, const float* boxes, int64_t P, const float* thresholds, int n_thr,
                             const int8_t* labels, int allow_low_quality, int64_t* matches, int8_t* match_labels, float* matched_vals) {
  if (G == 0) {
    for (int64_t p = 0; p < P; p++) { matches[p] = 0; match_labels[p] = labels[0]; if (matched_vals) matched_vals[p] = 0.f; }
    return;
  }
  for (int64_t p = 0; p < P; p++) {
    float best = -1.f; int64_t arg = 0;
    for (int64_t g = 0; g < G; g++) {
      float v = orc_iou_pair(gt + 4 * g, boxes + 4 * p);
      if (v > best) { best = v; arg = g; }
    }
    matches[p] = arg;
    if (matched_vals) matched_vals[p] = best;
    int8_t l = 1;                                   /* new_full(..., 1), then every (low <= v < high) band overwrites */
    for (int i = 0; i <= n_thr; i++) {
      float low = i == 0 ? -INFINITY : thresholds[i - 1], high = i == n_thr ? INFINITY : thresholds[i];
      if (best >= low && best < high) l = labels[i];
    }
    match_labels[p] = l;
  }
  if (allow_low_quality)
    for (int64_t g = 0; g < G; g++) {
      float hi = -1.f;
      for (int64_t p = 0; p < P; p++) { float v = orc_iou_pair(gt + 4 * g, boxes + 4 * p); if (v > hi) hi = v; }
      for (int64_t p = 0; p < P; p++) if (orc_iou_pair(gt + 4 * g, boxes + 4 * p) == hi) match_labels[p] = 1;
    }
}

/* RPN.losses (detectron2/modeling/proposal_generator/rpn.py:328-400) before normalisation: the two sums
 *   objectness = sum over anchors with label >= 0 of binary_cross_entropy_with_logits(logit, label)      (:387-392)
 *   localization = sum over anchors with label == 1 of smooth_l1(pred_delta - get_deltas(anchor, gt_box), beta)   (:367-376)
 * with Box2BoxTransform.get_deltas (box_regression.py:38-71) and fvcore's smooth_l1_loss (beta < 1e-5 -> L1).  Accumulated in
 * double; the reference sums fp32 tensors with torch's pairwise reduction, so parity is to ~1e-6 relative. */
ORC_API void orc_rpn_losses(const float* anchors, const float* logits, const float* deltas, const int8_t* labels,
                            const float* gt_boxes, int64_t N, int64_t A, const float* weights, float beta, double* out2) {
  double cls = 0.0, loc = 0.0;
  for (int64_t n = 0; n < N; n++)
    for (int64_t a = 0; a < A; a++) {
      const int8_t l = labels[n * A + a];
      if (l >= 0) {
        float x = logits[n * A + a], y = (float)l;
        float mx = x > 0.f ? x : 0.f;
        cls += (double)(mx - x * y + log1pf(expf(-fabsf(x))));
      }
      if (l == 1) {
        const float* s = anchors + 4 * a;
        const float* t = gt_boxes + 4 * (n * A + a);
        float sw = s[2] - s[0], sh = s[3] - s[1], sx = s[0] + 0.5f * sw, sy = s[1] + 0.5f * sh;
        float tw = t[2] - t[0], th = t[3] - t[1], tx = t[0] + 0.5f * tw, ty = t[1] + 0.5f * th;
        float d[4] = {weights[0] * (tx - sx) / sw, weights[1] * (ty - sy) / sh, weights[2] * logf(tw / sw), weights[3] * logf(th / sh)};
        for (int k = 0; k < 4; k++) {
          float nn = fabsf(deltas[4 * (n * A + a) + k] - d[k]);
          loc += (double)(beta < 1e-5f ? nn : (nn < beta ? 0.5f * nn * nn / beta : nn - 0.5f * beta));
        }
      }
    }
  out2[0] = cls; out2[1] = loc;
}

/* FastRCNNOutputs.losses (lvc/modeling/roi_heads/fast_rcnn.py:267-279, 296-358, 424-438) before the division by R:
 *   cls = sum over rows of cross_entropy(logits[r, :], gt_classes[r])          (F.cross_entropy, log-sum-exp with the row maximum)
 *   box = sum over foreground rows (0 <= gt < K) of smooth_l1(deltas[r, 4*gt : 4*gt+4] - get_deltas(proposal, gt_box), beta)
 * class-agnostic regression when n_delta_cols == 4.  Accumulated in double. */
ORC_API void orc_fast_rcnn_losses(const float* logits, const float* deltas, int n_delta_cols, const int64_t* gt_classes,
                                  const float* proposals, const float* gt_boxes, int64_t R, int K, const float* weights, float beta,
                                  double* out2) {
  double cls = 0.0, box = 0.0;
  for (int64_t r = 0; r < R; r++) {
    const float* x = logits + r * (K + 1);
    float mx = x[0];
    for (int k = 1; k <= K; k++) mx = x[k] > mx ? x[k] : mx;
    float se = 0.f;
    for (int k = 0; k <= K; k++) se += expf(x[k] - mx);
    const int64_t g = gt_classes[r];
    cls += (double)(logf(se) + mx - x[g]);
    if (g >= 0 && g < K) {
      const float* s = proposals + 4 * r;
      const float* t = gt_boxes + 4 * r;
      float sw = s[2] - s[0], sh = s[3] - s[1], sx = s[0] + 0.5f * sw, sy = s[1] + 0.5f * sh;
      float tw = t[2] - t[0], th = t[3] - t[1], tx = t[0] + 0.5f * tw, ty = t[1] + 0.5f * th;
      float d[4] = {weights[0] * (tx - sx) / sw, weights[1] * (ty - sy) / sh, weights[2] * logf(tw / sw), weights[3] * logf(th / sh)};
      const float* pd = deltas + r * n_delta_cols + (n_delta_cols == 4 ? 0 : 4 * g);
      for (int k = 0; k < 4; k++) {
        float nn = fabsf(pd[k] - d[k]);
        box += (double)(beta < 1e-5f ? nn : (nn < beta ? 0.5f * nn * nn / beta : nn - 0.5f * beta));
      }
    }
  }
  out2[0] = cls; out2[1] = box;
}
