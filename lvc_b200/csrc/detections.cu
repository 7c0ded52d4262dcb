// Box-head post-processing on device (no nonzero()/host sync):
//   det_score_decode_kernel : warp per RoI -- softmax over K+1 logits (fast_rcnn.py:460-468), score > thresh mask
//       (:116-125), class-specific apply_deltas + clip for the surviving (roi, class) pairs only (:440-458, :110-113;
//       the reference decodes all R x K boxes), appended to per-(image, class) candidate lists.
//   det_class_nms_kernel    : CTA per (image, class) -- in-CTA sort by score, greedy NMS (== batched_nms over class ids, :128)
//   det_rank_kernel         : CTA per (image, class) -- global rank of the class's first topk kept candidates among the first topk
//       of every other class list (all lists are score-sorted, so nothing beyond a list's first topk entries can reach the image's
//       top-k, :129-131): the K x topk score table sits in shared memory, ranks come from binary searches in it.
//   det_finalize_kernel     : warp per image -- detector_postprocess: rescale, clip, drop empty (postprocessing.py:10-79), ordered.
//   det_merge_kernel        : the one-CTA-per-image form of the two (global-memory searches); kept for K * topk tables that do
//       not fit in shared memory ("all detections" calls).
#include "nms_core.cuh"
#include "sort_core.cuh"

namespace lvcb200 {

constexpr int kMaxRois = 1024;  // per image (POST_NMS_TOPK_TEST = 1000)

struct DetWs {
  size_t off_hdr, off_ccount, off_cscore, off_crow, off_cbox, off_kcount, off_tmp_boxes, off_tmp_scores, off_tmp_cls,
      off_tmp_rows, total;
};

static DetWs det_layout(int n_images, int K, int topk) {
  DetWs w; size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o = align_up(o + b, 256); return r; };
  size_t slots = (size_t)n_images * K;
  w.off_hdr = take(sizeof(uint32_t) * 2 * n_images);
  w.off_ccount = take(slots * 4);
  w.off_cscore = take(slots * kMaxRois * 4);
  w.off_crow = take(slots * kMaxRois * 4);
  w.off_cbox = take(slots * kMaxRois * 16);
  w.off_kcount = take(slots * 4);
  w.off_tmp_boxes = take((size_t)n_images * topk * 16);
  w.off_tmp_scores = take((size_t)n_images * topk * 4);
  w.off_tmp_cls = take((size_t)n_images * topk * 4);
  w.off_tmp_rows = take((size_t)n_images * topk * 4);
  w.total = o;
  return w;
}

__global__ void __launch_bounds__(256)
det_score_decode_kernel(const float* __restrict__ logits, int64_t logit_pitch, const float* __restrict__ row_scale,
                        const float* __restrict__ deltas, int64_t delta_pitch, const float4* __restrict__ proposals,
                        const int32_t* __restrict__ roi_image, const int32_t* __restrict__ first_row, int64_t R, int K,
                        int class_agnostic, float wx, float wy, float ww, float wh, float score_thresh,
                        const int32_t* __restrict__ image_sizes, int* __restrict__ ccount, float* __restrict__ cscore,
                        int* __restrict__ crow, float4* __restrict__ cbox, uint32_t* __restrict__ hdr) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const int img = roi_image[r];
  if (img < 0) return;  // padded row (image produced fewer proposals than the static per-image capacity)
  const int row_in_img = (int)(r - first_row[img]);
  const float* x = logits + r * logit_pitch;
  const float sc = row_scale ? row_scale[r] : 1.0f;
  float v[4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int c = lane + 32 * i;
    v[i] = (c <= K) ? __fmul_rn(x[c], sc) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) { v[i] = (lane + 32 * i <= K) ? expf(v[i] - mx) : 0.f; sum += v[i]; }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float4 p = proposals[r];
  const int ih = image_sizes[img * 2], iw = image_sizes[img * 2 + 1];
  const float widths = __fsub_rn(p.z, p.x), heights = __fsub_rn(p.w, p.y);
  const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, widths)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, heights));
  const float clampv = 4.135166556742356f;
  uint32_t mxc = 0u; unsigned int ncand = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int c = lane + 32 * i;
    if (c >= K) continue;  // background column K is dropped (fast_rcnn.py:109)
    float prob = __fdiv_rn(v[i], sum);
    if (!(prob > score_thresh)) continue;
    const float* d = deltas + r * delta_pitch + (class_agnostic ? 0 : c * 4);
    float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
    float dw = fminf(__fdiv_rn(d[2], ww), clampv), dh = fminf(__fdiv_rn(d[3], wh), clampv);
    float pcx = __fadd_rn(__fmul_rn(dx, widths), cx), pcy = __fadd_rn(__fmul_rn(dy, heights), cy);
    float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
    float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    x1 = fminf(fmaxf(x1, 0.f), (float)iw); y1 = fminf(fmaxf(y1, 0.f), (float)ih);
    x2 = fminf(fmaxf(x2, 0.f), (float)iw); y2 = fminf(fmaxf(y2, 0.f), (float)ih);
    int slot = img * K + c;
    int pos = atomicAdd(&ccount[slot], 1);
    if (pos < kMaxRois) {
      cscore[(size_t)slot * kMaxRois + pos] = prob;
      crow[(size_t)slot * kMaxRois + pos] = row_in_img;
      cbox[(size_t)slot * kMaxRois + pos] = make_float4(x1, y1, x2, y2);
    }
    mxc = max(mxc, float_to_ordered(fmaxf(fmaxf(x1, y1), fmaxf(x2, y2))));
    ncand++;
  }
  for (int o = 16; o; o >>= 1) { mxc = max(mxc, __shfl_xor_sync(0xffffffffu, mxc, o)); ncand += __shfl_xor_sync(0xffffffffu, ncand, o); }
  if (lane == 0 && ncand) { atomicMax(&hdr[img * 2], mxc); atomicAdd(&hdr[img * 2 + 1], ncand); }
}

__global__ void __launch_bounds__(256)
det_class_nms_kernel(int K, float thr, int nms_mode, const uint32_t* __restrict__ hdr, int* __restrict__ ccount,
                     float* __restrict__ cscore, int* __restrict__ crow, float4* __restrict__ cbox, int* __restrict__ kcount) {
  const int slot = blockIdx.x;
  int n = ccount[slot];
  if (n > kMaxRois) n = kMaxRois;
  if (n == 0) { if (threadIdx.x == 0) kcount[slot] = 0; return; }
  __shared__ NmsShared sh;
  __shared__ unsigned long long keys[kMaxRois];
  __shared__ float4 sbox[kMaxRois];
  __shared__ float kx1[kMaxRois], ky1[kMaxRois], kx2[kMaxRois], ky2[kMaxRois], kar[kMaxRois];
  __shared__ unsigned char flags[kMaxRois];
  __shared__ int warp_cnt[8];
  const int img = slot / K, cls = slot - img * K;
  float* sc = cscore + (size_t)slot * kMaxRois;
  int* rw = crow + (size_t)slot * kMaxRois;
  float4* bx = cbox + (size_t)slot * kMaxRois;
  int np2 = 64; while (np2 < n) np2 <<= 1;
  // sort by (score desc, row asc): key = score bits << 32 | (~row) << 10 | position   (n, row <= 1024)
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < n) k = ((unsigned long long)float_to_ordered(sc[i]) << 32) | ((unsigned long long)((~(unsigned)rw[i]) & 0x3fffffu) << 10) | (unsigned)i;
    keys[i] = k;
  }
  __syncthreads();
  bitonic_sort_desc(keys, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) sbox[i] = bx[keys[i] & 1023ull];
  __syncthreads();
  int mode = nms_mode;
  if (mode < 0) mode = reference_cuda_nms_mode((long long)hdr[img * 2 + 1]);
  const float off = (mode == 0) ? __fmul_rn((float)cls, __fadd_rn(ordered_to_float(hdr[img * 2]), 1.0f)) : 0.f;
  auto get = [&](int j, float& x1, float& y1, float& x2, float& y2) {
    float4 b = sbox[j];
    x1 = __fadd_rn(b.x, off); y1 = __fadd_rn(b.y, off); x2 = __fadd_rn(b.z, off); y2 = __fadd_rn(b.w, off);
    return true;
  };
  segment_nms(sh, get, n, thr, kx1, ky1, kx2, ky2, kar, flags);
  __syncthreads();
  // gather scores / rows of the sorted order into registers before overwriting the lists in place
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int base = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    int j = j0 + threadIdx.x;
    bool kf = j < n && flags[j];
    float s = 0.f; int row = 0;
    if (j < n) { int src = (int)(keys[j] & 1023ull); s = sc[src]; row = rw[src]; }
    __syncthreads();  // all reads of this tile's sources done? (sources may lie anywhere) -> see below
    unsigned int b = __ballot_sync(0xffffffffu, kf);
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w2 = 0; w2 < 8; w2++) { if (w2 < wid) pre += warp_cnt[w2]; tot += warp_cnt[w2]; }
    if (kf) {
      int pos = base + pre + __popc(b & ((1u << lane) - 1u));
      // write the kept list into the *box* array region reinterpretation-free: separate kept arrays live in smem
      kx1[pos] = s; ky1[pos] = __int_as_float(row);   // reuse smem (NMS finished): kx1 = score, ky1 = row
      kx2[pos] = __int_as_float(j);                     // sorted position -> box
    }
    base += tot;
    __syncthreads();
  }
  // now overwrite the global candidate lists with the kept, score-sorted entries
  for (int i = threadIdx.x; i < base; i += blockDim.x) {
    sc[i] = kx1[i]; rw[i] = __float_as_int(ky1[i]); bx[i] = sbox[__float_as_int(kx2[i])];
  }
  if (threadIdx.x == 0) kcount[slot] = base;
}

__global__ void __launch_bounds__(256)
det_merge_kernel(int K, int topk, const int* __restrict__ kcount, const float* __restrict__ cscore, const int* __restrict__ crow,
                 const float4* __restrict__ cbox, const int32_t* __restrict__ image_sizes, const int32_t* __restrict__ out_sizes,
                 float4* __restrict__ tmp_boxes, float* __restrict__ tmp_scores, int* __restrict__ tmp_cls, int* __restrict__ tmp_rows,
                 float4* __restrict__ det_boxes, float* __restrict__ det_scores, int64_t* __restrict__ det_classes,
                 int64_t* __restrict__ det_rows, int32_t* __restrict__ det_counts) {
  const int img = blockIdx.x;
  extern __shared__ int s_cnt[];  // K counts
  __shared__ int s_total;
  for (int c = threadIdx.x; c < K; c += blockDim.x) s_cnt[c] = kcount[img * K + c];
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int c = 0; c < K; c++) t += s_cnt[c]; s_total = t; }
  __syncthreads();
  const int total = s_total;
  const int n_out = total < topk ? total : topk;
  // one warp per class list; each kept element finds its global rank
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = wid; c < K; c += nw) {
    const int n = s_cnt[c];
    if (n == 0) continue;
    const size_t base = (size_t)(img * K + c) * kMaxRois;
    for (int p = lane; p < n; p += 32) {
      const float s = cscore[base + p];
      const int row = crow[base + p];
      int rank = p;
      for (int c2 = 0; c2 < K; c2++) {
        const int n2 = s_cnt[c2];
        if (c2 == c || n2 == 0) continue;
        const float* a = cscore + (size_t)(img * K + c2) * kMaxRois;
        int g = count_greater_desc(a, n2, s);
        // ties: the reference's stable sort keeps nonzero() order = (row, class) ascending
        const int* r2 = crow + (size_t)(img * K + c2) * kMaxRois;
        while (g < n2 && a[g] == s && (r2[g] < row || (r2[g] == row && c2 < c))) g++;
        rank += g;
      }
      if (rank < topk) {
        tmp_boxes[(size_t)img * topk + rank] = cbox[base + p];
        tmp_scores[(size_t)img * topk + rank] = s;
        tmp_cls[(size_t)img * topk + rank] = c;
        tmp_rows[(size_t)img * topk + rank] = row;
      }
    }
  }
  __syncthreads();
  // detector_postprocess: scale, clip, drop empty -- ordered compaction by one warp
  const float ih = (float)image_sizes[img * 2], iw = (float)image_sizes[img * 2 + 1];
  const int oh = out_sizes[img * 2], ow = out_sizes[img * 2 + 1];
  const float sx = __fdiv_rn((float)ow, iw), sy = __fdiv_rn((float)oh, ih);
  if (wid == 0) {
    int outp = 0;
    for (int i0 = 0; i0 < n_out; i0 += 32) {
      int i = i0 + lane;
      bool ok = false; float4 b = make_float4(0, 0, 0, 0);
      if (i < n_out) {
        b = tmp_boxes[(size_t)img * topk + i];
        b.x = fminf(fmaxf(__fmul_rn(b.x, sx), 0.f), (float)ow); b.y = fminf(fmaxf(__fmul_rn(b.y, sy), 0.f), (float)oh);
        b.z = fminf(fmaxf(__fmul_rn(b.z, sx), 0.f), (float)ow); b.w = fminf(fmaxf(__fmul_rn(b.w, sy), 0.f), (float)oh);
        ok = (__fsub_rn(b.z, b.x) > 0.f) && (__fsub_rn(b.w, b.y) > 0.f);
      }
      unsigned int m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        size_t o = (size_t)img * topk + outp + __popc(m & ((1u << lane) - 1u));
        det_boxes[o] = b; det_scores[o] = tmp_scores[(size_t)img * topk + i];
        det_classes[o] = tmp_cls[(size_t)img * topk + i]; det_rows[o] = tmp_rows[(size_t)img * topk + i];
      }
      outp += __popc(m);
    }
    for (int i = outp + lane; i < topk; i += 32) {
      size_t o = (size_t)img * topk + i;
      det_boxes[o] = make_float4(0, 0, 0, 0); det_scores[o] = 0.f; det_classes[o] = -1; det_rows[o] = -1;
    }
    if (lane == 0) det_counts[img] = outp;
  }
}

// Parallel form of the merge.  With a realistic score distribution (or SCORE_THRESH_TEST 0.0) tens of thousands of candidates
// per image survive the per-class NMS; the single-CTA merge above then spends milliseconds in dependent global-memory binary
// searches (0.95 ms for 8 images of random logits, ~15 ms at threshold 0 -- profiles/r02_ops_ncu.md).  Only the first topk entries
// of each sorted class list can reach the image's top-k, so K * topk scores (32 KB for 80 x 100) are all that has to be searched.
__global__ void __launch_bounds__(128)
det_rank_kernel(int K, int topk, const int* __restrict__ kcount, const float* __restrict__ cscore, const int* __restrict__ crow,
                const float4* __restrict__ cbox, float4* __restrict__ tmp_boxes, float* __restrict__ tmp_scores, int* __restrict__ tmp_cls,
                int* __restrict__ tmp_rows) {
  extern __shared__ int dr_smem[];
  int* s_cnt = dr_smem;                                        // K counts, clamped to topk
  float* s_score = reinterpret_cast<float*>(dr_smem + K);      // [K][topk]
  const int img = blockIdx.y, c = blockIdx.x;
  const int n = min(kcount[img * K + c], topk);
  if (n == 0) return;
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_cnt[i] = min(kcount[img * K + i], topk);
  __syncthreads();
  for (int i = threadIdx.x; i < K * topk; i += blockDim.x) {
    const int c2 = i / topk, p = i - c2 * topk;
    s_score[i] = p < s_cnt[c2] ? cscore[(size_t)(img * K + c2) * kMaxRois + p] : -INFINITY;
  }
  __syncthreads();
  const size_t base = (size_t)(img * K + c) * kMaxRois;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const float s = s_score[c * topk + p];
    const int row = crow[base + p];
    int rank = p;
    for (int c2 = 0; c2 < K && rank < topk; c2++) {
      const int n2 = s_cnt[c2];
      if (c2 == c || n2 == 0) continue;
      const float* a = s_score + c2 * topk;
      int g = count_greater_desc(a, n2, s);
      if (g < n2 && a[g] == s) {   // ties: the reference's stable sort keeps nonzero() order = (row, class) ascending
        const int* r2 = crow + (size_t)(img * K + c2) * kMaxRois;
        while (g < n2 && a[g] == s && (r2[g] < row || (r2[g] == row && c2 < c))) g++;
      }
      rank += g;
    }
    if (rank < topk) {
      tmp_boxes[(size_t)img * topk + rank] = cbox[base + p];
      tmp_scores[(size_t)img * topk + rank] = s;
      tmp_cls[(size_t)img * topk + rank] = c;
      tmp_rows[(size_t)img * topk + rank] = row;
    }
  }
}

__global__ void __launch_bounds__(32)
det_finalize_kernel(int K, int topk, const int* __restrict__ kcount, const int32_t* __restrict__ image_sizes,
                    const int32_t* __restrict__ out_sizes, const float4* __restrict__ tmp_boxes, const float* __restrict__ tmp_scores,
                    const int* __restrict__ tmp_cls, const int* __restrict__ tmp_rows, float4* __restrict__ det_boxes,
                    float* __restrict__ det_scores, int64_t* __restrict__ det_classes, int64_t* __restrict__ det_rows,
                    int32_t* __restrict__ det_counts) {
  const int img = blockIdx.x, lane = threadIdx.x;
  int total = 0;
  for (int c = lane; c < K; c += 32) total += kcount[img * K + c];
  for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  const int n_out = total < topk ? total : topk;
  const float ih = (float)image_sizes[img * 2], iw = (float)image_sizes[img * 2 + 1];
  const int oh = out_sizes[img * 2], ow = out_sizes[img * 2 + 1];
  const float sx = __fdiv_rn((float)ow, iw), sy = __fdiv_rn((float)oh, ih);
  int outp = 0;
  for (int i0 = 0; i0 < n_out; i0 += 32) {
    int i = i0 + lane;
    bool ok = false; float4 b = make_float4(0, 0, 0, 0);
    if (i < n_out) {
      b = tmp_boxes[(size_t)img * topk + i];
      b.x = fminf(fmaxf(__fmul_rn(b.x, sx), 0.f), (float)ow); b.y = fminf(fmaxf(__fmul_rn(b.y, sy), 0.f), (float)oh);
      b.z = fminf(fmaxf(__fmul_rn(b.z, sx), 0.f), (float)ow); b.w = fminf(fmaxf(__fmul_rn(b.w, sy), 0.f), (float)oh);
      ok = (__fsub_rn(b.z, b.x) > 0.f) && (__fsub_rn(b.w, b.y) > 0.f);
    }
    unsigned int m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      size_t o = (size_t)img * topk + outp + __popc(m & ((1u << lane) - 1u));
      det_boxes[o] = b; det_scores[o] = tmp_scores[(size_t)img * topk + i];
      det_classes[o] = tmp_cls[(size_t)img * topk + i]; det_rows[o] = tmp_rows[(size_t)img * topk + i];
    }
    outp += __popc(m);
  }
  for (int i = outp + lane; i < topk; i += 32) {
    size_t o = (size_t)img * topk + i;
    det_boxes[o] = make_float4(0, 0, 0, 0); det_scores[o] = 0.f; det_classes[o] = -1; det_rows[o] = -1;
  }
  if (lane == 0) det_counts[img] = outp;
}

// Box2BoxTransform.apply_deltas (box_regression.py:73-110) for class-agnostic deltas + Boxes.clip: the per-stage box update of
// the box corrector (BoxOnlyLayersCascade.predict_boxes roi_heads_cascade.py:197-211; _create_proposals_from_boxes
// cascade_rcnn.py:348-369).
__global__ void apply_deltas_clip_kernel(const float4* __restrict__ boxes, const float* __restrict__ deltas, int64_t delta_pitch,
                                         const int32_t* __restrict__ roi_image, const int32_t* __restrict__ image_sizes, int64_t R,
                                         float wx, float wy, float ww, float wh, int clip, float4* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float4 p = boxes[r];
  const float* d = deltas + r * delta_pitch;
  const float clampv = 4.135166556742356f;
  const float widths = __fsub_rn(p.z, p.x), heights = __fsub_rn(p.w, p.y);
  const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, widths)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, heights));
  float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
  float dw = fminf(__fdiv_rn(d[2], ww), clampv), dh = fminf(__fdiv_rn(d[3], wh), clampv);
  float pcx = __fadd_rn(__fmul_rn(dx, widths), cx), pcy = __fadd_rn(__fmul_rn(dy, heights), cy);
  float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  if (clip) {
    const int img = roi_image[r];
    const float ih = (float)image_sizes[img * 2], iw = (float)image_sizes[img * 2 + 1];
    x1 = fminf(fmaxf(x1, 0.f), iw); y1 = fminf(fmaxf(y1, 0.f), ih);
    x2 = fminf(fmaxf(x2, 0.f), iw); y2 = fminf(fmaxf(y2, 0.f), ih);
  }
  out[r] = make_float4(x1, y1, x2, y2);
}

__global__ void det_first_row_kernel(const int32_t* __restrict__ roi_image, int64_t R, int32_t* __restrict__ first_row) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (roi_image[r] >= 0 && (r == 0 || roi_image[r] != roi_image[r - 1])) first_row[roi_image[r]] = (int32_t)r;
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" size_t lvcb200_detections_workspace(const lvcb200_det_params* p) {
  if (!p || p->n_images <= 0) return 256;
  return det_layout(p->n_images, p->num_classes, p->topk_per_image).total + align_up(sizeof(int32_t) * p->n_images, 256);
}

extern "C" int lvcb200_detections(const float* cls_logits, int64_t logit_pitch, const float* row_scale, const float* box_deltas,
                                  int64_t delta_pitch, const float* proposals, const int32_t* roi_image, int64_t R,
                                  const lvcb200_det_params* p, const int32_t* image_sizes, const int32_t* out_sizes,
                                  float* det_boxes, float* det_scores, int64_t* det_classes, int64_t* det_rows,
                                  int32_t* det_counts, void* workspace, size_t workspace_bytes, void* stream) {
  LVC_REQUIRE(p, "detections: NULL params");
  LVC_REQUIRE(p->num_classes >= 1 && p->num_classes + 1 <= 128, "detections: num_classes + 1 must be <= 128");
  LVC_REQUIRE(p->max_rois_per_image >= 1 && p->max_rois_per_image <= kMaxRois, "detections: max_rois_per_image must be <= 1024");
  LVC_REQUIRE(p->topk_per_image >= 1, "detections: topk_per_image must be positive (pass R for 'all')");
  LVC_REQUIRE(p->n_images >= 1, "detections: n_images must be >= 1");
  LVC_REQUIRE(det_boxes && det_scores && det_classes && det_rows && det_counts && workspace && image_sizes && out_sizes,
              "detections: NULL pointer");
  LVC_REQUIRE(((uintptr_t)proposals % 16) == 0 && ((uintptr_t)det_boxes % 16) == 0, "detections: box arrays must be 16-byte aligned");
  if (workspace_bytes < lvcb200_detections_workspace(p)) return set_error(LVCB200_EWORKSPACE, "detections: workspace too small");
  const int K = p->num_classes, topk = p->topk_per_image;
  DetWs w = det_layout(p->n_images, K, topk);
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  uint32_t* hdr = (uint32_t*)(ws + w.off_hdr);
  int* ccount = (int*)(ws + w.off_ccount);
  int32_t* first_row = (int32_t*)(ws + w.total);
  LVC_CUDA(cudaMemsetAsync(hdr, 0, sizeof(uint32_t) * 2 * p->n_images, s));
  LVC_CUDA(cudaMemsetAsync(ccount, 0, sizeof(int) * (size_t)p->n_images * K, s));
  LVC_CUDA(cudaMemsetAsync(first_row, 0, sizeof(int32_t) * p->n_images, s));
  int rc;
  if (R > 0) {
    LVC_REQUIRE(cls_logits && box_deltas && proposals && roi_image, "detections: NULL input");
    det_first_row_kernel<<<(unsigned)ceil_div64(R, 256), 256, 0, s>>>(roi_image, R, first_row);
    if ((rc = check_launch("det_first_row_kernel"))) return rc;
    det_score_decode_kernel<<<(unsigned)ceil_div64(R * 32, 256), 256, 0, s>>>(
        cls_logits, logit_pitch, row_scale, box_deltas, delta_pitch, (const float4*)proposals, roi_image, first_row, R, K,
        p->class_agnostic, p->weights[0], p->weights[1], p->weights[2], p->weights[3], p->score_thresh, image_sizes, ccount,
        (float*)(ws + w.off_cscore), (int*)(ws + w.off_crow), (float4*)(ws + w.off_cbox), hdr);
    if ((rc = check_launch("det_score_decode_kernel"))) return rc;
  }
  det_class_nms_kernel<<<p->n_images * K, 256, 0, s>>>(K, p->nms_thresh, p->nms_mode, hdr, ccount, (float*)(ws + w.off_cscore),
                                                       (int*)(ws + w.off_crow), (float4*)(ws + w.off_cbox), (int*)(ws + w.off_kcount));
  if ((rc = check_launch("det_class_nms_kernel"))) return rc;
  const size_t rank_smem = sizeof(int) * K + sizeof(float) * (size_t)K * topk;
  if (rank_smem <= 160 * 1024) {   // the usual case (80 x 100: 32 KB): parallel rank + finalize
    static size_t smem_set = 0;
    if (rank_smem > smem_set && rank_smem > 48 * 1024) {
      LVC_CUDA(cudaFuncSetAttribute(det_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_smem));
      smem_set = rank_smem;
    }
    det_rank_kernel<<<dim3(K, p->n_images), 128, rank_smem, s>>>(
        K, topk, (const int*)(ws + w.off_kcount), (const float*)(ws + w.off_cscore), (const int*)(ws + w.off_crow),
        (const float4*)(ws + w.off_cbox), (float4*)(ws + w.off_tmp_boxes), (float*)(ws + w.off_tmp_scores), (int*)(ws + w.off_tmp_cls),
        (int*)(ws + w.off_tmp_rows));
    if ((rc = check_launch("det_rank_kernel"))) return rc;
    det_finalize_kernel<<<p->n_images, 32, 0, s>>>(
        K, topk, (const int*)(ws + w.off_kcount), image_sizes, out_sizes, (const float4*)(ws + w.off_tmp_boxes),
        (const float*)(ws + w.off_tmp_scores), (const int*)(ws + w.off_tmp_cls), (const int*)(ws + w.off_tmp_rows), (float4*)det_boxes,
        det_scores, det_classes, det_rows, det_counts);
    return check_launch("det_finalize_kernel");
  }
  det_merge_kernel<<<p->n_images, 256, sizeof(int) * K, s>>>(
      K, topk, (const int*)(ws + w.off_kcount), (const float*)(ws + w.off_cscore), (const int*)(ws + w.off_crow),
      (const float4*)(ws + w.off_cbox), image_sizes, out_sizes, (float4*)(ws + w.off_tmp_boxes), (float*)(ws + w.off_tmp_scores),
      (int*)(ws + w.off_tmp_cls), (int*)(ws + w.off_tmp_rows), (float4*)det_boxes, det_scores, det_classes, det_rows, det_counts);
  return check_launch("det_merge_kernel");
}

extern "C" int lvcb200_apply_deltas_clip(const float* boxes, const float* deltas, int64_t delta_pitch, const int32_t* roi_image,
                                         const int32_t* image_sizes, int64_t R, const float* weights4 /*host*/, int clip, float* out,
                                         void* stream) {
  if (R == 0) return 0;
  LVC_REQUIRE(boxes && deltas && out && weights4, "apply_deltas_clip: NULL pointer");
  LVC_REQUIRE(!clip || (roi_image && image_sizes), "apply_deltas_clip: clip needs roi_image and image_sizes");
  LVC_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out % 16) == 0, "apply_deltas_clip: boxes must be 16-byte aligned");
  apply_deltas_clip_kernel<<<(unsigned)ceil_div64(R, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)boxes, deltas, delta_pitch, roi_image, image_sizes, R, weights4[0], weights4[1], weights4[2], weights4[3], clip,
      (float4*)out);
  return check_launch("apply_deltas_clip_kernel");
}
