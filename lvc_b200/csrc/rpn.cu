// RPN post-processing on device, no host round trips:
//   rpn_select_decode_kernel : per (image, level) CTA -- exact radix-select of the top-k objectness logits
//       (same set and order as the reference's full descending sort + slice, proposal_utils.py:59-74),
//       in-CTA sort, anchors regenerated in registers (anchor_generator.py:157-208), Box2BoxTransform.apply_deltas
//       on the selected anchors only (box_regression.py:73-110), finite / clip / non-empty filters
//       (proposal_utils.py:88-102).  The reference decodes all 268 569 anchors per image to keep 4 819.
//   rpn_nms_kernel           : per (image, level) CTA -- greedy NMS (nms_core.cuh) == batched_nms(boxes, scores, lvl)
//   rpn_merge_kernel         : per image CTA -- merge the per-level kept lists by score, first post_nms_topk.
#include "nms_core.cuh"
#include "sort_core.cuh"

namespace lvcb200 {

constexpr int kMaxTopk = 1024;
constexpr int kMaxLevels = 8;
constexpr int kScanChunk = 2048;   // anchors per CTA in the grid-wide scan kernels
constexpr int kCandCap = 4096;     // candidate list capacity per (image, level)

struct RpnLevels {
  lvcb200_rpn_level lv[kMaxLevels];
  int chunk_prefix[kMaxLevels + 1];  // CTA decomposition of one image's anchors over levels (scan kernels)
};

struct RpnWs {  // element offsets inside the workspace (per (image, level) slot of kMaxTopk entries)
  size_t off_hdr, off_boxes, off_scores, off_valid, off_kboxes, off_kscores, off_kcount, off_hist1, off_hist2, off_ccount, off_cand,
      off_mask, zero_bytes, total;
};

static RpnWs rpn_layout(int n_images, int n_levels) {
  RpnWs w; size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o = align_up(o + b, 256); return r; };
  size_t slots = (size_t)n_images * n_levels;
  // zero-initialised header region (one memset): per-image hdr, histograms, candidate counters
  w.off_hdr = take(sizeof(uint32_t) * 2 * n_images);  // per image: max coordinate (ordered), valid count
  w.off_hist1 = take(slots * 2048 * 4);
  w.off_hist2 = take(slots * 2048 * 4);
  w.off_ccount = take(slots * 4);
  w.zero_bytes = o;
  w.off_cand = take(slots * kCandCap * 8);
  w.off_mask = take(slots * (size_t)kMaxTopk * (kMaxTopk / 64) * 8);
  w.off_boxes = take(slots * kMaxTopk * 16);
  w.off_scores = take(slots * kMaxTopk * 4);
  w.off_valid = take(slots * kMaxTopk);
  w.off_kboxes = take(slots * kMaxTopk * 16);
  w.off_kscores = take(slots * kMaxTopk * 4);
  w.off_kcount = take(slots * 4);
  w.total = o;
  return w;
}

__device__ __forceinline__ int64_t rpn_addr(int i, int A, int W, int64_t row_stride, int64_t pix_stride, int mult) {
  int pix = i / A, a = i - pix * A;
  if (row_stride == 0) return (int64_t)pix * pix_stride + a * mult;
  int y = pix / W, x = pix - y * W;
  return (int64_t)y * row_stride + (int64_t)x * pix_stride + a * mult;
}


// ---- grid-wide top-k selection, phase 1-3 (the single-CTA select in rpn_select_decode_kernel is the fallback for
// degenerate inputs whose 22-bit key prefix does not separate the k-th value, e.g. an all-equal logit map).
__device__ __forceinline__ bool rpn_scan_range(const RpnLevels& L, int n_levels, int& img, int& lvl, int& beg, int& end) {
  const int per_img = L.chunk_prefix[n_levels];
  img = blockIdx.x / per_img;
  const int c = blockIdx.x - img * per_img;
  lvl = 0;
  while (lvl + 1 < n_levels && c >= L.chunk_prefix[lvl + 1]) lvl++;
  const int n = L.lv[lvl].H * L.lv[lvl].W * L.lv[lvl].A;
  beg = (c - L.chunk_prefix[lvl]) * kScanChunk;
  end = min(beg + kScanChunk, n);
  return beg < n;
}

// digit (from the top) at which the suffix count of hist reaches `need`; returns digit, writes count strictly above it.
// The 2048-bin histogram is staged in shared memory first (coalesced): every CTA of the scan kernels runs this, and walking the bins in
// global memory (64 strided loads per lane, then a serial walk) cost ~15 us per call -- more than the CTA's own 2048-anchor scan.
__device__ __forceinline__ int rpn_find_digit(const unsigned int* __restrict__ hist, unsigned int need, unsigned int* above_out,
                                              unsigned int* s_tmp /* 2 words smem */, unsigned int* s_hist /* 2048 words smem */) {
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_hist[i] = hist[i];
  __syncthreads();
  if (threadIdx.x < 32) {
    const int per = 64;
    unsigned int local = 0;
    for (int b = 0; b < per; b++) local += s_hist[lane * per + ((b + lane) & (per - 1))];   // rotated: bank-conflict free
    unsigned int suffix = local;
    for (int o = 1; o < 32; o <<= 1) { unsigned int v = __shfl_down_sync(0xffffffffu, suffix, o); if (lane + o < 32) suffix += v; }
    unsigned int above = suffix - local;
    if (above < need && suffix >= need) {
      unsigned int acc = above; int d = 0;
      for (int b = per - 1; b >= 0; b--) {
        unsigned int h = s_hist[lane * per + b];
        if (acc + h >= need) { d = lane * per + b; break; }
        acc += h;
      }
      s_tmp[0] = (unsigned)d; s_tmp[1] = acc;
    }
  }
  __syncthreads();
  *above_out = s_tmp[1];
  return (int)s_tmp[0];
}

__global__ void __launch_bounds__(256)
rpn_hist1_kernel(RpnLevels L, int n_levels, unsigned int* __restrict__ hist1) {
  __shared__ unsigned int h[2048];
  int img, lvl, beg, end;
  if (!rpn_scan_range(L, n_levels, img, lvl, beg, end)) return;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const lvcb200_rpn_level& lv = L.lv[lvl];
  const float* logits = lv.logits + (int64_t)img * lv.img_stride_l;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x)
    atomicAdd(&h[float_to_ordered(__ldg(logits + rpn_addr(i, lv.A, lv.W, lv.row_stride_l, lv.pix_stride_l, 1))) >> 21], 1u);
  __syncthreads();
  unsigned int* g = hist1 + (size_t)(img * n_levels + lvl) * 2048;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) if (h[i]) atomicAdd(&g[i], h[i]);
}

__global__ void __launch_bounds__(256)
rpn_hist2_kernel(RpnLevels L, int n_levels, int topk, const unsigned int* __restrict__ hist1, unsigned int* __restrict__ hist2) {
  __shared__ unsigned int h[2048], sh[2048];
  __shared__ unsigned int tmp[2];
  int img, lvl, beg, end;
  if (!rpn_scan_range(L, n_levels, img, lvl, beg, end)) return;
  const lvcb200_rpn_level& lv = L.lv[lvl];
  const int n = lv.H * lv.W * lv.A;
  const unsigned int k = (unsigned)(topk < n ? topk : n);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) h[i] = 0u;
  unsigned int above;
  const unsigned int d1 = (unsigned)rpn_find_digit(hist1 + (size_t)(img * n_levels + lvl) * 2048, k, &above, tmp, sh);
  const float* logits = lv.logits + (int64_t)img * lv.img_stride_l;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    unsigned int key = float_to_ordered(__ldg(logits + rpn_addr(i, lv.A, lv.W, lv.row_stride_l, lv.pix_stride_l, 1)));
    if ((key >> 21) == d1) atomicAdd(&h[(key >> 10) & 2047u], 1u);
  }
  __syncthreads();
  unsigned int* g = hist2 + (size_t)(img * n_levels + lvl) * 2048;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) if (h[i]) atomicAdd(&g[i], h[i]);
}

__global__ void __launch_bounds__(256)
rpn_collect_kernel(RpnLevels L, int n_levels, int topk, const unsigned int* __restrict__ hist1, const unsigned int* __restrict__ hist2,
                   unsigned int* __restrict__ ccount, unsigned long long* __restrict__ cand) {
  __shared__ unsigned int tmp[2], sh[2048];
  int img, lvl, beg, end;
  if (!rpn_scan_range(L, n_levels, img, lvl, beg, end)) return;
  const lvcb200_rpn_level& lv = L.lv[lvl];
  const int n = lv.H * lv.W * lv.A;
  const unsigned int k = (unsigned)(topk < n ? topk : n);
  const int slot = img * n_levels + lvl;
  unsigned int above1, above2;
  const unsigned int d1 = (unsigned)rpn_find_digit(hist1 + (size_t)slot * 2048, k, &above1, tmp, sh);
  __syncthreads();
  const unsigned int d2 = (unsigned)rpn_find_digit(hist2 + (size_t)slot * 2048, k - above1, &above2, tmp, sh);
  const unsigned int t22 = (d1 << 11) | d2;   // every key with a 22-bit prefix >= t22 is a candidate (>= k of them)
  const float* logits = lv.logits + (int64_t)img * lv.img_stride_l;
  const int lane = threadIdx.x & 31;
  for (int i0 = beg; i0 < end; i0 += blockDim.x) {
    int i = i0 + threadIdx.x;
    unsigned int key = 0; bool take = false;
    if (i < end) {
      key = float_to_ordered(__ldg(logits + rpn_addr(i, lv.A, lv.W, lv.row_stride_l, lv.pix_stride_l, 1)));
      take = (key >> 10) >= t22;
    }
    unsigned int m = __ballot_sync(0xffffffffu, take);
    if (m) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(&ccount[slot], (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) {
        unsigned int pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < (unsigned)kCandCap) cand[(size_t)slot * kCandCap + pos] = ((unsigned long long)key << 32) | (unsigned int)(~(unsigned int)i);
      }
    }
  }
}

__global__ void __launch_bounds__(1024)
rpn_select_decode_kernel(RpnLevels L, int n_levels, int topk, float min_box_size, float wx, float wy, float ww, float wh,
                         const int32_t* __restrict__ image_sizes, float4* __restrict__ ws_boxes, float* __restrict__ ws_scores,
                         unsigned char* __restrict__ ws_valid, uint32_t* __restrict__ hdr, const unsigned int* __restrict__ ccount,
                         const unsigned long long* __restrict__ cand) {
  __shared__ unsigned int hist[2048];
  __shared__ unsigned long long sel[kCandCap];
  __shared__ unsigned int warp_gt[32], warp_eq[32];
  __shared__ unsigned int s_prefix, s_need;
  const int img = blockIdx.x / n_levels, lvl = blockIdx.x % n_levels;
  const lvcb200_rpn_level& lv = L.lv[lvl];
  const int A = lv.A, W = lv.W, n = lv.H * lv.W * lv.A;
  const int k = topk < n ? topk : n;
  const float* logits = lv.logits + (int64_t)img * lv.img_stride_l;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  const unsigned int n_cand = ccount ? ccount[blockIdx.x] : 0xffffffffu;
  if (n_cand <= (unsigned)kCandCap) {
    // ---- fast path: the grid-wide scan kernels left <= 4096 candidates (all keys >= the 22-bit threshold prefix);
    //      sort them (key desc, index asc): the first k are the reference's sorted top-k
    for (int i = tid; i < kCandCap; i += blockDim.x) sel[i] = (i < (int)n_cand) ? cand[(size_t)blockIdx.x * kCandCap + i] : 0ull;
    __syncthreads();
    int np2 = 1024; while (np2 < (int)n_cand) np2 <<= 1;
    bitonic_sort_desc(sel, np2);
  } else {
  // ---- fallback: exact radix select of the k-th largest key over all anchors by this CTA alone: 11 + 11 + 10 bits
  unsigned int prefix = 0u, mask = 0u, need = (unsigned)k;
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
    const int sh = shifts[pass];
    const unsigned int nb = 1u << bits[pass];
    for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      unsigned int key = float_to_ordered(__ldg(logits + rpn_addr(i, A, W, lv.row_stride_l, lv.pix_stride_l, 1)));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> sh) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (wid == 0) {  // find the digit where the suffix count (from the top) reaches `need`
      const int per = 2048 / 32;
      unsigned int local = 0;
      for (int b = 0; b < per; b++) { int bin = lane * per + b; if (bin < (int)nb) local += hist[bin]; }
      unsigned int suffix = local;  // inclusive suffix over lanes (lane 31 is the top)
      for (int o = 1; o < 32; o <<= 1) { unsigned int v = __shfl_down_sync(0xffffffffu, suffix, o); if (lane + o < 32) suffix += v; }
      unsigned int above = suffix - local;  // keys in lanes above mine
      bool mine = (above < need) && (suffix >= need);
      if (mine) {
        unsigned int acc = above;
        int d = -1;
        for (int b = per - 1; b >= 0; b--) {
          int bin = lane * per + b;
          unsigned int h = bin < (int)nb ? hist[bin] : 0u;
          if (acc + h >= need) { d = bin; break; }
          acc += h;
        }
        s_prefix = prefix | ((unsigned)d << sh);
        s_need = need - acc;
      }
    }
    __syncthreads();
    prefix = s_prefix; need = s_need; mask |= (nb - 1) << sh;
    __syncthreads();
  }
  const unsigned int T = prefix;  // key of the k-th largest; take all > T and the first `need` == T in index order

  // ---- ordered compaction
  for (int i = tid; i < kMaxTopk; i += blockDim.x) sel[i] = 0ull;
  __syncthreads();
  unsigned int base_gt = 0, base_eq = 0;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    int i = i0 + tid;
    unsigned int key = 0; bool gt = false, eq = false;
    if (i < n) {
      key = float_to_ordered(__ldg(logits + rpn_addr(i, A, W, lv.row_stride_l, lv.pix_stride_l, 1)));
      gt = key > T; eq = key == T;
    }
    unsigned int bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { warp_gt[wid] = __popc(bg); warp_eq[wid] = __popc(be); }
    __syncthreads();
    unsigned int pg = 0, pe = 0, tg = 0, te = 0;
    for (int w2 = 0; w2 < 32; w2++) {
      unsigned int g = warp_gt[w2], e = warp_eq[w2];
      if (w2 < wid) { pg += g; pe += e; }
      tg += g; te += e;
    }
    pg += __popc(bg & ((1u << lane) - 1u)); pe += __popc(be & ((1u << lane) - 1u));
    unsigned int eq_before = base_eq + pe;
    if (gt || (eq && eq_before < need)) {
      unsigned int slot = base_gt + pg + min(eq_before, need);
      sel[slot] = ((unsigned long long)key << 32) | (unsigned int)(~(unsigned int)i);
    }
    base_gt += tg; base_eq += te;
    __syncthreads();
  }
  bitonic_sort_desc(sel, kMaxTopk);
  }

  // ---- decode the selected anchors
  const int slot0 = (img * n_levels + lvl) * kMaxTopk;
  const int ih = image_sizes[img * 2], iw = image_sizes[img * 2 + 1];
  const float* deltas = lv.deltas + (int64_t)img * lv.img_stride_d;
  const float clampv = 4.135166556742356f;  // log(1000/16)
  if (tid < kMaxTopk) {
    bool ok = false;
    float4 bx = make_float4(0, 0, 0, 0);
    float score = 0.f;
    if (tid < k) {
      unsigned long long e = sel[tid];
      int i = (int)(~(unsigned int)(e & 0xffffffffull));
      score = ordered_to_float((unsigned int)(e >> 32));
      int pix = i / A, a = i - pix * A, y = pix / W, x = pix - y * W;
      float sx = (float)(x * lv.stride), sy = (float)(y * lv.stride);
      float ax1 = __fadd_rn(sx, lv.cell_anchors[a * 4]), ay1 = __fadd_rn(sy, lv.cell_anchors[a * 4 + 1]);
      float ax2 = __fadd_rn(sx, lv.cell_anchors[a * 4 + 2]), ay2 = __fadd_rn(sy, lv.cell_anchors[a * 4 + 3]);
      const float* d = deltas + rpn_addr(i, A, W, lv.row_stride_d, lv.pix_stride_d, 4);
      float widths = __fsub_rn(ax2, ax1), heights = __fsub_rn(ay2, ay1);
      float cx = __fadd_rn(ax1, __fmul_rn(0.5f, widths)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, heights));
      float dx = __fdiv_rn(__ldg(d), wx), dy = __fdiv_rn(__ldg(d + 1), wy);
      float dw = fminf(__fdiv_rn(__ldg(d + 2), ww), clampv), dh = fminf(__fdiv_rn(__ldg(d + 3), wh), clampv);
      // torch.clamp(max=) propagates NaN; fminf would drop it
      if (__ldg(d + 2) != __ldg(d + 2)) dw = __ldg(d + 2);
      if (__ldg(d + 3) != __ldg(d + 3)) dh = __ldg(d + 3);
      float pcx = __fadd_rn(__fmul_rn(dx, widths), cx), pcy = __fadd_rn(__fmul_rn(dy, heights), cy);
      float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
      float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
      ok = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(score);
      x1 = fminf(fmaxf(x1, 0.f), (float)iw); y1 = fminf(fmaxf(y1, 0.f), (float)ih);
      x2 = fminf(fmaxf(x2, 0.f), (float)iw); y2 = fminf(fmaxf(y2, 0.f), (float)ih);
      ok = ok && (__fsub_rn(x2, x1) > min_box_size) && (__fsub_rn(y2, y1) > min_box_size);
      bx = make_float4(x1, y1, x2, y2);
    }
    ws_boxes[slot0 + tid] = bx;
    ws_scores[slot0 + tid] = score;
    ws_valid[slot0 + tid] = ok ? 1 : 0;
    // per-image max coordinate over the boxes that reach batched_nms (trick offset), and their count
    uint32_t m = ok ? float_to_ordered(fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w))) : 0u;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    unsigned int cnt = __popc(__ballot_sync(0xffffffffu, ok));
    if (lane == 0 && cnt) { atomicMax(&hdr[img * 2], m); atomicAdd(&hdr[img * 2 + 1], cnt); }
  }
}

// ---- per-(image, level) NMS as suppression bit-matrix + in-smem sweep (n <= 1024):
//   rpn_nms_mask_kernel : grid = slots x 136 upper-triangular 64x64 tile pairs, fully parallel IoU tests (same arithmetic as
//                         nms_core.cuh / torchvision), mask[slot][i][w] bit j set <=> box j (> i, lower score) overlaps box i
//   rpn_nms_sweep_kernel: CTA per slot: mask -> shared memory (128 KB), one warp resolves 64 candidates per step
//                         (diagonal word: serial, branch-free; remaining words: lanes OR the rows of the kept boxes),
//                         then ordered compaction of the kept boxes.  Greedy result == sequential NMS.
constexpr int kMaskWords = kMaxTopk / 64;       // 16
constexpr int kTilePairs = kMaskWords * (kMaskWords + 1) / 2;   // 136

__global__ void __launch_bounds__(64)
rpn_nms_mask_kernel(int n_levels, int topk, float thr, int nms_mode, const float4* __restrict__ ws_boxes,
                    const unsigned char* __restrict__ ws_valid, const uint32_t* __restrict__ hdr, unsigned long long* __restrict__ mask) {
  __shared__ float cx1[64], cy1[64], cx2[64], cy2[64], car[64];
  __shared__ int cvalid[64];
  const int slot = blockIdx.x / kTilePairs;
  int pair = blockIdx.x - slot * kTilePairs, tr = 0;
  while (pair >= kMaskWords - tr) { pair -= kMaskWords - tr; tr++; }
  const int tc = tr + pair;
  const int n = topk < kMaxTopk ? topk : kMaxTopk;
  if (tr * 64 >= n || tc * 64 >= n) return;
  const int img = slot / n_levels, lvl = slot - img * n_levels;
  int mode = nms_mode;
  if (mode < 0) mode = reference_cuda_nms_mode((long long)hdr[img * 2 + 1]);
  const float off = (mode == 0) ? __fmul_rn((float)lvl, __fadd_rn(ordered_to_float(hdr[img * 2]), 1.0f)) : 0.f;
  const int t = threadIdx.x;
  const int slot0 = slot * kMaxTopk;
  {
    const int j = tc * 64 + t;
    float4 b = make_float4(0, 0, 0, 0); bool v = false;
    if (j < n) { b = ws_boxes[slot0 + j]; v = ws_valid[slot0 + j] != 0; }
    cx1[t] = __fadd_rn(b.x, off); cy1[t] = __fadd_rn(b.y, off); cx2[t] = __fadd_rn(b.z, off); cy2[t] = __fadd_rn(b.w, off);
    car[t] = box_area(cx1[t], cy1[t], cx2[t], cy2[t]);
    cvalid[t] = v ? 1 : 0;
  }
  __syncthreads();
  const int i = tr * 64 + t;
  unsigned long long bits = 0ull;
  if (i < n && ws_valid[slot0 + i]) {
    float4 b = ws_boxes[slot0 + i];
    const float x1 = __fadd_rn(b.x, off), y1 = __fadd_rn(b.y, off), x2 = __fadd_rn(b.z, off), y2 = __fadd_rn(b.w, off);
    const float ar = box_area(x1, y1, x2, y2);
    const bool skip_zero = thr >= 0.f;
    for (int jj = 0; jj < 64; jj++) {
      const int j = tc * 64 + jj;
      if (j <= i || !cvalid[jj]) continue;
      if (skip_zero && (fminf(cx2[jj], x2) <= fmaxf(cx1[jj], x1) || fminf(cy2[jj], y2) <= fmaxf(cy1[jj], y1))) continue;
      if (iou_gt(x1, y1, x2, y2, ar, cx1[jj], cy1[jj], cx2[jj], cy2[jj], car[jj], thr)) bits |= (1ull << jj);
    }
  }
  if (i < n) mask[((size_t)slot0 + i) * kMaskWords + tc] = bits;
}

__global__ void __launch_bounds__(256)
rpn_nms_sweep_kernel(int topk, const unsigned long long* __restrict__ mask, const float4* __restrict__ ws_boxes,
                     const float* __restrict__ ws_scores, const unsigned char* __restrict__ ws_valid,
                     float4* __restrict__ k_boxes, float* __restrict__ k_scores, int* __restrict__ k_count) {
  extern __shared__ unsigned long long smask[];    // [n][16]
  __shared__ unsigned long long validw[kMaskWords], keptw[kMaskWords];
  __shared__ int warp_cnt[8];
  const int slot0 = blockIdx.x * kMaxTopk;
  const int n = topk < kMaxTopk ? topk : kMaxTopk;
  const int nchunks = (n + 63) / 64;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  {
    const uint4* src = reinterpret_cast<const uint4*>(mask + (size_t)slot0 * kMaskWords);
    uint4* dst = reinterpret_cast<uint4*>(smask);
    const int nv = n * kMaskWords / 2;
    for (int i0 = 0; i0 < nv; i0 += 8 * 256) {           // eight 16-byte loads in flight per thread (the copy is latency bound otherwise)
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int i = i0 + u * 256 + tid; if (i < nv) v[u] = src[i]; }
#pragma unroll
      for (int u = 0; u < 8; u++) { const int i = i0 + u * 256 + tid; if (i < nv) dst[i] = v[u]; }
    }
  }
  {   // validity bits by warp ballots (a 64-bit shared-memory atomicOr per box serialised 64 ways per word)
    unsigned int* validw32 = reinterpret_cast<unsigned int*>(validw);
    for (int j0 = 0; j0 < kMaxTopk; j0 += 256) {
      const int j = j0 + tid;
      const unsigned int b = __ballot_sync(0xffffffffu, j < n && ws_valid[slot0 + j] != 0);
      if (lane == 0) validw32[j >> 5] = b;
    }
  }
  __syncthreads();
  if (wid == 0) {
    unsigned long long removed = 0ull;   // lane w (< 16) owns suppression word w
    for (int c = 0; c < nchunks; c++) {
      unsigned long long rc = __shfl_sync(0xffffffffu, removed, c);
      unsigned long long kept = 0ull;
      if (lane == 0) {
        unsigned long long alive = validw[c] & ~rc;
        const unsigned long long* diag = smask + (size_t)(c * 64) * kMaskWords + c;
        const int rows = (n - c * 64) < 64 ? (n - c * 64) : 64;
#pragma unroll 8
        for (int i = 0; i < rows; i++) {
          unsigned long long bit = (alive >> i) & 1ull;
          kept |= bit << i;
          alive &= ~(diag[(size_t)i * kMaskWords] & (0ull - bit));
        }
        keptw[c] = kept;
      }
      kept = __shfl_sync(0xffffffffu, kept, 0);
      if (lane > c && lane < nchunks) {
        // rows of the kept boxes OR-ed into this lane's word: the addresses depend on `kept` only, so four loads are issued per step
        const unsigned long long* colw = smask + (size_t)(c * 64) * kMaskWords + lane;
        unsigned long long k = kept;
        while (k) {
          const int i0 = __ffsll((long long)k) - 1; k &= k - 1;
          const int i1 = k ? __ffsll((long long)k) - 1 : i0; k &= k - 1;      // (k & (k - 1) of 0 is 0: exhausted lanes repeat i0)
          const int i2 = k ? __ffsll((long long)k) - 1 : i0; k &= k - 1;
          const int i3 = k ? __ffsll((long long)k) - 1 : i0; k &= k - 1;
          const unsigned long long r0 = colw[(size_t)i0 * kMaskWords], r1 = colw[(size_t)i1 * kMaskWords],
                                   r2 = colw[(size_t)i2 * kMaskWords], r3 = colw[(size_t)i3 * kMaskWords];
          removed |= (r0 | r1) | (r2 | r3);
        }
      }
    }
  }
  __syncthreads();
  // ordered compaction of the kept boxes
  int base = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    int j = j0 + tid;
    bool kf = j < n && ((keptw[j >> 6] >> (j & 63)) & 1ull);
    unsigned int b = __ballot_sync(0xffffffffu, kf);
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w2 = 0; w2 < 8; w2++) { if (w2 < wid) pre += warp_cnt[w2]; tot += warp_cnt[w2]; }
    if (kf) {
      int pos = base + pre + __popc(b & ((1u << lane) - 1u));
      k_boxes[slot0 + pos] = ws_boxes[slot0 + j];
      k_scores[slot0 + pos] = ws_scores[slot0 + j];
    }
    base += tot;
    __syncthreads();
  }
  if (tid == 0) k_count[blockIdx.x] = base;
}

__global__ void __launch_bounds__(256)
rpn_merge_kernel(int n_levels, int post_topk, const float4* __restrict__ k_boxes, const float* __restrict__ k_scores,
                 const int* __restrict__ k_count, float4* __restrict__ proposals, float* __restrict__ prop_logits,
                 int32_t* __restrict__ counts) {
  // grid (n_levels, n_images): CTA (l, img) ranks level l's kept boxes among all levels of the image.  The (descending) score lists of
  // the image are staged in shared memory once, so the four binary searches per box never leave the SM.
  extern __shared__ float s_sc[];                         // [n_levels][kMaxTopk]
  __shared__ int cnt[kMaxLevels];
  const int l = blockIdx.x, img = blockIdx.y;
  if (threadIdx.x < n_levels) cnt[threadIdx.x] = k_count[img * n_levels + threadIdx.x];
  __syncthreads();
  int total = 0;
  for (int l2 = 0; l2 < n_levels; l2++) {
    total += cnt[l2];
    const float* a = k_scores + (img * n_levels + l2) * kMaxTopk;
    for (int p = threadIdx.x; p < cnt[l2]; p += blockDim.x) s_sc[l2 * kMaxTopk + p] = a[p];
  }
  __syncthreads();
  const int out_n = total < post_topk ? total : post_topk;
  const int slot = (img * n_levels + l) * kMaxTopk;
  for (int p = threadIdx.x; p < cnt[l]; p += blockDim.x) {
    const float s = s_sc[l * kMaxTopk + p];
    int rank = p;
    for (int l2 = 0; l2 < n_levels; l2++) {
      if (l2 == l) continue;
      const float* a = s_sc + l2 * kMaxTopk;
      int g = count_greater_desc(a, cnt[l2], s);
      if (l2 < l) { while (g < cnt[l2] && a[g] == s) g++; }  // ties: lower level (lower concatenated index) first
      rank += g;
    }
    if (rank < post_topk) {
      proposals[(int64_t)img * post_topk + rank] = k_boxes[slot + p];
      prop_logits[(int64_t)img * post_topk + rank] = s;
    }
  }
  if (l == 0) {
    for (int r = out_n + threadIdx.x; r < post_topk; r += blockDim.x) {
      proposals[(int64_t)img * post_topk + r] = make_float4(0, 0, 0, 0);
      prop_logits[(int64_t)img * post_topk + r] = 0.f;
    }
    if (threadIdx.x == 0) counts[img] = out_n;
  }
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" size_t lvcb200_rpn_proposals_workspace(const lvcb200_rpn_params* p) {
  if (!p || p->n_images <= 0 || p->n_levels <= 0) return 256;
  return rpn_layout(p->n_images, p->n_levels).total;
}

extern "C" int lvcb200_rpn_proposals(const lvcb200_rpn_level* levels, const lvcb200_rpn_params* p,
                                     const int32_t* image_sizes, float* proposals, float* prop_logits, int32_t* counts,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  LVC_REQUIRE(levels && p, "rpn_proposals: NULL descriptor");
  LVC_REQUIRE(p->n_levels >= 1 && p->n_levels <= kMaxLevels, "rpn_proposals: 1..8 levels");
  LVC_REQUIRE(p->pre_nms_topk >= 1 && p->pre_nms_topk <= kMaxTopk, "rpn_proposals: pre_nms_topk must be in [1,1024]");
  LVC_REQUIRE(p->post_nms_topk >= 1, "rpn_proposals: post_nms_topk must be positive");
  if (p->n_images == 0) return 0;
  LVC_REQUIRE(image_sizes && proposals && prop_logits && counts && workspace, "rpn_proposals: NULL pointer");
  LVC_REQUIRE(((uintptr_t)proposals % 16) == 0, "rpn_proposals: proposals must be 16-byte aligned");
  RpnWs w = rpn_layout(p->n_images, p->n_levels);
  if (workspace_bytes < w.total) return set_error(LVCB200_EWORKSPACE, "rpn_proposals: workspace too small");
  RpnLevels L;
  L.chunk_prefix[0] = 0;
  for (int i = 0; i < p->n_levels; i++) {
    L.lv[i] = levels[i];
    L.chunk_prefix[i + 1] = L.chunk_prefix[i] + (levels[i].H * levels[i].W * levels[i].A + kScanChunk - 1) / kScanChunk;
    LVC_REQUIRE(levels[i].A >= 1 && levels[i].A <= 3, "rpn_proposals: A must be 1..3");
    LVC_REQUIRE(levels[i].logits && levels[i].deltas, "rpn_proposals: NULL level pointer");
    LVC_REQUIRE((int64_t)levels[i].H * levels[i].W * levels[i].A < (1ll << 31), "rpn_proposals: level too large");
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  uint32_t* hdr = (uint32_t*)(ws + w.off_hdr);
  LVC_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, s));
  const int grid = p->n_images * p->n_levels;
  const int scan_grid = p->n_images * L.chunk_prefix[p->n_levels];
  unsigned int* hist1 = (unsigned int*)(ws + w.off_hist1);
  unsigned int* hist2 = (unsigned int*)(ws + w.off_hist2);
  unsigned int* ccount = (unsigned int*)(ws + w.off_ccount);
  unsigned long long* cand = (unsigned long long*)(ws + w.off_cand);
  int rc;
  rpn_hist1_kernel<<<scan_grid, 256, 0, s>>>(L, p->n_levels, hist1);
  if ((rc = check_launch("rpn_hist1_kernel"))) return rc;
  rpn_hist2_kernel<<<scan_grid, 256, 0, s>>>(L, p->n_levels, p->pre_nms_topk, hist1, hist2);
  if ((rc = check_launch("rpn_hist2_kernel"))) return rc;
  rpn_collect_kernel<<<scan_grid, 256, 0, s>>>(L, p->n_levels, p->pre_nms_topk, hist1, hist2, ccount, cand);
  if ((rc = check_launch("rpn_collect_kernel"))) return rc;
  rpn_select_decode_kernel<<<grid, 1024, 0, s>>>(L, p->n_levels, p->pre_nms_topk, p->min_box_size, p->weights[0], p->weights[1],
                                                 p->weights[2], p->weights[3], image_sizes, (float4*)(ws + w.off_boxes),
                                                 (float*)(ws + w.off_scores), (unsigned char*)(ws + w.off_valid), hdr, ccount, cand);
  rc = check_launch("rpn_select_decode_kernel");
  if (rc) return rc;
  unsigned long long* mask = (unsigned long long*)(ws + w.off_mask);
  rpn_nms_mask_kernel<<<grid * kTilePairs, 64, 0, s>>>(p->n_levels, p->pre_nms_topk, p->nms_thresh, p->nms_mode,
                                                      (const float4*)(ws + w.off_boxes), (const unsigned char*)(ws + w.off_valid), hdr, mask);
  if ((rc = check_launch("rpn_nms_mask_kernel"))) return rc;
  {
    static bool attr_set = false;
    const int smem = kMaxTopk * kMaskWords * 8;
    if (!attr_set) {
      LVC_CUDA(cudaFuncSetAttribute(rpn_nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    rpn_nms_sweep_kernel<<<grid, 256, smem, s>>>(p->pre_nms_topk, mask, (const float4*)(ws + w.off_boxes), (const float*)(ws + w.off_scores),
                                                 (const unsigned char*)(ws + w.off_valid), (float4*)(ws + w.off_kboxes),
                                                 (float*)(ws + w.off_kscores), (int*)(ws + w.off_kcount));
  }
  rc = check_launch("rpn_nms_sweep_kernel");
  if (rc) return rc;
  rpn_merge_kernel<<<dim3(p->n_levels, p->n_images), 256, p->n_levels * kMaxTopk * sizeof(float), s>>>(p->n_levels, p->post_nms_topk, (const float4*)(ws + w.off_kboxes),
                                               (const float*)(ws + w.off_kscores), (const int*)(ws + w.off_kcount),
                                               (float4*)proposals, prop_logits, counts);
  return check_launch("rpn_merge_kernel");
}
