// kNN label verification (tools/run_nearest_neighbours.py:142-162, 214-227).
//   knn_prepare: bank mean (crop_mean, :144) and the centred, L2-normalised bank rows (F.cosine_similarity's
//                clamp of the norm at eps = 1e-8), done once after the support bank is gathered.
//   knn_verify : CTA per 32 queries; centred query tile and bank tile staged in shared memory, 4x4 register
//                tiles, fp32 FMA; running top-k per query in shared memory; votes / mode / keep fused at the end.
// Two paths:
//   exact SIMT (knn_verify_kernel)  : fp32 FMA, any S <= 4096; also the fallback of the tensor-core path.
//   tensor core (lvcb200_knn_verify_tc): (1) approximate scores (q . bhat_s - mu . bhat_s) for all (query, bank row) pairs with the
//       tcgen05 shift-GEMM in kind::tf32 straight from the fp32 queries (gemm_tc.cu), fp16 out; (2) knn_rerank_kernel, warp per
//       query: rigorous candidate set {s : approx_s >= (k-th largest approx) - 2 eps}, eps bounding TF32 + fp16 rounding by
//       Cauchy-Schwarz, then EXACT fp32 centred-cosine re-scoring of the candidates, top-k, votes, mode, keep.  The result is
//       the exact fp32 top-k (not a TF32 approximation); queries whose candidate set overflows are redone by the SIMT kernel.
#include <cuda_fp16.h>
#include <string.h>

#include "common.cuh"

namespace lvcb200 {

constexpr int QT = 32, BT = 128, KT = 32, KNN_MAXK = 16;

// column mean of the bank (crop_mean): 32 columns per CTA, 8 warps stride the rows, double accumulation
__global__ void __launch_bounds__(256)
knn_mean_kernel(const float* __restrict__ bank, int S, int D, float* __restrict__ mean) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  double a = 0.0;
  if (d < D)
    for (int s = w; s < S; s += 8) a += (double)bank[(size_t)s * D + d];
  part[w][lane] = a;
  __syncthreads();
  if (w == 0 && d < D) {
    double t = 0.0;
    for (int i = 0; i < 8; i++) t += part[i][lane];
    mean[d] = (float)(t / (double)S);
  }
}

__global__ void __launch_bounds__(256)
knn_normalize_kernel(const float* __restrict__ bank, int S, int D, const float* __restrict__ mean, float* __restrict__ bhat,
                     float* __restrict__ negc) {
  const int s = blockIdx.x;
  if (s >= S) {   // zero padding rows (the score GEMM reads S rounded up to a multiple of 8)
    for (int d = threadIdx.x; d < D; d += blockDim.x) bhat[(size_t)s * D + d] = 0.f;
    if (threadIdx.x == 0) negc[s] = 0.f;
    return;
  }
  __shared__ double red[8];
  __shared__ double red2[8];
  __shared__ float s_inv;
  double a = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = __fsub_rn(bank[(size_t)s * D + d], mean[d]);
    a += (double)v * (double)v;
  }
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int i = 0; i < 8; i++) t += red[i];
    float n = (float)sqrt(t);
    s_inv = n > 1e-8f ? n : 1e-8f;
  }
  __syncthreads();
  const float nrm = s_inv;
  double c = 0.0;   // c_s = mu . bhat_s  (so that q . bhat_s - c_s == (q - mu) . bhat_s)
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = __fdiv_rn(__fsub_rn(bank[(size_t)s * D + d], mean[d]), nrm);
    bhat[(size_t)s * D + d] = v;
    c += (double)v * (double)mean[d];
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red2[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < 8; i++) t += red2[i]; negc[s] = (float)(-t); }
}

// Euclidean variant of the bank preparation (QUERY_EXPAND.COSINE_SIM = False, run_nearest_neighbours.py:154-159: ranking by
// -cdist(bank, query)): no centring, rows kept as they are, bnorm2[s] = |b_s|^2; the verify kernel then scores
// -sqrt(max(|q|^2 + |b|^2 - 2 q.b, 0)), the matmul form torch.cdist itself uses for more than 25 rows.
__global__ void __launch_bounds__(256)
knn_prepare_euclid_kernel(const float* __restrict__ bank, int S, int D, float* __restrict__ mean, float* __restrict__ bhat,
                          float* __restrict__ bnorm2) {
  const int s = blockIdx.x;
  if (s == 0) for (int d = threadIdx.x; d < D; d += blockDim.x) mean[d] = 0.f;
  __shared__ double red[8];
  double a = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = s < S ? bank[(size_t)s * D + d] : 0.f;
    bhat[(size_t)s * D + d] = v;
    a += (double)v * (double)v;
  }
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < 8; i++) t += red[i]; bnorm2[s] = (float)t; }
}

__global__ void __launch_bounds__(256)
knn_verify_kernel(const float* __restrict__ mean, const float* __restrict__ bhat, const int64_t* __restrict__ bank_cls, int S, int D,
                  const float* __restrict__ queries, const int64_t* __restrict__ query_cls, int64_t Q, int topk, int knn,
                  int64_t* __restrict__ top_idx, float* __restrict__ top_sim, int64_t* __restrict__ votes, uint8_t* __restrict__ keep,
                  const uint8_t* __restrict__ only_flagged, const float* __restrict__ bnorm2 = nullptr) {
  __shared__ __align__(16) float As[KT][QT];
  __shared__ __align__(16) float Bs[KT][BT];
  __shared__ float Ss[QT][BT + 1];
  __shared__ float qn[QT];
  __shared__ float qs2[QT];
  __shared__ float tk_sim[QT][KNN_MAXK];
  __shared__ int tk_idx[QT][KNN_MAXK];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * QT;
  if (only_flagged != nullptr) {   // fallback pass of the tensor-core path: skip tiles without an overflowed query
    int any = 0;
    if (tid < QT && q0 + tid < Q) any = only_flagged[q0 + tid];
    if (!__syncthreads_or(any)) return;
  }
  if (tid < QT) for (int k = 0; k < KNN_MAXK; k++) { tk_sim[tid][k] = -INFINITY; tk_idx[tid][k] = -1; }
  const int lq = tid >> 3, lk = (tid & 7) * 4;  // loader mapping for the query tile
  float qss = 0.f;                              // partial sum of squares of the centred query (first bank tile only)
  for (int b0 = 0; b0 < S; b0 += BT) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += KT) {
      {  // query tile, centred
        float4 v = make_float4(0, 0, 0, 0);
        if (q0 + lq < Q && k0 + lk < D) {
          v = __ldg(reinterpret_cast<const float4*>(queries + (size_t)(q0 + lq) * D + k0 + lk));
          float4 m = __ldg(reinterpret_cast<const float4*>(mean + k0 + lk));
          v.x = __fsub_rn(v.x, m.x); v.y = __fsub_rn(v.y, m.y); v.z = __fsub_rn(v.z, m.z); v.w = __fsub_rn(v.w, m.w);
        }
        As[lk][lq] = v.x; As[lk + 1][lq] = v.y; As[lk + 2][lq] = v.z; As[lk + 3][lq] = v.w;
        if (b0 == 0) qss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {  // bank tile
        int r = (tid >> 3) + 32 * j;
        float4 v = make_float4(0, 0, 0, 0);
        if (b0 + r < S && k0 + lk < D) v = __ldg(reinterpret_cast<const float4*>(bhat + (size_t)(b0 + r) * D + k0 + lk));
        Bs[lk][r] = v.x; Bs[lk + 1][r] = v.y; Bs[lk + 2][r] = v.z; Bs[lk + 3][r] = v.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KT; k++) {
        float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
        acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
        acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
        acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
      }
      __syncthreads();
    }
    if (b0 == 0) {  // finish the query norms: 8 consecutive lanes share a query
      float s = qss;
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((tid & 7) == 0) { float n = sqrtf(s); qn[lq] = n > 1e-8f ? n : 1e-8f; qs2[lq] = s; }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int r = b0 + tx * 4 + j;
        float sc = -INFINITY;
        if (r < S) sc = bnorm2 == nullptr ? __fdiv_rn(acc[i][j], qn[ty * 4 + i])
                                          : -sqrtf(fmaxf(__fadd_rn(__fadd_rn(qs2[ty * 4 + i], bnorm2[r]), -2.f * acc[i][j]), 0.f));
        Ss[ty * 4 + i][tx * 4 + j] = sc;
      }
    __syncthreads();
    if (tid < QT) {  // running top-k, ascending bank index so that ties keep the lower index
      float worst = tk_sim[tid][topk - 1];
      for (int j = 0; j < BT; j++) {
        float s = Ss[tid][j];
        if (s > worst) {
          int p = topk - 1;
          while (p > 0 && tk_sim[tid][p - 1] < s) { tk_sim[tid][p] = tk_sim[tid][p - 1]; tk_idx[tid][p] = tk_idx[tid][p - 1]; p--; }
          tk_sim[tid][p] = s; tk_idx[tid][p] = b0 + j;
          worst = tk_sim[tid][topk - 1];
        }
      }
    }
    __syncthreads();
  }
  if (tid < QT && q0 + tid < Q && (only_flagged == nullptr || only_flagged[q0 + tid])) {
    const int64_t q = q0 + tid;
    int64_t v[KNN_MAXK];
    for (int k = 0; k < topk; k++) {
      int idx = tk_idx[tid][k];
      v[k] = idx >= 0 ? bank_cls[idx] : -1;
      top_idx[q * topk + k] = idx;
      votes[q * topk + k] = v[k];
      if (top_sim) top_sim[q * topk + k] = tk_sim[tid][k];
    }
    int64_t best_v = 0; int best_c = 0;   // torch.mode: most frequent, smallest value on ties
    const int kk = knn < topk ? knn : topk;
    for (int a = 0; a < kk; a++) {
      int c = 0;
      for (int b = 0; b < kk; b++) c += (v[b] == v[a]);
      if (c > best_c || (c == best_c && v[a] < best_v)) { best_c = c; best_v = v[a]; }
    }
    keep[q] = (query_cls[q] == best_v) ? 1 : 0;
  }
}


constexpr int RR_WARPS = 4;
constexpr int RR_CAND = 96;

// warp per query; dynamic smem per warp: S_pad halfs (scores) + D floats (centred query) + candidate arrays
__global__ void __launch_bounds__(RR_WARPS * 32)
knn_rerank_kernel(const float* __restrict__ mean, const float* __restrict__ bhat, const int64_t* __restrict__ bank_cls, int S, int S_pad,
                  int D, const float* __restrict__ queries, const int64_t* __restrict__ query_cls, int64_t Q,
                  const __half* __restrict__ scores, int topk, int knn, int64_t* __restrict__ top_idx, float* __restrict__ top_sim,
                  int64_t* __restrict__ votes, uint8_t* __restrict__ keep, uint8_t* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t per_warp = (size_t)S_pad * 2 + (size_t)D * 4 + RR_CAND * 8;
  unsigned char* base = rr_smem + w * ((per_warp + 15) / 16 * 16);
  __half* ssc = reinterpret_cast<__half*>(base);
  float* qc = reinterpret_cast<float*>(base + (((size_t)S_pad * 2 + 15) / 16 * 16));
  int* cidx = reinterpret_cast<int*>(qc + D);
  float* csim = reinterpret_cast<float*>(cidx + RR_CAND);
  for (int64_t q = (int64_t)blockIdx.x * RR_WARPS + w; q < Q; q += (int64_t)gridDim.x * RR_WARPS) {
    __syncwarp();
    // (a) approximate scores of this query -> smem
    const uint4* srow = reinterpret_cast<const uint4*>(scores + q * S_pad);
    for (int i = lane; i < S_pad / 8; i += 32) reinterpret_cast<uint4*>(ssc)[i] = __ldg(srow + i);
    // (b) centred query -> smem, norms
    float nq = 0.f, nqc = 0.f;
    for (int k = lane * 4; k < D; k += 128) {
      float4 v = __ldg(reinterpret_cast<const float4*>(queries + q * D + k));
      float4 m = __ldg(reinterpret_cast<const float4*>(mean + k));
      nq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      v.x = __fsub_rn(v.x, m.x); v.y = __fsub_rn(v.y, m.y); v.z = __fsub_rn(v.z, m.z); v.w = __fsub_rn(v.w, m.w);
      nqc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      *reinterpret_cast<float4*>(qc + k) = v;
    }
    for (int o = 16; o; o >>= 1) { nq += __shfl_xor_sync(0xffffffffu, nq, o); nqc += __shfl_xor_sync(0xffffffffu, nqc, o); }
    const float norm_q = sqrtf(nq), norm_qc = sqrtf(nqc);
    // rigorous error bound of an approximate score: TF32 operand rounding (2^-10 each) by Cauchy-Schwarz with |bhat| <= 1,
    // fp32 accumulation, fp16 output rounding (2^-11 |score|, |score| <= |q - mu|); 30 % slack
    const float eps = 0.0026f * norm_q + 0.0007f * norm_qc + 1e-6f;
    __syncwarp();
    // (c) a lower bound of the k-th largest approximate score: the k-th largest of the 32 per-lane maxima (the k-th largest
    //     of any subset is <= the k-th largest of the whole row), found by rank counting over the warp.  Any lower bound keeps
    //     the candidate rule rigorous; this one typically sits 2-3 ranks below the exact k-th value.
    float cur_v;
    {
      float lm = -INFINITY;
      for (int i = lane; i < S; i += 32) lm = fmaxf(lm, __half2float(ssc[i]));
      int rank = 0;
      for (int o = 0; o < 32; o++) {
        float ov = __shfl_sync(0xffffffffu, lm, o);
        rank += (ov > lm) || (ov == lm && o < lane);
      }
      const int want = (topk - 1) < 31 ? (topk - 1) : 31;
      unsigned int m = __ballot_sync(0xffffffffu, rank == want);
      cur_v = __shfl_sync(0xffffffffu, lm, __ffs(m) - 1);
    }
    // (d) two-phase candidate set.  Phase A: every row whose approximate score reaches the lower bound T0 (>= k rows by
    //     construction: the k largest lane maxima) is re-scored EXACTLY; the k-th largest exact score among them, TL, is a lower
    //     bound of the true k-th largest score.  Phase B: a row outside A can only belong to the true top-k if its exact score
    //     reaches TL, i.e. if approx_s + eps >= TL -- one eps against an exact threshold, instead of 2 eps against an approximate
    //     one: ~13 exact rows per query instead of ~21 on the bench workload, and the bank-row gathers are what this kernel costs.
    const float inv_n = 1.0f / (norm_qc > 1e-8f ? norm_qc : 1e-8f);
    int nc = 0;
    for (int i0 = 0; i0 < S; i0 += 32) {
      int i = i0 + lane;
      bool c = i < S && __half2float(ssc[i]) >= cur_v;
      unsigned int m = __ballot_sync(0xffffffffu, c);
      if (c) { int pos = nc + __popc(m & ((1u << lane) - 1u)); if (pos < RR_CAND) cidx[pos] = i; }
      nc += __popc(m);
    }
    if (nc > RR_CAND) {   // pathological ties / huge norms: hand the query to the exact SIMT kernel
      if (lane == 0) overflow[q] = 1;
      continue;
    }
    __syncwarp();
    // exact fp32 centred dot products of candidates [j_begin, j_end) -> csim (un-normalised; scaled by inv_n when ranked)
    auto score_range = [&](int j_begin, int j_end) {
      for (int j0 = j_begin; j0 < j_end; j0 += 4) {   // four candidates per pass: 4x the loads in flight per lane
        const float* br[4];
#pragma unroll
        for (int u = 0; u < 4; u++) br[u] = bhat + (size_t)cidx[(j0 + u < j_end) ? j0 + u : j0] * D;
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane * 4; k < D; k += 128) {
          const float4 a = *reinterpret_cast<const float4*>(qc + k);
          float4 b[4];
#pragma unroll
          for (int u = 0; u < 4; u++) b[u] = __ldg(reinterpret_cast<const float4*>(br[u] + k));
#pragma unroll
          for (int u = 0; u < 4; u++) dot[u] += a.x * b[u].x + a.y * b[u].y + a.z * b[u].z + a.w * b[u].w;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          for (int o = 16; o; o >>= 1) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
        if (lane < 4 && j0 + lane < j_end) csim[j0 + lane] = (lane == 0 ? dot[0] : lane == 1 ? dot[1] : lane == 2 ? dot[2] : dot[3]);
      }
      __syncwarp();
    };
    score_range(0, nc);
    const int nA = nc;
    if (nA >= topk) {   // always, unless fewer than k lanes own a row
      float tl = -INFINITY;   // k-th largest exact score of phase A (rank counting; ties by position)
      for (int j = lane; j < nA; j += 32) {
        const float v = csim[j];
        int rank = 0;
        for (int o = 0; o < nA; o++) { const float ov = csim[o]; rank += (ov > v) || (ov == v && o < j); }
        if (rank == topk - 1) tl = v;
      }
      for (int o = 16; o; o >>= 1) tl = fmaxf(tl, __shfl_xor_sync(0xffffffffu, tl, o));
      for (int i0 = 0; i0 < S; i0 += 32) {
        int i = i0 + lane;
        float a = i < S ? __half2float(ssc[i]) : -INFINITY;
        bool c = i < S && a < cur_v && a + eps >= tl;
        unsigned int m = __ballot_sync(0xffffffffu, c);
        if (c) { int pos = nc + __popc(m & ((1u << lane) - 1u)); if (pos < RR_CAND) cidx[pos] = i; }
        nc += __popc(m);
      }
      if (nc > RR_CAND) {
        if (lane == 0) overflow[q] = 1;
        continue;
      }
      __syncwarp();
      if (nc > nA) score_range(nA, nc);
    }
    for (int j = lane; j < nc; j += 32) csim[j] *= inv_n;
    __syncwarp();
    // (f) exact top-k among the candidates (sim desc, index asc), votes, mode, keep
    int64_t myvote = -1;   // lane r keeps the vote of rank r
    if (nc <= 32) {        // one candidate per lane: rank by counting, rank r writes output slot r
      const float mv = lane < nc ? csim[lane] : -INFINITY;
      const int mi = lane < nc ? cidx[lane] : 0x7fffffff;
      int rank = 0;
      for (int o = 0; o < nc; o++) {
        float ov = __shfl_sync(0xffffffffu, mv, o); int oi = __shfl_sync(0xffffffffu, mi, o);
        rank += (ov > mv) || (ov == mv && oi < mi);
      }
      const bool out_lane = lane < nc && rank < topk;
      const int64_t vote = out_lane ? bank_cls[mi] : -1;
      if (out_lane) {
        top_idx[q * topk + rank] = mi;
        votes[q * topk + rank] = vote;
        if (top_sim) top_sim[q * topk + rank] = mv;
      }
      // move the vote of rank r to lane r
      for (int r2 = 0; r2 < topk; r2++) {
        unsigned int m = __ballot_sync(0xffffffffu, out_lane && rank == r2);
        int64_t v = __shfl_sync(0xffffffffu, vote, m ? __ffs(m) - 1 : 0);
        if (lane == r2) myvote = m ? v : -1;
      }
    } else {
      float pv = INFINITY; int pi = -1;
      for (int round = 0; round < topk; round++) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int j = lane; j < nc; j += 32) {
          float v = csim[j]; int i = cidx[j];
          bool after = (v < pv) || (v == pv && i > pi);
          if (after && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
        }
        for (int o = 16; o; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        pv = bv; pi = bi;
        const int64_t vote = (bi != 0x7fffffff) ? bank_cls[bi] : -1;
        if (lane == round) myvote = vote;
        if (lane == 0) {
          top_idx[q * topk + round] = (bi != 0x7fffffff) ? bi : -1;
          votes[q * topk + round] = vote;
          if (top_sim) top_sim[q * topk + round] = bv;
        }
      }
    }
    // torch.mode over the first knn votes: most frequent, smallest value on ties
    const int kk = knn < topk ? knn : topk;
    int cnt = 0;
    for (int r2 = 0; r2 < kk; r2++) { int64_t v = __shfl_sync(0xffffffffu, myvote, r2); cnt += (lane < kk && v == myvote); }
    int bc = (lane < kk) ? cnt : -1; int64_t bvv = myvote;
    for (int o = 16; o; o >>= 1) {
      int oc = __shfl_xor_sync(0xffffffffu, bc, o); int64_t ov = __shfl_xor_sync(0xffffffffu, bvv, o);
      if (oc > bc || (oc == bc && ov < bvv)) { bc = oc; bvv = ov; }
    }
    if (lane == 0) keep[q] = (query_cls[q] == bvv) ? 1 : 0;
  }
}

}  // namespace lvcb200

using namespace lvcb200;

static inline int knn_s_pad(int S) { return (S + 7) / 8 * 8; }
static inline int knn_s_al(int S) { return (S + 63) / 64 * 64; }
// knn_tc3.cu: tensor-core path v2 (bf16 pair operands, top-k in the epilogue)
int knn3_bank_rows(int S);
size_t knn3_workspace_bytes(int64_t Q, int D);
int knn3_split_bank(const float* bhat, int S, int D, const float* mean, void* bpair, float* mu2, cudaStream_t st);
int knn3_verify(const float* mean, const float* negc, const float* bhat, const void* bpair, const float* mu2, const int64_t* bank_cls, int S,
                int D, const float* queries, const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                int64_t* votes, uint8_t* keep, void* workspace, cudaStream_t st, int (*simt_fallback)(const uint8_t*, cudaStream_t, void*),
                void* fb_ctx);
// prepared layout: mean[D] | negc[S_al] | bhat[S_pad][D] | (256-byte aligned) mu2 | bank bf16 pair [2 * rows3][D]
static inline size_t knn_f32_part(int S, int D) { return align_up(sizeof(float) * ((size_t)D + knn_s_al(S) + (size_t)knn_s_pad(S) * D), 256); }
extern "C" size_t lvcb200_knn_prepared_bytes(int S, int D) {
  return knn_f32_part(S, D) + 256 + (size_t)2 * knn3_bank_rows(S) * D * 2;
}

extern "C" int lvcb200_knn_prepare(const float* bank, int S, int D, void* bank_prepared, void* stream) {
  LVC_REQUIRE(S >= 1 && D >= 4 && D % 4 == 0, "knn_prepare: need S >= 1 and D a positive multiple of 4");
  LVC_REQUIRE(bank && bank_prepared, "knn_prepare: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  float* mean = (float*)bank_prepared;
  float* negc = mean + D;
  float* bhat = negc + knn_s_al(S);
  knn_mean_kernel<<<(D + 31) / 32, 256, 0, s>>>(bank, S, D, mean);
  int rc = check_launch("knn_mean_kernel");
  if (rc) return rc;
  knn_normalize_kernel<<<knn_s_pad(S), 256, 0, s>>>(bank, S, D, mean, bhat, negc);
  if ((rc = check_launch("knn_normalize_kernel"))) return rc;
  if (D % 8 == 0) {   // operands of the tensor-core path: the normalised bank as a bf16 hi/lo pair, |mu|^2
    uint8_t* tail = (uint8_t*)bank_prepared + knn_f32_part(S, D);
    return knn3_split_bank(bhat, S, D, mean, tail + 256, (float*)tail, s);
  }
  return 0;
}

extern "C" int lvcb200_knn_prepare_euclid(const float* bank, int S, int D, void* bank_prepared, void* stream) {
  LVC_REQUIRE(S >= 1 && D >= 4 && D % 4 == 0, "knn_prepare_euclid: need S >= 1 and D a positive multiple of 4");
  LVC_REQUIRE(bank && bank_prepared, "knn_prepare_euclid: NULL pointer");
  float* mean = (float*)bank_prepared;
  float* bn2 = mean + D;
  float* bhat = bn2 + knn_s_al(S);
  knn_prepare_euclid_kernel<<<knn_s_pad(S), 256, 0, (cudaStream_t)stream>>>(bank, S, D, mean, bhat, bn2);
  return check_launch("knn_prepare_euclid_kernel");
}

extern "C" int lvcb200_knn_verify_euclid(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                                         const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                                         int64_t* votes, uint8_t* keep, void* stream) {
  LVC_REQUIRE(S >= 1 && S <= 4096 && D >= 4 && D % 4 == 0, "knn_verify_euclid: need 1 <= S <= 4096 and D a multiple of 4");
  LVC_REQUIRE(topk >= 1 && topk <= KNN_MAXK && topk <= S && knn >= 1, "knn_verify_euclid: need 1 <= topk <= min(16, S), knn >= 1");
  if (Q == 0) return 0;
  LVC_REQUIRE(bank_prepared && bank_cls && queries && query_cls && top_idx && votes && keep, "knn_verify_euclid: NULL pointer");
  LVC_REQUIRE(((uintptr_t)queries % 16) == 0, "knn_verify_euclid: queries must be 16-byte aligned");
  const float* mean = (const float*)bank_prepared;
  knn_verify_kernel<<<(unsigned)ceil_div64(Q, QT), 256, 0, (cudaStream_t)stream>>>(
      mean, mean + D + knn_s_al(S), bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx, top_sim, votes, keep, nullptr, mean + D);
  return check_launch("knn_verify_kernel<euclid>");
}

extern "C" int lvcb200_knn_verify(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                                  const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                                  int64_t* votes, uint8_t* keep, void* stream) {
  LVC_REQUIRE(S >= 1 && S <= 4096 && D >= 4 && D % 4 == 0, "knn_verify: need 1 <= S <= 4096 and D a multiple of 4");
  LVC_REQUIRE(topk >= 1 && topk <= KNN_MAXK && topk <= S && knn >= 1, "knn_verify: need 1 <= topk <= min(16, S), knn >= 1");
  if (Q == 0) return 0;
  LVC_REQUIRE(bank_prepared && bank_cls && queries && query_cls && top_idx && votes && keep, "knn_verify: NULL pointer");
  LVC_REQUIRE(((uintptr_t)queries % 16) == 0, "knn_verify: queries must be 16-byte aligned");
  const float* mean = (const float*)bank_prepared;
  knn_verify_kernel<<<(unsigned)ceil_div64(Q, QT), 256, 0, (cudaStream_t)stream>>>(
      mean, mean + D + knn_s_al(S), bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx, top_sim, votes, keep, nullptr);
  return check_launch("knn_verify_kernel");
}

static int g_knn_tc_version = -1;
static int knn_tc_version() {   // LVCB200_KNN_TC=1: the round-1 path (TF32 scores + per-query exact re-rank); default 3: bf16-pair scores, top-k in the epilogue
  if (g_knn_tc_version < 0) {
    const char* e = getenv("LVCB200_KNN_TC");
    g_knn_tc_version = e ? atoi(e) : 3;
  }
  return g_knn_tc_version;
}
extern "C" int lvcb200_knn_tc_select(int version) {
  LVC_REQUIRE(version == 1 || version == 3, "knn_tc_select: version must be 1 (TF32 + exact re-rank) or 3 (bf16 pairs, top-k in the epilogue)");
  g_knn_tc_version = version;
  return 0;
}

extern "C" size_t lvcb200_knn_tc_workspace(int64_t Q, int S, int D) {
  const size_t v1 = align_up((size_t)Q * knn_s_pad(S) * 2, 256) + align_up((size_t)Q, 256);
  const size_t v3 = knn3_workspace_bytes(Q, D);
  return v1 > v3 ? v1 : v3;
}

namespace {
struct SimtFallbackCtx {
  const float* mean; const float* bhat; const int64_t* bank_cls; int S, D; const float* queries; const int64_t* query_cls; int64_t Q;
  int topk, knn; int64_t* top_idx; float* top_sim; int64_t* votes; uint8_t* keep;
};
int simt_fallback_launch(const uint8_t* overflow, cudaStream_t st, void* vctx) {
  const SimtFallbackCtx& c = *(const SimtFallbackCtx*)vctx;
  knn_verify_kernel<<<(unsigned)ceil_div64(c.Q, QT), 256, 0, st>>>(c.mean, c.bhat, c.bank_cls, c.S, c.D, c.queries, c.query_cls, c.Q, c.topk, c.knn,
                                                                   c.top_idx, c.top_sim, c.votes, c.keep, overflow);
  return check_launch("knn_verify_kernel");
}
}  // namespace

extern "C" int lvcb200_knn_verify_tc(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                                     const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                                     int64_t* votes, uint8_t* keep, void* workspace, size_t workspace_bytes, void* stream) {
  LVC_REQUIRE(S >= 64 && S <= 4096 && D >= 32 && D % 8 == 0 && D <= 4096, "knn_verify_tc: need 64 <= S <= 4096, D a multiple of 8 in [32, 4096]");
  LVC_REQUIRE(topk >= 1 && topk <= KNN_MAXK && topk <= S && knn >= 1, "knn_verify_tc: need 1 <= topk <= min(16, S), knn >= 1");
  if (Q == 0) return 0;
  LVC_REQUIRE(bank_prepared && bank_cls && queries && query_cls && top_idx && votes && keep && workspace, "knn_verify_tc: NULL pointer");
  LVC_REQUIRE(((uintptr_t)queries % 16) == 0, "knn_verify_tc: queries must be 16-byte aligned");
  if (workspace_bytes < lvcb200_knn_tc_workspace(Q, S, D)) return set_error(LVCB200_EWORKSPACE, "knn_verify_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int S_pad = knn_s_pad(S);
  const float* mean = (const float*)bank_prepared;
  const float* negc = mean + D;
  const float* bhat = negc + knn_s_al(S);
  if (knn_tc_version() >= 3 && topk <= 10 && D >= 64) {
    const uint8_t* tail = (const uint8_t*)bank_prepared + knn_f32_part(S, D);
    SimtFallbackCtx ctx{mean, bhat, bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx, top_sim, votes, keep};
    return knn3_verify(mean, negc, bhat, tail + 256, (const float*)tail, bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx, top_sim, votes,
                       keep, workspace, st, simt_fallback_launch, &ctx);
  }
  __half* scores = (__half*)workspace;
  uint8_t* overflow = (uint8_t*)workspace + align_up((size_t)Q * S_pad * 2, 256);
  LVC_CUDA(cudaMemsetAsync(overflow, 0, (size_t)Q, st));
  // (1) approximate scores on the tensor cores: scores[q, s] = q . bhat_s - mu . bhat_s   (TF32 operands, fp16 out)
  lvcb200_gemm_desc g;
  memset(&g, 0, sizeof(g));
  g.a_dtype = LVCB200_F32;
  g.A = queries; g.lda = D; g.M_rows = Q;
  g.W = bhat; g.ldw = D;
  g.bias = negc;
  g.D = scores; g.ldd = S_pad; g.d_dtype = LVCB200_F16;
  g.M = Q; g.N = S_pad; g.K = D; g.taps = 1;
  int rc = lvcb200_gemm_bf16(&g, stream);
  if (rc) return rc;
  // (2) rigorous candidate selection + exact fp32 re-scoring
  const size_t per_warp = (((size_t)S_pad * 2 + 15) / 16 * 16 + (size_t)D * 4 + RR_CAND * 8 + 15) / 16 * 16;
  const size_t smem = per_warp * RR_WARPS;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    LVC_CUDA(cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int64_t blocks = ceil_div64(Q, RR_WARPS);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  knn_rerank_kernel<<<(unsigned)blocks, RR_WARPS * 32, smem, st>>>(mean, bhat, bank_cls, S, S_pad, D, queries, query_cls, Q, scores, topk,
                                                                 knn, top_idx, top_sim, votes, keep, overflow);
  if ((rc = check_launch("knn_rerank_kernel"))) return rc;
  // (3) exact SIMT pass over the (normally zero) queries whose candidate set overflowed
  knn_verify_kernel<<<(unsigned)ceil_div64(Q, QT), 256, 0, st>>>(mean, bhat, bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx,
                                                                top_sim, votes, keep, overflow);
  return check_launch("knn_verify_kernel");
}
