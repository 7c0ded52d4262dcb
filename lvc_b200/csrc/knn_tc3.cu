// kNN label verification, tensor-core path v2 (tools/run_nearest_neighbours.py:142-162, 214-227): fp32-grade scores straight
// from the bf16 tensor pipe, top-k selected in the epilogue out of TMEM -- no score matrix, no per-query re-rank gathers.
//
//   knn_split_queries_kernel : q (fp32) -> bf16 pair q = q_hi + q_lo (16 significant bits), |q|^2 and q . mu per query
//   knn_tc3_kernel           : persistent, one 128-query tile per CTA pass.  The bank (centred, normalised, split into a bf16
//       pair by knn_prepare) is swept in chunks of 160 rows: per 64-wide K block the three products q_hi b_hi + q_lo b_hi +
//       q_hi b_lo accumulate in one fp32 TMEM tile (tcgen05.mma kind::f16, operands by TMA, 3-stage mbarrier ring); two TMEM
//       buffers let the epilogue of chunk i overlap the MMAs of chunk i+1.  Epilogue: one thread per query reads its 160 scores
//       (tcgen05.ld), adds -mu . bhat_s, and keeps the 12 best (score, index) keys in registers across the chunks.
//       Error of a score: operand representation 3 * 2^-18 |q| (rigorous, Cauchy-Schwarz, |bhat| = 1) + fp32 accumulation in
//       the tensor core; eps = 2e-5 |q| covers both with a wide margin.  If all gaps among the 11 best exceed 2 eps the order
//       is certain and the thread writes top_idx / votes / mode / keep itself.  Otherwise the query is flagged:
//   knn_resolve_kernel       : warp per flagged query, EXACT fp32 re-scoring of its 11 candidates (they contain the true top-10
//       whenever the 12th approximate score is more than 2 eps below the 10th; if not, the query goes to the exact SIMT kernel).
// The results are the exact-fp32 top-k of the SIMT kernel (same tie rule: lower bank index first).
#include <cuda_fp16.h>
#include <string.h>

#include "tc_ptx.cuh"

namespace lvcb200 {

constexpr int K3_CN = 160;            // bank rows per chunk (UMMA N)
constexpr int K3_STAGES = 3;
constexpr int K3_STAGE_A = BLOCK_M * BLOCK_K * 2;          // 16 KB (one of hi / lo)
constexpr int K3_STAGE_B = K3_CN * BLOCK_K * 2;            // 20 KB
constexpr int K3_STAGE = 2 * K3_STAGE_A + 2 * K3_STAGE_B;  // 72 KB
constexpr int K3_LIST = 12;
constexpr int K3_THREADS = 256;
constexpr int K3_SMEM = K3_STAGES * K3_STAGE + 1024 /*ctrl*/ + 1024 /*align*/;

// one warp per query row: bf16 pair, |q|^2, q . mu
__global__ void __launch_bounds__(256)
knn_split_queries_kernel(const float* __restrict__ q, int64_t Q, int D, const float* __restrict__ mean, __nv_bfloat16* __restrict__ pair,
                         int64_t lo_rows, float* __restrict__ qn2, float* __restrict__ qmu) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const float* src = q + row * D;
  __nv_bfloat16* hi = pair + row * D;
  __nv_bfloat16* lo = pair + (row + lo_rows) * D;
  float s2 = 0.f, sm = 0.f;
  for (int k = lane * 8; k < D; k += 256) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + k)), b = __ldg(reinterpret_cast<const float4*>(src + k + 4));
    const float4 ma = __ldg(reinterpret_cast<const float4*>(mean + k)), mb = __ldg(reinterpret_cast<const float4*>(mean + k + 4));
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const float m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
    uint4 h, l;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
    __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      hh[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      const float2 f = __bfloat1622float2(hh[i]);
      ll[i] = __floats2bfloat162_rn(__fsub_rn(v[2 * i], f.x), __fsub_rn(v[2 * i + 1], f.y));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { s2 = fmaf(v[i], v[i], s2); sm = fmaf(v[i], m[i], sm); }
    *reinterpret_cast<uint4*>(hi + k) = h;
    *reinterpret_cast<uint4*>(lo + k) = l;
  }
  for (int o = 16; o; o >>= 1) { s2 += __shfl_xor_sync(0xffffffffu, s2, o); sm += __shfl_xor_sync(0xffffffffu, sm, o); }
  if (lane == 0) { qn2[row] = s2; qmu[row] = sm; }
}

// bank pair from the prepared fp32 bhat: rows >= S are zero; mu2 = |mu|^2
__global__ void __launch_bounds__(256)
knn_split_bank_kernel(const float* __restrict__ bhat, int S, int D, int rows_alloc, const float* __restrict__ mean,
                      __nv_bfloat16* __restrict__ pair, float* __restrict__ mu2) {
  const int s = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = s < S ? bhat[(size_t)s * D + d] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    pair[(size_t)s * D + d] = h;
    pair[(size_t)(s + rows_alloc) * D + d] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(h)));
  }
  if (s == 0) {
    __shared__ double red[8];
    double a = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) a += (double)mean[d] * (double)mean[d];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < 8; i++) t += red[i]; *mu2 = (float)t; }
  }
}

struct Knn3Params {
  int64_t Q; int S; int D;
  int q_lo_rows, b_lo_rows;          // row offset of the lo halves in the pair matrices
  int n_chunks, k_blocks, m_tiles;
  int topk, knn;
  const float* negc; const float* qn2; const float* qmu; const float* mu2;
  const int64_t* bank_cls; const int64_t* query_cls;
  int64_t* top_idx; float* top_sim; int64_t* votes; uint8_t* keep;
  int32_t* cand; uint8_t* flag;      // [Q, 11] candidate rows and 0 = done / 1 = resolve exactly / 2 = exact SIMT fallback
};

__device__ __forceinline__ void k3_insert(unsigned long long (&t)[K3_LIST], unsigned long long key) {
  if (key <= t[K3_LIST - 1]) return;
  t[K3_LIST - 1] = key;
#pragma unroll
  for (int i = K3_LIST - 1; i > 0; i--) {
    const unsigned long long a = t[i - 1], b = t[i];
    const bool sw = b > a;
    t[i - 1] = sw ? b : a;
    t[i] = sw ? a : b;
  }
}
__device__ __forceinline__ float k3_score(unsigned long long key) { return ordered_to_float((uint32_t)(key >> 32)); }
__device__ __forceinline__ int k3_index(unsigned long long key) { return (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)); }

__global__ void __launch_bounds__(K3_THREADS, 1)
knn_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_b, const Knn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  uint8_t* ctrl = smem_al + K3_STAGES * K3_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);   // full[3], empty[3], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 128);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * K3_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * K3_STAGES, bar_tempty = bar_tfull + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < K3_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_chunks = p.n_chunks, k_blocks = p.k_blocks;

  if (warp == 0) {
    if (lane == 0) {  // ===================================== TMA producer
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        const int m0 = tile * BLOCK_M;
        for (int c = 0; c < n_chunks; c++) {
          const int b0 = c * K3_CN;
          for (int kb = 0; kb < k_blocks; kb++) {
            const uint32_t fb = bar_full + 8 * stage;
            const uint32_t sa = smem_base + stage * K3_STAGE;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            mbar_arrive_expect_tx(fb, K3_STAGE);
            tma_load_2d(sa, &tmap_q, fb, kb * BLOCK_K, m0);
            tma_load_2d(sa + K3_STAGE_A, &tmap_q, fb, kb * BLOCK_K, m0 + p.q_lo_rows);
            tma_load_2d(sa + 2 * K3_STAGE_A, &tmap_b, fb, kb * BLOCK_K, b0);
            tma_load_2d(sa + 2 * K3_STAGE_A + K3_STAGE_B, &tmap_b, fb, kb * BLOCK_K, b0 + p.b_lo_rows);
            if (++stage == K3_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================================== MMA issuer
      uint32_t stage = 0, phase = 0, cc = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int c = 0; c < n_chunks; c++, cc++) {
          const uint32_t b = cc & 1u, bph = (cc >> 1) & 1u;
          int n_eff = p.S - c * K3_CN;
          n_eff = n_eff >= K3_CN ? K3_CN : ((n_eff + 15) & ~15);
          const uint32_t idesc = make_idesc_bf16(BLOCK_M, n_eff);
          mbar_wait(bar_tempty + 8 * b, bph ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + b * K3_CN;
          for (int kb = 0; kb < k_blocks; kb++) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * K3_STAGE;
            const uint64_t ahi = make_smem_desc_sw128(sa), alo = make_smem_desc_sw128(sa + K3_STAGE_A);
            const uint64_t bhi = make_smem_desc_sw128(sa + 2 * K3_STAGE_A), blo = make_smem_desc_sw128(sa + 2 * K3_STAGE_A + K3_STAGE_B);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
              umma_bf16(tmem_d, ahi + 2 * k, bhi + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_bf16(tmem_d, alo + 2 * k, bhi + 2 * k, idesc, 1u);
              umma_bf16(tmem_d, ahi + 2 * k, blo + 2 * k, idesc, 1u);
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == K3_STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * b);
        }
      }
    }
  } else if (warp >= 4) {  // ================================= epilogue: one thread per query of the tile
    const int qd = warp & 3;
    uint32_t cc = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const int64_t q = (int64_t)tile * BLOCK_M + qd * 32 + lane;
      unsigned long long top[K3_LIST];
#pragma unroll
      for (int i = 0; i < K3_LIST; i++) top[i] = 0ull;
      for (int c = 0; c < n_chunks; c++, cc++) {
        const uint32_t b = cc & 1u, bph = (cc >> 1) & 1u;
        const int s0 = c * K3_CN;
        int n_here = p.S - s0; n_here = n_here > K3_CN ? K3_CN : n_here;
        if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + b * K3_CN;
#pragma unroll 1
        for (int c0 = 0; c0 < n_here; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int s = s0 + c0 + j;
            if (c0 + j < n_here) {
              const float sc = __fadd_rn(__uint_as_float(v[j]), __ldg(p.negc + s));
              k3_insert(top, ((unsigned long long)float_to_ordered(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)s));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
      }
      if (q >= p.Q) continue;
      // certainty check on the 12 best approximate scores (dot-product units)
      const float nq = sqrtf(fmaxf(p.qn2[q], 0.f));
      const float eps2 = 4e-5f * nq + 1e-30f;                       // 2 eps
      const int topk = p.topk;
      float sc[K3_LIST];
#pragma unroll
      for (int i = 0; i < K3_LIST; i++) sc[i] = k3_score(top[i]);
      int state = 0;
      const int avail = p.S < K3_LIST ? p.S : K3_LIST;              // list entries that are real rows
      if (avail > topk + 1 && sc[topk + 1] >= sc[topk - 1] - eps2) state = 2;   // the 12th could belong to the top-10: exact SIMT fallback
      else {
#pragma unroll
        for (int i = 0; i < K3_LIST - 2; i++)
          if (i < topk && i + 1 < avail && sc[i] - sc[i + 1] < eps2) state = 1;
      }
      if (state != 0) {
        p.flag[q] = (uint8_t)state;
#pragma unroll
        for (int i = 0; i < K3_LIST - 1; i++) p.cand[q * (K3_LIST - 1) + i] = i < avail ? k3_index(top[i]) : -1;
        continue;
      }
      const float ncq2 = p.qn2[q] - 2.f * p.qmu[q] + *p.mu2;        // |q - mu|^2
      const float ncq = sqrtf(fmaxf(ncq2, 0.f));
      const float inv_n = 1.0f / (ncq > 1e-8f ? ncq : 1e-8f);
      int64_t vt[K3_LIST];
#pragma unroll
      for (int i = 0; i < K3_LIST; i++) {
        if (i < topk) {
          const int idx = k3_index(top[i]);
          vt[i] = p.bank_cls[idx];
          p.top_idx[q * topk + i] = idx;
          p.votes[q * topk + i] = vt[i];
          if (p.top_sim) p.top_sim[q * topk + i] = sc[i] * inv_n;
        } else vt[i] = -1;
      }
      const int kk = p.knn < topk ? p.knn : topk;
      int64_t best_v = 0; int best_c = 0;                            // torch.mode: most frequent, smallest value on ties
#pragma unroll
      for (int a = 0; a < K3_LIST; a++) {
        if (a >= kk) continue;
        int cnt = 0;
#pragma unroll
        for (int b2 = 0; b2 < K3_LIST; b2++) cnt += (b2 < kk && vt[b2] == vt[a]);
        if (cnt > best_c || (cnt == best_c && vt[a] < best_v)) { best_c = cnt; best_v = vt[a]; }
      }
      p.keep[q] = (p.query_cls[q] == best_v) ? 1 : 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// warp per query with flag == 1: exact fp32 centred-cosine scores of its 11 candidates, then top-k / votes / mode / keep
__global__ void __launch_bounds__(256)
knn_resolve_kernel(const float* __restrict__ mean, const float* __restrict__ bhat, const int64_t* __restrict__ bank_cls, int S, int D,
                   const float* __restrict__ queries, const int64_t* __restrict__ query_cls, int64_t Q, const int32_t* __restrict__ cand,
                   const uint8_t* __restrict__ flag, int topk, int knn, int64_t* __restrict__ top_idx, float* __restrict__ top_sim,
                   int64_t* __restrict__ votes, uint8_t* __restrict__ keep) {
  constexpr int NC = K3_LIST - 1;
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q || flag[q] != 1) return;
  int ci[NC];
#pragma unroll
  for (int j = 0; j < NC; j++) ci[j] = cand[q * NC + j];
  float dot[NC];
#pragma unroll
  for (int j = 0; j < NC; j++) dot[j] = 0.f;
  float nqc = 0.f;
  for (int k = lane * 4; k < D; k += 128) {
    float4 v = __ldg(reinterpret_cast<const float4*>(queries + q * D + k));
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + k));
    v.x = __fsub_rn(v.x, m.x); v.y = __fsub_rn(v.y, m.y); v.z = __fsub_rn(v.z, m.z); v.w = __fsub_rn(v.w, m.w);
    nqc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int j = 0; j < NC; j++) {
      if (ci[j] < 0) continue;
      const float4 b = __ldg(reinterpret_cast<const float4*>(bhat + (size_t)ci[j] * D + k));
      dot[j] += v.x * b.x + v.y * b.y + v.z * b.z + v.w * b.w;
    }
  }
  for (int o = 16; o; o >>= 1) {
    nqc += __shfl_xor_sync(0xffffffffu, nqc, o);
#pragma unroll
    for (int j = 0; j < NC; j++) dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
  }
  const float nrm = sqrtf(nqc);
  const float inv_n = 1.0f / (nrm > 1e-8f ? nrm : 1e-8f);
  // lane j holds candidate j; rank by (sim desc, index asc)
  float mv = -INFINITY; int mi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NC; j++) if (lane == j && ci[j] >= 0) { mv = dot[j] * inv_n; mi = ci[j]; }
  int rank = 0;
  for (int o = 0; o < NC; o++) {
    const float ov = __shfl_sync(0xffffffffu, mv, o); const int oi = __shfl_sync(0xffffffffu, mi, o);
    rank += (ov > mv) || (ov == mv && oi < mi);
  }
  const bool out_lane = lane < NC && mi != 0x7fffffff && rank < topk;
  const int64_t vote = out_lane ? bank_cls[mi] : -1;
  if (out_lane) {
    top_idx[q * topk + rank] = mi;
    votes[q * topk + rank] = vote;
    if (top_sim) top_sim[q * topk + rank] = mv;
  }
  int64_t myvote = -1;   // lane r <- vote of rank r
  for (int r2 = 0; r2 < topk; r2++) {
    const unsigned int m = __ballot_sync(0xffffffffu, out_lane && rank == r2);
    const int64_t v = __shfl_sync(0xffffffffu, vote, m ? __ffs(m) - 1 : 0);
    if (lane == r2) myvote = m ? v : -1;
  }
  const int kk = knn < topk ? knn : topk;
  int cnt = 0;
  for (int r2 = 0; r2 < kk; r2++) { const int64_t v = __shfl_sync(0xffffffffu, myvote, r2); cnt += (lane < kk && v == myvote); }
  int bc = (lane < kk) ? cnt : -1; int64_t bvv = myvote;
  for (int o = 16; o; o >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, bc, o); const int64_t ov = __shfl_xor_sync(0xffffffffu, bvv, o);
    if (oc > bc || (oc == bc && ov < bvv)) { bc = oc; bvv = ov; }
  }
  if (lane == 0) keep[q] = (query_cls[q] == bvv) ? 1 : 0;
}

__global__ void knn_flag_to_overflow_kernel(const uint8_t* __restrict__ flag, int64_t Q, uint8_t* __restrict__ overflow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) overflow[i] = flag[i] == 2 ? 1 : 0;
}

}  // namespace lvcb200

using namespace lvcb200;

// ---- shared with knn.cu (layout of the prepared bank)
int knn3_bank_rows(int S) { return (S + K3_CN - 1) / K3_CN * K3_CN; }

size_t knn3_workspace_bytes(int64_t Q, int D) {
  const int64_t Sq = (Q + 127) / 128 * 128;
  return align_up((size_t)2 * Sq * D * 2, 256) + 2 * align_up((size_t)Q * 4, 256) + align_up((size_t)Q * (K3_LIST - 1) * 4, 256) +
         2 * align_up((size_t)Q, 256);
}

int knn3_split_bank(const float* bhat, int S, int D, const float* mean, void* bpair, float* mu2, cudaStream_t st) {
  const int rows = knn3_bank_rows(S);
  knn_split_bank_kernel<<<rows, 256, 0, st>>>(bhat, S, D, rows, mean, (__nv_bfloat16*)bpair, mu2);
  return check_launch("knn_split_bank_kernel");
}

// the whole v2 path; `simt_fallback` runs knn_verify_kernel over the queries whose overflow byte is set (defined in knn.cu)
int knn3_verify(const float* mean, const float* negc, const float* bhat, const void* bpair, const float* mu2, const int64_t* bank_cls, int S,
                int D, const float* queries, const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                int64_t* votes, uint8_t* keep, void* workspace, cudaStream_t st, int (*simt_fallback)(const uint8_t*, cudaStream_t, void*),
                void* fb_ctx) {
  const int64_t Sq = (Q + 127) / 128 * 128;
  uint8_t* ws = (uint8_t*)workspace;
  __nv_bfloat16* qpair = (__nv_bfloat16*)ws; ws += align_up((size_t)2 * Sq * D * 2, 256);
  float* qn2 = (float*)ws; ws += align_up((size_t)Q * 4, 256);
  float* qmu = (float*)ws; ws += align_up((size_t)Q * 4, 256);
  int32_t* cand = (int32_t*)ws; ws += align_up((size_t)Q * (K3_LIST - 1) * 4, 256);
  uint8_t* flag = ws; ws += align_up((size_t)Q, 256);
  uint8_t* overflow = ws;
  LVC_CUDA(cudaMemsetAsync(flag, 0, (size_t)Q, st));
  knn_split_queries_kernel<<<(unsigned)ceil_div64(Q * 32, 256), 256, 0, st>>>(queries, Q, D, mean, qpair, Sq, qn2, qmu);
  int rc = check_launch("knn_split_queries_kernel");
  if (rc) return rc;
  const int brows = knn3_bank_rows(S);
  CUtensorMap tq, tb;
  if ((rc = make_tmap_2d(&tq, qpair, 2 * Sq, D, D, BLOCK_M))) return rc;
  if ((rc = make_tmap_2d(&tb, bpair, 2 * (long long)brows, D, D, K3_CN))) return rc;
  Knn3Params p;
  p.Q = Q; p.S = S; p.D = D;
  p.q_lo_rows = (int)Sq; p.b_lo_rows = brows;
  p.n_chunks = brows / K3_CN; p.k_blocks = (D + BLOCK_K - 1) / BLOCK_K; p.m_tiles = (int)(Sq / BLOCK_M);
  p.topk = topk; p.knn = knn;
  p.negc = negc; p.qn2 = qn2; p.qmu = qmu; p.mu2 = mu2;
  p.bank_cls = bank_cls; p.query_cls = query_cls;
  p.top_idx = top_idx; p.top_sim = top_sim; p.votes = votes; p.keep = keep;
  p.cand = cand; p.flag = flag;
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(knn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM));
    attr_set = true;
  }
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  knn_tc3_kernel<<<grid, K3_THREADS, K3_SMEM, st>>>(tq, tb, p);
  if ((rc = check_launch("knn_tc3_kernel"))) return rc;
  knn_resolve_kernel<<<(unsigned)ceil_div64(Q * 32, 256), 256, 0, st>>>(mean, bhat, bank_cls, S, D, queries, query_cls, Q, cand, flag, topk, knn,
                                                                       top_idx, top_sim, votes, keep);
  if ((rc = check_launch("knn_resolve_kernel"))) return rc;
  knn_flag_to_overflow_kernel<<<(unsigned)ceil_div64(Q, 256), 256, 0, st>>>(flag, Q, overflow);
  if ((rc = check_launch("knn_flag_to_overflow_kernel"))) return rc;
  return simt_fallback(overflow, st, fb_ctx);
}
