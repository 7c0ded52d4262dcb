// kNN label verification, tensor-core path v2 (tools/run_nearest_neighbours.py:142-162, 214-227): fp32-grade scores straight
// from the bf16 tensor pipe, top-k selected in the epilogue out of TMEM -- no score matrix, no per-query re-rank gathers.
//
//   knn_split_queries_kernel : q (fp32) -> bf16 pair q = q_hi + q_lo (16 significant bits), |q|^2 and q . mu per query
//   knn_tc3_kernel           : persistent, one 128-query tile per CTA pass.  The bank (centred, normalised, split into a bf16
//       pair by knn_prepare) is swept in chunks of 160 rows: per 64-wide K block the three products q_hi b_hi + q_lo b_hi +
//       q_hi b_lo accumulate in one fp32 TMEM tile (tcgen05.mma kind::f16, operands by TMA, 3-stage mbarrier ring); two TMEM
//       buffers let the epilogue of chunk i overlap the MMAs of chunk i+1.  Epilogue: one thread per query reads its 160 scores
//       (tcgen05.ld), adds -mu . bhat_s, and keeps the 14 best (score, index) entries in registers across the chunks.
//       Error of a score: operand representation 3 * 2^-18 |q| (rigorous, Cauchy-Schwarz, |bhat| = 1) + fp32 accumulation in
//       the tensor core; eps = 2e-5 |q| covers both with a wide margin.  If all gaps that matter among the 14 best exceed 2 eps the order
//       is certain and the thread writes top_idx / votes / mode / keep itself.  Otherwise the query is flagged:
//   knn_resolve_kernel       : warp per flagged query, EXACT fp32 re-scoring of the uncertain ones among its 13 candidates (they contain the true top-10
//       whenever the 14th approximate score is more than 2 eps below the 10th; if not, the query goes to the exact SIMT kernel).
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
constexpr int K3_LIST = 14;             // best (score, index) entries kept per query: top-10 + 3 boundary candidates + 1 sentinel
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
  int32_t* nflag; int32_t* flist;                              // compact list of the flagged queries (the resolve kernel runs dense warps)
  int32_t* cand; float* csc; uint16_t* amask; uint8_t* flag;   // per flagged query: 11 candidate rows, their approximate scores, which of them
                                                               // sit in an uncertain cluster; flag 0 = done / 1 = re-score the clusters / 2 = exact full scan
};

// Sorted insertion into the per-query list (score as order-preserving uint32, bank row).  Bank rows reach a thread in ascending order,
// so a strict comparison on the score alone leaves an equal score behind the earlier (lower) row: exactly the reference's tie rule.
__device__ __forceinline__ void k3_insert(uint32_t (&ts)[K3_LIST], int (&ti)[K3_LIST], uint32_t key, int idx) {
  if (key <= ts[K3_LIST - 1]) return;
  ts[K3_LIST - 1] = key; ti[K3_LIST - 1] = idx;
#pragma unroll
  for (int i = K3_LIST - 1; i > 0; i--) {
    const uint32_t a = ts[i - 1], b = ts[i];
    const int ia = ti[i - 1], ib = ti[i];
    const bool sw = b > a;
    ts[i - 1] = sw ? b : a; ts[i] = sw ? a : b;
    ti[i - 1] = sw ? ib : ia; ti[i] = sw ? ia : ib;
  }
}

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
      uint32_t ts[K3_LIST]; int ti[K3_LIST];
#pragma unroll
      for (int i = 0; i < K3_LIST; i++) { ts[i] = 0u; ti[i] = -1; }
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
          // two phases, because every thread owns a different query: (1) branch-free, which of my 32 scores beat my current 12th best;
          // (2) insert only those.  A per-element `if (beats) insert` makes the whole warp walk the 90-instruction insertion whenever
          // ANY of its 32 queries accepts the element -- which is nearly always (first version: 352 M instructions, tensor pipe 29 %).
          const uint32_t kth = ts[K3_LIST - 1];
          uint32_t mask = 0u;
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int sj = s0 + c0 + j < p.S ? s0 + c0 + j : p.S - 1;
            const uint32_t o = float_to_ordered(__fadd_rn(__uint_as_float(v[j]), __ldg(p.negc + sj)));
            v[j] = o;
            mask |= (uint32_t)(c0 + j < n_here && o > kth) << j;
          }
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1u;
            k3_insert(ts, ti, v[j], s0 + c0 + j);
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
      for (int i = 0; i < K3_LIST; i++) sc[i] = ordered_to_float(ts[i]);
      int state = 0;
      const int avail = p.S < K3_LIST ? p.S : K3_LIST;              // list entries that are real rows
      // candidates that sit in a cluster of scores closer than 2 eps which reaches into the top-k: only these need exact scores
      uint32_t am = 0u;
#pragma unroll
      for (int i = 0; i < K3_LIST - 2; i++)
        if (i + 1 < avail && sc[i] - sc[i + 1] < eps2 && (i < topk || ((am >> i) & 1u))) am |= 3u << i;
      if (am) state = 1;
      // the true top-k lies within {approx >= k-th approx - 2 eps}; if even the last list entry is that close the list may miss a member
      if (avail == K3_LIST && sc[K3_LIST - 1] >= sc[topk - 1] - eps2) state = 2;
      if (state != 0) {
        p.flag[q] = (uint8_t)state;
        p.flist[atomicAdd(p.nflag, 1)] = (int32_t)q;
        p.amask[q] = (uint16_t)am;
#pragma unroll
        for (int i = 0; i < K3_LIST - 1; i++) {
          p.cand[q * (K3_LIST - 1) + i] = i < avail ? ti[i] : -1;
          p.csc[q * (K3_LIST - 1) + i] = sc[i];
        }
        continue;
      }
      const float ncq2 = p.qn2[q] - 2.f * p.qmu[q] + *p.mu2;        // |q - mu|^2
      const float ncq = sqrtf(fmaxf(ncq2, 0.f));
      const float inv_n = 1.0f / (ncq > 1e-8f ? ncq : 1e-8f);
      int64_t vt[K3_LIST];
#pragma unroll
      for (int i = 0; i < K3_LIST; i++) {
        if (i < topk) {
          const int idx = ti[i];
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

// warp per flagged query.  flag 1: exact fp32 centred dot products for the candidates inside uncertain clusters (amask); the others
// keep their approximate score -- clusters are separated by more than 2 eps, so sorting the mixed values gives the exact order.
// flag 2 (the candidate set itself is uncertain, ~5e-4 of the queries): exact scan of the whole bank by the warp.
__global__ void __launch_bounds__(256)
knn_resolve_kernel(const float* __restrict__ mean, const float* __restrict__ bhat, const int64_t* __restrict__ bank_cls, int S, int D,
                   const float* __restrict__ queries, const int64_t* __restrict__ query_cls, int64_t Q, const int32_t* __restrict__ cand,
                   const float* __restrict__ csc, const uint16_t* __restrict__ amask, const uint8_t* __restrict__ flag,
                   const int32_t* __restrict__ nflag, const int32_t* __restrict__ flist, int topk, int knn,
                   int64_t* __restrict__ top_idx, float* __restrict__ top_sim, int64_t* __restrict__ votes, uint8_t* __restrict__ keep) {
  constexpr int NC = K3_LIST - 1;
  extern __shared__ float rs_q[];                      // [8 warps][D] centred queries
  const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (slot >= Q || slot >= *nflag) return;
  const int64_t q = flist[slot];
  const int state = flag[q];
  float* qc = rs_q + (size_t)wib * D;
  float nqc = 0.f;
  for (int k0 = 0; k0 < D; k0 += 1024) {               // 8 independent 16-byte loads in flight per lane
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int k = k0 + u * 128 + lane * 4;
      v[u] = k < D ? __ldg(reinterpret_cast<const float4*>(queries + q * D + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int k = k0 + u * 128 + lane * 4;
      if (k < D) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean + k));
        float4 w = v[u];
        w.x = __fsub_rn(w.x, m.x); w.y = __fsub_rn(w.y, m.y); w.z = __fsub_rn(w.z, m.z); w.w = __fsub_rn(w.w, m.w);
        nqc += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
        *reinterpret_cast<float4*>(qc + k) = w;
      }
    }
  }
  for (int o = 16; o; o >>= 1) nqc += __shfl_xor_sync(0xffffffffu, nqc, o);
  const float nrm = sqrtf(nqc);
  const float inv_n = 1.0f / (nrm > 1e-8f ? nrm : 1e-8f);
  __syncwarp();
  auto exact_dot = [&](int row) {                      // all lanes return the full sum
    const float* br = bhat + (size_t)row * D;
    float d = 0.f;
    for (int k0 = 0; k0 < D; k0 += 1024) {
      float4 b[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + u * 128 + lane * 4;
        b[u] = k < D ? __ldg(reinterpret_cast<const float4*>(br + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = k0 + u * 128 + lane * 4;
        if (k < D) {
          const float4 a = *reinterpret_cast<const float4*>(qc + k);
          d += a.x * b[u].x + a.y * b[u].y + a.z * b[u].z + a.w * b[u].w;
        }
      }
    }
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    return d;
  };
  float mv = -INFINITY; int mi = 0x7fffffff;           // lane j: value and bank row of list entry j (sorted or not)
  int n_list = 0;
  if (state == 1) {
    const uint32_t am = amask[q];
    if (lane < NC) { mi = cand[q * NC + lane]; if (mi < 0) mi = 0x7fffffff; else mv = csc[q * NC + lane] * inv_n; }
    for (int j = 0; j < NC; j++) {
      if (!((am >> j) & 1u)) continue;                 // uniform across the warp
      const int row = __shfl_sync(0xffffffffu, mi, j);
      if (row == 0x7fffffff) continue;
      const float d = exact_dot(row) * inv_n;
      if (lane == j) mv = d;
    }
    n_list = NC;
  } else {
    // exact scan: lane r (< topk) holds the r-th best so far
    for (int s0 = 0; s0 < S; s0++) {
      const float d = exact_dot(s0) * inv_n;
      const bool mine_worse = lane < topk && (d > mv || (d == mv && s0 < mi));     // the new entry ranks before this lane's entry
      const unsigned int w = __ballot_sync(0xffffffffu, mine_worse);
      if (w) {
        const int pos = __ffs(w) - 1;                  // insertion position: entries at pos.. shift down by one
        const float uv = __shfl_up_sync(0xffffffffu, mv, 1); const int ui = __shfl_up_sync(0xffffffffu, mi, 1);
        if (lane > pos && lane < topk) { mv = uv; mi = ui; }
        if (lane == pos) { mv = d; mi = s0; }
      }
    }
    n_list = topk;
  }
  int rank = 0;
  for (int o = 0; o < n_list; o++) {
    const float ov = __shfl_sync(0xffffffffu, mv, o); const int oi = __shfl_sync(0xffffffffu, mi, o);
    rank += (ov > mv) || (ov == mv && oi < mi);
  }
  const bool out_lane = lane < n_list && mi != 0x7fffffff && rank < topk;
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

}  // namespace lvcb200

using namespace lvcb200;

// ---- shared with knn.cu (layout of the prepared bank)
int knn3_bank_rows(int S) { return (S + K3_CN - 1) / K3_CN * K3_CN; }

size_t knn3_workspace_bytes(int64_t Q, int D) {
  const int64_t Sq = (Q + 127) / 128 * 128;
  return align_up((size_t)2 * Sq * D * 2, 256) + 2 * align_up((size_t)Q * 4, 256) + 2 * align_up((size_t)Q * (K3_LIST - 1) * 4, 256) +
         align_up((size_t)Q * 2, 256) + align_up((size_t)Q, 256) + 256 + align_up((size_t)Q * 4, 256);
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
  float* csc = (float*)ws; ws += align_up((size_t)Q * (K3_LIST - 1) * 4, 256);
  uint16_t* amask = (uint16_t*)ws; ws += align_up((size_t)Q * 2, 256);
  uint8_t* flag = ws; ws += align_up((size_t)Q, 256);
  int32_t* nflag = (int32_t*)ws; ws += 256;
  int32_t* flist = (int32_t*)ws;
  LVC_CUDA(cudaMemsetAsync(flag, 0, align_up((size_t)Q, 256) + 256, st));   // flags and the list counter
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
  p.cand = cand; p.csc = csc; p.amask = amask; p.flag = flag; p.nflag = nflag; p.flist = flist;
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(knn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM));
    attr_set = true;
  }
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  knn_tc3_kernel<<<grid, K3_THREADS, K3_SMEM, st>>>(tq, tb, p);
  if ((rc = check_launch("knn_tc3_kernel"))) return rc;
  const size_t rs_smem = (size_t)8 * D * sizeof(float);
  static size_t rs_set = 0;
  if (rs_smem > rs_set && rs_smem > 48 * 1024) {
    LVC_CUDA(cudaFuncSetAttribute(knn_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    rs_set = rs_smem;
  }
  knn_resolve_kernel<<<(unsigned)ceil_div64(Q * 32, 256), 256, rs_smem, st>>>(mean, bhat, bank_cls, S, D, queries, query_cls, Q, cand, csc, amask, flag,
                                                                             nflag, flist, topk, knn, top_idx, top_sim, votes, keep);
  (void)simt_fallback; (void)fb_ctx;   // the uncertain-candidate-set case is an exact scan inside knn_resolve_kernel
  return check_launch("knn_resolve_kernel");
}
