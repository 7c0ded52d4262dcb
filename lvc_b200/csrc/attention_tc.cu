// Fused attention of the descriptor front end (the DINO ViT-S/8 forward, head_dim 64) on the 5th-generation tensor cores:
// softmax(Q K^T / sqrt(d)) V per (crop, head), flash style, with both products on tcgen05.mma and the accumulators in TMEM.
//
//   CTA = 128 queries of one (crop, head); 320 threads: warp 0 = TMA producer, warp 1 = single-lane MMA issuer (+ TMEM allocation),
//   warps 2-9 = softmax / epilogue, two threads per query row (TMEM lane = row; each thread owns 64 of the block's 128 key columns and 32 of
//   the 64 output columns, the pair agrees on the row maximum through shared memory once per block).
//   Per block of 128 keys:
//     S[128 x 128] = Q K^T            tcgen05.mma, A = Q tile, B = K tile (both K-major SWIZZLE_128B tiles loaded by TMA straight out of the
//                                     qkv matrix), fp32 in TMEM columns [0, 128)
//     softmax warps: two passes over S with tcgen05.ld (row max, then p = exp2(s - max)); P is written as bf16 into shared memory in the
//                                     K-major SWIZZLE_128B operand layout, the running max / sum / rescale factor stay in registers
//     O_blk[128 x 64] = P V           tcgen05.mma, A = P, B = the V tile exactly as TMA lands it ([128 keys x 64 d], d contiguous): an MN-major
//                                     SWIZZLE_128B operand (instruction-descriptor bit 16; 16 keys = 2 048 bytes per K step), so V needs no
//                                     transposed copy; TMEM columns [128, 192)
//     softmax warps: o = o * rescale + O_blk (tcgen05.ld; the running output lives in registers, so no TMEM read-modify-write)
//   K / V tiles cycle through a three-slot ring (K_j, V_j, K_j+1 in flight); two CTAs per SM (99 KB of shared memory, 256 TMEM columns
//   each) overlap one CTA's softmax with the other's MMAs and loads.
// Keys >= N are masked to -inf (the tile rows behind a crop's last token belong to the next crop, or are zero-filled by TMA at the end of
// the matrix): their P is exactly 0, and the V rows they multiply are finite.  Query rows >= N are computed and dropped.
#include "tc_ptx.cuh"

namespace lvcb200 {

constexpr int FA_TILE = 128 * 128;                 // bytes of one [128 x 64] bf16 tile (Q, K, V blocks; one 64-key half of P)
constexpr int FA_RING = 3;
constexpr int FA_THREADS = 320;                    // TMA warp, MMA warp, eight softmax warps (two threads per query row)
constexpr int FA_SMEM = (1 + FA_RING + 2) * FA_TILE + 1024 + 256 + 2048;
constexpr int FA_TMEM_COLS = 256;

#ifdef LVCB200_FA_TRACE
__device__ unsigned long long g_fa_trace[256];
#define FA_STAMP(slot)                                                                        \
  do {                                                                                        \
    if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) {                              \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
      g_fa_trace[slot] = t_;                                                                  \
    }                                                                                         \
  } while (0)
#else
#define FA_STAMP(slot) do {} while (0)
#endif

__device__ __forceinline__ uint32_t fa_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float fa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One thread's half row of S (64 scores in TMEM at `ts`), read in four 16-column pieces through two register buffers: the next piece's
// tcgen05.ld is in flight while the current one is processed (tcgen05.wait::ld waits for everything outstanding, so the load is issued
// right after the wait).  MASK: keys >= valid (relative to the half row) count as -inf.
template <bool MASK>
__device__ __forceinline__ float fa_row_max(uint32_t ts, int valid) {      // max of the raw scores (the scale is positive: applied by the caller)
  float mx = -INFINITY;
  uint32_t v[2][16];
  tmem_ld16(ts, v[0]);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    tmem_ld_wait();
    if (c < 3) tmem_ld16(ts + (c + 1) * 16, v[(c + 1) & 1]);
#pragma unroll
    for (int i = 0; i < 16; i++)
      if (!MASK || c * 16 + i < valid) mx = fmaxf(mx, __uint_as_float(v[c & 1][i]));
  }
  return mx;
}
// p = 2^(s * scale - mx) as bf16 into the row's slots of the P operand chunk (64 keys = one 128-byte swizzled row); returns the sum.
template <bool MASK>
__device__ __forceinline__ float fa_row_exp(uint32_t ts, uint32_t p_row, uint32_t sw, float scale_log2e, float mx, int valid) {
  float sum = 0.f;
  uint32_t v[2][16];
  tmem_ld16(ts, v[0]);
#pragma unroll
  for (int c = 0; c < 4; c++) {
    tmem_ld_wait();
    if (c < 3) tmem_ld16(ts + (c + 1) * 16, v[(c + 1) & 1]);
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float s0 = (!MASK || c * 16 + i < valid) ? fmaf(__uint_as_float(v[c & 1][i]), scale_log2e, -mx) : -INFINITY;
      const float s1 = (!MASK || c * 16 + i + 1 < valid) ? fmaf(__uint_as_float(v[c & 1][i + 1]), scale_log2e, -mx) : -INFINITY;
      const float p0 = fa_ex2(s0), p1 = fa_ex2(s1);
      sum += p0 + p1;
      pk[i >> 1] = fa_pack(p0, p1);
    }
#pragma unroll
    for (int g = 0; g < 2; g++) {       // 16-byte unit u = c * 2 + g of the row, stored at u ^ (row & 7)
      const uint32_t addr = p_row + ((((uint32_t)c * 2 + g) ^ sw) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * g]), "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3]) : "memory");
    }
  }
  return sum;
}

__global__ void __launch_bounds__(FA_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qk, int N, int H, float scale_log2e,
                    __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_q = smem_u32(smem), s_ring = s_q + FA_TILE, s_p = s_ring + FA_RING * FA_TILE;
  uint8_t* ctrl = smem + (1 + FA_RING + 2) * FA_TILE;
  const uint32_t bars = smem_u32(ctrl);
  const uint32_t bar_full = bars, bar_empty = bars + 8 * FA_RING, bar_q = bars + 16 * FA_RING, bar_s = bar_q + 8, bar_p = bar_q + 16,
                 bar_o = bar_q + 24, bar_oread = bar_q + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 16 * FA_RING + 48);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nkb = (N + 127) >> 7;
  const int row_base = b * N;                    // first token row of this crop in the qkv matrix

  if (threadIdx.x == 0) {
    for (int s = 0; s < FA_RING; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1); mbar_init(bar_oread, 8);   // one arrival per softmax warp: every arrival wakes the waiting MMA thread
    fence_barrier_init();
    tma_prefetch_desc(&tm_qk);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), FA_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {   // ============================================================ TMA producer
      mbar_arrive_expect_tx(bar_q, FA_TILE);
      tma_load_2d(s_q, &tm_qk, bar_q, h * 64, row_base + q0);
      uint32_t slot = 0, ph = 0;
      for (int j = 0; j < nkb; j++) {
#pragma unroll
        for (int kv = 0; kv < 2; kv++) {           // K block j, then V block j: same rows of qkv, columns (1 + kv) * H * 64 + h * 64
          mbar_wait(bar_empty + 8 * slot, ph ^ 1u);
          mbar_arrive_expect_tx(bar_full + 8 * slot, FA_TILE);
          tma_load_2d(s_ring + slot * FA_TILE, &tm_qk, bar_full + 8 * slot, ((1 + kv) * H + h) * 64, row_base + j * 128);
          if (++slot == FA_RING) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ============================================================ MMA issuer
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128), idesc_o = make_idesc_bf16(128, 64) | (1u << 16);   // bit 16: B is MN-major
      // One thread issues every MMA of the CTA and sits on the critical path of each key block (softmax -> P V -> next S): descriptors are
      // built once and advanced by adding to their address field (16-byte units), ring slot / phase counters advance incrementally.
      const uint64_t dq = make_smem_desc_sw128(s_q), dp = make_smem_desc_sw128(s_p), dr = make_smem_desc_sw128(s_ring);
      uint32_t slot = 0, ph = 0;
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nkb; j++) {
        {   // S = Q K^T.  S is free: P(j - 1) was complete (bar_p) before the previous P V product was issued
          FA_STAMP(16 * j + 8);
          mbar_wait(bar_full + 8 * slot, ph);
          tc_fence_after();
          FA_STAMP(16 * j + 9);
          const uint64_t dk = dr + (uint64_t)slot * (FA_TILE >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ks++) umma_bf16(t_s, dq + 2 * ks, dk + 2 * ks, idesc_s, ks > 0);      // 32 bytes per K step
          umma_commit(bar_empty + 8 * slot);
          umma_commit(bar_s);
          if (++slot == FA_RING) { slot = 0; ph ^= 1u; }
        }
        {   // O_blk = P V
          FA_STAMP(16 * j + 10);
          mbar_wait(bar_p, j & 1);
          FA_STAMP(16 * j + 11);
          mbar_wait(bar_full + 8 * slot, ph);
          if (j > 0) mbar_wait(bar_oread, (j - 1) & 1);
          tc_fence_after();
          FA_STAMP(16 * j + 12);
          const uint64_t dv = dr + (uint64_t)slot * (FA_TILE >> 4);
#pragma unroll
          for (int kc = 0; kc < 2; kc++)
#pragma unroll
            for (int ks = 0; ks < 4; ks++)     // A: 64-key half kc of P, 32 bytes per K step; B: 16 keys = 2 048 bytes per K step (MN-major)
              umma_bf16(t_o, dp + kc * (FA_TILE >> 4) + 2 * ks, dv + (kc * 4 + ks) * 128, idesc_o, (kc | ks) > 0);
          umma_commit(bar_empty + 8 * slot);
          umma_commit(bar_o);
          FA_STAMP(16 * j + 13);
          if (++slot == FA_RING) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else {             // ============================================================ softmax / epilogue: two threads per query row
    // warps 2-5 own key columns 0-63 of the block (P chunk 0) and output columns 0-31, warps 6-9 columns 64-127 and outputs 32-63; the
    // two threads of a row meet once per block to agree on the row maximum (shared-memory exchange + a 256-thread named barrier)
    const int quarter = warp & 3;                       // TMEM lanes this warp may read: 32 * (warp % 4) ..
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t ts = t_s + ((uint32_t)(quarter * 32) << 16) + half * 64, to = t_o + ((uint32_t)(quarter * 32) << 16) + half * 32;
    const uint32_t p_row = s_p + half * FA_TILE + row * 128;
    const uint32_t sw = (uint32_t)(row & 7);
    float* xch = reinterpret_cast<float*>(ctrl + 256);  // [2 (block parity)][2 (half)][128 (row)]
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; i++) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nkb; j++) {
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 0);
      if (lane == 0) mbar_wait(bar_s, j & 1);
      __syncwarp();
      tc_fence_after();
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 1);
      const int valid = N - j * 128 - half * 64;        // keys of this half row below N
      const float own = (valid >= 64 ? fa_row_max<false>(ts, valid) : fa_row_max<true>(ts, valid)) * scale_log2e;
      float* x = xch + (j & 1) * 256;
      x[half * 128 + row] = own;
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 3);
      const float mx = fmaxf(m, fmaxf(own, x[(half ^ 1) * 128 + row]));
      const float c0 = fa_ex2(m - mx);                  // 0 on the first block
      const float sum = valid >= 64 ? fa_row_exp<false>(ts, p_row, sw, scale_log2e, mx, valid) : fa_row_exp<true>(ts, p_row, sw, scale_log2e, mx, valid);
      l = l * c0 + sum;
      m = mx;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 4);
      // o = o * c0 + P V
      if (lane == 0) mbar_wait(bar_o, j & 1);
      __syncwarp();
      tc_fence_after();
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 5);
      {
        uint32_t v[32];
        tmem_ld32(to, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = fmaf(o[i], c0, __uint_as_float(v[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_oread);
      if (warp == 2 && lane == 0) FA_STAMP(16 * j + 6);
    }
    float* x = xch + (nkb & 1) * 256;                   // the buffer the last block did not use
    x[half * 128 + row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += x[(half ^ 1) * 128 + row];
    const int q = q0 + row;
    if (q < N) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(out + ((long long)(row_base + q) * H + h) * 64 + half * 32);
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint4 u;
        u.x = fa_pack(o[8 * g] * inv, o[8 * g + 1] * inv);
        u.y = fa_pack(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
        u.z = fa_pack(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
        u.w = fa_pack(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
        dst[g] = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_attention_tc(const void* qkv, int B, int N, int H, int head_dim, float scale, void* out, void* stream) {
  if (B == 0 || N == 0) return 0;
  LVC_REQUIRE(qkv && out && H >= 1 && head_dim == 64, "attention_tc: head_dim must be 64");
  LVC_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "attention_tc: 16-byte aligned pointers");
  LVC_REQUIRE((long long)B * N < (1ll << 31), "attention_tc: too many rows");
  CUtensorMap tm_qk;
  int rc = make_tmap_2d(&tm_qk, qkv, (long long)B * N, 3LL * H * 64, 3LL * H * 64, 128);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    attr_set = true;
  }
  attention_tc_kernel<<<dim3((N + 127) / 128, H, B), FA_THREADS, FA_SMEM, (cudaStream_t)stream>>>(tm_qk, N, H, scale * 1.4426950408889634f,
                                                                                                   (__nv_bfloat16*)out);
  return check_launch("attention_tc_kernel");
}

#ifdef LVCB200_FA_TRACE
extern "C" int lvcb200_debug_attention_trace(unsigned long long* host256) {
  LVC_CUDA(cudaMemcpyFromSymbol(host256, g_fa_trace, sizeof(unsigned long long) * 256));
  return 0;
}
#endif
