// Shift-GEMM on the 5th-generation tensor cores (sm_100a): tcgen05.mma with TMEM accumulators, operands
// staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), mbarrier producer/consumer pipeline, persistent CTAs.
//
//   D[m, n] = act( sum_t sum_k A[m + shift[t], k] * W[n, t*K + k] + bias[n] + residual[m, n] ) * rowmask[m]
//
// One kernel runs every dense layer of the mining path (see include/lvcb200.h): a KxK conv over a zero-bordered
// channels-last plane is this GEMM with one row shift per tap, so the im2col matrix is never materialised --
// TMA fetches the shifted [128 x 64] bf16 activation box straight from the plane, rows outside the tensor are
// zero-filled by the TMA unit.  FrozenBN is folded into W / bias on the host; bias + residual + ReLU + border
// re-zeroing are fused in the epilogue, which reads the fp32 accumulator out of TMEM (tcgen05.ld).
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warp 2 TMEM allocator,
// warps 4..7 epilogue (TMEM lane quadrant = warp % 4).  Two TMEM accumulator buffers let the epilogue of tile i
// overlap the main loop of tile i+1.  Tile = 128 x BLOCK_N, BLOCK_K = 64 (one 128-byte swizzle row).
#include "tc_ptx.cuh"

namespace lvcb200 {

constexpr int kGemmThreads = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-11: epilogue (two per TMEM lane quadrant)
constexpr int kEpiThreads = 256;
constexpr int kEpiWarp0 = 4;


struct GemmParams {
  const float* bias;
  void* D; long long ldd; int d_f32; int d_f16;
  long long M; int N; int K;
  int k_elems;                     // operand elements per 128-byte K block: 64 (bf16) or 32 (fp32 consumed as tf32)
  int taps; int shift[9];
  int relu;
  int plane_h, plane_w;
  int m_tiles, n_tiles, k_blocks;  // k_blocks per tap
  int has_res;                     // residual tile added on the tensor core: D += R_tile * I (identity B operand)
  int num_stages;                  // smem pipeline depth (runtime: deep for big-K layers, shallow + big staging for small-K)
  int phase_cols;                  // MODE 1: output columns staged per TMA-store phase (64 or 128)
  const __nv_bfloat16* up; long long ldu; int up_ph, up_pw;   // UPS instantiation: D += nearest-2x-upsample(up) (FPN top-down add)
  int warp_epi;                    // MODE 1: warp-private staging + one TMA store per warp per 64 columns (no CTA-wide barriers)
  int debug;                       // experiments only, bit mask: 1 = issue no MMAs, 2 = issue no TMA loads, 4 = issue no TMA stores (results are garbage)
  int split_rows;                  // SPLIT instantiations: row offset of the lo half of every hi/lo bf16 pair matrix (A, residual, D)
};

// MODE 0: epilogue stores straight from registers (fp32 heads, tiny N).
// MODE 1: bf16 output: TMEM -> registers -> (bias, ReLU, border zero) -> 128B-swizzled smem staging (double buffered) -> TMA
//         store; one fence + two CTA-local barriers per 64/128 output columns, stores complete asynchronously.
// The residual never touches the epilogue: its [128 x 64] tiles ride the operand pipeline as extra K-steps against an
// identity matrix (exact: bf16 * 1.0 accumulated in fp32), so D = A*W + R leaves the tensor core already summed.
constexpr int kStageBytesA = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int kIdentBytes = 64 * 64 * 2;              // 8 KB
constexpr int kCtrlBytes = 2048;
constexpr int kMaxStages = 8;

template <int BLOCK_N> struct GemmCfg {
  static constexpr int kStageBytesB = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
};

__host__ __device__ inline int gemm_smem_bytes(int block_n, int mode, int num_stages, int phase_cols, int has_res) {
  int b = num_stages * (kStageBytesA + block_n * BLOCK_K * 2);
  if (mode == 1) b += 2 * BLOCK_M * phase_cols * 2;
  if (has_res) b += kIdentBytes;
  return b + kCtrlBytes + 1024 /*alignment slack*/;
}

// KIND 0: bf16 operands (kind::f16); KIND 1: fp32 operands as tf32 (kind::tf32).  UPS 1: the epilogue adds the 2x-upsampled coarser FPN
// level (its own instantiation: the other layers' code is untouched).
// TAP3 1 (3x3 convs on narrow tiles): the three kw taps of a kernel row read ONE [136 x 64] A tile through row-shifted descriptors
// (tools/desc_probe.cu: exact for any 128-byte row offset inside a SWIZZLE_128B tile); a stage = that A tile + the three taps' B tiles.
constexpr int kTap3BytesA = 136 * 128;   // 17 KB: rows m0 + shift(kh, kw = 0) .. + 135 cover the 130 rows the three taps touch
// SPLIT 1 (strict engine mode, fp32-grade results on the bf16 tensor pipe): every real operand x is carried as a bf16 pair
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 significant bits, |x - hi - lo| <= 2^-18 |x|).  A, the residual and a bf16 D are
// [2 * split_rows, cols] matrices with the hi half at row 0 and the lo half at row split_rows; W is [N, taps * 2K] with [hi | lo]
// per tap.  The contraction is the classic three-term product A_hi W_hi + A_lo W_hi + A_hi W_lo (the dropped lo * lo term is
// 2^-18 relative), all accumulated in the same fp32 TMEM tile: three K loops per tap instead of one.
// erf GELU for the fused epilogue: 0.5 x (1 + erf(x / sqrt 2)) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2), erfc by Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7 in erf: a five-term polynomial in t = 1 / (1 + p z) times exp(-z^2)).  16 instructions, two of them MUFU, against ~30
// for erff(): the fc1 GEMM of the ViT MLP is bound by its epilogue (1.23 G output elements per layer through four epilogue warps per SM).
// The output is rounded to bf16 (2^-9 relative): the approximation error is three orders of magnitude below that.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  return fmaxf(x, 0.f) - 0.5f * ax * (q * t) * e;
}
// The same for two values with the packed fp32x2 instructions of sm_100 (FMUL2 / FFMA2 / FADD2): the polynomial and the products cost half
// the issue slots; the two MUFU per value and the max stay scalar.  Bit-identical to gelu_erf on each lane (the packed ops round each half
// like their scalar forms, and no product here is contracted differently).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = __fmul2_rn(ax, make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  float2 t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  float2 q = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  q = __ffma2_rn(q, t, make_float2(1.421413741f, 1.421413741f));
  q = __ffma2_rn(q, t, make_float2(-0.284496736f, -0.284496736f));
  q = __ffma2_rn(q, t, make_float2(0.254829592f, 0.254829592f));
  const float2 a = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
  const float2 h = __fmul2_rn(__fmul2_rn(__fmul2_rn(ax, make_float2(0.5f, 0.5f)), __fmul2_rn(q, t)), e);
  return make_float2(fmaxf(x.x, 0.f) - h.x, fmaxf(x.y, 0.f) - h.y);
}

// ACT 1: exact (erf) GELU instead of the ReLU flag (fc1 of the ViT MLP, csrc/vit.cu), its own instantiation.
template <int BLOCK_N, int MODE, int KIND, int UPS = 0, int TAP3 = 0, int SPLIT = 0, int ACT = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  const int kStages = p.num_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kA = TAP3 ? kTap3BytesA : kStageBytesA;
  constexpr int kB = TAP3 ? 3 * Cfg::kStageBytesB : Cfg::kStageBytesB;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + kStages * kA;
  const uint32_t off_staging = kStages * (kA + kB);
  const uint32_t staging_bytes = (MODE == 1) ? 2u * BLOCK_M * p.phase_cols * 2u : 0u;
  const uint32_t off_ident = off_staging + staging_bytes;
  const uint32_t off_ctrl = off_ident + (p.has_res ? kIdentBytes : 0);
  uint8_t* ctrl = smem_al + off_ctrl;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);   // full[8], empty[8], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 8 * (2 * kMaxStages + 4));
  float* s_bias = reinterpret_cast<float*>(ctrl + 1024);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages, bar_tempty = bar_tfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_w);
    if constexpr (MODE == 1) tma_prefetch_desc(&tmap_d);
    if (p.has_res) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, kEpiThreads / 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  if (p.has_res && warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // 64x64 bf16 identity, K-major, SWIZZLE_128B: row n = 128 bytes, 16-byte chunk c stored at position c ^ (n & 7)
    uint8_t* ident = smem_al + off_ident;
    const int t = threadIdx.x - kEpiWarp0 * 32;
    for (int i = t; i < kIdentBytes / 16; i += 128) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t < 64) {
      const int n = t, c = n >> 3;
      *reinterpret_cast<__nv_bfloat16*>(ident + n * 128 + ((c ^ (n & 7)) << 4) + (n & 7) * 2) = __float2bfloat16_rn(1.0f);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, identity tile) ran while
  // the previous kernel of the stream was still draining; its results are needed only from here on.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_blocks * (SPLIT ? 3 : 1);
  constexpr int kSplitPasses = SPLIT ? 2 : 1;        // residual / output halves

  if (warp == 0) {
    if (lane == 0) {  // ===================================== TMA producer
      // stage / phase advance incrementally: the stage count is a run-time value, and an integer modulo per K step was
      // the critical path of this single-thread loop (profiles: ~500 cycles per K iteration regardless of tile width)
      uint32_t stage = 0, phase = 0;
      const int n_tiles = p.n_tiles, k_blocks = p.k_blocks, taps = p.taps, k_elems = p.k_elems, Kdim = p.K, Ndim = p.N;
      const bool has_res = p.has_res != 0, no_tma = (p.debug & 2) != 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
        if constexpr (TAP3) {
          for (int kh = 0; kh < 3; kh++) {
            const int row = m0 + p.shift[3 * kh];             // the kw = 0 tap; kw = 1, 2 are the next two rows of the same tile
            for (int kb = 0; kb < k_blocks; kb++) {
              const uint32_t fb = bar_full + 8 * stage;
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              mbar_arrive_expect_tx(fb, kA + kB);
              tma_load_2d(smem_a0 + stage * kA, &tmap_a, fb, kb * k_elems, row);
#pragma unroll
              for (int kw = 0; kw < 3; kw++)
                tma_load_2d(smem_b0 + stage * kB + kw * Cfg::kStageBytesB, &tmap_w, fb, (3 * kh + kw) * Kdim + kb * k_elems, n0);
              if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
            }
          }
          continue;
        }
        for (int t = 0; t < taps; t++) {
#pragma unroll 1
          for (int sp = 0; sp < (SPLIT ? 3 : 1); sp++) {   // SPLIT: (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo)
            const int row = m0 + p.shift[t] + ((SPLIT && sp == 1) ? p.split_rows : 0);
            const int wcol0 = SPLIT ? (2 * t + (sp == 2 ? 1 : 0)) * Kdim : t * Kdim;
            for (int kb = 0; kb < k_blocks; kb++) {
              const uint32_t fb = bar_full + 8 * stage;
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              if (no_tma) { mbar_arrive(fb); }
              else {
                mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
                tma_load_2d(smem_a0 + stage * kStageBytesA, &tmap_a, fb, kb * k_elems, row);
                tma_load_2d(smem_b0 + stage * Cfg::kStageBytesB, &tmap_w, fb, wcol0 + kb * k_elems, n0);
              }
              if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        if (has_res) {   // residual [128 x 64] tiles as extra A operands; a stage carries two of them (A slot + B slot)
          constexpr int kResPerStage = (BLOCK_N >= 128) ? 2 : 1;
          for (int hp = 0; hp < kSplitPasses; hp++) {      // SPLIT: the hi tiles, then the lo tiles
            const int rrow = m0 + (SPLIT ? hp * p.split_rows : 0);
            for (int j = 0; j < BLOCK_N / 64 && n0 + j * 64 < Ndim; j += kResPerStage) {
              const uint32_t fb = bar_full + 8 * stage;
              const bool two = kResPerStage == 2 && (n0 + (j + 1) * 64 < Ndim);
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              mbar_arrive_expect_tx(fb, two ? 2 * kStageBytesA : kStageBytesA);
              tma_load_2d(smem_a0 + stage * kStageBytesA, &tmap_r, fb, n0 + j * 64, rrow);
              if (two) tma_load_2d(smem_b0 + stage * Cfg::kStageBytesB, &tmap_r, fb, n0 + (j + 1) * 64, rrow);
              if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===================================== MMA issuer
      constexpr uint32_t idesc = KIND == 1 ? make_idesc_tf32(BLOCK_M, BLOCK_N) : make_idesc_bf16(BLOCK_M, BLOCK_N);
      constexpr uint32_t idesc_res = make_idesc_bf16(BLOCK_M, 64);
      const uint64_t ident_desc = make_smem_desc_sw128(smem_base + off_ident);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_a0), bdesc0 = make_smem_desc_sw128(smem_b0);
      uint32_t stage = 0, phase = 0, tc = 0;
      const int n_tiles = p.n_tiles, Ndim = p.N;
      const bool has_res = p.has_res != 0, no_mma = (p.debug & 1) != 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tc++) {
        const int n0 = (tile % n_tiles) * BLOCK_N;
        const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
        mbar_wait(bar_tempty + 8 * b, bph ^ 1u);   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * BLOCK_N;
        // a partial last N tile issues a narrower MMA (N rounded up to 16): no tensor-pipe time for padding columns
        const int n_rem = Ndim - n0;
        const int n_eff = n_rem >= BLOCK_N ? BLOCK_N : ((n_rem + 15) & ~15);
        const uint32_t idesc_t = (idesc & ~(0x3fu << 17)) | ((uint32_t)(n_eff >> 3) << 17);
        if constexpr (TAP3) {
          const int groups = 3 * p.k_blocks;                  // (kh, kb) groups of three taps
          for (int gi = 0; gi < groups; gi++) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(stage * (kA >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (kB >> 4));
#pragma unroll
            for (int kw = 0; kw < 3; kw++)
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; k++)      // A rows shifted by kw: + kw * 128 bytes = + 8 * kw in the (address >> 4) field
                umma_bf16(tmem_d, adesc + 8 * kw + 2 * k, bdesc + (uint64_t)(kw * (Cfg::kStageBytesB >> 4)) + 2 * k, idesc_t,
                          (gi > 0 || kw > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_empty + 8 * stage);
            if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * b);
          continue;
        }
        for (int ki = 0; ki < k_iters; ki++) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          // descriptors advance by whole stages in the (address >> 4) field
          const uint64_t adesc = adesc0 + (uint64_t)(stage * (kStageBytesA >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (Cfg::kStageBytesB >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++) {  // +32 bytes per K step inside the swizzle row => +2 in the >>4 address field
            if (no_mma && (ki > 0 || k > 0)) continue;
            if constexpr (KIND == 1) umma_tf32(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * stage);             // frees the smem slot once these MMAs have read it
          if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
        }
        if (has_res) {
          if constexpr (BLOCK_N >= 64) {
            constexpr int kResPerStage = (BLOCK_N >= 128) ? 2 : 1;
            for (int hp = 0; hp < kSplitPasses; hp++)
            for (int j = 0; j < BLOCK_N / 64 && n0 + j * 64 < Ndim; j += kResPerStage) {
              const bool two = kResPerStage == 2 && (n0 + (j + 1) * 64 < Ndim);
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              const uint64_t adesc = adesc0 + (uint64_t)(stage * (kStageBytesA >> 4));
              const uint64_t adesc2 = bdesc0 + (uint64_t)(stage * (Cfg::kStageBytesB >> 4));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                umma_bf16(tmem_d + j * 64, adesc + 2 * k, ident_desc + 2 * k, idesc_res, 1u);   // D[:, 64j:64j+64] += R_tile * I
              if (two) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                  umma_bf16(tmem_d + (j + 1) * 64, adesc2 + 2 * k, ident_desc + 2 * k, idesc_res, 1u);
              }
              umma_commit(bar_empty + 8 * stage);
              if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
        umma_commit(bar_tfull + 8 * b);               // accumulator complete -> epilogue
      }
    }
  } else if (warp >= kEpiWarp0) {  // ========================= epilogue warps (8: two per TMEM lane quadrant, split by columns)
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int half = (warp - kEpiWarp0) >> 2;         // which half of the columns of a phase / chunk pair this warp handles
    const int et = threadIdx.x - kEpiWarp0 * 32;      // 0..255
    constexpr int CH = BLOCK_N < 32 ? BLOCK_N : 32;
    uint32_t tc = 0;
    [[maybe_unused]] uint32_t gphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tc++) {
      const int m0 = (tile / p.n_tiles) * BLOCK_M, n0 = (tile % p.n_tiles) * BLOCK_N;
      const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
      if constexpr (MODE == 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // previous tile's bias reads are done
      if (MODE == 0 || !p.warp_epi)
        for (int j = et; j < BLOCK_N; j += kEpiThreads) s_bias[j] = (p.bias != nullptr && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      if constexpr (MODE == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      const long long m = (long long)m0 + q * 32 + lane;
      bool zero_row = false;
      [[maybe_unused]] const __nv_bfloat16* up_row = nullptr;
      if (p.plane_h > 0) {
        unsigned int plane = (unsigned)(p.plane_h * p.plane_w);
        unsigned int rem = (unsigned int)((unsigned long long)m % plane);
        unsigned int y = rem / (unsigned)p.plane_w, x = rem - y * (unsigned)p.plane_w;
        zero_row = (y == 0) || (y == (unsigned)p.plane_h - 1) || (x == 0) || (x == (unsigned)p.plane_w - 1);
        if constexpr (UPS) {   // row of the coarser plane under this pixel: interior (y, x) -> ((y - 1) / 2 + 1, (x - 1) / 2 + 1)
          const long long img = m / plane;
          // rows past M (the tail of the last tile) belong to no image: they must not index the coarse plane (they would read past its end)
          up_row = (zero_row || m >= p.M) ? nullptr
                            : p.up + ((img * p.up_ph + (long long)(((y - 1) >> 1) + 1)) * p.up_pw + (((x - 1) >> 1) + 1)) * p.ldu + n0;
        }
      }
      if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);   // one poller per warp keeps the mbarrier unit free for the TMA / MMA threads
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * BLOCK_N;
      if constexpr (MODE == 0) {
#pragma unroll 1
        for (int c = half * CH; c < BLOCK_N; c += 2 * CH) {
          if (n0 + c >= p.N) break;                    // the rest of this warp's chunks lie beyond N
          uint32_t v[32];
          if constexpr (CH == 32) tmem_ld32(taddr + c, v); else tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          if (m < p.M) {
            const int ncols = (p.N - (n0 + c)) < CH ? (p.N - (n0 + c)) : CH;   // multiple of 8
            float f[CH];
#pragma unroll
            for (int j = 0; j < CH; j++) {
              f[j] = __uint_as_float(v[j]) + s_bias[c + j];
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
              if (zero_row) f[j] = 0.f;
            }
            if (p.d_f32) {
              float* dp = reinterpret_cast<float*>(p.D) + m * p.ldd + n0 + c;
#pragma unroll
              for (int j = 0; j < CH; j += 4)
                if (j < ncols) *reinterpret_cast<float4*>(dp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.D) + m * p.ldd + n0 + c;
#pragma unroll
              for (int j = 0; j < CH; j += 8) {
                if (j < ncols) {
                  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                  for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(f[j + 2 * e], f[j + 2 * e + 1]);
                  *reinterpret_cast<uint4*>(dp + j) = u;
                }
              }
            }
          }
        }
      } else if (p.warp_epi) {
        // Warp-private epilogue: every warp owns its 32 rows x (BLOCK_N / 2) columns, stages 64 columns at a time in its own 4 KB
        // swizzled buffer(s) and issues its own [32 x 64] TMA store -- no CTA-wide barrier anywhere, so the eight warps' TMEM
        // loads, converts and stores overlap freely.  (The shared-phase version below serialised two bar.syncs and a store
        // hand-off per phase: ~3.3 us per 128 x 256 tile, which bound every K <= 512 layer; profiles/r01_gemm_modes.md.)
        const int nbuf = p.phase_cols >> 6;             // private buffers per warp: 1 (deep K loops) or 2
        const uint32_t my_stg = off_staging + (uint32_t)(warp - kEpiWarp0) * (uint32_t)(nbuf * 4096);
        constexpr int kColsW = BLOCK_N >= 128 ? BLOCK_N / 2 : BLOCK_N;
        const int cbeg = BLOCK_N >= 128 ? half * kColsW : 0;
        const bool works = BLOCK_N >= 128 || half == 0;
        const bool rows_in = (long long)m0 + q * 32 < p.M;
        bool released = false;
        if (works) {
#pragma unroll 1
          for (int c = cbeg; c < cbeg + kColsW; c += 64) {
            if (n0 + c >= p.N) break;
            const uint32_t buf = nbuf == 2 ? (gphase & 1u) : 0u;
            if (lane == 0) { if (nbuf == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
            __syncwarp();
            uint32_t v[64];
            tmem_ld32(taddr + c, v);
            tmem_ld32(taddr + c + 32, v + 32);
            tmem_ld_wait();
            if (c + 64 >= cbeg + kColsW || n0 + c + 64 >= p.N) {   // last TMEM read of this tile: hand the accumulator back early
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
              released = true;
            }
            uint8_t* rowp = smem_al + my_stg + buf * 4096u + lane * 128;
            const int sw = lane & 7;
#pragma unroll
            for (int j = 0; j < 8; j++) {
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (p.bias != nullptr && n0 + c + 8 * j < p.N) {
                b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + 8 * j));
                b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + 8 * j + 4));
              }
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint4 o; __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; e++) {
                float a0 = __uint_as_float(v[8 * j + 2 * e]) + bb[2 * e];
                float a1 = __uint_as_float(v[8 * j + 2 * e + 1]) + bb[2 * e + 1];
                if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                if (zero_row) { a0 = 0.f; a1 = 0.f; }
                if (p.d_f16) { __half2 hh = __floats2half2_rn(a0, a1); ho[e] = *reinterpret_cast<__nv_bfloat162*>(&hh); }
                else ho[e] = __floats2bfloat162_rn(a0, a1);
              }
              *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = o;
            }
            fence_proxy_async();                          // generic-proxy smem writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0 && rows_in && !(p.debug & 4)) {
              tma_store_2d(&tmap_d, smem_base + my_stg + buf * 4096u, n0 + c, m0 + q * 32);   // rows >= M / cols >= N clipped by the TMA unit
              tma_store_commit();
            }
            gphase++;
          }
        }
        if (!released) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
        }
        continue;
      } else {
        const int r = q * 32 + lane;
        const int sw = r & 7;                           // SWIZZLE_128B: 16-byte chunk index ^= row & 7
        const bool relu_f = ACT == 0 && p.relu != 0;
        const int cols_per_warp = p.phase_cols >> 1;    // 32 (phase 64) or 64 (phase 128)
#pragma unroll 1
        for (int pcs = 0; pcs < BLOCK_N * kSplitPasses; pcs += p.phase_cols) {
          // SPLIT: every phase runs twice -- first the hi half bf16(v), then the lo half bf16(v - hi), stored split_rows further down
          const int pc = SPLIT ? (pcs / (2 * p.phase_cols)) * p.phase_cols : pcs;
          [[maybe_unused]] const bool lo_pass = SPLIT && ((pcs / p.phase_cols) & 1);
          if (n0 + pc >= p.N) break;
          const uint32_t buf_off = off_staging + (gphase & 1u) * (BLOCK_M * p.phase_cols * 2);
          if (et == 0) tma_store_wait_read<1>();        // the TMA store issued two phases ago (this buffer) has read its smem
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int c0 = pc + half * cols_per_warp;     // this warp's columns of the phase: [c0, c0 + cols_per_warp)
          if (n0 + c0 < p.N) {
            uint32_t v[64];
            tmem_ld32(taddr + c0, v);
            if (cols_per_warp == 64) tmem_ld32(taddr + c0 + 32, v + 32);
            tmem_ld_wait();
#pragma unroll
            for (int g2 = 0; g2 < 2; g2++) {
              if (g2 * 32 >= cols_per_warp) break;
              const int c = c0 + g2 * 32;
              // 64-column slot (16 KB, 128-byte rows) inside the phase buffer; this 32-column group is its low or high half
              uint8_t* rowp = smem_al + buf_off + ((c - pc) >> 6) * (BLOCK_M * 128) + r * 128;
              const int hs = ((c - pc) >> 5) & 1;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c + 8 * j);
                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c + 8 * j + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint4 o; __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
                [[maybe_unused]] uint4 ut = make_uint4(0, 0, 0, 0);
                if constexpr (UPS) { if (up_row != nullptr) ut = __ldg(reinterpret_cast<const uint4*>(up_row + c + 8 * j)); }
                [[maybe_unused]] const uint32_t* uw = &ut.x;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const float2 ab = __fadd2_rn(make_float2(__uint_as_float(v[g2 * 32 + 8 * j + 2 * e]), __uint_as_float(v[g2 * 32 + 8 * j + 2 * e + 1])),
                                               make_float2(bb[2 * e], bb[2 * e + 1]));      // one FADD2 (rounds each half like FADD)
                  float a0 = ab.x, a1 = ab.y;
                  if constexpr (UPS) { a0 += __uint_as_float(uw[e] << 16); a1 += __uint_as_float(uw[e] & 0xffff0000u); }
                  if constexpr (ACT == 1) {
                    const float2 gg = gelu_erf2(make_float2(a0, a1));
                    a0 = gg.x; a1 = gg.y;
                  }
                  if constexpr (!SPLIT && ACT == 0) {
                    // plain bf16 / fp16 output: ReLU rides the conversion (cvt.rn.relu.bf16x2.f32), border rows are zeroed on the packed word
                    uint32_t pk;
                    if (p.d_f16) {
                      if (relu_f) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                      const __half2 hh = __floats2half2_rn(a0, a1);
                      pk = *reinterpret_cast<const uint32_t*>(&hh);
                    } else if (relu_f) {
                      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a1), "f"(a0));
                    } else {
                      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(a1), "f"(a0));
                    }
                    (&o.x)[e] = zero_row ? 0u : pk;
                    continue;
                  }
                  if (relu_f) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                  if (zero_row) { a0 = 0.f; a1 = 0.f; }
                  if constexpr (SPLIT) {
                    const __nv_bfloat162 hi2 = __floats2bfloat162_rn(a0, a1);
                    if (lo_pass) {   // v - hi is exact in fp32 (hi holds the leading bits of v)
                      const float2 hf = __bfloat1622float2(hi2);
                      ho[e] = __floats2bfloat162_rn(__fsub_rn(a0, hf.x), __fsub_rn(a1, hf.y));
                    } else ho[e] = hi2;
                  } else if (p.d_f16) { __half2 hh = __floats2half2_rn(a0, a1); ho[e] = *reinterpret_cast<__nv_bfloat162*>(&hh); }
                  else ho[e] = __floats2bfloat162_rn(a0, a1);
                }
                *reinterpret_cast<uint4*>(rowp + (((hs * 4 + j) ^ sw) << 4)) = o;
              }
            }
          }
          fence_proxy_async();                          // generic-proxy smem writes -> visible to the TMA store
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
            const int srow = m0 + (lo_pass ? p.split_rows : 0);
            for (int c = pc; c < pc + p.phase_cols && n0 + c < p.N; c += 64)   // rows >= M / cols >= N are clipped by the TMA unit
              tma_store_2d(&tmap_d, smem_base + buf_off + ((c - pc) >> 6) * (BLOCK_M * 128), n0 + c, srow);
            tma_store_commit();
          }
          gphase++;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
    }
    if constexpr (MODE == 1) {
      if (p.warp_epi ? lane == 0 : et == 0) tma_store_wait_read<0>();   // smem must stay valid until the last stores have read it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

// ------------------------------------------------------------------------------------------ 2-CTA variant
// CTA pairs (cluster of 2, same TPC) compute 256 x BLOCK_N tiles with tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of
// A and HALF of the B tile, the leader CTA's elected thread issues one MMA for both SMs.  One instruction now carries 256 rows, so
// narrow-N layers (the single-thread issue rate is ~128 cycles per MMA whatever N is -- tools/gemm_sweep.py) do twice the work per
// issue slot, and every layer moves half the B bytes per SM.  bf16 operands, bf16 output (staged TMA stores), residual as identity MMA.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's barrier (local for the leader itself)
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

__host__ __device__ inline int gemm2_smem_bytes(int block_n, int num_stages, int phase_cols, int has_res) {
  int b = num_stages * (kStageBytesA + (block_n / 2) * BLOCK_K * 2) + 2 * BLOCK_M * phase_cols * 2;
  if (has_res) b += kIdentBytes / 2;
  return b + kCtrlBytes + 1024;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  constexpr int kHalfN = BLOCK_N / 2;
  constexpr int kStageBytesBh = kHalfN * BLOCK_K * 2;
  constexpr int kStageBytesCta = kStageBytesA + kStageBytesBh;
  constexpr int kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  const int kStages = p.num_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + kStages * kStageBytesA;
  const uint32_t off_staging = kStages * kStageBytesCta;
  const uint32_t staging_bytes = 2u * BLOCK_M * p.phase_cols * 2u;
  const uint32_t off_ident = off_staging + staging_bytes;
  const uint32_t off_ctrl = off_ident + (p.has_res ? kIdentBytes / 2 : 0);
  uint8_t* ctrl = smem_al + off_ctrl;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 8 * (2 * kMaxStages + 4));
  float* s_bias = reinterpret_cast<float*>(ctrl + 1024);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kMaxStages;
  const uint32_t bar_tfull = bar_empty + 8 * kMaxStages, bar_tempty = bar_tfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_w); tma_prefetch_desc(&tmap_d);
    if (p.has_res) tma_prefetch_desc(&tmap_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; b++) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 2 * (kEpiThreads / 32)); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(smem_u32(tmem_slot), kTmemCols);
  if (p.has_res && warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // this CTA's half of the 64x64 identity B operand: rows n = 32*rank + nl, K-major, SWIZZLE_128B
    uint8_t* ident = smem_al + off_ident;
    const int t = threadIdx.x - kEpiWarp0 * 32;
    for (int i = t; i < kIdentBytes / 2 / 16; i += 128) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (t < 32) {
      const int nl = t, k = 32 * (int)rank + nl, c = k >> 3;
      *reinterpret_cast<__nv_bfloat16*>(ident + nl * 128 + ((c ^ (nl & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(1.0f);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // peer barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;       // m_tiles counts 256-row tiles here
  const int k_iters = p.taps * p.k_blocks;

  if (warp == 0) {
    if (lane == 0) {  // ===================================== TMA producer (both CTAs: own A rows, own half of B)
      uint32_t stage = 0, phase = 0;
      const int n_tiles = p.n_tiles, k_blocks = p.k_blocks, taps = p.taps, Kdim = p.K, Ndim = p.N;
      const bool has_res = p.has_res != 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
        for (int t = 0; t < taps; t++) {
          const int row = m0 + p.shift[t];
          const int wcol0 = t * Kdim;
          for (int kb = 0; kb < k_blocks; kb++) {
            const uint32_t fb = bar_full + 8 * stage;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(fb, 2 * kStageBytesCta);     // bytes of BOTH CTAs land on the leader's barrier
            tma_load_2d_2sm(smem_a0 + stage * kStageBytesA, &tmap_a, fb, kb * BLOCK_K, row);
            tma_load_2d_2sm(smem_b0 + stage * kStageBytesBh, &tmap_w, fb, wcol0 + kb * BLOCK_K, n0 + (int)rank * kHalfN);
            if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
          }
        }
        if (has_res) {
          for (int j = 0; j < BLOCK_N / 64 && n0 + j * 64 < Ndim; j++) {
            const uint32_t fb = bar_full + 8 * stage;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(fb, 2 * kStageBytesA);
            tma_load_2d_2sm(smem_a0 + stage * kStageBytesA, &tmap_r, fb, n0 + j * 64, m0);
            if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ======================== MMA issuer (leader CTA only)
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N);
      constexpr uint32_t idesc_res = make_idesc_bf16(2 * BLOCK_M, 64);
      const uint64_t ident_desc = make_smem_desc_sw128(smem_base + off_ident);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_a0), bdesc0 = make_smem_desc_sw128(smem_b0);
      uint32_t stage = 0, phase = 0, tc = 0;
      const int n_tiles = p.n_tiles, Ndim = p.N;
      const bool has_res = p.has_res != 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, tc++) {
        const int n0 = (tile % n_tiles) * BLOCK_N;
        const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
        mbar_wait(bar_tempty + 8 * b, bph ^ 1u);   // both CTAs' epilogues have drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * BLOCK_N;
        for (int ki = 0; ki < k_iters; ki++) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(stage * (kStageBytesA >> 4));
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (kStageBytesBh >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++)
            umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(bar_empty + 8 * stage);     // frees this stage in both CTAs
          if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
        }
        if (has_res) {
          for (int j = 0; j < BLOCK_N / 64 && n0 + j * 64 < Ndim; j++) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint64_t adesc = adesc0 + (uint64_t)(stage * (kStageBytesA >> 4));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++)
              umma_bf16_2sm(tmem_d + j * 64, adesc + 2 * k, ident_desc + 2 * k, idesc_res, 1u);
            umma_commit_2sm(bar_empty + 8 * stage);
            if (++stage == (uint32_t)kStages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit_2sm(bar_tfull + 8 * b);           // accumulator complete -> epilogue warps of both CTAs
      }
    }
  } else if (warp >= kEpiWarp0) {  // ========================= epilogue warps: this CTA's 128 rows
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const int et = threadIdx.x - kEpiWarp0 * 32;
    uint32_t tc = 0, gphase = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, tc++) {
      const int m0 = (tile / p.n_tiles) * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = (tile % p.n_tiles) * BLOCK_N;
      const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
      for (int j = et; j < BLOCK_N; j += kEpiThreads) s_bias[j] = (p.bias != nullptr && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      const long long m = (long long)m0 + q * 32 + lane;
      bool zero_row = false;
      if (p.plane_h > 0) {
        unsigned int plane = (unsigned)(p.plane_h * p.plane_w);
        unsigned int rem = (unsigned int)((unsigned long long)m % plane);
        unsigned int y = rem / (unsigned)p.plane_w, x = rem - y * (unsigned)p.plane_w;
        zero_row = (y == 0) || (y == (unsigned)p.plane_h - 1) || (x == 0) || (x == (unsigned)p.plane_w - 1);
      }
      if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * BLOCK_N;
      const int r = q * 32 + lane;
      const int sw = r & 7;
      const int cols_per_warp = p.phase_cols >> 1;
#pragma unroll 1
      for (int pc = 0; pc < BLOCK_N; pc += p.phase_cols) {
        if (n0 + pc >= p.N) break;
        const uint32_t buf_off = off_staging + (gphase & 1u) * (BLOCK_M * p.phase_cols * 2);
        if (et == 0) tma_store_wait_read<1>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int c0 = pc + half * cols_per_warp;
        if (n0 + c0 < p.N) {
          uint32_t v[64];
          tmem_ld32(taddr + c0, v);
          if (cols_per_warp == 64) tmem_ld32(taddr + c0 + 32, v + 32);
          tmem_ld_wait();
#pragma unroll
          for (int g2 = 0; g2 < 2; g2++) {
            if (g2 * 32 >= cols_per_warp) break;
            const int c = c0 + g2 * 32;
            uint8_t* rowp = smem_al + buf_off + ((c - pc) >> 6) * (BLOCK_M * 128) + r * 128;
            const int hs = ((c - pc) >> 5) & 1;
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c + 8 * j + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint4 o; __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; e++) {
                float a0 = __uint_as_float(v[g2 * 32 + 8 * j + 2 * e]) + bb[2 * e];
                float a1 = __uint_as_float(v[g2 * 32 + 8 * j + 2 * e + 1]) + bb[2 * e + 1];
                if (p.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                if (zero_row) { a0 = 0.f; a1 = 0.f; }
                ho[e] = __floats2bfloat162_rn(a0, a1);
              }
              *reinterpret_cast<uint4*>(rowp + (((hs * 4 + j) ^ sw) << 4)) = o;
            }
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          for (int c = pc; c < pc + p.phase_cols && n0 + c < p.N; c += 64)
            tma_store_2d(&tmap_d, smem_base + buf_off + ((c - pc) >> 6) * (BLOCK_M * 128), n0 + c, m0);
          tma_store_commit();
        }
        gphase++;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_tempty + 8 * b);
    }
    if (et == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // nobody leaves while the pair may still signal into its shared memory
  if (warp == 2) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, kTmemCols); }
}


template <int BLOCK_N, int MODE, int KIND = 0, int UPS = 0, int TAP3 = 0, int SPLIT = 0, int ACT = 0>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& td, const CUtensorMap& tr, const GemmParams& p,
                       cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BLOCK_N, MODE, KIND, UPS, TAP3, SPLIT, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int smem = gemm_smem_bytes(BLOCK_N, MODE, p.num_stages, p.phase_cols, p.has_res) +
                   (TAP3 ? p.num_stages * (kTap3BytesA - kStageBytesA + 2 * BLOCK_N * BLOCK_K * 2) : 0);
  if (smem > 232448) return set_error(LVCB200_EINVAL, "gemm: internal shared-memory budget exceeded");
  int tiles = p.m_tiles * p.n_tiles;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  static const char* e_pdl = getenv("LVCB200_GEMM_PDL");
  if (e_pdl == nullptr || atoi(e_pdl) != 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    LVC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<BLOCK_N, MODE, KIND, UPS, TAP3, SPLIT, ACT>, ta, tw, td, tr, p));
  } else {
    gemm_bf16_tc_kernel<BLOCK_N, MODE, KIND, UPS, TAP3, SPLIT, ACT><<<grid, kGemmThreads, smem, s>>>(ta, tw, td, tr, p);
  }
  return check_launch("gemm_bf16_tc_kernel");
}

template <int BLOCK_N>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& td, const CUtensorMap& tr, const GemmParams& p,
                        cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int smem = gemm2_smem_bytes(BLOCK_N, p.num_stages, p.phase_cols, p.has_res);
  if (smem > 232448) return set_error(LVCB200_EINVAL, "gemm: internal shared-memory budget exceeded (2-CTA)");
  int tiles = p.m_tiles * p.n_tiles;
  int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  LVC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc2_kernel<BLOCK_N>, ta, tw, td, tr, p));
  return check_launch("gemm_bf16_tc2_kernel");
}

template <int BLOCK_N, int SPLIT = 0>
static int launch_gemm_mode(int mode, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& td, const CUtensorMap& tr,
                            const GemmParams& p, cudaStream_t s) {
  if constexpr (BLOCK_N >= 64) {
    if (mode == 1) return launch_gemm<BLOCK_N, 1, 0, 0, 0, SPLIT>(ta, tw, td, tr, p, s);
  }
  return launch_gemm<BLOCK_N, 0, 0, 0, 0, SPLIT>(ta, tw, td, tr, p, s);
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_gemm_bf16(const lvcb200_gemm_desc* d, void* stream) {
  LVC_REQUIRE(d, "gemm: NULL descriptor");
  LVC_REQUIRE(d->M >= 0 && d->N > 0 && d->K > 0 && d->taps >= 1 && d->taps <= 9, "gemm: bad extents");
  if (d->M == 0) return 0;
  LVC_REQUIRE(d->A && d->W && d->D, "gemm: NULL pointer");
  LVC_REQUIRE(d->N % 8 == 0, "gemm: N must be a multiple of 8");
  const bool tf32 = d->a_dtype == LVCB200_F32;   // fp32 operands consumed by the tensor core as TF32
  const bool split = d->split_rows != 0;         // strict mode: hi/lo bf16 pair operands, three-term product
  LVC_REQUIRE(d->relu >= 0 && d->relu <= 2 && !(d->relu == 2 && split), "gemm: relu must be 0, 1 (ReLU) or 2 (erf GELU; not with split operands)");
  LVC_REQUIRE(!split || (!tf32 && d->split_rows >= d->M && d->split_rows % BLOCK_M == 0 && d->split_rows < (1ll << 30) &&
                         !d->upsample_add && d->d_dtype != LVCB200_F16 && d->K % BLOCK_K == 0),
              "gemm: split mode: bf16 pairs, split_rows a multiple of 128 and >= M, K % 64 == 0, no upsample_add, bf16 (pair) or fp32 output");
  LVC_REQUIRE(d->a_dtype == LVCB200_BF16 || tf32, "gemm: a_dtype must be LVCB200_BF16 or LVCB200_F32");
  LVC_REQUIRE(!tf32 || (d->taps == 1 && !d->residual && d->d_dtype != LVCB200_F32 && d->N >= 64), "gemm: tf32 path: single tap, no residual, 16-bit output, N >= 64");
  LVC_REQUIRE(d->K % 8 == 0 && (d->taps == 1 || d->K % BLOCK_K == 0), "gemm: K must be a multiple of 8 (64 when taps > 1)");
  LVC_REQUIRE(d->lda % 8 == 0 && d->ldw % 8 == 0, "gemm: lda / ldw must be multiples of 8 elements");
  LVC_REQUIRE(((uintptr_t)d->A % 16) == 0 && ((uintptr_t)d->W % 16) == 0 && ((uintptr_t)d->D % 16) == 0, "gemm: pointers must be 16-byte aligned");
  LVC_REQUIRE(d->ldd % (d->d_dtype == LVCB200_F32 ? 4 : 8) == 0, "gemm: ldd alignment");
  LVC_REQUIRE(!d->residual || (d->ldr % 8 == 0 && ((uintptr_t)d->residual % 16) == 0), "gemm: residual alignment");
  LVC_REQUIRE(d->M < (1ll << 31) && d->M_rows < (1ll << 31), "gemm: M too large");
  int bn = d->N >= 256 ? 256 : (d->N > 64 ? 128 : (d->N > 32 ? 64 : (d->N > 16 ? 32 : 16)));
  if (d->N > 128 && d->N < 256) bn = 256;
  if (tf32) bn = d->N > 128 ? 256 : 128;
  {
    static const char* e_bn = getenv("LVCB200_GEMM_BN");
    if (e_bn && !tf32 && d->N >= atoi(e_bn) && atoi(e_bn) >= 64) bn = atoi(e_bn);
  }   // fp32 operands double the smem bytes per MMA: wide tiles keep B traffic per flop low
  // epilogue mode: fp32 output -> direct stores from registers; bf16 output -> smem staging + TMA stores
  const int mode = d->d_dtype == LVCB200_F32 ? 0 : 1;
  if ((mode == 1 || d->residual) && bn < 64) bn = 64;
  GemmParams p;
  p.bias = d->bias;
  p.D = d->D; p.ldd = d->ldd; p.d_f32 = d->d_dtype == LVCB200_F32; p.d_f16 = d->d_dtype == LVCB200_F16;
  p.k_elems = tf32 ? 32 : BLOCK_K;
  p.M = d->M; p.N = d->N; p.K = d->K; p.taps = d->taps;
  for (int i = 0; i < 9; i++) p.shift[i] = i < d->taps ? d->shift[i] : 0;
  p.relu = d->relu; p.plane_h = d->plane_h; p.plane_w = d->plane_w;
  p.m_tiles = (int)((d->M + BLOCK_M - 1) / BLOCK_M);
  p.n_tiles = (d->N + bn - 1) / bn;
  p.k_blocks = (d->K + p.k_elems - 1) / p.k_elems;
  p.has_res = d->residual ? 1 : 0;
  // smem split: deep operand pipeline for long K loops, shallow pipeline + wide staging when the epilogue dominates
  const int k_iters = p.taps * p.k_blocks * (split ? 3 : 1);
  p.split_rows = (int)d->split_rows;
  const int stage_bytes = kStageBytesA + bn * BLOCK_K * 2;
  const bool deep = k_iters >= 12 && !p.has_res;
  if (mode == 1) {
    p.phase_cols = deep ? 64 : (bn >= 128 ? 128 : 64);
    int budget = 232448 - (kCtrlBytes + 1024) - 2 * BLOCK_M * p.phase_cols * 2 - (p.has_res ? kIdentBytes : 0);
    p.num_stages = budget / stage_bytes;
  } else {
    p.phase_cols = 64;
    p.num_stages = (232448 - (kCtrlBytes + 1024) - (p.has_res ? kIdentBytes : 0)) / stage_bytes;
  }
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  if (!deep && p.num_stages > 4) p.num_stages = 4;
  p.debug = 0;
  p.up = (const __nv_bfloat16*)d->upsample_add; p.ldu = d->ldu; p.up_ph = d->up_plane_h; p.up_pw = d->up_plane_w;
  if (p.up) {
    LVC_REQUIRE(mode == 1 && !tf32 && !d->residual && d->N % 64 == 0 && d->plane_h > 2 && d->plane_w > 2 && d->d_dtype == LVCB200_BF16,
                "gemm: upsample_add needs a bf16 plane output, N % 64 == 0, no residual");
    LVC_REQUIRE(d->plane_h - 2 == 2 * (d->up_plane_h - 2) && d->plane_w - 2 == 2 * (d->up_plane_w - 2) && d->ldu % 8 == 0 &&
                ((uintptr_t)d->upsample_add % 16) == 0, "gemm: upsample_add: the coarse plane must be exactly half the size (fpn.py:131), 16-byte aligned");
    LVC_REQUIRE(bn == 256, "gemm: upsample_add is built for N >= 256 (FPN laterals)");
  }
  {
    static const char* e_we = getenv("LVCB200_GEMM_WEPI");
    p.warp_epi = (mode == 1 && !p.up && !split && e_we != nullptr && atoi(e_we) != 0 && ((uintptr_t)d->bias % 16) == 0) ? 1 : 0;   // experiment: same speed as the shared-phase epilogue (profiles/r01_gemm_modes.md)
  }
  {  // tuning overrides for experiments (tools/gemm_sweep.py); not used by the engine
    static const char* e_dbg = getenv("LVCB200_GEMM_DEBUG");
    if (e_dbg) p.debug = atoi(e_dbg);
    static const char* e_st = getenv("LVCB200_GEMM_STAGES");
    static const char* e_ph = getenv("LVCB200_GEMM_PHASE");
    if (e_ph && mode == 1) p.phase_cols = atoi(e_ph) > bn ? bn : atoi(e_ph);
    if (e_st) p.num_stages = atoi(e_st);
    while (p.num_stages > 2 && gemm_smem_bytes(bn, mode, p.num_stages, p.phase_cols, p.has_res) > 232448) p.num_stages--;
  }
  LVC_REQUIRE(p.num_stages >= 2, "gemm: internal: pipeline too shallow");
  CUtensorMap ta, tw, td, tr;
  int rc = make_tmap_2d(&ta, d->A, d->M_rows, d->K, d->lda, BLOCK_M, tf32);
  if (rc) return rc;
  rc = make_tmap_2d(&tw, d->W, d->N, (long long)d->taps * d->K * (split ? 2 : 1), d->ldw, bn, tf32);
  if (rc) return rc;
  td = ta; tr = ta;  // placeholders when unused (a valid map must still be passed by value)
  const long long rows_d = split ? d->split_rows + d->M : d->M;   // a pair matrix ends M rows into its lo half
  if (mode == 1 && (rc = make_tmap_2d(&td, d->D, rows_d, d->N, d->ldd, BLOCK_M))) return rc;
  CUtensorMap td32 = td;   // [32 x 64] store box of the warp-private epilogue
  if (mode == 1 && p.warp_epi && (rc = make_tmap_2d(&td32, d->D, d->M, d->N, d->ldd, 32))) return rc;
  if (p.has_res && (rc = make_tmap_2d(&tr, d->residual, rows_d, d->N, d->ldr, BLOCK_M))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (split) {   // per-layer launches of the SPLIT instantiations (no 2-CTA / TAP3 / chain variants in strict mode)
    switch (bn) {
      case 256: return launch_gemm_mode<256, 1>(mode, ta, tw, td, tr, p, s);
      case 128: return launch_gemm_mode<128, 1>(mode, ta, tw, td, tr, p, s);
      case 64: return launch_gemm_mode<64, 1>(mode, ta, tw, td, tr, p, s);
      case 32: return launch_gemm_mode<32, 1>(mode, ta, tw, td, tr, p, s);
      default: return launch_gemm_mode<16, 1>(mode, ta, tw, td, tr, p, s);
    }
  }
  {
    static const char* e_2 = getenv("LVCB200_GEMM_2CTA");
    const int two_cta = e_2 ? atoi(e_2) : 2;   // default: 2-CTA tiles for the big 3x3 convs only (same-box A/B: dense stack -1.7 % sustained)
    // 1: every eligible layer; 2 (default): long K loops on wide tiles -- the big 3x3 convs and the fc layers (per-shape table in
    // profiles/r01_gemm_modes.md: the pair handshake is amortised and half the B bytes per SM buy two more ring stages); 3: big 3x3 only
    const bool want2 = (p.up || d->relu == 2) ? false : two_cta == 1 || (two_cta == 3 && d->taps == 9 && bn == 256 && d->M >= 100000) ||
                       (two_cta == 2 && bn == 256 && !p.has_res && k_iters >= 16 && (d->M >= 100000 || (d->taps == 1 && d->M >= 8000)));
    if (want2 && mode == 1 && !tf32 && bn >= 64 && d->M >= 256) {
      GemmParams p2 = p;
      p2.m_tiles = (int)((d->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
      static const char* e_2p = getenv("LVCB200_GEMM_2CTA_PHASE");
      p2.phase_cols = e_2p ? atoi(e_2p) : (bn >= 128 ? 128 : 64);
      const int stage_bytes2 = kStageBytesA + (bn / 2) * BLOCK_K * 2;
      p2.num_stages = (232448 - (kCtrlBytes + 1024) - 2 * BLOCK_M * p2.phase_cols * 2 - (p2.has_res ? kIdentBytes / 2 : 0)) / stage_bytes2;
      if (p2.num_stages > kMaxStages) p2.num_stages = kMaxStages;
      CUtensorMap tw2;
      if ((rc = make_tmap_2d(&tw2, d->W, d->N, (long long)d->taps * d->K, d->ldw, bn / 2))) return rc;
      switch (bn) {
        case 256: return launch_gemm2<256>(ta, tw2, td, tr, p2, s);
        case 128: return launch_gemm2<128>(ta, tw2, td, tr, p2, s);
        default: return launch_gemm2<64>(ta, tw2, td, tr, p2, s);
      }
    }
  }
  if (p.up) return launch_gemm<256, 1, 0, 1>(ta, tw, td, tr, p, s);
  if (d->relu == 2) {   // GELU epilogue: its own instantiation (wide bf16 layers without residual: the ViT's fc1)
    LVC_REQUIRE(mode == 1 && bn == 256 && !tf32 && !p.has_res && !p.warp_epi, "gemm: relu = 2 (GELU) needs a bf16 output with N >= 129 and no residual");
    return launch_gemm<256, 1, 0, 0, 0, 0, 1>(ta, tw, td, tr, p, s);
  }
  {  // 3x3 conv on a 64-wide tile (res2 conv2): the three kw taps share one A tile (TAP3 instantiation)
    static const char* e_t3 = getenv("LVCB200_GEMM_TAP3");
    bool tap3 = (e_t3 == nullptr || atoi(e_t3) != 0) && d->taps == 9 && bn == 64 && mode == 1 && !tf32 && !p.has_res && !p.warp_epi &&
                d->M >= 4096 && d->M_rows >= 136;   // small planes gain nothing and the 136-row TMA box needs that many rows
    for (int kh = 0; kh < 3 && tap3; kh++)
      tap3 = d->shift[3 * kh + 1] == d->shift[3 * kh] + 1 && d->shift[3 * kh + 2] == d->shift[3 * kh] + 2;
    if (tap3) {
      GemmParams p3 = p;
      const int stage3 = kTap3BytesA + 3 * 64 * BLOCK_K * 2;
      p3.num_stages = (232448 - (kCtrlBytes + 1024) - 2 * BLOCK_M * p3.phase_cols * 2) / stage3;
      if (p3.num_stages > kMaxStages) p3.num_stages = kMaxStages;
      if (p3.num_stages >= 2) {
        CUtensorMap ta136;
        if ((rc = make_tmap_2d(&ta136, d->A, d->M_rows, d->K, d->lda, 136))) return rc;
        return launch_gemm<64, 1, 0, 0, 1>(ta136, tw, td, tr, p3, s);
      }
    }
  }
  const CUtensorMap& tdu = p.warp_epi ? td32 : td;
  if (tf32) return bn == 256 ? launch_gemm<256, 1, 1>(ta, tw, tdu, tr, p, s) : launch_gemm<128, 1, 1>(ta, tw, tdu, tr, p, s);
  switch (bn) {
    case 256: return launch_gemm_mode<256>(mode, ta, tw, tdu, tr, p, s);
    case 128: return launch_gemm_mode<128>(mode, ta, tw, tdu, tr, p, s);
    case 64: return launch_gemm_mode<64>(mode, ta, tw, tdu, tr, p, s);
    case 32: return launch_gemm_mode<32>(mode, ta, tw, tdu, tr, p, s);
    default: return launch_gemm_mode<16>(mode, ta, tw, tdu, tr, p, s);
  }
}
