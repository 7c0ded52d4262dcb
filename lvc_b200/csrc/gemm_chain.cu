// Layer-chain shift-GEMM: ONE persistent launch runs a whole sequence of dependent dense layers (a ResNet stage: 1x1 / 3x3 /
// 1x1 + residual bottlenecks) with TILE-granular dependencies instead of kernel boundaries.
//
// Why: at batch 8 the res3..res5 layers are 10-40 us kernels, and each launch pays a fixed ~6 us of pipeline fill, last-tile
// epilogue and tail imbalance (profiles/r01_gemm_layers_v0.md: 125 launches, 1.8 ms over the per-layer max(tensor, HBM) bounds).
// Here the CTAs walk ONE tile list covering all layers round-robin (tile g -> CTA g % grid).  A scheduler warp per CTA runs up to
// four tiles ahead of the pipeline: it reads the tile entry and the layer's scalar block, prefetches the layer's tensor maps, and
// waits on per-(layer, 128-row block) completion counters of the producing layers (the row block itself for a 1x1 conv, the blocks
// covering [m0 - PW - 1, m0 + 127 + PW + 1] for a 3x3) before it hands the tile to the TMA producer / MMA issuer / epilogue
// warps through a 4-slot shared-memory queue -- so no global-memory latency sits on the operand pipeline, which never drains
// between layers: a layer's tail tiles overlap the next layer's first tiles.
//
// Same arithmetic as gemm_tc.cu (identical MMA order per tile => bit-identical outputs): tcgen05.mma kind::f16 128 x BLOCK_N x 16,
// fp32 accumulators double-buffered in TMEM, TMA/mbarrier operand ring of 24 x 8 KB units whose block size follows the layer's
// BLOCK_N (8 blocks in flight at BLOCK_N = 64, 4 at 256), residual added on the tensor core (D += R * I with a 16 x 16 identity: four
// N = 16 MMAs per 64 residual columns), epilogue TMEM -> bias/ReLU/border-zero -> bf16 -> swizzled staging -> TMA store issued by a
// dedicated store thread, which also publishes the tile's completion counter once its stores have landed.
// Weight-stationary layers (conv3: 1x1, K <= 256, several N tiles): a CTA's tiles of the layer share one N tile, so its weight tile is
// loaded once into the top of the operand buffer and only A / residual blocks stream (same-box A/B: dense stack -1.4 % sustained).
#include <string.h>
#include <vector>

#include "tc_ptx.cuh"

namespace lvcb200 {

constexpr int kChainThreads = 384;       // warp 0 TMA producer, 1 MMA issuer, 2 TMEM alloc, 3 store/publish, 4-11 epilogue
constexpr int kChainEpiWarp0 = 4;
constexpr int kChainEpiWarps = 8;
// Operand ring: 24 units of 8 KB.  One K block takes A [128 x 64] bf16 = 2 units + B [BLOCK_N x 64] = BLOCK_N / 64 units, contiguous
// (a block that would wrap skips to unit 0).  Blocks are signalled by sequence number (full / empty barrier k % 8), so the
// geometry may change from tile to tile: BLOCK_N = 64 layers keep 8 blocks (192 KB) in flight, BLOCK_N = 256 layers 4.
constexpr int kUnits = 24;
constexpr int kUnitBytes = 8192;
constexpr int kRingBars = 8;
constexpr int kStagingBytes = 16384;     // one 64-column output phase: [128 x 64] bf16, 128-byte swizzled rows
constexpr int kOffStaging = kUnits * kUnitBytes;
constexpr int kOffIdent = kOffStaging + 2 * kStagingBytes;
constexpr int kIdent16Bytes = 2048;      // 16 rows x 128 B (16 x 16 bf16 identity in the first 32 bytes of each row, swizzled)
constexpr int kOffCtrl = kOffIdent + kIdent16Bytes;
constexpr int kChainSmem = kOffCtrl + 1024;   // = 232448 = 227 KB exactly
static_assert(kChainSmem == 232448, "chain kernel shared-memory budget");
// control KB: [0,256) mbarriers, [256,260) TMEM base, [512,1024) tile queue (4 slots x 32 words)
constexpr int kQueueDepth = 4;
constexpr int kQueueConsumers = 3 + kChainEpiWarps;   // producer, MMA issuer, store thread, 8 epilogue warps

// layer scalars as one 128-byte block (word indices): the scheduler warp copies it into the tile queue with one coalesced load
enum : int { W_BIAS_LO = 0, W_BIAS_HI, W_M, W_N, W_K, W_TAPS, W_SHIFT0, W_RELU = 15, W_PLANE_H, W_PLANE_W, W_BN, W_MTILES, W_NTILES, W_KBLOCKS,
             W_HAS_RES, W_CNT_OFF, W_DEP_A_OFF, W_DEP_A_NEED, W_DEP_A_MTILES, W_MIN_SHIFT, W_MAX_SHIFT, W_DEP_R_OFF, W_DEP_R_NEED, W_TILE = 31 };

struct alignas(128) ChainLayer {
  CUtensorMap ta, tw, td, tr;
  int w[32];
};

// tile list entry: layer (8 bits) | n tile (4 bits) | m tile (20 bits)
__host__ __device__ inline uint32_t pack_tile(int layer, int m, int n) { return ((uint32_t)layer << 24) | ((uint32_t)n << 20) | (uint32_t)m; }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// whole warp: lanes poll counters [lo, hi] of one layer until each has reached `need`
__device__ __forceinline__ void wait_rows_ready(const uint32_t* cnt, int lo, int hi, uint32_t need, int lane) {
  for (int base = lo; base <= hi; base += 32) {
    const int i = base + lane;
    const bool mine = i <= hi;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while (true) {
      const bool ok = !mine || ld_acquire_u32(cnt + i) >= need;
      if (__all_sync(0xffffffffu, ok)) break;
      if ((++spins & 0x3ffu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ull) __trap();   // a dependency bug must surface as a launch error, never as a hang
      }
    }
  }
}

// ring allocation rule shared by the producer and the MMA issuer: block of u units at `pos`, skipping to 0 instead of wrapping
struct RingPos {
  uint32_t pos = 0, k = 0, units = kUnits;   // `units`: ring size; the top of the buffer is lent to a resident weight tile in stationary mode
  __device__ __forceinline__ uint32_t place(uint32_t u, uint32_t& pad) {   // returns the first unit of the block
    pad = (pos + u > units) ? units - pos : 0u;
    const uint32_t start = pad ? 0u : pos;
    pos = start + u; if (pos == units) pos = 0;
    return start;
  }
};

__global__ void __launch_bounds__(kChainThreads, 1)
gemm_chain_kernel(const ChainLayer* __restrict__ layers, const uint32_t* __restrict__ tiles, int total_tiles,
                  uint32_t* __restrict__ counters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();               // SWIZZLE_128B operands need 1024-byte alignment; the budget has no slack
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffCtrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffCtrl + 256);
  volatile int* queue = reinterpret_cast<volatile int*>(smem + kOffCtrl + 512);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kRingBars;
  const uint32_t bar_tfull = bar_empty + 8 * kRingBars, bar_tempty = bar_tfull + 16;
  const uint32_t bar_sfull = bar_tempty + 16, bar_sempty = bar_sfull + 16;
  const uint32_t bar_qfull = bar_sempty + 16, bar_qempty = bar_qfull + 8 * kQueueDepth;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRingBars; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; b++) {
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, kChainEpiWarps);
      mbar_init(bar_sfull + 8 * b, kChainEpiWarps); mbar_init(bar_sempty + 8 * b, 1);
    }
    for (int s = 0; s < kQueueDepth; s++) { mbar_init(bar_qfull + 8 * s, 1); mbar_init(bar_qempty + 8 * s, kQueueConsumers); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  if (warp == 3) {
    // 16 x 16 bf16 identity, K-major, SWIZZLE_128B: row n = 128 bytes, 16-byte chunk c stored at position c ^ (n & 7)
    uint8_t* ident = smem + kOffIdent;
    for (int i = lane; i < kIdent16Bytes / 16; i += 32) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if (lane < 16) {
      const int n = lane, c = n >> 3;
      *reinterpret_cast<__nv_bfloat16*>(ident + n * 128 + ((c ^ (n & 7)) << 4) + (n & 7) * 2) = __float2bfloat16_rn(1.0f);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int my_tiles = cta < total_tiles ? (total_tiles - cta + G - 1) / G : 0;

  if (warp == 2) {   // ============================================== scheduler warp: tile list -> dependency wait -> tile queue
    // Everything with global-memory latency happens here, up to kQueueDepth tiles ahead of the pipeline: the tile entry, the layer's
    // scalar block (one coalesced 128-byte load), the completion counters of the producing layers, and a tensor-map prefetch.
    uint32_t e_next = my_tiles > 0 ? __ldg(tiles + cta) : 0u;
    for (int i = 0; i < my_tiles; i++) {
      const uint32_t e = e_next;
      if (i + 1 < my_tiles) e_next = __ldg(tiles + cta + (size_t)(i + 1) * G);
      const ChainLayer* ly = layers + (e >> 24);
      const int wv = __ldg(&ly->w[lane]);
      if (lane < 4) tma_prefetch_desc(reinterpret_cast<const CUtensorMap*>(ly) + lane);
      const int m = (int)(e & 0xfffffu), m0 = m * BLOCK_M;
      const int dep_a_off = __shfl_sync(0xffffffffu, wv, W_DEP_A_OFF), dep_r_off = __shfl_sync(0xffffffffu, wv, W_DEP_R_OFF);
      const int has_res = __shfl_sync(0xffffffffu, wv, W_HAS_RES) & 1;
      bool waited = false;
      if (dep_a_off >= 0) {
        const int min_shift = __shfl_sync(0xffffffffu, wv, W_MIN_SHIFT), max_shift = __shfl_sync(0xffffffffu, wv, W_MAX_SHIFT);
        const int dep_mt = __shfl_sync(0xffffffffu, wv, W_DEP_A_MTILES), need = __shfl_sync(0xffffffffu, wv, W_DEP_A_NEED);
        long long r_lo = (long long)m0 + min_shift, r_hi = (long long)m0 + BLOCK_M - 1 + max_shift;
        int lo = r_lo < 0 ? 0 : (int)(r_lo / BLOCK_M), hi = (int)(r_hi / BLOCK_M);
        if (hi > dep_mt - 1) hi = dep_mt - 1;
        if (lo <= hi) { wait_rows_ready(counters + dep_a_off, lo, hi, (uint32_t)need, lane); waited = true; }
      }
      if (has_res && dep_r_off >= 0) {
        const int need = __shfl_sync(0xffffffffu, wv, W_DEP_R_NEED);
        wait_rows_ready(counters + dep_r_off, m, m, (uint32_t)need, lane); waited = true;
      }
      const uint32_t slot = (uint32_t)i % kQueueDepth, use = (uint32_t)i / kQueueDepth;
      if (lane == 0) mbar_wait(bar_qempty + 8 * slot, (use & 1u) ^ 1u);
      __syncwarp();
      queue[slot * 32 + lane] = lane == W_TILE ? (int)e : wv;
      if (waited) fence_proxy_async_all();   // acquire (generic proxy) -> the producer's TMA loads (async proxy), ordered through qfull
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_qfull + 8 * slot);
    }
  } else if (warp == 0) {
    if (lane == 0) {   // ============================================ TMA producer
      RingPos rp;
      uint32_t rel_k = 0, used = 0, lens = 0;             // oldest unreleased block, units held by unreleased blocks, 4-bit lengths
      int cur_mode = -1;
      for (int i = 0; i < my_tiles; i++) {
        const uint32_t slot = (uint32_t)i % kQueueDepth, use = (uint32_t)i / kQueueDepth;
        mbar_wait(bar_qfull + 8 * slot, use & 1u);
        const volatile int* q = queue + slot * 32;
        const uint32_t e = (uint32_t)q[W_TILE];
        const ChainLayer* ly = layers + (e >> 24);
        const int bn = q[W_BN], k_blocks = q[W_KBLOCKS], taps = q[W_TAPS], Kdim = q[W_K], Ndim = q[W_N];
        const int hr = q[W_HAS_RES];
        const bool has_res = (hr & 1) != 0, stat = (hr & 2) != 0;
        int shift[9];
#pragma unroll
        for (int t = 0; t < 9; t++) shift[t] = q[W_SHIFT0 + t];
        mbar_arrive(bar_qempty + 8 * slot);
        fence_proxy_async_all();               // the scheduler's dependency acquire is ordered before this thread's TMA loads
        const int m0 = (int)(e & 0xfffffu) * BLOCK_M, n0 = (int)((e >> 20) & 15u) * bn;
        // Weight-stationary layers (1x1, K <= 256, BLOCK_N = 256, several N tiles: conv3 of a bottleneck): this CTA's tiles of the layer all
        // have the same N tile, so its weight tile (K x 256 bf16 <= 128 KB) is loaded ONCE into the top of the operand buffer and the
        // ring shrinks to the units below it; per tile only A (and the residual) stream through L2 -> SM, which is what paces these
        // layers (profiles/r01_gemm_timeline.md).  Entering / leaving the mode drains the ring.
        const int mode = stat ? (int)(e >> 20) : -1;               // (layer, N tile) of a stationary tile
        bool load_w = false;
        if (mode != cur_mode) {
          while (rel_k != rp.k) {                                   // drain: every outstanding block released (so the old weight tile is dead too)
            mbar_wait(bar_empty + 8 * (rel_k & 7u), (rel_k >> 3) & 1u);
            used -= (lens >> (4 * (rel_k & 7u))) & 15u;
            rel_k++;
          }
          rp.pos = 0;
          rp.units = stat ? (uint32_t)(kUnits - taps * k_blocks * (bn >> 6)) : (uint32_t)kUnits;
          cur_mode = mode;
          load_w = stat;
        }
        const uint32_t u = stat ? 2u : 2u + (uint32_t)(bn >> 6);
        const uint32_t u_res = stat ? 4u : u;
        const uint32_t w_base = smem_base + (uint32_t)(kUnits - taps * k_blocks * (bn >> 6)) * kUnitBytes;   // resident weight tile: one [bn x 64] block per K block
        const uint32_t stage_bytes = 16384u + ((stat && !load_w) ? 0u : (uint32_t)bn * 128u);
        const int per = bn >= 128 ? 2 : 1;
        const int n_res = has_res ? (((Ndim - n0 < bn ? Ndim - n0 : bn) / 64 + per - 1) / per) : 0;
        const int n_main = taps * k_blocks;
        const int iters = n_main + n_res;
        int tp = 0, kb = 0;
        for (int it = 0; it < iters; it++) {
          const uint32_t ub = it < n_main ? u : u_res;
          uint32_t pad;
          const uint32_t start = rp.place(ub, pad);
          const uint32_t need = pad + ub;
          while (used + need > rp.units || rp.k - rel_k >= (uint32_t)kRingBars) {   // oldest blocks release their units in order
            mbar_wait(bar_empty + 8 * (rel_k & 7u), (rel_k >> 3) & 1u);
            used -= (lens >> (4 * (rel_k & 7u))) & 15u;
            rel_k++;
          }
          lens = (lens & ~(15u << (4 * (rp.k & 7u)))) | (need << (4 * (rp.k & 7u)));
          used += need;
          const uint32_t fb = bar_full + 8 * (rp.k & 7u);
          const uint32_t dst = smem_base + start * kUnitBytes;
          rp.k++;
          if (it < n_main) {
            int sh = shift[0];
#pragma unroll
            for (int t = 1; t < 9; t++) sh = tp == t ? shift[t] : sh;
            mbar_arrive_expect_tx(fb, stage_bytes);
            tma_load_2d(dst, &ly->ta, fb, kb * BLOCK_K, m0 + sh);
            if (!stat) tma_load_2d(dst + 16384, &ly->tw, fb, tp * Kdim + kb * BLOCK_K, n0);
            else if (load_w) tma_load_2d(w_base + (uint32_t)it * ((uint32_t)bn * 128u), &ly->tw, fb, tp * Kdim + kb * BLOCK_K, n0);
            if (++kb == k_blocks) { kb = 0; tp++; }
          } else {   // residual [128 x 64] tiles as extra A operands, two per block when BLOCK_N >= 128
            const int j = (it - n_main) * per;
            const bool two = per == 2 && (n0 + (j + 1) * 64 < Ndim);
            mbar_arrive_expect_tx(fb, two ? 32768u : 16384u);
            tma_load_2d(dst, &ly->tr, fb, n0 + j * 64, m0);
            if (two) tma_load_2d(dst + 16384, &ly->tr, fb, n0 + (j + 1) * 64, m0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ============================================ MMA issuer
      const uint64_t ident_desc = make_smem_desc_sw128(smem_base + kOffIdent);
      constexpr uint32_t idesc_res = make_idesc_bf16(BLOCK_M, 16);
      RingPos rp;
      int cur_mode = -1;
      for (int i = 0; i < my_tiles; i++) {
        const uint32_t slot = (uint32_t)i % kQueueDepth, use = (uint32_t)i / kQueueDepth;
        mbar_wait(bar_qfull + 8 * slot, use & 1u);
        const volatile int* q = queue + slot * 32;
        const uint32_t e = (uint32_t)q[W_TILE];
        const int bn = q[W_BN], Ndim = q[W_N], k_blocks = q[W_KBLOCKS], k_iters = q[W_TAPS] * k_blocks;
        const int hr = q[W_HAS_RES];
        const bool has_res = (hr & 1) != 0, stat = (hr & 2) != 0;
        mbar_arrive(bar_qempty + 8 * slot);
        const int n0 = (int)((e >> 20) & 15u) * bn;
        const int mode = stat ? (int)(e >> 20) : -1;
        if (mode != cur_mode) {                 // same rule as the producer: ring restarts at unit 0 with the new size
          rp.pos = 0;
          rp.units = stat ? (uint32_t)(kUnits - k_iters * (bn >> 6)) : (uint32_t)kUnits;
          cur_mode = mode;
        }
        const uint32_t u = stat ? 2u : 2u + (uint32_t)(bn >> 6);
        const uint32_t u_res = stat ? 4u : u;
        const uint64_t wdesc0 = make_smem_desc_sw128(smem_base + (uint32_t)(kUnits - k_iters * (bn >> 6)) * kUnitBytes);
        const uint32_t b = (uint32_t)i & 1u, bph = ((uint32_t)i >> 1) & 1u;
        mbar_wait(bar_tempty + 8 * b, bph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * 256;
        const int n_rem = Ndim - n0;
        const int n_eff = n_rem >= bn ? bn : ((n_rem + 15) & ~15);
        const uint32_t idesc_t = make_idesc_bf16(BLOCK_M, n_eff);
        for (int ki = 0; ki < k_iters; ki++) {
          uint32_t pad;
          const uint32_t start = rp.place(u, pad);
          mbar_wait(bar_full + 8 * (rp.k & 7u), (rp.k >> 3) & 1u);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_base + start * kUnitBytes);
          const uint64_t bdesc = stat ? wdesc0 + (uint64_t)ki * (uint64_t)((bn * 128) >> 4) : adesc + (16384u >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar_empty + 8 * (rp.k & 7u));
          rp.k++;
        }
        if (has_res) {
          const int per = bn >= 128 ? 2 : 1;
          for (int j = 0; j < bn / 64 && n0 + j * 64 < Ndim; j += per) {
            const bool two = per == 2 && (n0 + (j + 1) * 64 < Ndim);
            uint32_t pad;
            const uint32_t start = rp.place(u_res, pad);
            mbar_wait(bar_full + 8 * (rp.k & 7u), (rp.k >> 3) & 1u);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc_sw128(smem_base + start * kUnitBytes);
            const uint64_t adesc2 = adesc + (16384u >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++)   // D[:, 64j + 16k : +16] += R[:, 16k : 16k + 16] * I16
              umma_bf16(tmem_d + j * 64 + 16 * k, adesc + 2 * k, ident_desc, idesc_res, 1u);
            if (two) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                umma_bf16(tmem_d + (j + 1) * 64 + 16 * k, adesc2 + 2 * k, ident_desc, idesc_res, 1u);
            }
            umma_commit(bar_empty + 8 * (rp.k & 7u));
            rp.k++;
          }
        }
        umma_commit(bar_tfull + 8 * b);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // ============================================ store thread: staging -> TMA store, publish tile completion
      uint32_t gphase = 0;
      for (int i = 0; i < my_tiles; i++) {
        const uint32_t slot = (uint32_t)i % kQueueDepth, use = (uint32_t)i / kQueueDepth;
        mbar_wait(bar_qfull + 8 * slot, use & 1u);
        const volatile int* q = queue + slot * 32;
        const uint32_t e = (uint32_t)q[W_TILE];
        const int bn = q[W_BN], Ndim = q[W_N], cnt_off = q[W_CNT_OFF];
        mbar_arrive(bar_qempty + 8 * slot);
        const ChainLayer* ly = layers + (e >> 24);
        const int m = (int)(e & 0xfffffu), m0 = m * BLOCK_M, n0 = (int)((e >> 20) & 15u) * bn;
        for (int pc = 0; pc < bn && n0 + pc < Ndim; pc += 64, gphase++) {
          const uint32_t buf = gphase & 1u;
          mbar_wait(bar_sfull + 8 * buf, (gphase >> 1) & 1u);
          tma_store_2d(&ly->td, smem_base + kOffStaging + buf * kStagingBytes, n0 + pc, m0);   // rows >= M clipped by the TMA unit
          tma_store_commit();
          tma_store_wait_read<0>();                    // the staging buffer has been read: hand it back to the epilogue warps
          mbar_arrive(bar_sempty + 8 * buf);
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this tile's stores have landed
        fence_proxy_async_all();
        red_release_add_u32(counters + cnt_off + m, 1u);
      }
    }
  } else if (warp >= kChainEpiWarp0) {   // ============================ epilogue warps (two per TMEM lane quadrant)
    const int q4 = warp & 3;
    const int half = (warp - kChainEpiWarp0) >> 2;
    const int r = q4 * 32 + lane;
    const int sw = r & 7;
    uint32_t gphase = 0;
    for (int i = 0; i < my_tiles; i++) {
      const uint32_t slot = (uint32_t)i % kQueueDepth, use = (uint32_t)i / kQueueDepth;
      if (lane == 0) mbar_wait(bar_qfull + 8 * slot, use & 1u);
      __syncwarp();
      const volatile int* q = queue + slot * 32;
      const uint32_t e = (uint32_t)q[W_TILE];
      const int bn = q[W_BN], Ndim = q[W_N];
      const int relu = q[W_RELU], plane_h = q[W_PLANE_H], plane_w = q[W_PLANE_W];
      const float* __restrict__ bias = reinterpret_cast<const float*>(((unsigned long long)(uint32_t)q[W_BIAS_HI] << 32) | (uint32_t)q[W_BIAS_LO]);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_qempty + 8 * slot);
      const int m0 = (int)(e & 0xfffffu) * BLOCK_M, n0 = (int)((e >> 20) & 15u) * bn;
      const uint32_t b = (uint32_t)i & 1u, bph = ((uint32_t)i >> 1) & 1u;
      const long long mrow = (long long)m0 + r;
      bool zero_row = false;
      if (plane_h > 0) {
        unsigned int plane = (unsigned)(plane_h * plane_w);
        unsigned int rem = (unsigned int)((unsigned long long)mrow % plane);
        unsigned int y = rem / (unsigned)plane_w, x = rem - y * (unsigned)plane_w;
        zero_row = (y == 0) || (y == (unsigned)plane_h - 1) || (x == 0) || (x == (unsigned)plane_w - 1);
      }
      if (lane == 0) mbar_wait(bar_tfull + 8 * b, bph);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + b * 256;
#pragma unroll 1
      for (int pc = 0; pc < bn && n0 + pc < Ndim; pc += 64, gphase++) {
        const uint32_t buf = gphase & 1u;
        if (lane == 0) mbar_wait(bar_sempty + 8 * buf, ((gphase >> 1) & 1u) ^ 1u);
        __syncwarp();
        const int c = pc + half * 32;
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        uint8_t* rowp = smem + kOffStaging + buf * kStagingBytes + r * 128;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (bias != nullptr) {
            b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + 8 * j));
            b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + 8 * j + 4));
          }
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint4 o;
#pragma unroll
          for (int e2 = 0; e2 < 4; e2++) {
            const float2 ab = __fadd2_rn(make_float2(__uint_as_float(v[8 * j + 2 * e2]), __uint_as_float(v[8 * j + 2 * e2 + 1])),
                                         make_float2(bb[2 * e2], bb[2 * e2 + 1]));      // one FADD2 (rounds each half like FADD)
            float a0 = ab.x, a1 = ab.y;
            if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
            (&o.x)[e2] = zero_row ? 0u : *reinterpret_cast<const uint32_t*>(&h2);      // border rows zeroed on the packed word
          }
          *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ sw) << 4)) = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sfull + 8 * buf);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static inline size_t chain_counter_bytes(int total_m_tiles) { return align_up((size_t)total_m_tiles * 4, 256); }

static int chain_validate(const lvcb200_gemm_desc* descs, int n, int* total_m_tiles) {
  LVC_REQUIRE(descs && n >= 1 && n <= 255, "gemm_chain: need 1..255 layer descriptors");
  int tm = 0;
  for (int i = 0; i < n; i++) {
    const lvcb200_gemm_desc* d = descs + i;
    LVC_REQUIRE(d->a_dtype == LVCB200_BF16 && d->d_dtype == LVCB200_BF16, "gemm_chain: bf16 operands and bf16 outputs only");
    LVC_REQUIRE(d->M >= 1 && d->M < (1ll << 31) && d->M_rows < (1ll << 31), "gemm_chain: bad M");
    LVC_REQUIRE(d->N >= 64 && d->N % 64 == 0 && d->K >= 64 && d->K % 64 == 0, "gemm_chain: N and K must be multiples of 64");
    LVC_REQUIRE(d->taps >= 1 && d->taps <= 9, "gemm_chain: taps");
    LVC_REQUIRE(d->split_rows == 0, "gemm_chain: split (strict-mode) layers run as per-layer launches");
    LVC_REQUIRE(d->relu == 0 || d->relu == 1, "gemm_chain: activation must be none or ReLU");
    LVC_REQUIRE(d->A && d->W && d->D, "gemm_chain: NULL pointer");
    LVC_REQUIRE(((uintptr_t)d->bias % 16) == 0, "gemm_chain: bias must be 16-byte aligned");
    LVC_REQUIRE(d->lda % 8 == 0 && d->ldw % 8 == 0 && d->ldd % 8 == 0 && (!d->residual || d->ldr % 8 == 0), "gemm_chain: leading dimensions must be multiples of 8");
    LVC_REQUIRE(((uintptr_t)d->A % 16) == 0 && ((uintptr_t)d->W % 16) == 0 && ((uintptr_t)d->D % 16) == 0 && ((uintptr_t)d->residual % 16) == 0, "gemm_chain: pointers must be 16-byte aligned");
    for (int j = 0; j < i; j++) {   // buffers are identical (a dependency) or disjoint; no write-after-read / write-after-write inside a chain
      LVC_REQUIRE(d->D != descs[j].D && d->D != descs[j].A && d->D != descs[j].residual, "gemm_chain: a layer may not overwrite a buffer an earlier layer of the chain reads or writes");
    }
    LVC_REQUIRE(d->D != d->A && d->D != d->residual, "gemm_chain: in-place layers are not supported");
    tm += (int)((d->M + BLOCK_M - 1) / BLOCK_M);
  }
  *total_m_tiles = tm;
  return 0;
}

}  // namespace lvcb200

using namespace lvcb200;

static int chain_sms() {
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

static long long chain_total_tiles(const lvcb200_gemm_desc* descs, int n) {
  long long t = 0;
  for (int i = 0; i < n; i++) {
    const int bn = descs[i].N >= 256 ? 256 : (descs[i].N > 64 ? (descs[i].N > 128 ? 256 : 128) : 64);
    t += ((descs[i].M + BLOCK_M - 1) / BLOCK_M) * ((descs[i].N + bn - 1) / bn);
  }
  return t;
}

extern "C" size_t lvcb200_gemm_chain_workspace(const lvcb200_gemm_desc* descs, int n) {
  int tm = 0;
  if (chain_validate(descs, n, &tm)) return 0;
  return chain_counter_bytes(tm) + (size_t)n * sizeof(ChainLayer) + align_up((size_t)chain_total_tiles(descs, n) * 4, 256);
}

extern "C" int lvcb200_gemm_chain_plan(const lvcb200_gemm_desc* descs, int n, void* workspace, size_t workspace_bytes,
                                       lvcb200_chain_plan* plan) {
  int tm = 0;
  int rc = chain_validate(descs, n, &tm);
  if (rc) return rc;
  LVC_REQUIRE(workspace && plan, "gemm_chain_plan: NULL pointer");
  LVC_REQUIRE(n <= 255, "gemm_chain_plan: at most 255 layers per chain");
  LVC_REQUIRE(((uintptr_t)workspace % 256) == 0, "gemm_chain_plan: workspace must be 256-byte aligned");
  const size_t cbytes = chain_counter_bytes(tm);
  const long long total = chain_total_tiles(descs, n);
  LVC_REQUIRE(total < (1ll << 30), "gemm_chain_plan: too many tiles");
  if (workspace_bytes < cbytes + (size_t)n * sizeof(ChainLayer) + align_up((size_t)total * 4, 256))
    return set_error(LVCB200_EWORKSPACE, "gemm_chain_plan: workspace too small");
  const int sms = chain_sms();
  std::vector<ChainLayer> tab((size_t)n);
  std::vector<int> dep_a(n, -1), dep_r(n, -1);
  int cnt_off = 0;
  bool same_m = true;
  for (int i = 0; i < n; i++) {
    const lvcb200_gemm_desc* d = descs + i;
    ChainLayer& L = tab[i];
    memset(&L, 0, sizeof(L));
    const int bn = d->N >= 256 ? 256 : (d->N > 64 ? (d->N > 128 ? 256 : 128) : 64);
    const unsigned long long bp = (unsigned long long)(uintptr_t)d->bias;
    L.w[W_BIAS_LO] = (int)(uint32_t)(bp & 0xffffffffull); L.w[W_BIAS_HI] = (int)(uint32_t)(bp >> 32);
    L.w[W_M] = (int)d->M; L.w[W_N] = d->N; L.w[W_K] = d->K; L.w[W_TAPS] = d->taps;
    int mn = 0, mx = 0;
    for (int t = 0; t < 9; t++) {
      L.w[W_SHIFT0 + t] = t < d->taps ? d->shift[t] : 0;
      if (t < d->taps) { if (d->shift[t] < mn) mn = d->shift[t]; if (d->shift[t] > mx) mx = d->shift[t]; }
    }
    L.w[W_MIN_SHIFT] = mn; L.w[W_MAX_SHIFT] = mx;
    L.w[W_RELU] = d->relu; L.w[W_PLANE_H] = d->plane_h; L.w[W_PLANE_W] = d->plane_w;
    L.w[W_BN] = bn;
    L.w[W_MTILES] = (int)((d->M + BLOCK_M - 1) / BLOCK_M);
    L.w[W_NTILES] = (d->N + bn - 1) / bn;
    LVC_REQUIRE(L.w[W_MTILES] < (1 << 20) && L.w[W_NTILES] <= 15, "gemm_chain_plan: layer too large for the tile list encoding");
    L.w[W_KBLOCKS] = d->K / BLOCK_K;
    L.w[W_HAS_RES] = d->residual ? 1 : 0;
    {  // bit 1: weight-stationary (see the producer): 1x1, K <= 256, BLOCK_N = 256, >= 2 N tiles, and every CTA keeps its N tile
      static const char* e_ws = getenv("LVCB200_CHAIN_WSTAT");
      static const char* e_ord0 = getenv("LVCB200_CHAIN_ORDER");
      const int grid0 = (int)(total < sms ? total : sms);
      const bool layer_order = !(e_ord0 && atoi(e_ord0) != 0);
      const int w_units = d->taps * L.w[W_KBLOCKS] * (bn >> 6);        // resident weight tile, in 8 KB units; >= 8 units stay for the ring
      const int ws_mode = e_ws ? atoi(e_ws) : 1;                        // 1: conv3-type layers only (several N tiles); 2: every layer whose tile fits
      if (ws_mode != 0 && layer_order && w_units <= 16 && d->N % bn == 0 && grid0 % L.w[W_NTILES] == 0 &&
          (ws_mode == 2 || (d->taps == 1 && bn == 256 && L.w[W_NTILES] >= 2)))
        L.w[W_HAS_RES] |= 2;
    }
    L.w[W_CNT_OFF] = cnt_off;
    L.w[W_DEP_A_OFF] = -1; L.w[W_DEP_R_OFF] = -1;
    if (L.w[W_MTILES] != tab[0].w[W_MTILES]) same_m = false;
    for (int j = i - 1; j >= 0; j--) {
      if (L.w[W_DEP_A_OFF] < 0 && descs[j].D == d->A) {
        LVC_REQUIRE(descs[j].ldd == d->lda && descs[j].N >= d->K, "gemm_chain_plan: a layer reads an earlier output with a different geometry");
        L.w[W_DEP_A_OFF] = tab[j].w[W_CNT_OFF]; L.w[W_DEP_A_NEED] = tab[j].w[W_NTILES]; L.w[W_DEP_A_MTILES] = tab[j].w[W_MTILES];
        dep_a[i] = j;
      }
      if (d->residual && L.w[W_DEP_R_OFF] < 0 && descs[j].D == d->residual) {
        LVC_REQUIRE(descs[j].ldd == d->ldr && descs[j].N >= d->N && descs[j].M >= d->M, "gemm_chain_plan: residual produced with a different geometry");
        L.w[W_DEP_R_OFF] = tab[j].w[W_CNT_OFF]; L.w[W_DEP_R_NEED] = tab[j].w[W_NTILES];
        dep_r[i] = j;
      }
    }
    if ((rc = make_tmap_2d(&L.ta, d->A, d->M_rows, d->K, d->lda, BLOCK_M))) return rc;
    if ((rc = make_tmap_2d(&L.tw, d->W, d->N, (long long)d->taps * d->K, d->ldw, bn))) return rc;
    if ((rc = make_tmap_2d(&L.td, d->D, d->M, d->N, d->ldd, BLOCK_M))) return rc;
    L.tr = L.ta;
    if (d->residual && (rc = make_tmap_2d(&L.tr, d->residual, d->M, d->N, d->ldr, BLOCK_M))) return rc;
    cnt_off += L.w[W_MTILES];
  }
  // Tile order.  Every tile's dependencies must precede it in the list (CTAs take tiles round-robin and spin on unfinished
  // dependencies).  Default: layer after layer (all CTAs then stream the same weight tiles at the same time, which the L2 serves
  // as one broadcast).  LVCB200_CHAIN_ORDER=1 selects a skewed wavefront over the row blocks instead (layer L handles row block m at
  // step m + lag[L], lag[L] = lag[producer] + halo + slack) so that a layer's output is consumed out of L2 a few hundred tiles
  // after it is written; measured slower on B200 (res4: 2.55 ms vs 1.85 ms) because concurrent CTAs then pull DIFFERENT weight
  // tiles through the L2 -> SM fabric -- kept as an experiment, see DESIGN.md.
  static const char* e_ord = getenv("LVCB200_CHAIN_ORDER");
  static const char* e_slack = getenv("LVCB200_CHAIN_SLACK");
  static const char* e_l2 = getenv("LVCB200_CHAIN_L2_MB");
  const int order = e_ord ? atoi(e_ord) : 0;
  const bool wavefront = same_m && order == 1;
  const double slack_rounds = e_slack ? atof(e_slack) : 3.0;
  const double l2_budget = (e_l2 ? atof(e_l2) : 72.0) * 1e6;
  std::vector<uint32_t> tiles;
  tiles.reserve((size_t)total);
  // cumulative halo (in 128-row blocks) of each layer behind its producers: the skew that keeps a blocked order dependency-safe
  std::vector<int> lagc(n, 0);
  int max_lagc = 0;
  for (int i = 0; i < n; i++) {
    int l = 0;
    if (dep_a[i] >= 0) {
      const int halo = tab[i].w[W_MAX_SHIFT] > 0 ? (tab[i].w[W_MAX_SHIFT] + BLOCK_M - 1) / BLOCK_M + 1 : 0;
      l = lagc[dep_a[i]] + halo;
    }
    if (dep_r[i] >= 0 && lagc[dep_r[i]] > l) l = lagc[dep_r[i]];
    lagc[i] = l;
    if (l > max_lagc) max_lagc = l;
  }
  // working set of the chain per 128-row block: input + output of the widest layer, twice (a block's input and output tensors are
  // alive together), in bf16
  int widest = 0;
  for (int i = 0; i < n; i++) if (tab[i].w[W_K] + tab[i].w[W_N] > widest) widest = tab[i].w[W_K] + tab[i].w[W_N];
  const double ws_per_block = (double)BLOCK_M * 2.0 * widest * 2.0;
  const int mt0 = tab[0].w[W_MTILES];
  int chunks = same_m ? (int)((ws_per_block * mt0 + l2_budget - 1) / l2_budget) : 1;
  if (chunks < 1) chunks = 1;
  if (order == 0 || !same_m || chunks == 1) {
    if (!wavefront)
      for (int i = 0; i < n; i++)
        for (int m = 0; m < tab[i].w[W_MTILES]; m++)
          for (int nt = 0; nt < tab[i].w[W_NTILES]; nt++) tiles.push_back(pack_tile(i, m, nt));
  } else if (order == 2) {
    // LVCB200_CHAIN_ORDER=2, L2-blocked layer order (experiment): the rows are cut into `chunks` bands whose working set fits the L2
    // budget; band c runs through ALL layers before band c + 1 starts, so every inter-layer read is served by L2, while inside a
    // (band, layer) the CTAs still stream the same weight tiles together.  Layer i's band boundaries are shifted up by its
    // cumulative halo lagc[i] so that the rows a 3x3 needs from the band above were produced in this band.  Measured slower: a
    // (band, layer) segment of 94-217 tiles is about one round of the 148 CTAs, so every segment waits for the full latency of the
    // previous one (res4 4.35 ms vs 1.87 ms, res3 0.81 vs 0.64 ms at a 72 MB budget; res2 0.91 vs 0.97 ms at 120 MB).
    const int Mc = (mt0 + chunks - 1) / chunks;
    const int last = (mt0 - 1 + max_lagc) / Mc;
    for (int c = 0; c <= last; c++)
      for (int i = 0; i < n; i++) {
        int lo = c * Mc - lagc[i], hi = (c + 1) * Mc - lagc[i];
        if (lo < 0) lo = 0;
        if (hi > mt0) hi = mt0;
        for (int m = lo; m < hi; m++)
          for (int nt = 0; nt < tab[i].w[W_NTILES]; nt++) tiles.push_back(pack_tile(i, m, nt));
      }
  }
  if (wavefront) {
    int per_step = 0;
    for (int i = 0; i < n; i++) per_step += tab[i].w[W_NTILES];
    const int slack = (int)((slack_rounds * sms + per_step - 1) / per_step) + 1;
    std::vector<int> lag(n, 0);
    int max_lag = 0;
    for (int i = 0; i < n; i++) {
      int l = 0;
      if (dep_a[i] >= 0) {
        const int halo = (tab[i].w[W_MAX_SHIFT] + BLOCK_M - 1) / BLOCK_M + (tab[i].w[W_MAX_SHIFT] > 0 ? 1 : 0);
        l = lag[dep_a[i]] + halo + slack;
      }
      if (dep_r[i] >= 0 && lag[dep_r[i]] + slack > l) l = lag[dep_r[i]] + slack;
      lag[i] = l;
      if (l > max_lag) max_lag = l;
    }
    for (int step = 0; step < mt0 + max_lag; step++)
      for (int i = 0; i < n; i++) {
        const int m = step - lag[i];
        if (m < 0 || m >= mt0) continue;
        for (int nt = 0; nt < tab[i].w[W_NTILES]; nt++) tiles.push_back(pack_tile(i, m, nt));
      }
  }
  if ((long long)tiles.size() != total) return set_error(LVCB200_EINVAL, "gemm_chain_plan: internal: tile list size mismatch");
  uint8_t* wsb = (uint8_t*)workspace;
  LVC_CUDA(cudaMemcpy(wsb + cbytes, tab.data(), (size_t)n * sizeof(ChainLayer), cudaMemcpyHostToDevice));
  LVC_CUDA(cudaMemcpy(wsb + cbytes + (size_t)n * sizeof(ChainLayer), tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice));
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
    attr_set = true;
  }
  plan->workspace = workspace;
  plan->n_layers = n;
  plan->counter_bytes = (int64_t)cbytes;
  plan->grid = (int)(total < sms ? total : sms);   // every CTA must be resident: dependencies are resolved by spinning
  plan->total_tiles = (int64_t)total;
  return 0;
}

extern "C" int lvcb200_gemm_chain_run(const lvcb200_chain_plan* plan, void* stream) {
  LVC_REQUIRE(plan && plan->workspace && plan->n_layers >= 1 && plan->grid >= 1, "gemm_chain_run: bad plan");
  cudaStream_t s = (cudaStream_t)stream;
  LVC_CUDA(cudaMemsetAsync(plan->workspace, 0, (size_t)plan->counter_bytes, s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = kChainSmem;
  cfg.stream = s;
  const uint8_t* wsb = (const uint8_t*)plan->workspace;
  const ChainLayer* layers = reinterpret_cast<const ChainLayer*>(wsb + plan->counter_bytes);
  const uint32_t* tiles = reinterpret_cast<const uint32_t*>(wsb + plan->counter_bytes + (size_t)plan->n_layers * sizeof(ChainLayer));
  uint32_t* counters = reinterpret_cast<uint32_t*>(plan->workspace);
  LVC_CUDA(cudaLaunchKernelEx(&cfg, gemm_chain_kernel, layers, tiles, (int)plan->total_tiles, counters));
  return check_launch("gemm_chain_kernel");
}
