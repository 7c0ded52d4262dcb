// Layer-chain shift-GEMM: ONE persistent launch runs a whole sequence of dependent dense layers (a ResNet stage: 1x1 / 3x3 /
// 1x1 + residual bottlenecks) with TILE-granular dependencies instead of kernel boundaries.
//
// Why: at batch 8 the res3..res5 layers are 10-40 us kernels, and each launch pays a fixed ~6 us of pipeline fill, last-tile
// epilogue and tail imbalance (profiles/r02_gemm_layers_v0.md: 125 launches, 1.8 ms over the per-layer max(tensor, HBM) bounds).
// Here every CTA walks the concatenated tile list of all layers round-robin (global tile g -> CTA g % grid); before the TMA
// producer loads the A rows of a tile it waits on per-(layer, 128-row block) completion counters of the producing layer (the
// row block itself for a 1x1 conv, the blocks covering [m0 - PW - 1, m0 + 127 + PW + 1] for a 3x3), so the operand pipeline,
// the tensor core and the epilogue never drain between layers and a layer's tail tiles overlap the next layer's first tiles.
//
// Same arithmetic as gemm_tc.cu (identical MMA order per tile => bit-identical outputs): tcgen05.mma kind::f16 128 x BLOCK_N x 16,
// fp32 accumulators double-buffered in TMEM, 4-slot TMA/mbarrier operand ring, residual added on the tensor core (D += R * I with a
// 16x16 identity: four N=16 MMAs per 64 residual columns), epilogue TMEM -> bias/ReLU/border-zero -> bf16 -> swizzled staging ->
// TMA store issued by a dedicated store thread, which also publishes the tile's completion counter once its stores have landed.
#include <vector>

#include "tc_ptx.cuh"

namespace lvcb200 {

constexpr int kChainThreads = 384;       // warp 0 TMA producer, 1 MMA issuer, 2 TMEM alloc, 3 store/publish, 4-11 epilogue
constexpr int kChainEpiWarp0 = 4;
constexpr int kChainEpiWarps = 8;
constexpr int kSlots = 4;
constexpr int kSlotBytes = 49152;        // A [128 x 64] bf16 (16 KB) + B [<=256 x 64] bf16 (32 KB)
constexpr int kSlotBOff = 16384;
constexpr int kStagingBytes = 16384;     // one 64-column output phase: [128 x 64] bf16, 128-byte swizzled rows
constexpr int kOffStaging = kSlots * kSlotBytes;
constexpr int kOffIdent = kOffStaging + 2 * kStagingBytes;
constexpr int kIdent16Bytes = 2048;      // 16 rows x 128 B (16 x 16 bf16 identity in the first 32 bytes of each row, swizzled)
constexpr int kOffCtrl = kOffIdent + kIdent16Bytes;
constexpr int kChainSmem = kOffCtrl + 1024;   // = 232448 = 227 KB exactly
static_assert(kChainSmem == 232448, "chain kernel shared-memory budget");

struct alignas(128) ChainLayer {
  CUtensorMap ta, tw, td, tr;
  const float* bias;
  int M, N, K;
  int taps; int shift[9];
  int relu, plane_h, plane_w;
  int block_n, m_tiles, n_tiles, k_blocks, has_res;
  int tile_base;                  // global index of this layer's tile 0 (round-robin CTA assignment over the whole chain)
  int cnt_off;                    // first completion counter of this layer (one per 128-row block)
  int dep_a_off, dep_a_need, dep_a_mtiles, min_shift, max_shift;   // producer of A inside the chain (dep_a_off < 0: external)
  int dep_r_off, dep_r_need;      // producer of the residual inside the chain
  int pad_[3];
};

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

__global__ void __launch_bounds__(kChainThreads, 1)
gemm_chain_kernel(const ChainLayer* __restrict__ layers, int n_layers, uint32_t* __restrict__ counters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();               // SWIZZLE_128B operands need 1024-byte alignment; the budget has no slack
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffCtrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffCtrl + 256);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kSlots;
  const uint32_t bar_tfull = bar_empty + 8 * kSlots, bar_tempty = bar_tfull + 16;
  const uint32_t bar_sfull = bar_tempty + 16, bar_sempty = bar_sfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kSlots; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; b++) {
      mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, kChainEpiWarps);
      mbar_init(bar_sfull + 8 * b, kChainEpiWarps); mbar_init(bar_sempty + 8 * b, 1);
    }
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

  if (warp == 0) {   // ============================================== TMA producer (whole warp polls dependencies, lane 0 issues)
    uint32_t stage = 0, phase = 0;
    for (int L = 0; L < n_layers; L++) {
      const ChainLayer* ly = layers + L;
      const int n_tiles = ly->n_tiles, k_blocks = ly->k_blocks, taps = ly->taps, Kdim = ly->K, Ndim = ly->N, bn = ly->block_n;
      const int T = ly->m_tiles * n_tiles;
      const bool has_res = ly->has_res != 0;
      const int dep_a_off = ly->dep_a_off, dep_r_off = ly->dep_r_off;
      const uint32_t stage_bytes = 16384u + (uint32_t)bn * 128u;
      int t = (cta - ly->tile_base) % G; if (t < 0) t += G;
      int last_m = -1;
      for (; t < T; t += G) {
        const int m = t / n_tiles, m0 = m * BLOCK_M, n0 = (t % n_tiles) * bn;
        if (m != last_m) {
          last_m = m;
          bool waited = false;
          if (dep_a_off >= 0) {
            long long r_lo = (long long)m0 + ly->min_shift, r_hi = (long long)m0 + BLOCK_M - 1 + ly->max_shift;
            int lo = r_lo < 0 ? 0 : (int)(r_lo / BLOCK_M), hi = (int)(r_hi / BLOCK_M);
            if (hi > ly->dep_a_mtiles - 1) hi = ly->dep_a_mtiles - 1;
            if (lo <= hi) { wait_rows_ready(counters + dep_a_off, lo, hi, (uint32_t)ly->dep_a_need, lane); waited = true; }
          }
          if (has_res && dep_r_off >= 0) { wait_rows_ready(counters + dep_r_off, m, m, (uint32_t)ly->dep_r_need, lane); waited = true; }
          if (waited) fence_proxy_async_all();   // acquire (generic proxy) -> the TMA loads below (async proxy)
        }
        if (lane == 0) {
          for (int tp = 0; tp < taps; tp++) {
            const int row = m0 + ly->shift[tp];
            const int wcol0 = tp * Kdim;
            for (int kb = 0; kb < k_blocks; kb++) {
              const uint32_t fb = bar_full + 8 * stage;
              const uint32_t slot = smem_base + stage * kSlotBytes;
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              mbar_arrive_expect_tx(fb, stage_bytes);
              tma_load_2d(slot, &ly->ta, fb, kb * BLOCK_K, row);
              tma_load_2d(slot + kSlotBOff, &ly->tw, fb, wcol0 + kb * BLOCK_K, n0);
              if (++stage == kSlots) { stage = 0; phase ^= 1u; }
            }
          }
          if (has_res) {   // residual [128 x 64] tiles as extra A operands, two per slot (A region + first 16 KB of the B region)
            const int per = bn >= 128 ? 2 : 1;
            for (int j = 0; j < bn / 64 && n0 + j * 64 < Ndim; j += per) {
              const uint32_t fb = bar_full + 8 * stage;
              const uint32_t slot = smem_base + stage * kSlotBytes;
              const bool two = per == 2 && (n0 + (j + 1) * 64 < Ndim);
              mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
              mbar_arrive_expect_tx(fb, two ? 32768u : 16384u);
              tma_load_2d(slot, &ly->tr, fb, n0 + j * 64, m0);
              if (two) tma_load_2d(slot + kSlotBOff, &ly->tr, fb, n0 + (j + 1) * 64, m0);
              if (++stage == kSlots) { stage = 0; phase ^= 1u; }
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ============================================ MMA issuer
      const uint64_t ident_desc = make_smem_desc_sw128(smem_base + kOffIdent);
      constexpr uint32_t idesc_res = make_idesc_bf16(BLOCK_M, 16);
      uint32_t stage = 0, phase = 0, tc = 0;
      for (int L = 0; L < n_layers; L++) {
        const ChainLayer* ly = layers + L;
        const int n_tiles = ly->n_tiles, Ndim = ly->N, bn = ly->block_n;
        const int k_iters = ly->taps * ly->k_blocks;
        const int T = ly->m_tiles * n_tiles;
        const bool has_res = ly->has_res != 0;
        int t = (cta - ly->tile_base) % G; if (t < 0) t += G;
        for (; t < T; t += G, tc++) {
          const int n0 = (t % n_tiles) * bn;
          const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
          mbar_wait(bar_tempty + 8 * b, bph ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + b * 256;
          const int n_rem = Ndim - n0;
          const int n_eff = n_rem >= bn ? bn : ((n_rem + 15) & ~15);
          const uint32_t idesc_t = make_idesc_bf16(BLOCK_M, n_eff);
          for (int ki = 0; ki < k_iters; ki++) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc_sw128(smem_base + stage * kSlotBytes);
            const uint64_t bdesc = make_smem_desc_sw128(smem_base + stage * kSlotBytes + kSlotBOff);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; k++)
              umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc_t, (ki > 0 || k > 0) ? 1u : 0u);
            umma_commit(bar_empty + 8 * stage);
            if (++stage == kSlots) { stage = 0; phase ^= 1u; }
          }
          if (has_res) {
            const int per = bn >= 128 ? 2 : 1;
            for (int j = 0; j < bn / 64 && n0 + j * 64 < Ndim; j += per) {
              const bool two = per == 2 && (n0 + (j + 1) * 64 < Ndim);
              mbar_wait(bar_full + 8 * stage, phase);
              tc_fence_after();
              const uint64_t adesc = make_smem_desc_sw128(smem_base + stage * kSlotBytes);
              const uint64_t adesc2 = make_smem_desc_sw128(smem_base + stage * kSlotBytes + kSlotBOff);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; k++)   // D[:, 64j + 16k : +16] += R[:, 16k : 16k + 16] * I16
                umma_bf16(tmem_d + j * 64 + 16 * k, adesc + 2 * k, ident_desc, idesc_res, 1u);
              if (two) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; k++)
                  umma_bf16(tmem_d + (j + 1) * 64 + 16 * k, adesc2 + 2 * k, ident_desc, idesc_res, 1u);
              }
              umma_commit(bar_empty + 8 * stage);
              if (++stage == kSlots) { stage = 0; phase ^= 1u; }
            }
          }
          umma_commit(bar_tfull + 8 * b);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {   // ============================================ store thread: staging -> TMA store, publish tile completion
      uint32_t gphase = 0;
      for (int L = 0; L < n_layers; L++) {
        const ChainLayer* ly = layers + L;
        const int n_tiles = ly->n_tiles, Ndim = ly->N, bn = ly->block_n;
        const int T = ly->m_tiles * n_tiles;
        uint32_t* cnt = counters + ly->cnt_off;
        int t = (cta - ly->tile_base) % G; if (t < 0) t += G;
        for (; t < T; t += G) {
          const int m = t / n_tiles, m0 = m * BLOCK_M, n0 = (t % n_tiles) * bn;
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
          red_release_add_u32(cnt + m, 1u);
        }
      }
    }
  } else if (warp >= kChainEpiWarp0) {   // ============================ epilogue warps (two per TMEM lane quadrant)
    const int q = warp & 3;
    const int half = (warp - kChainEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const int sw = r & 7;
    uint32_t tc = 0, gphase = 0;
    for (int L = 0; L < n_layers; L++) {
      const ChainLayer* ly = layers + L;
      const int n_tiles = ly->n_tiles, Ndim = ly->N, bn = ly->block_n;
      const int T = ly->m_tiles * n_tiles;
      const int relu = ly->relu, plane_h = ly->plane_h, plane_w = ly->plane_w;
      const float* __restrict__ bias = ly->bias;
      int t = (cta - ly->tile_base) % G; if (t < 0) t += G;
      for (; t < T; t += G, tc++) {
        const int m0 = (t / n_tiles) * BLOCK_M, n0 = (t % n_tiles) * bn;
        const uint32_t b = tc & 1u, bph = (tc >> 1) & 1u;
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
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * 256;
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
            uint4 o; __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; e++) {
              float a0 = __uint_as_float(v[8 * j + 2 * e]) + bb[2 * e];
              float a1 = __uint_as_float(v[8 * j + 2 * e + 1]) + bb[2 * e + 1];
              if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
              if (zero_row) { a0 = 0.f; a1 = 0.f; }
              ho[e] = __floats2bfloat162_rn(a0, a1);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static inline size_t chain_counter_bytes(int total_m_tiles) { return align_up((size_t)total_m_tiles * 4, 256); }

static int chain_validate(const lvcb200_gemm_desc* descs, int n, int* total_m_tiles) {
  LVC_REQUIRE(descs && n >= 1 && n <= 4096, "gemm_chain: need 1..4096 layer descriptors");
  int tm = 0;
  for (int i = 0; i < n; i++) {
    const lvcb200_gemm_desc* d = descs + i;
    LVC_REQUIRE(d->a_dtype == LVCB200_BF16 && d->d_dtype == LVCB200_BF16, "gemm_chain: bf16 operands and bf16 outputs only");
    LVC_REQUIRE(d->M >= 1 && d->M < (1ll << 31) && d->M_rows < (1ll << 31), "gemm_chain: bad M");
    LVC_REQUIRE(d->N >= 64 && d->N % 64 == 0 && d->K >= 64 && d->K % 64 == 0, "gemm_chain: N and K must be multiples of 64");
    LVC_REQUIRE(d->taps >= 1 && d->taps <= 9, "gemm_chain: taps");
    LVC_REQUIRE(d->A && d->W && d->D, "gemm_chain: NULL pointer");
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

extern "C" size_t lvcb200_gemm_chain_workspace(const lvcb200_gemm_desc* descs, int n) {
  int tm = 0;
  if (chain_validate(descs, n, &tm)) return 0;
  return chain_counter_bytes(tm) + (size_t)n * sizeof(ChainLayer);
}

extern "C" int lvcb200_gemm_chain_plan(const lvcb200_gemm_desc* descs, int n, void* workspace, size_t workspace_bytes,
                                       lvcb200_chain_plan* plan) {
  int tm = 0;
  int rc = chain_validate(descs, n, &tm);
  if (rc) return rc;
  LVC_REQUIRE(workspace && plan, "gemm_chain_plan: NULL pointer");
  LVC_REQUIRE(((uintptr_t)workspace % 256) == 0, "gemm_chain_plan: workspace must be 256-byte aligned");
  const size_t cbytes = chain_counter_bytes(tm);
  if (workspace_bytes < cbytes + (size_t)n * sizeof(ChainLayer)) return set_error(LVCB200_EWORKSPACE, "gemm_chain_plan: workspace too small");
  std::vector<ChainLayer> tab((size_t)n);
  long long tile_base = 0;
  int cnt_off = 0;
  for (int i = 0; i < n; i++) {
    const lvcb200_gemm_desc* d = descs + i;
    ChainLayer& L = tab[i];
    memset(&L, 0, sizeof(L));
    const int bn = d->N >= 256 ? 256 : (d->N > 64 ? (d->N > 128 ? 256 : 128) : 64);
    L.bias = d->bias;
    L.M = (int)d->M; L.N = d->N; L.K = d->K; L.taps = d->taps;
    int mn = 0, mx = 0;
    for (int t = 0; t < 9; t++) {
      L.shift[t] = t < d->taps ? d->shift[t] : 0;
      if (t < d->taps) { if (d->shift[t] < mn) mn = d->shift[t]; if (d->shift[t] > mx) mx = d->shift[t]; }
    }
    L.min_shift = mn; L.max_shift = mx;
    L.relu = d->relu; L.plane_h = d->plane_h; L.plane_w = d->plane_w;
    L.block_n = bn;
    L.m_tiles = (int)((d->M + BLOCK_M - 1) / BLOCK_M);
    L.n_tiles = (d->N + bn - 1) / bn;
    L.k_blocks = d->K / BLOCK_K;
    L.has_res = d->residual ? 1 : 0;
    L.tile_base = (int)(tile_base % (1ll << 30));
    L.cnt_off = cnt_off;
    L.dep_a_off = -1; L.dep_r_off = -1;
    for (int j = i - 1; j >= 0; j--) {
      if (L.dep_a_off < 0 && descs[j].D == d->A) {
        LVC_REQUIRE(descs[j].ldd == d->lda && descs[j].N >= d->K, "gemm_chain_plan: a layer reads an earlier output with a different geometry");
        L.dep_a_off = tab[j].cnt_off; L.dep_a_need = tab[j].n_tiles; L.dep_a_mtiles = tab[j].m_tiles;
      }
      if (d->residual && L.dep_r_off < 0 && descs[j].D == d->residual) {
        LVC_REQUIRE(descs[j].ldd == d->ldr && descs[j].N >= d->N && descs[j].M >= d->M, "gemm_chain_plan: residual produced with a different geometry");
        L.dep_r_off = tab[j].cnt_off; L.dep_r_need = tab[j].n_tiles;
      }
    }
    if ((rc = make_tmap_2d(&L.ta, d->A, d->M_rows, d->K, d->lda, BLOCK_M))) return rc;
    if ((rc = make_tmap_2d(&L.tw, d->W, d->N, (long long)d->taps * d->K, d->ldw, bn))) return rc;
    if ((rc = make_tmap_2d(&L.td, d->D, d->M, d->N, d->ldd, BLOCK_M))) return rc;
    L.tr = L.ta;
    if (d->residual && (rc = make_tmap_2d(&L.tr, d->residual, d->M, d->N, d->ldr, BLOCK_M))) return rc;
    tile_base += (long long)L.m_tiles * L.n_tiles;
    cnt_off += L.m_tiles;
  }
  LVC_REQUIRE(tile_base < (1ll << 30), "gemm_chain_plan: too many tiles");
  LVC_CUDA(cudaMemcpy((uint8_t*)workspace + cbytes, tab.data(), (size_t)n * sizeof(ChainLayer), cudaMemcpyHostToDevice));
  static bool attr_set = false;
  if (!attr_set) {
    LVC_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem));
    attr_set = true;
  }
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  plan->workspace = workspace;
  plan->n_layers = n;
  plan->counter_bytes = (int64_t)cbytes;
  plan->grid = (int)(tile_base < sms ? tile_base : sms);   // every CTA must be resident: dependencies are resolved by spinning
  plan->total_tiles = (int64_t)tile_base;
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
  const ChainLayer* layers = reinterpret_cast<const ChainLayer*>((const uint8_t*)plan->workspace + plan->counter_bytes);
  uint32_t* counters = reinterpret_cast<uint32_t*>(plan->workspace);
  LVC_CUDA(cudaLaunchKernelEx(&cfg, gemm_chain_kernel, layers, plan->n_layers, counters));
  return check_launch("gemm_chain_kernel");
}
