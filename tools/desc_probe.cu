// Micro-experiment for DESIGN section 8 item 4: does a tcgen05 A-operand descriptor whose start address is shifted by r rows (r * 128 B)
// inside a SWIZZLE_128B tile read rows r .. r+127 correctly, and does it need the descriptor's base_offset field ((addr >> 7) & 7)?
// One CTA: TMA-load a [136 x 64] bf16 tile, B = 64 x 64 identity, D = A_shift * I, compare on the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -I lvc_b200/csrc tools/desc_probe.cu -o tools/build/desc_probe -lcudart
#include <cstdio>
#include <vector>
#include "tc_ptx.cuh"

namespace lvcb200 { thread_local char g_last_error[512] = ""; std::atomic<long long> g_launch_count{0}; }
using namespace lvcb200;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, int r, int use_base_offset, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  uint8_t* ident = smem + 18432;                       // A tile: 136 rows x 128 B = 17 408 B, padded to 18 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 18432 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 18432 + 8192 + 64);
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) { mbar_init(smem_u32(bars), 1); mbar_init(smem_u32(bars) + 8, 1); fence_barrier_init(); }
  for (int i = t; i < 8192 / 16; i += 128) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (t < 64) {
    const int n = t, c = n >> 3;
    *reinterpret_cast<__nv_bfloat16*>(ident + n * 128 + ((c ^ (n & 7)) << 4) + (n & 7) * 2) = __float2bfloat16_rn(1.0f);
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (t == 0) {
    mbar_arrive_expect_tx(smem_u32(bars), 136 * 128);
    tma_load_2d(base, &tmap_a, smem_u32(bars), 0, 0);
    mbar_wait(smem_u32(bars), 0);
    tc_fence_after();
    const uint32_t a_addr = base + (uint32_t)r * 128u;
    uint64_t adesc = make_smem_desc_sw128(a_addr);
    if (use_base_offset) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(ident));
    const uint32_t idesc = make_idesc_bf16(128, 64);
    for (int k = 0; k < 4; k++) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_commit(smem_u32(bars) + 8);
  }
  mbar_wait(smem_u32(bars) + 8, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++) out[t * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  const int R = 136, C = 64;
  std::vector<__nv_bfloat16> ha(R * C);
  for (int i = 0; i < R; i++) for (int k = 0; k < C; k++) ha[i * C + k] = __float2bfloat16_rn((float)(i * 64 + k) * 0.25f);   // exact in bf16? values < 2^8 steps: use small ints
  for (int i = 0; i < R; i++) for (int k = 0; k < C; k++) ha[i * C + k] = __float2bfloat16_rn((float)((i * 7 + k * 3) % 251));
  __nv_bfloat16* da; float* dout;
  cudaMalloc(&da, R * C * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), R * C * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta;
  if (make_tmap_2d(&ta, da, R, C, C, R)) { printf("tmap failed: %s\n", g_last_error); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> ho(128 * 64);
  for (int bo = 0; bo < 2; bo++)
    for (int r = 0; r <= 8; r++) {
      cudaMemset(dout, 0, 128 * 64 * 4);
      probe_kernel<<<1, 128, 32768>>>(ta, r, bo, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("r=%d base_offset=%d: CUDA error %s\n", r, bo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int i = 0; i < 128; i++) for (int k = 0; k < 64; k++) {
        const float want = __bfloat162float(ha[(i + r) * C + k]);
        if (ho[i * 64 + k] != want) { if (first < 0) first = i * 64 + k; bad++; }
      }
      printf("row shift r=%d base_offset_field=%d: %s (%d of 8192 wrong%s)\n", r, bo, bad ? "MISMATCH" : "exact", bad,
             bad ? "" : "");
      if (bad && first >= 0) printf("   first mismatch at row %d col %d: got %.1f want %.1f\n", first / 64, first % 64, ho[first], __bfloat162float(ha[(first / 64 + r) * C + first % 64]));
    }
  return 0;
}
