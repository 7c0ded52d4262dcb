// Descriptor front end of the label-verification step: the DINO ViT-S/8 forward that turns candidate crops into the
// 384-d descriptors the kNN consumes (tools/run_nearest_neighbours.py:108-128: `crop_features = model(crops)`, model =
// torch.hub 'facebookresearch/dino:main' dino_vits8 -- a third-party dependency of the reference, restated from its published
// architecture: patch 8, dim 384, depth 12, 6 heads, MLP ratio 4, LayerNorm eps 1e-6, erf GELU, CLS token of the final norm).
// Every linear layer (patch embedding = the 8x8 / stride-8 conv as a GEMM over patch rows, qkv, proj, fc1, fc2) runs on the tcgen05
// GEMM of gemm_tc.cu with bias / residual fused; this file holds the kernels in between:
//   vit_patchify_kernel   : crops [B,3,S,S] fp32 -> patch rows [B * (S/8)^2, 192] bf16, column order (c, iy, ix) = the conv weight's
//   vit_assemble_kernel   : [cls | patch tokens] + pos_embed -> token matrix [B * (1 + Np), D] bf16
//   layernorm_kernel      : warp per row, fp32 statistics, bf16 or fp32 output, arbitrary row pitch (the final norm reads CLS rows only)
//   gelu_kernel           : exact (erf) GELU in place on bf16
//   attention_kernel      : fused softmax(Q K^T / sqrt(d)) V for head_dim 64 on mma.sync (flash-style: 64 queries per CTA, K / V streamed in
//                           64-key blocks through cp.async double buffers, online softmax in registers, P re-used as the A fragment).
//                           The default attention of DinoViT is the tcgen05 / TMEM kernel of attention_tc.cu; this one stays selectable.
//   *8 / *_reg variants   : 16-byte fast paths of patchify / assemble (patch 8, D % 8 == 0) and LayerNorm with the row in registers (D = 384)
#include <cuda_fp16.h>

#include "common.cuh"

namespace lvcb200 {

__global__ void vit_patchify_kernel(const float* __restrict__ crops, int B, int S, int P, __nv_bfloat16* __restrict__ out) {
  const int gp = S / P, K = 3 * P * P;
  const long long total = (long long)B * gp * gp * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long row = i / K;
    const int px = (int)(row % gp), py = (int)((row / gp) % gp), b = (int)(row / ((long long)gp * gp));
    const int c = k / (P * P), iy = (k / P) % P, ix = k % P;
    out[i] = __float2bfloat16_rn(crops[(((long long)b * 3 + c) * S + py * P + iy) * S + px * P + ix]);
  }
}

__global__ void vit_assemble_kernel(const __nv_bfloat16* __restrict__ patch_tokens, const float* __restrict__ cls_token,
                                    const float* __restrict__ pos_embed, int B, int Np, int D, __nv_bfloat16* __restrict__ x) {
  const long long total = (long long)B * (Np + 1) * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long tok = i / D;
    const int t = (int)(tok % (Np + 1));
    const long long b = tok / (Np + 1);
    const float v = t == 0 ? cls_token[d] : __bfloat162float(patch_tokens[(b * Np + t - 1) * D + d]);
    x[i] = __float2bfloat16_rn(v + pos_embed[(long long)t * D + d]);
  }
}

// patch 8 fast path: one thread per (patch, channel, patch row): eight contiguous pixels in (two float4), eight bf16 out (one 16-byte store);
// consecutive threads walk the 24 (channel, row) pairs of a patch, so a patch's 384-byte output row is written by 24 neighbouring threads.
__global__ void __launch_bounds__(256)
vit_patchify8_kernel(const float* __restrict__ crops, int B, int S, __nv_bfloat16* __restrict__ out) {
  const int gp = S >> 3;
  const long long total = (long long)B * gp * gp * 24;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ciy = (int)(i % 24);
    const long long row = i / 24;
    const int c = ciy >> 3, iy = ciy & 7;
    const int px = (int)(row % gp), py = (int)((row / gp) % gp);
    const long long b = row / ((long long)gp * gp);
    const float4* src = reinterpret_cast<const float4*>(crops + ((b * 3 + c) * S + py * 8 + iy) * S + px * 8);
    const float4 a = __ldg(src), d = __ldg(src + 1);
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
    h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
    h[2] = __floats2bfloat162_rn(d.x, d.y); h[3] = __floats2bfloat162_rn(d.z, d.w);
    *reinterpret_cast<uint4*>(out + row * 192 + ciy * 8) = o;
  }
}

// D % 8 == 0 fast path of vit_assemble: eight channels (16 bytes) per thread.
__global__ void __launch_bounds__(256)
vit_assemble8_kernel(const __nv_bfloat16* __restrict__ patch_tokens, const float* __restrict__ cls_token, const float* __restrict__ pos_embed,
                     int B, int Np, int D, __nv_bfloat16* __restrict__ x) {
  const int dv = D >> 3;
  const long long total = (long long)B * (Np + 1) * dv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dv) * 8;
    const long long tok = i / dv;
    const int t = (int)(tok % (Np + 1));
    const long long b = tok / (Np + 1);
    float v[8];
    if (t == 0) {
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = cls_token[d + e];
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(patch_tokens + (b * Np + t - 1) * D + d);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
    }
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos_embed + (long long)t * D + d));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos_embed + (long long)t * D + d + 4));
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
    ho[0] = __floats2bfloat162_rn(v[0] + p0.x, v[1] + p0.y); ho[1] = __floats2bfloat162_rn(v[2] + p0.z, v[3] + p0.w);
    ho[2] = __floats2bfloat162_rn(v[4] + p1.x, v[5] + p1.y); ho[3] = __floats2bfloat162_rn(v[6] + p1.z, v[7] + p1.w);
    *reinterpret_cast<uint4*>(x + tok * D + d) = o;
  }
}

template <typename TO>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int D, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, TO* __restrict__ out, long long ldo) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* p = x + row * ldx;
  float s = 0.f, s2 = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + d));
    s += v.x + v.y;
    s2 += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float mean = s / (float)D;
  const float var = fmaxf(s2 / (float)D - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  TO* o = out + row * ldo;
  for (int d = lane * 2; d < D; d += 64) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + d));
    const float a = (v.x - mean) * rstd * gamma[d] + beta[d], b = (v.y - mean) * rstd * gamma[d + 1] + beta[d + 1];
    if constexpr (sizeof(TO) == 2) *reinterpret_cast<__nv_bfloat162*>(o + d) = __floats2bfloat162_rn(a, b);
    else { o[d] = a; o[d + 1] = b; }
  }
}

// Same, for D = 64 * NP with the row held in registers (one read of x instead of two): warp per row, lane l owns the bf16 pairs at
// d = 2 l + 64 i.  gamma / beta are read through the read-only path (the same 3 KB for every row).
template <typename TO, int NP>
__global__ void __launch_bounds__(256)
layernorm_reg_kernel(const __nv_bfloat16* __restrict__ x, long long rows, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, TO* __restrict__ out, long long ldo) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* p = x + row * ldx + 2 * lane;
  float2 v[NP];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NP; i++) v[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + 64 * i));
#pragma unroll
  for (int i = 0; i < NP; i++) { s += v[i].x + v[i].y; s2 += v[i].x * v[i].x + v[i].y * v[i].y; }
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  constexpr float inv_d = 1.0f / (float)(64 * NP);
  const float mean = s * inv_d;
  const float var = fmaxf(s2 * inv_d - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  TO* o = out + row * ldo + 2 * lane;
#pragma unroll
  for (int i = 0; i < NP; i++) {
    const float2 g = __ldg(reinterpret_cast<const float2*>(gamma + 2 * lane + 64 * i));
    const float2 b = __ldg(reinterpret_cast<const float2*>(beta + 2 * lane + 64 * i));
    const float a0 = (v[i].x - mean) * rstd * g.x + b.x, a1 = (v[i].y - mean) * rstd * g.y + b.y;
    if constexpr (sizeof(TO) == 2) *reinterpret_cast<__nv_bfloat162*>(o + 64 * i) = __floats2bfloat162_rn(a0, a1);
    else { o[64 * i] = a0; o[64 * i + 1] = a1; }
  }
}

__global__ void gelu_kernel(uint4* __restrict__ x, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    uint4 u = x[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float2 f = __bfloat1622float2(h[j]);
      f.x = 0.5f * f.x * (1.0f + erff(f.x * 0.70710678118654752f));
      f.y = 0.5f * f.y * (1.0f + erff(f.y * 0.70710678118654752f));
      h[j] = __floats2bfloat162_rn(f.x, f.y);
    }
    x[i] = u;
  }
}

// ------------------------------------------------------------------------------------------ fused attention, head_dim 64
constexpr int AT_Q = 64, AT_K = 64, AT_D = 64, AT_LD = AT_D + 8;   // padded rows (144 bytes): conflict-free ldmatrix
constexpr int AT_TILE = AT_K * AT_LD * 2;                          // one K or V block in shared memory (9216 bytes)

__device__ __forceinline__ void at_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t at_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void at_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;     // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// qkv [B*N, 3*H*64] (q | k | v, each [H][64]); out [B*N, H*64].  grid (ceil(N / 64), H, B), 128 threads: warp w owns queries 16 w .. 16 w + 15.
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, int N, int H, float scale_log2e, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) uint8_t smem[5 * AT_TILE];               // Q | K0 | V0 | K1 | V1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * AT_Q, h = blockIdx.y, b = blockIdx.z;
  const long long ld = 3LL * H * AT_D;
  const __nv_bfloat16* base = qkv + (long long)b * N * ld + (long long)h * AT_D;
  const uint32_t s_q = (uint32_t)__cvta_generic_to_shared(smem);
  auto load_tile = [&](uint32_t dst, const __nv_bfloat16* src, int row0) {   // 64 rows x 64 bf16 (8 x 16-byte chunks per row)
    for (int i = threadIdx.x; i < AT_K * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      const bool ok = row0 + r < N;
      at_cp16(dst + r * (AT_LD * 2) + c * 16, src + (long long)(ok ? row0 + r : 0) * ld + c * 8, ok);
    }
  };
  load_tile(s_q, base, q0);
  const int nkb = (N + AT_K - 1) / AT_K;
  load_tile(s_q + AT_TILE, base + H * AT_D, 0);
  load_tile(s_q + 2 * AT_TILE, base + 2 * H * AT_D, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) { o[i][0] = 0.f; o[i][1] = 0.f; o[i][2] = 0.f; o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;         // running max / sum of rows g and g + 8 (this thread's partial columns)
  uint32_t qa[4][4];
  for (int kb = 0; kb < nkb; kb++) {
    const uint32_t s_k = s_q + (1 + 2 * (kb & 1)) * AT_TILE, s_v = s_k + AT_TILE;
    if (kb + 1 < nkb) {
      load_tile(s_q + (1 + 2 * ((kb + 1) & 1)) * AT_TILE, base + H * AT_D, (kb + 1) * AT_K);
      load_tile(s_q + (2 + 2 * ((kb + 1) & 1)) * AT_TILE, base + 2 * H * AT_D, (kb + 1) * AT_K);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (kb == 0) {   // Q fragments of this warp's 16 rows, four k-steps of 16
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const uint32_t addr = s_q + (warp * 16 + (lane & 15)) * (AT_LD * 2) + ks * 32 + (lane >> 4) * 16;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(qa[ks][0]), "=r"(qa[ks][1]), "=r"(qa[ks][2]), "=r"(qa[ks][3]) : "r"(addr));
      }
    }
    // S = Q K^T : 16 x 64 scores per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { s[i][0] = 0.f; s[i][1] = 0.f; s[i][2] = 0.f; s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
#pragma unroll
      for (int np = 0; np < 4; np++) {   // two key tiles of 8 per ldmatrix.x4: (keys 16 np .. +7, k lo / hi), (keys 16 np + 8 .. +15, k lo / hi)
        uint32_t k0, k1, k2, k3;
        const uint32_t addr = s_k + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * (AT_LD * 2) + ks * 32 + ((lane >> 3) & 1) * 16;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(addr));
        at_mma(s[2 * np], qa[ks], k0, k1);
        at_mma(s[2 * np + 1], qa[ks], k2, k3);
      }
    }
    // scale, mask keys >= N, online softmax (base 2)
    const int key0 = kb * AT_K;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int key = key0 + i * 8 + 2 * t + (e & 1);
        s[i][e] = key < N ? s[i][e] * scale_log2e : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = exp2f(m0 - mx0), c1 = exp2f(m1 - mx1);          // rescale of what was accumulated so far (0 on the first block)
    m0 = mx0; m1 = mx1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int i = 0; i < 8; i++) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
    uint32_t pa[4][4];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float p0 = exp2f(s[i][0] - mx0), p1 = exp2f(s[i][1] - mx0), p2 = exp2f(s[i][2] - mx1), p3 = exp2f(s[i][3] - mx1);
      l0 += p0 + p1; l1 += p2 + p3;
      pa[i >> 1][(i & 1) * 2] = at_pack(p0, p1);          // A fragment of the P V product: k-step i / 2, (row g | row g + 8) x (keys lo | hi)
      pa[i >> 1][(i & 1) * 2 + 1] = at_pack(p2, p3);
    }
    // O += P V : V block rows = keys (k), columns = d (n) -> ldmatrix.trans
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
#pragma unroll
      for (int np = 0; np < 4; np++) {
        uint32_t v0, v1, v2, v3;
        const uint32_t addr = s_v + (ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * (AT_LD * 2) + np * 32 + (lane >> 4) * 16;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        at_mma(o[2 * np], pa[ks], v0, v1);
        at_mma(o[2 * np + 1], pa[ks], v2, v3);
      }
    }
    __syncthreads();   // everybody is done with this K / V buffer before the next iteration's loads overwrite the other one's predecessor
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* ob = out + (long long)b * N * H * AT_D + (long long)h * AT_D;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (r0 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * H * AT_D + i * 8 + 2 * t) = at_pack(o[i][0] * i0, o[i][1] * i0);
    if (r1 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * H * AT_D + i * 8 + 2 * t) = at_pack(o[i][2] * i1, o[i][3] * i1);
  }
}

}  // namespace lvcb200

using namespace lvcb200;

static inline unsigned vit_grid(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 32;
  return (unsigned)(b < cap ? b : cap);
}

extern "C" int lvcb200_vit_patchify(const float* crops, int B, int S, int patch, void* out, void* stream) {
  if (B == 0) return 0;
  LVC_REQUIRE(crops && out && S > 0 && patch > 0 && S % patch == 0, "vit_patchify: bad argument");
  if (patch == 8 && S % 8 == 0 && ((uintptr_t)crops % 16) == 0 && ((uintptr_t)out % 16) == 0) {
    const long long n = (long long)B * (S / 8) * (S / 8) * 24;
    vit_patchify8_kernel<<<vit_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(crops, B, S, (__nv_bfloat16*)out);
    return check_launch("vit_patchify8_kernel");
  }
  const long long total = (long long)B * (S / patch) * (S / patch) * 3 * patch * patch;
  vit_patchify_kernel<<<vit_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(crops, B, S, patch, (__nv_bfloat16*)out);
  return check_launch("vit_patchify_kernel");
}

extern "C" int lvcb200_vit_assemble(const void* patch_tokens, const float* cls_token, const float* pos_embed, int B, int Np, int D, void* x,
                                    void* stream) {
  if (B == 0) return 0;
  LVC_REQUIRE(patch_tokens && cls_token && pos_embed && x && Np > 0 && D > 0, "vit_assemble: bad argument");
  if (D % 8 == 0 && ((uintptr_t)patch_tokens % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)pos_embed % 16) == 0) {
    const long long n = (long long)B * (Np + 1) * (D / 8);
    vit_assemble8_kernel<<<vit_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)patch_tokens, cls_token, pos_embed, B, Np, D,
                                                                          (__nv_bfloat16*)x);
    return check_launch("vit_assemble8_kernel");
  }
  const long long total = (long long)B * (Np + 1) * D;
  vit_assemble_kernel<<<vit_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)patch_tokens, cls_token, pos_embed, B, Np, D,
                                                                            (__nv_bfloat16*)x);
  return check_launch("vit_assemble_kernel");
}

extern "C" int lvcb200_layernorm(const void* x, int64_t rows, int D, int64_t ldx, const float* gamma, const float* beta, float eps, void* out,
                                 int out_dtype, int64_t ldo, void* stream) {
  if (rows == 0) return 0;
  LVC_REQUIRE(x && gamma && beta && out && D > 0 && D % 2 == 0 && ldx % 2 == 0 && ldo % 2 == 0, "layernorm: bad argument (D, pitches must be even)");
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (D == 384 && ((uintptr_t)gamma % 8) == 0 && ((uintptr_t)beta % 8) == 0 && (out_dtype == LVCB200_BF16 || out_dtype == LVCB200_F32)) {   // ViT-S
    if (out_dtype == LVCB200_BF16)
      layernorm_reg_kernel<__nv_bfloat16, 6><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, ldx, gamma, beta, eps, (__nv_bfloat16*)out, ldo);
    else
      layernorm_reg_kernel<float, 6><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, ldx, gamma, beta, eps, (float*)out, ldo);
    return check_launch("layernorm_reg_kernel");
  }
  if (out_dtype == LVCB200_BF16)
    layernorm_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, D, ldx, gamma, beta, eps, (__nv_bfloat16*)out, ldo);
  else if (out_dtype == LVCB200_F32)
    layernorm_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, D, ldx, gamma, beta, eps, (float*)out, ldo);
  else
    return set_error(LVCB200_EINVAL, "layernorm: out_dtype must be LVCB200_BF16 or LVCB200_F32");
  return check_launch("layernorm_kernel");
}

extern "C" int lvcb200_gelu(void* x, int64_t n, void* stream) {
  if (n == 0) return 0;
  LVC_REQUIRE(x && n % 8 == 0 && ((uintptr_t)x % 16) == 0, "gelu: n must be a multiple of 8, x 16-byte aligned");
  gelu_kernel<<<vit_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((uint4*)x, n / 8);
  return check_launch("gelu_kernel");
}

extern "C" int lvcb200_attention(const void* qkv, int B, int N, int H, int head_dim, float scale, void* out, void* stream) {
  if (B == 0 || N == 0) return 0;
  LVC_REQUIRE(qkv && out && H >= 1 && head_dim == AT_D, "attention: head_dim must be 64");
  LVC_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 4) == 0, "attention: qkv must be 16-byte aligned");
  dim3 grid((N + AT_Q - 1) / AT_Q, H, B);
  attention_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, N, H, scale * 1.4426950408889634f, (__nv_bfloat16*)out);
  return check_launch("attention_kernel");
}
