// Memory-bound helpers of the conv engine.  Activations are "zero-bordered channels-last planes":
// bf16 [n, H+2, W+2, C] with a one-pixel zero frame, so that a 3x3 conv is a GEMM over row-shifted views of the
// same matrix (gemm_tc.cu) and no kernel ever needs an explicit halo test.  All kernels move 16-byte vectors.
#include "common.cuh"

namespace lvcb200 {

// (x - mean) / std + zero pad (rcnn.py:324-333, image_list.py:57-119) fused with the patch gather of the 7x7/2
// stem conv (resnet.py:588-590): row = output pixel of the half-resolution plane, k = c*49 + kh*7 + kw.
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* const* __restrict__ images, const int32_t* __restrict__ image_sizes, int n, int Ho, int Wo,
                   const float* __restrict__ mean, const float* __restrict__ inv_std, __nv_bfloat16* __restrict__ out, int Kpad) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int PW = Wo + 2, PH = Ho + 2;
  const long long rows = (long long)n * PH * PW;
  if (warp >= rows) return;
  const int img = (int)(warp / (PH * PW));
  const int rem = (int)(warp - (long long)img * PH * PW);
  const int py = rem / PW, px = rem - py * PW;
  __nv_bfloat16* o = out + warp * Kpad;
  const bool border = py == 0 || py == PH - 1 || px == 0 || px == PW - 1;
  const int H = image_sizes[img * 2], W = image_sizes[img * 2 + 1];
  const float* im = images[img];
  const int oy = py - 1, ox = px - 1;
  for (int k = lane; k < Kpad; k += 32) {
    float v = 0.f;
    if (!border && k < 147) {
      int c = k / 49, r = k - c * 49, kh = r / 7, kw = r - kh * 7;
      int y = 2 * oy - 3 + kh, x = 2 * ox - 3 + kw;
      if (y >= 0 && y < H && x >= 0 && x < W) v = (__ldg(im + ((long long)c * H + y) * W + x) - mean[c]) * inv_std[c];
    }
    o[k] = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) z[i] = __hmax2(x[i], y[i]);
  return r;
}
__device__ __forceinline__ uint4 bf16x8_add(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) {  // add in fp32, round once
    float2 f = __bfloat1622float2(x[i]), g = __bfloat1622float2(y[i]);
    z[i] = __floats2bfloat162_rn(f.x + g.x, f.y + g.y);
  }
  return r;
}

// F.max_pool2d(k=3, s=2, p=1) (resnet.py:591).  Input is post-ReLU (>= 0), so the zero frame is equivalent to the
// reference's -inf padding.
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, int n, int H, int W, int CV, uint4* __restrict__ out, int Ho, int Wo) {
  const long long total = (long long)n * (Ho + 2) * (Wo + 2) * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int px = (int)(pix % (Wo + 2)), py = (int)((pix / (Wo + 2)) % (Ho + 2)), img = (int)(pix / ((long long)(Wo + 2) * (Ho + 2)));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (py >= 1 && py <= Ho && px >= 1 && px <= Wo) {
      int oy = py - 1, ox = px - 1;
      const uint4* base = in + ((long long)img * (H + 2) * (W + 2)) * CV + cv;
      // window rows 2*oy-1 .. 2*oy+1 in image coords = 2*oy .. 2*oy+2 in plane coords (always inside the plane)
      bool first = true;
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
          int yy = 2 * oy + dy, xx = 2 * ox + dx;
          if (yy > H + 1 || xx > W + 1) continue;
          uint4 v = __ldg(base + ((long long)yy * (W + 2) + xx) * CV);
          r = first ? v : bf16x8_max(r, v);
          first = false;
        }
    }
    out[i] = r;
  }
}

// out[n, oy, ox] = in[n, 2*oy, 2*ox]  (1x1 stride-2 conv input side; LastLevelMaxPool k=1 s=2)
__global__ void subsample2_kernel(const uint4* __restrict__ in, int n, int H, int W, int CV, uint4* __restrict__ out, int Ho, int Wo) {
  const long long total = (long long)n * (Ho + 2) * (Wo + 2) * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int px = (int)(pix % (Wo + 2)), py = (int)((pix / (Wo + 2)) % (Ho + 2)), img = (int)(pix / ((long long)(Wo + 2) * (Ho + 2)));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (py >= 1 && py <= Ho && px >= 1 && px <= Wo) {
      int y = 2 * (py - 1) + 1, x = 2 * (px - 1) + 1;  // plane coords of the source pixel
      r = __ldg(in + (((long long)img * (H + 2) + y) * (W + 2) + x) * CV + cv);
    }
    out[i] = r;
  }
}

// inout += nearest_upsample_2x(top)  (fpn.py:131-133)
__global__ void upsample2_add_kernel(const uint4* __restrict__ top, int n, int Ht, int Wt, int CV, uint4* __restrict__ io, int H, int W) {
  const long long total = (long long)n * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int x = (int)(pix % W), y = (int)((pix / W) % H), img = (int)(pix / ((long long)W * H));
    long long dst = (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * CV + cv;
    long long src = (((long long)img * (Ht + 2) + (y >> 1) + 1) * (Wt + 2) + (x >> 1) + 1) * CV + cv;
    io[dst] = bf16x8_add(io[dst], __ldg(top + src));
  }
}

}  // namespace lvcb200

using namespace lvcb200;

static inline unsigned grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 32;
  return (unsigned)(b < cap ? b : cap);
}

extern "C" int lvcb200_stem_im2col(const float* const* images, const int32_t* image_sizes, int n, int Hpad, int Wpad,
                                   const float* mean, const float* inv_std, void* out, int Kpad, void* stream) {
  LVC_REQUIRE(n >= 1 && Hpad % 2 == 0 && Wpad % 2 == 0 && Kpad >= 147 && Kpad % 8 == 0, "stem_im2col: bad shape");
  LVC_REQUIRE(images && image_sizes && mean && inv_std && out, "stem_im2col: NULL pointer");
  const int Ho = Hpad / 2, Wo = Wpad / 2;
  long long rows = (long long)n * (Ho + 2) * (Wo + 2);
  stem_im2col_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      images, image_sizes, n, Ho, Wo, mean, inv_std, (__nv_bfloat16*)out, Kpad);
  return check_launch("stem_im2col_kernel");
}

extern "C" int lvcb200_maxpool3x3s2(const void* in, int n, int H, int W, int C, void* out, void* stream) {
  LVC_REQUIRE(n >= 1 && H >= 2 && W >= 2 && C % 8 == 0 && in && out, "maxpool3x3s2: bad argument");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  long long total = (long long)n * (Ho + 2) * (Wo + 2) * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, n, H, W, C / 8, (uint4*)out, Ho, Wo);
  return check_launch("maxpool3x3s2_kernel");
}

extern "C" int lvcb200_subsample2(const void* in, int n, int H, int W, int C, void* out, void* stream) {
  LVC_REQUIRE(n >= 1 && H >= 1 && W >= 1 && C % 8 == 0 && in && out, "subsample2: bad argument");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  long long total = (long long)n * (Ho + 2) * (Wo + 2) * (C / 8);
  subsample2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, n, H, W, C / 8, (uint4*)out, Ho, Wo);
  return check_launch("subsample2_kernel");
}

extern "C" int lvcb200_upsample2_add(const void* top, int n, int Ht, int Wt, int C, void* inout, int H, int W, void* stream) {
  LVC_REQUIRE(n >= 1 && C % 8 == 0 && top && inout, "upsample2_add: bad argument");
  LVC_REQUIRE(H == 2 * Ht && W == 2 * Wt, "upsample2_add: fine level must be exactly 2x the coarse level (fpn.py:131)");
  long long total = (long long)n * H * W * (C / 8);
  upsample2_add_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)top, n, Ht, Wt, C / 8, (uint4*)inout, H, W);
  return check_launch("upsample2_add_kernel");
}
