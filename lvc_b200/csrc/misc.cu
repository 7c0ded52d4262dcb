// Memory-bound helpers of the conv engine.  Activations are "zero-bordered channels-last planes":
// bf16 [n, H+2, W+2, C] with a one-pixel zero frame, so that a 3x3 conv is a GEMM over row-shifted views of the
// same matrix (gemm_tc.cu) and no kernel ever needs an explicit halo test.  All kernels move 16-byte vectors.
#include "common.cuh"

namespace lvcb200 {

// (x - mean) / std + zero pad to /32 (rcnn.py:324-333, image_list.py:57-119) fused with a 4x4 space-to-depth:
//   X4[n, y4, x4, (iy*4 + ix)*3 + c] = norm(img[c, 4*y4 + iy, 4*x4 + ix]),  48 channels padded to 64, zero-bordered plane.
// On X4 the 7x7/stride-2/pad-3 stem conv (resnet.py:588-590) is a 3x3 stride-1 conv with 64 input channels per tap and
// 4 x 64 output channels (the 2x2 output pixels of each X4 cell), i.e. exactly the shift-GEMM the rest of the network
// uses -- no im2col matrix (the 832 MB / batch the first version wrote) is ever materialised.
// PAIR (strict mode): the normalised pixel is written as a bf16 pair, hi plane at `out`, lo plane lo_off uint4s further on.
template <typename T, bool PAIR = false>
__global__ void __launch_bounds__(256)
stem_s2d4_kernel(const T* const* __restrict__ images, const int32_t* __restrict__ image_sizes, int n, int H4, int W4,
                 const float* __restrict__ mean, const float* __restrict__ inv_std, uint4* __restrict__ out, long long lo_off = 0) {
  const int PW = W4 + 2, PH = H4 + 2;
  const long long total = (long long)n * PH * PW;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int img = (int)(i / (PH * PW));
  const int rem = (int)(i - (long long)img * PH * PW);
  const int py = rem / PW, px = rem - py * PW;
  uint4* o = out + i * 8;   // 64 bf16 = 8 x 16 bytes
  __nv_bfloat16 v[64];
  [[maybe_unused]] __nv_bfloat16 vl[PAIR ? 64 : 1];
#pragma unroll
  for (int k = 0; k < 64; k++) v[k] = __float2bfloat16_rn(0.f);
  if constexpr (PAIR) {
#pragma unroll
    for (int k = 0; k < 64; k++) vl[k] = __float2bfloat16_rn(0.f);
  }
  if (py >= 1 && py <= H4 && px >= 1 && px <= W4) {
    const int H = image_sizes[img * 2], W = image_sizes[img * 2 + 1];
    const T* im = images[img];
    const int y0 = 4 * (py - 1), x0 = 4 * (px - 1);
    const float m0 = mean[0], m1 = mean[1], m2 = mean[2], s0 = inv_std[0], s1 = inv_std[1], s2 = inv_std[2];
#pragma unroll
    for (int iy = 0; iy < 4; iy++) {
      const int y = y0 + iy;
      if (y >= H) continue;
#pragma unroll
      for (int ix = 0; ix < 4; ix++) {
        const int x = x0 + ix;
        if (x >= W) continue;
        const long long off = (long long)y * W + x;
        const long long cs = (long long)H * W;
        const float f0 = ((float)im[off] - m0) * s0, f1 = ((float)im[off + cs] - m1) * s1, f2 = ((float)im[off + 2 * cs] - m2) * s2;
        const int k0 = (iy * 4 + ix) * 3;
        v[k0 + 0] = __float2bfloat16_rn(f0);
        v[k0 + 1] = __float2bfloat16_rn(f1);
        v[k0 + 2] = __float2bfloat16_rn(f2);
        if constexpr (PAIR) {
          vl[k0 + 0] = __float2bfloat16_rn(f0 - __bfloat162float(v[k0 + 0]));
          vl[k0 + 1] = __float2bfloat16_rn(f1 - __bfloat162float(v[k0 + 1]));
          vl[k0 + 2] = __float2bfloat16_rn(f2 - __bfloat162float(v[k0 + 2]));
        }
      }
    }
  }
  const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int k = 0; k < 8; k++) o[k] = src[k];
  if constexpr (PAIR) {
    const uint4* srcl = reinterpret_cast<const uint4*>(vl);
#pragma unroll
    for (int k = 0; k < 8; k++) o[lo_off + k] = srcl[k];
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

// F.max_pool2d(k=3, s=2, p=1) (resnet.py:591) reading the stem output in its space-to-depth layout
//   S2[n, y2, x2, ((Y&1)*2 + (X&1))*C + c] = stem[n, Y, X, c],  y2 = Y>>1, x2 = X>>1  (zero-bordered plane of the pooled grid)
// out[n, y, x, c] = max_{dy,dx in -1..1} stem[2y+dy, 2x+dx, c].  Input is post-ReLU (>= 0): the zero frame equals -inf padding.
__global__ void maxpool_s2d_kernel(const uint4* __restrict__ in, int n, int Ho, int Wo, int CV, uint4* __restrict__ out) {
  const long long total = (long long)n * (Ho + 2) * (Wo + 2) * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int px = (int)(pix % (Wo + 2)), py = (int)((pix / (Wo + 2)) % (Ho + 2)), img = (int)(pix / ((long long)(Wo + 2) * (Ho + 2)));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (py >= 1 && py <= Ho && px >= 1 && px <= Wo) {
      const uint4* base = in + ((long long)img * (Ho + 2) * (Wo + 2)) * (4 * CV);
      bool first = true;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          // stem pixel (2*(py-1)+dy, 2*(px-1)+dx): cell = floor(/2) (+1 for the frame), sub-position = parity
          int Y = 2 * (py - 1) + dy, X = 2 * (px - 1) + dx;
          int cy = (Y + 2) / 2, cx = (X + 2) / 2;      // (Y>>1) + 1 for Y >= -1
          int sub = ((Y & 1) * 2 + (X & 1));
          uint4 v = __ldg(base + ((long long)cy * (Wo + 2) + cx) * (4 * CV) + sub * CV + cv);
          r = first ? v : bf16x8_max(r, v);
          first = false;
        }
    }
    out[i] = r;
  }
}

// out[n, oy, ox] = in[n, 2*oy, 2*ox]  (1x1 stride-2 conv input side; LastLevelMaxPool k=1 s=2)
// grid (ceil((Wo + 2) * CV / 256), n * (Ho + 2)): a block row per output plane row, 32-bit index arithmetic only (the grid-stride form
// below spends ~100 instructions of 64-bit division per 16-byte vector).
__global__ void __launch_bounds__(256)
subsample2_rows_kernel(const uint4* __restrict__ in, int H, int W, int CV, uint4* __restrict__ out, int Ho, int Wo) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (Wo + 2) * CV) return;
  const int px = t / CV, cv = t - px * CV;
  const int row = blockIdx.y, img = row / (Ho + 2), py = row - img * (Ho + 2);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (py >= 1 && py <= Ho && px >= 1 && px <= Wo)
    r = __ldg(in + (((long long)img * (H + 2) + 2 * py - 1) * (W + 2) + 2 * px - 1) * CV + cv);
  out[((long long)row * (Wo + 2) + px) * CV + cv] = r;
}

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

// ------------------------------------------------------------------------------------------ strict mode: bf16 hi/lo pairs
// A pair plane stores x as hi = bf16(x) and lo = bf16(x - hi) in two bf16 planes `lo_off` vectors apart (gemm_tc.cu, SPLIT).
__device__ __forceinline__ void pair_to_f32(const uint4& h, const uint4& l, float* v) {
  const uint32_t* hw = &h.x; const uint32_t* lw = &l.x;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void f32_to_pair(const float* v, uint4& h, uint4& l) {
  __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&h);
  __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    hh[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __bfloat1622float2(hh[i]);
    ll[i] = __floats2bfloat162_rn(__fsub_rn(v[2 * i], f.x), __fsub_rn(v[2 * i + 1], f.y));
  }
}

__global__ void maxpool_s2d_pair_kernel(const uint4* __restrict__ in, long long in_lo, int n, int Ho, int Wo, int CV,
                                        uint4* __restrict__ out, long long out_lo) {
  const long long total = (long long)n * (Ho + 2) * (Wo + 2) * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int px = (int)(pix % (Wo + 2)), py = (int)((pix / (Wo + 2)) % (Ho + 2)), img = (int)(pix / ((long long)(Wo + 2) * (Ho + 2)));
    float best[8];
#pragma unroll
    for (int k = 0; k < 8; k++) best[k] = 0.f;       // post-ReLU input: 0 == the -inf padding of F.max_pool2d
    if (py >= 1 && py <= Ho && px >= 1 && px <= Wo) {
      const uint4* base = in + ((long long)img * (Ho + 2) * (Wo + 2)) * (4 * CV);
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          int Y = 2 * (py - 1) + dy, X = 2 * (px - 1) + dx;
          int cy = (Y + 2) / 2, cx = (X + 2) / 2;
          int sub = ((Y & 1) * 2 + (X & 1));
          const uint4* q = base + ((long long)cy * (Wo + 2) + cx) * (4 * CV) + sub * CV + cv;
          float v[8];
          pair_to_f32(__ldg(q), __ldg(q + in_lo), v);
#pragma unroll
          for (int k = 0; k < 8; k++) best[k] = fmaxf(best[k], v[k]);
        }
    }
    uint4 h, l;
    f32_to_pair(best, h, l);
    out[i] = h;
    out[i + out_lo] = l;
  }
}

__global__ void upsample2_add_pair_kernel(const uint4* __restrict__ top, long long top_lo, int n, int Ht, int Wt, int CV,
                                          uint4* __restrict__ io, long long io_lo, int H, int W) {
  const long long total = (long long)n * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long pix = i / CV;
    int x = (int)(pix % W), y = (int)((pix / W) % H), img = (int)(pix / ((long long)W * H));
    long long dst = (((long long)img * (H + 2) + y + 1) * (W + 2) + x + 1) * CV + cv;
    long long src = (((long long)img * (Ht + 2) + (y >> 1) + 1) * (Wt + 2) + (x >> 1) + 1) * CV + cv;
    float a[8], b[8];
    pair_to_f32(io[dst], io[dst + io_lo], a);
    pair_to_f32(__ldg(top + src), __ldg(top + src + top_lo), b);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] += b[k];
    uint4 h, l;
    f32_to_pair(a, h, l);
    io[dst] = h;
    io[dst + io_lo] = l;
  }
}

// pair -> fp32 and fp32 -> pair over a [rows, cols] matrix (8 elements per thread)
__global__ void pair_merge_kernel(const uint4* __restrict__ pair, long long lo_off, long long nvec, float4* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    pair_to_f32(__ldg(pair + i), __ldg(pair + i + lo_off), v);
    out[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
    out[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
__global__ void pair_split_kernel(const float4* __restrict__ in, long long nvec, uint4* __restrict__ pair, long long lo_off) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 h, l;
    f32_to_pair(v, h, l);
    pair[i] = h;
    pair[i + lo_off] = l;
  }
}

// out[r] = scale / (||x_r||_2 + eps): the per-row factor of CosineSimOutputLayers (fast_rcnn.py:826-829), one warp per row.
// x is bf16 (lo_off == 0), a bf16 pair (lo_off != 0, elements) or fp32.
template <typename T>
__global__ void row_inv_norm_kernel(const T* __restrict__ x, long long lo_off, long long R, int C, long long ld, float scale, float eps,
                                    float* __restrict__ out) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const T* p = x + row * ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = (float)p[c];
    if (lo_off != 0) v += (float)p[c + lo_off];
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = scale / (sqrtf(s) + eps);
}

// rois [n*P, 5] = (image, x1, y1, x2, y2) and roi_image [n*P] (image index, -1 for padding rows >= counts[image]) from the
// padded proposal block [n, P, 4] (convert_boxes_to_pooler_format, poolers.py:69-96, on the RPN's fixed-size output)
__global__ void make_rois_kernel(const float4* __restrict__ props, const int32_t* __restrict__ counts, int n, int P,
                                 float* __restrict__ rois, int32_t* __restrict__ roi_image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * P) return;
  const int img = i / P, j = i - img * P;
  const float4 b = props[i];
  float* r = rois + (long long)i * 5;
  r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
  roi_image[i] = j < counts[img] ? img : -1;
}

// get_crops_qe (lvc/data/utils.py:485-519) + preprocess_crops (tools/run_nearest_neighbours.py:102-105): per box, crop
// (optionally with square context), zero-pad to a square, nearest-resize to S x S, (x - mean) / std.  Pure index arithmetic:
// the crop geometry (python slicing clamps, get_padding's half-pixel rule) is computed per box on the host side of the call.
struct CropGeom { int y0, x0, ah, aw, tp, lp, Hp, Wp; };   // source origin, actual crop extent, top/left padding, padded extent

template <typename T>
__global__ void crops_qe_kernel(const T* __restrict__ img, int H, int W, const CropGeom* __restrict__ geom, int n, int S,
                                const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ out) {
  const long long total = (long long)n * 3 * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % S), oy = (int)((i / S) % S), c = (int)((i / ((long long)S * S)) % 3), b = (int)(i / ((long long)3 * S * S));
    const CropGeom g = geom[b];
    // F.interpolate(mode='nearest'): src = min(floor(dst * (in / out)), in - 1), scale in fp32
    const float sy = (float)g.Hp / (float)S, sx = (float)g.Wp / (float)S;
    int py = (int)floorf((float)oy * sy); py = py < g.Hp - 1 ? py : g.Hp - 1;
    int px = (int)floorf((float)ox * sx); px = px < g.Wp - 1 ? px : g.Wp - 1;
    const int cy = py - g.tp, cx = px - g.lp;          // position inside the (unpadded) crop
    float v = 0.f;
    if (cy >= 0 && cy < g.ah && cx >= 0 && cx < g.aw) v = (float)img[((long long)c * H + g.y0 + cy) * W + g.x0 + cx];
    if (mean != nullptr) v = (v - mean[c]) * inv_std[c];
    out[i] = v;
  }
}

// S % 4 == 0 fast path: grid (ceil(S * S / 4 / 256), 3, boxes): a thread produces four neighbouring output pixels of one (box, channel)
// plane (one 16-byte store); the box geometry, the two fp32 scale divisions and the row's source line are computed once per thread instead of
// once per pixel, and no 64-bit index arithmetic is left.  Same arithmetic per pixel as crops_qe_kernel: identical output.
template <typename T>
__global__ void __launch_bounds__(256)
crops_qe4_kernel(const T* __restrict__ img, int H, int W, const CropGeom* __restrict__ geom, int n, int S,
                 const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ out) {
  const int q4 = S >> 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S * q4) return;
  const int oy = t / q4, ox0 = (t - oy * q4) * 4;
  const int c = blockIdx.y;
  const float m = mean != nullptr ? mean[c] : 0.f, is = mean != nullptr ? inv_std[c] : 1.f;
  for (int b = blockIdx.z; b < n; b += gridDim.z) {
    const CropGeom g = geom[b];
    const float sy = (float)g.Hp / (float)S, sx = (float)g.Wp / (float)S;
    int py = (int)floorf((float)oy * sy); py = py < g.Hp - 1 ? py : g.Hp - 1;
    const int cy = py - g.tp;
    const bool row_in = cy >= 0 && cy < g.ah;
    const T* line = img + ((long long)c * H + g.y0 + (row_in ? cy : 0)) * W + g.x0;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      int px = (int)floorf((float)(ox0 + e) * sx); px = px < g.Wp - 1 ? px : g.Wp - 1;
      const int cx = px - g.lp;
      float x = 0.f;
      if (row_in && cx >= 0 && cx < g.aw) x = (float)line[cx];
      if (mean != nullptr) x = (x - m) * is;
      v[e] = x;
    }
    *reinterpret_cast<float4*>(out + (((long long)b * 3 + c) * S + oy) * S + ox0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// The bank exchange of the kNN (run_nearest_neighbours.py:303-309) by our own kernel over NVLink peer memory: every rank has written
// its padded shard [1 + cap, row_floats] (header row first) into a buffer that is mapped into all ranks (torch symmetric memory); after the
// device-side barrier this kernel reads the peers' rows with plain loads through NVLink / NVSwitch and writes the rank-major concatenation
// [total, row_floats] locally -- no NCCL launch, no staging copy.  blockIdx.y = source rank.
struct PeerRows { const float* buf[16]; int rows[16]; int offset[16]; };
__global__ void __launch_bounds__(256)
gather_rows_p2p_kernel(PeerRows pr, int row_floats, float* __restrict__ out) {
  const int r = blockIdx.y;
  const long long nv = (long long)pr.rows[r] * row_floats / 2;                 // float2 vectors (row_floats is even: D + 2)
  const float2* src = reinterpret_cast<const float2*>(pr.buf[r] + row_floats);   // skip the header row
  float2* dst = reinterpret_cast<float2*>(out + (long long)pr.offset[r] * row_floats);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace lvcb200

using namespace lvcb200;

static inline unsigned grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 32;
  return (unsigned)(b < cap ? b : cap);
}

extern "C" int lvcb200_stem_s2d4(const void* const* images, int image_dtype, const int32_t* image_sizes, int n, int Hpad, int Wpad,
                                 const float* mean, const float* inv_std, void* out, void* stream) {
  LVC_REQUIRE(n >= 1 && Hpad % 4 == 0 && Wpad % 4 == 0, "stem_s2d4: padded size must be a multiple of 4");
  LVC_REQUIRE(images && image_sizes && mean && inv_std && out, "stem_s2d4: NULL pointer");
  const int H4 = Hpad / 4, W4 = Wpad / 4;
  long long total = (long long)n * (H4 + 2) * (W4 + 2);
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (image_dtype == LVCB200_F32)
    stem_s2d4_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float* const*)images, image_sizes, n, H4, W4, mean, inv_std, (uint4*)out);
  else if (image_dtype == LVCB200_U8)
    stem_s2d4_kernel<unsigned char><<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char* const*)images, image_sizes, n, H4, W4, mean, inv_std, (uint4*)out);
  else
    return set_error(LVCB200_EINVAL, "stem_s2d4: image dtype must be LVCB200_F32 or LVCB200_U8");
  return check_launch("stem_s2d4_kernel");
}

extern "C" int lvcb200_stem_s2d4_pair(const void* const* images, int image_dtype, const int32_t* image_sizes, int n, int Hpad, int Wpad,
                                      const float* mean, const float* inv_std, void* out, int64_t split_rows, void* stream) {
  LVC_REQUIRE(n >= 1 && Hpad % 4 == 0 && Wpad % 4 == 0, "stem_s2d4_pair: padded size must be a multiple of 4");
  LVC_REQUIRE(images && image_sizes && mean && inv_std && out, "stem_s2d4_pair: NULL pointer");
  const int H4 = Hpad / 4, W4 = Wpad / 4;
  long long total = (long long)n * (H4 + 2) * (W4 + 2);
  LVC_REQUIRE(split_rows >= total, "stem_s2d4_pair: split_rows smaller than the plane");
  unsigned blocks = (unsigned)((total + 255) / 256);
  const long long lo = split_rows * 8;   // 64 bf16 per row = 8 uint4
  if (image_dtype == LVCB200_F32)
    stem_s2d4_kernel<float, true><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float* const*)images, image_sizes, n, H4, W4, mean, inv_std, (uint4*)out, lo);
  else if (image_dtype == LVCB200_U8)
    stem_s2d4_kernel<unsigned char, true><<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char* const*)images, image_sizes, n, H4, W4, mean, inv_std, (uint4*)out, lo);
  else
    return set_error(LVCB200_EINVAL, "stem_s2d4_pair: image dtype must be LVCB200_F32 or LVCB200_U8");
  return check_launch("stem_s2d4_kernel<pair>");
}

extern "C" int lvcb200_maxpool_s2d_pair(const void* in, int64_t in_split_rows, int n, int Ho, int Wo, int C, void* out, int64_t out_split_rows,
                                        void* stream) {
  LVC_REQUIRE(n >= 1 && Ho >= 1 && Wo >= 1 && C % 8 == 0 && in && out, "maxpool_s2d_pair: bad argument");
  long long total = (long long)n * (Ho + 2) * (Wo + 2) * (C / 8);
  maxpool_s2d_pair_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, in_split_rows * (4 * C / 8), n, Ho, Wo, C / 8,
                                                                                  (uint4*)out, out_split_rows * (C / 8));
  return check_launch("maxpool_s2d_pair_kernel");
}

extern "C" int lvcb200_upsample2_add_pair(const void* top, int64_t top_split_rows, int n, int Ht, int Wt, int C, void* inout,
                                          int64_t io_split_rows, int H, int W, void* stream) {
  LVC_REQUIRE(n >= 1 && C % 8 == 0 && top && inout, "upsample2_add_pair: bad argument");
  LVC_REQUIRE(H == 2 * Ht && W == 2 * Wt, "upsample2_add_pair: fine level must be exactly 2x the coarse level (fpn.py:131)");
  long long total = (long long)n * H * W * (C / 8);
  upsample2_add_pair_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)top, top_split_rows * (C / 8), n, Ht, Wt, C / 8,
                                                                                    (uint4*)inout, io_split_rows * (C / 8), H, W);
  return check_launch("upsample2_add_pair_kernel");
}

extern "C" int lvcb200_pair_merge(const void* pair, int64_t split_rows, int64_t rows, int cols, float* out, void* stream) {
  if (rows == 0) return 0;
  LVC_REQUIRE(pair && out && cols % 8 == 0 && split_rows >= rows, "pair_merge: bad argument");
  long long nvec = rows * (cols / 8);
  pair_merge_kernel<<<grid_for(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)pair, split_rows * (cols / 8), nvec, (float4*)out);
  return check_launch("pair_merge_kernel");
}

extern "C" int lvcb200_pair_split(const float* in, int64_t rows, int cols, void* pair, int64_t split_rows, void* stream) {
  if (rows == 0) return 0;
  LVC_REQUIRE(pair && in && cols % 8 == 0 && split_rows >= rows, "pair_split: bad argument");
  long long nvec = rows * (cols / 8);
  pair_split_kernel<<<grid_for(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)in, nvec, (uint4*)pair, split_rows * (cols / 8));
  return check_launch("pair_split_kernel");
}

extern "C" int lvcb200_row_inv_norm(const void* x, int dtype, int64_t lo_off, int64_t R, int C, int64_t ld, float scale, float eps, float* out,
                                    void* stream) {
  if (R == 0) return 0;
  LVC_REQUIRE(x && out && C > 0 && ld >= C, "row_inv_norm: bad argument");
  const unsigned blocks = (unsigned)((R * 32 + 255) / 256);
  if (dtype == LVCB200_BF16)
    row_inv_norm_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, lo_off, R, C, ld, scale, eps, out);
  else if (dtype == LVCB200_F32 && lo_off == 0)
    row_inv_norm_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, 0, R, C, ld, scale, eps, out);
  else
    return set_error(LVCB200_EINVAL, "row_inv_norm: dtype must be LVCB200_BF16 (optionally a pair) or LVCB200_F32");
  return check_launch("row_inv_norm_kernel");
}

extern "C" int lvcb200_make_rois(const float* proposals, const int32_t* counts, int n, int P, float* rois, int32_t* roi_image, void* stream) {
  if (n * P == 0) return 0;
  LVC_REQUIRE(proposals && counts && rois && roi_image && ((uintptr_t)proposals % 16) == 0, "make_rois: bad argument");
  make_rois_kernel<<<(n * P + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)proposals, counts, n, P, rois, roi_image);
  return check_launch("make_rois_kernel");
}

extern "C" int lvcb200_maxpool_s2d(const void* in, int n, int Ho, int Wo, int C, void* out, void* stream) {
  LVC_REQUIRE(n >= 1 && Ho >= 1 && Wo >= 1 && C % 8 == 0 && in && out, "maxpool_s2d: bad argument");
  long long total = (long long)n * (Ho + 2) * (Wo + 2) * (C / 8);
  maxpool_s2d_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, n, Ho, Wo, C / 8, (uint4*)out);
  return check_launch("maxpool_s2d_kernel");
}

extern "C" int lvcb200_subsample2(const void* in, int n, int H, int W, int C, void* out, void* stream) {
  LVC_REQUIRE(n >= 1 && H >= 1 && W >= 1 && C % 8 == 0 && in && out, "subsample2: bad argument");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if ((long long)n * (Ho + 2) <= 65535) {
    const dim3 grid((unsigned)(((Wo + 2) * (C / 8) + 255) / 256), (unsigned)(n * (Ho + 2)));
    subsample2_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)in, H, W, C / 8, (uint4*)out, Ho, Wo);
    return check_launch("subsample2_rows_kernel");
  }
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

extern "C" int lvcb200_crops_qe(const void* image, int image_dtype, int H, int W, const int32_t* geom /*device [n,8]*/, int n, int S,
                                const float* mean, const float* inv_std, float* out, void* stream) {
  if (n == 0) return 0;
  LVC_REQUIRE(image && geom && out && H > 0 && W > 0 && S > 0, "crops_qe: bad argument");
  LVC_REQUIRE((mean == nullptr) == (inv_std == nullptr), "crops_qe: mean and inv_std go together");
  if (S % 4 == 0 && S <= 4096 && ((uintptr_t)out % 16) == 0 && (image_dtype == LVCB200_U8 || image_dtype == LVCB200_F32)) {
    const dim3 grid((unsigned)((S * (S / 4) + 255) / 256), 3, (unsigned)(n < 32768 ? n : 32768));
    if (image_dtype == LVCB200_U8)
      crops_qe4_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)image, H, W, (const CropGeom*)geom, n, S, mean, inv_std, out);
    else
      crops_qe4_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)image, H, W, (const CropGeom*)geom, n, S, mean, inv_std, out);
    return check_launch("crops_qe4_kernel");
  }
  long long total = (long long)n * 3 * S * S;
  if (image_dtype == LVCB200_U8)
    crops_qe_kernel<unsigned char><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)image, H, W, (const CropGeom*)geom, n, S, mean, inv_std, out);
  else if (image_dtype == LVCB200_F32)
    crops_qe_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)image, H, W, (const CropGeom*)geom, n, S, mean, inv_std, out);
  else
    return set_error(LVCB200_EINVAL, "crops_qe: image dtype must be LVCB200_F32 or LVCB200_U8");
  return check_launch("crops_qe_kernel");
}

extern "C" int lvcb200_gather_rows_p2p(const void* const* peer_bufs /*host array of W device pointers*/, int W, const int32_t* rows /*host [W]*/,
                                       int row_floats, void* out, void* stream) {
  LVC_REQUIRE(W >= 1 && W <= 16 && peer_bufs && rows && out && row_floats >= 2 && row_floats % 2 == 0, "gather_rows_p2p: bad argument (at most 16 ranks)");
  PeerRows pr;
  int off = 0, most = 0;
  for (int r = 0; r < 16; r++) { pr.buf[r] = nullptr; pr.rows[r] = 0; pr.offset[r] = 0; }
  for (int r = 0; r < W; r++) {
    LVC_REQUIRE(peer_bufs[r] && rows[r] >= 0 && ((uintptr_t)peer_bufs[r] % 8) == 0, "gather_rows_p2p: NULL / misaligned peer buffer");
    pr.buf[r] = (const float*)peer_bufs[r]; pr.rows[r] = rows[r]; pr.offset[r] = off;
    off += rows[r];
    most = rows[r] > most ? rows[r] : most;
  }
  if (off == 0) return 0;
  LVC_REQUIRE(((uintptr_t)out % 8) == 0, "gather_rows_p2p: out must be 8-byte aligned");
  const long long nv = (long long)most * row_floats / 2;
  unsigned gx = (unsigned)((nv + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  gather_rows_p2p_kernel<<<dim3(gx, W), 256, 0, (cudaStream_t)stream>>>(pr, row_floats, (float*)out);
  return check_launch("gather_rows_p2p_kernel");
}
