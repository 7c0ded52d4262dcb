// RoIAlign kernels.
//  (1) reference-layout op: NCHW fp32 in, [R,C,ph,pw] fp32 out  -- the detectron2.layers.roi_align drop-in.
//  (2) fused multi-level pooler over channels-last planes (bf16 / fp32): level assignment + bilinear
//      gather + scatter in one launch; one warp per (roi, bin), channels across lanes as 16-byte vectors,
//      so every corner fetch of a warp is one contiguous 512-byte segment.
// Arithmetic follows detectron2/layers/csrc/ROIAlign/ROIAlign_cuda.cu:12-139 (bilinear_interpolate,
// RoIAlignForward) with torchvision's handling of empty rois (grid 0 -> zeros).
#include "common.cuh"

namespace lvcb200 {

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* roi, float scale, int ph, int pw, int sampling_ratio, bool aligned) {
  RoiGeom g;
  float off = aligned ? 0.5f : 0.0f;
  g.start_w = roi[1] * scale - off;
  g.start_h = roi[2] * scale - off;
  float end_w = roi[3] * scale - off, end_h = roi[4] * scale - off;
  float rw = end_w - g.start_w, rh = end_h - g.start_h;
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  g.bin_h = rh / (float)ph;
  g.bin_w = rw / (float)pw;
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)ph);
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)pw);
  int c = g.grid_h * g.grid_w;
  g.count = (float)(c > 1 ? c : 1);
  return g;
}

struct Tap { int y0, y1, x0, x1; float w1, w2, w3, w4; bool valid; };

__device__ __forceinline__ Tap make_tap(float y, float x, int H, int W) {
  Tap t;
  t.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  t.y0 = yl; t.y1 = yh; t.x0 = xl; t.x1 = xh;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

// ---------------------------------------------------------------- (1) NCHW fp32
__global__ void roi_align_nchw_kernel(const float* __restrict__ in, int C, int H, int W, const float* __restrict__ rois,
                                      int64_t total, int ph_n, int pw_n, float scale, int sampling_ratio, bool aligned,
                                      float* __restrict__ out) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int pw = idx % pw_n;
    int ph = (idx / pw_n) % ph_n;
    int c = (idx / ((int64_t)pw_n * ph_n)) % C;
    int64_t n = idx / ((int64_t)pw_n * ph_n * C);
    const float* roi = rois + n * 5;
    int b = (int)roi[0];
    RoiGeom g = roi_geom(roi, scale, ph_n, pw_n, sampling_ratio, aligned);
    const float* p = in + ((int64_t)b * C + c) * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < g.grid_h; iy++) {
      float y = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
      for (int ix = 0; ix < g.grid_w; ix++) {
        float x = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
        Tap t = make_tap(y, x, H, W);
        if (!t.valid) continue;
        acc += t.w1 * __ldg(p + t.y0 * W + t.x0) + t.w2 * __ldg(p + t.y0 * W + t.x1) +
               t.w3 * __ldg(p + t.y1 * W + t.x0) + t.w4 * __ldg(p + t.y1 * W + t.x1);
      }
    }
    out[idx] = acc / g.count;
  }
}

// ---------------------------------------------------------------- level assignment (poolers.py:51-59)
__device__ __forceinline__ int assign_level(float x1, float y1, float x2, float y2, int min_level, int max_level,
                                            int canon_size, int canon_level) {
  float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  float size = sqrtf(area);
  float lvl = floorf(__fadd_rn((float)canon_level, log2f(__fadd_rn(__fdiv_rn(size, (float)canon_size), 1e-8f))));
  if (lvl != lvl) return -1;  // negative area (box inverted on one axis) -> sqrt NaN: matches no level in the reference
  if (lvl < (float)min_level) lvl = (float)min_level;
  if (lvl > (float)max_level) lvl = (float)max_level;
  return (int)lvl - min_level;
}

__global__ void assign_levels_kernel(const float* __restrict__ boxes, int64_t R, int min_level, int max_level,
                                     int canon_size, int canon_level, int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  float4 b = reinterpret_cast<const float4*>(boxes)[i];
  float area = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  float size = sqrtf(area);
  float lvl = floorf(__fadd_rn((float)canon_level, log2f(__fadd_rn(__fdiv_rn(size, (float)canon_size), 1e-8f))));
  if (lvl < (float)min_level) lvl = (float)min_level;
  if (lvl > (float)max_level) lvl = (float)max_level;
  // NaN follows torch: clamp keeps NaN, .to(int64) of NaN is INT64_MIN on x86; we do the same conversion
  out[i] = (lvl != lvl) ? (int64_t)INT64_MIN - min_level : (int64_t)lvl - min_level;
}

// ---------------------------------------------------------------- (2) fused multi-level pooler, channels-last
struct PoolLevels {
  lvcb200_fmap lv[4];
  int n_levels;
};

template <typename T> struct Vec;
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static uint4 load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
  }
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static uint4 load_raw(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void unpack(const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  __device__ static void load(const float* p, float* v) {
    float4 u = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
};

template <typename TO, int N> __device__ __forceinline__ void store_vec(TO* p, const float* v);
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16, 8>(__nv_bfloat16* p, const float* v) {
  uint4 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <> __device__ __forceinline__ void store_vec<float, 8>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store_vec<float, 4>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store_vec<__nv_bfloat16, 4>(__nv_bfloat16* p, const float* v) {
  uint2 u; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]); h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// 1-D bilinear footprint of one sample coordinate (the y or x half of bilinear_interpolate, ROIAlign_cuda.cu:12-62):
// rows lo / hi with weights w_lo / w_hi; valid = inside [-1, size].
struct Tap1 { int lo, hi; float wlo, whi; bool valid; };
__device__ __forceinline__ Tap1 make_tap1(float y, int size) {
  Tap1 t;
  t.valid = !(y < -1.0f || y > (float)size);
  if (y <= 0) y = 0;
  int yl = (int)y, yh;
  if (yl >= size - 1) { yh = yl = size - 1; y = (float)yl; } else yh = yl + 1;
  float l = y - yl;
  t.lo = yl; t.hi = yh; t.wlo = 1.f - l; t.whi = l;
  return t;
}

constexpr int kMaxSepGrid = 32;
constexpr int kPoolTbl = 128;   // (offset, weight) table entries per warp

// Bilinear weights are separable (w(y,x) = wy(y) * wx(x)) and the sample grid of a bin is a product grid, so
//   sum_{iy,ix} sum_{corners} w * f  ==  sum_{rows} sum_{cols} Wy[row] * Wx[col] * f[row, col]
// with Wy / Wx the per-row / per-column sums of the 1-D weights.  A bin with a g x g sample grid then reads (g+1)^2
// pixel vectors instead of 4 g^2.  One warp owns one (RoI, bin-row): the RoI geometry, the level assignment, Wy and the
// seven Wx tables are computed once and reused for the seven bins of the row (the first version recomputed them per bin and
// was instruction-issue bound at 2 100 instructions per bin, see profiles/).
// Slow generic path for sample grids larger than the separable tables (RoIs wider than 224 feature pixels at their level).
template <typename TI, int V>
__device__ __noinline__ void roi_bin_generic(const TI* base, const lvcb200_fmap& fm, const RoiGeom& g, int gh, int gw, int ph, int pw,
                                             int c0, float* acc) {
  const int H = fm.H, W = fm.W;
  for (int iy = 0; iy < gh; iy++) {
    float y = g.start_h + ph * g.bin_h + ((float)iy + .5f) * g.bin_h / (float)g.grid_h;
    for (int ix = 0; ix < gw; ix++) {
      float x = g.start_w + pw * g.bin_w + ((float)ix + .5f) * g.bin_w / (float)g.grid_w;
      Tap t = make_tap(y, x, H, W);
      if (!t.valid) continue;
      float v1[V], v2[V], v3[V], v4[V];
      Vec<TI>::load(base + ((int64_t)t.y0 * fm.row_stride + t.x0) * fm.c_stride + c0, v1);
      Vec<TI>::load(base + ((int64_t)t.y0 * fm.row_stride + t.x1) * fm.c_stride + c0, v2);
      Vec<TI>::load(base + ((int64_t)t.y1 * fm.row_stride + t.x0) * fm.c_stride + c0, v3);
      Vec<TI>::load(base + ((int64_t)t.y1 * fm.row_stride + t.x1) * fm.c_stride + c0, v4);
#pragma unroll
      for (int i = 0; i < V; i++) acc[i] += t.w1 * v1[i] + t.w2 * v2[i] + t.w3 * v3[i] + t.w4 * v4[i];
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256, 2)
roi_pool_fpn_kernel(PoolLevels L, int C, const float* __restrict__ rois, int64_t R, int P, int sampling_ratio,
                    int canon_size, int canon_level, int min_level, TO* __restrict__ out, int out_layout,
                    int64_t out_pitch, int64_t* __restrict__ levels_out) {
  constexpr int V = Vec<TI>::N;
  constexpr int MAXP = 8, WS = kMaxSepGrid + 4;
  __shared__ float sWy[8][WS], sWx[8][MAXP][WS];
  __shared__ int2 sTbl[8][kPoolTbl];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int bins = P * P;
  for (int64_t item = warp; item < R * P; item += nwarps) {
    const int64_t r = item / P;
    const int ph = (int)(item - r * P);
    const float* roi = rois + r * 5;
    float rx1 = roi[1], ry1 = roi[2], rx2 = roi[3], ry2 = roi[4];
    int lvl = 0;
    if (L.n_levels > 1)
      lvl = assign_level(rx1, ry1, rx2, ry2, min_level, min_level + L.n_levels - 1, canon_size, canon_level);
    if (levels_out != nullptr && ph == 0 && lane == 0) levels_out[r] = lvl;
    const bool no_level = lvl < 0;  // reference: `level_assignments == level` never true -> row stays zero
    if (no_level) lvl = 0;
    const lvcb200_fmap fm = L.lv[lvl];
    const int H = fm.H, W = fm.W;
    RoiGeom g = roi_geom(roi, fm.spatial_scale, P, P, sampling_ratio, true);
    const int gh = no_level ? 0 : g.grid_h, gw = no_level ? 0 : g.grid_w;
    const TI* base = reinterpret_cast<const TI*>(fm.base) + (int64_t)roi[0] * fm.img_stride * fm.c_stride;
    const bool separable = gh >= 1 && gw >= 1 && gh <= kMaxSepGrid && gw <= kMaxSepGrid && P <= MAXP;
    const float inv_count = 1.0f / g.count;
    int ybase = 0, ny = 0;
    const float ystep = gh > 0 ? g.bin_h / (float)gh : 0.f, xstep = gw > 0 ? g.bin_w / (float)gw : 0.f;
    const float y_first = g.start_h + ph * g.bin_h;
    if (separable) {
      __syncwarp();
      for (int i = lane; i < WS; i += 32) sWy[wib][i] = 0.f;
      for (int i = lane; i < MAXP * WS; i += 32) (&sWx[wib][0][0])[i] = 0.f;
      ybase = make_tap1(y_first + .5f * ystep, H).lo;
      ny = make_tap1(y_first + ((float)(gh - 1) + .5f) * ystep, H).hi - ybase + 1;
      __syncwarp();
      if (lane < gh) {
        Tap1 t = make_tap1(y_first + ((float)lane + .5f) * ystep, H);
        if (t.valid) { atomicAdd(&sWy[wib][t.lo - ybase], t.wlo); atomicAdd(&sWy[wib][t.hi - ybase], t.whi); }
      }
      for (int sidx = lane; sidx < P * gw; sidx += 32) {
        const int pw = sidx / gw, i = sidx - pw * gw;
        const float x_first = g.start_w + pw * g.bin_w;
        const int xb = make_tap1(x_first + .5f * xstep, W).lo;
        Tap1 t = make_tap1(x_first + ((float)i + .5f) * xstep, W);
        if (t.valid) { atomicAdd(&sWx[wib][pw][t.lo - xb], t.wlo); atomicAdd(&sWx[wib][pw][t.hi - xb], t.whi); }
      }
      __syncwarp();
    }
    for (int cb = 0; cb < C; cb += 32 * V) {
      const int c0 = cb + lane * V;
      const bool act = c0 < C;              // all lanes stay in the loop (table building and __syncwarp are warp-wide)
      for (int pw = 0; pw < P; pw++) {
        const int bin = ph * P + pw;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; i++) acc[i] = 0.f;
        if (separable) {
          const float x_first = g.start_w + pw * g.bin_w;
          const int xbase = make_tap1(x_first + .5f * xstep, W).lo;
          const int nx = make_tap1(x_first + ((float)(gw - 1) + .5f) * xstep, W).hi - xbase + 1;
          // the ny x nx footprint of the bin, flattened: the lanes first build a table of (element offset, weight) per footprint
          // pixel in shared memory (one entry per lane per pass), then every lane walks it with LB independent 16-byte loads in
          // flight -- one LDS.64 + one address add per load.  (Tracking (row, col) per load in registers cost ~25 integer /
          // predicate instructions per load and made this kernel issue-bound: 1 380 instructions per bin, profiles/.)
          const TI* binp = base + ((int64_t)ybase * fm.row_stride + xbase) * fm.c_stride + c0;
          const int total = ny * nx;
          const int cs = (int)fm.c_stride, rs = (int)fm.row_stride * cs;
          const float* wyp = sWy[wib];
          const float* wxp = sWx[wib][pw];
          constexpr int LB = 8;    // loads in flight per lane
          for (int ch0 = 0; ch0 < total; ch0 += kPoolTbl) {
            const int n = total - ch0 < kPoolTbl ? total - ch0 : kPoolTbl;
            const int npad = (n + LB - 1) / LB * LB;
            __syncwarp();
            for (int e = lane; e < npad; e += 32) {
              const int i = ch0 + (e < n ? e : n - 1);      // padding entries re-read the last pixel with weight 0
              const int ry = i / nx, rx = i - ry * nx;
              sTbl[wib][e] = make_int2(ry * rs + rx * cs, __float_as_int(e < n ? wyp[ry] * wxp[rx] : 0.f));
            }
            __syncwarp();
            float2 acc2[V / 2];                 // packed fp32x2 FMAs (FFMA2, sm_100): half the FMA issue slots of this issue-bound loop
#pragma unroll
            for (int i = 0; i < V / 2; i++) acc2[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
            for (int i0 = 0; act && i0 < npad; i0 += LB) {
              uint4 raw[LB];
              float w[LB];
#pragma unroll
              for (int u = 0; u < LB; u++) {
                const int2 t = sTbl[wib][i0 + u];
                w[u] = __int_as_float(t.y);
                raw[u] = Vec<TI>::load_raw(binp + t.x);
              }
#pragma unroll
              for (int u = 0; u < LB; u++) {
                float v[V];
                Vec<TI>::unpack(raw[u], v);
                const float2 ww = make_float2(w[u], w[u]);
#pragma unroll
                for (int i = 0; i < V / 2; i++) acc2[i] = __ffma2_rn(ww, make_float2(v[2 * i], v[2 * i + 1]), acc2[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < V / 2; i++) { acc[2 * i] = acc2[i].x; acc[2 * i + 1] = acc2[i].y; }
          }
        } else if (act) {
          roi_bin_generic<TI, V>(base, fm, g, gh, gw, ph, pw, c0, acc);
        }
        if (!act) continue;
#pragma unroll
        for (int i = 0; i < V; i++) acc[i] = acc[i] * inv_count;
        if (out_layout == LVCB200_OUT_NHWC) {
          store_vec<TO, V>(out + r * out_pitch + (int64_t)bin * C + c0, acc);
        } else {
#pragma unroll
          for (int i = 0; i < V; i++) out[r * out_pitch + (int64_t)(c0 + i) * bins + bin] = (TO)acc[i];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_roi_align_nchw_f32(const float* input, int N, int C, int H, int W, const float* rois, int R,
                                          int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio,
                                          int aligned, float* output, void* stream) {
  LVC_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, "roi_align: bad shape");
  int64_t total = (int64_t)R * C * pooled_h * pooled_w;
  if (total == 0) return 0;
  LVC_REQUIRE(input && rois && output, "roi_align: NULL pointer");
  int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  if (blocks > kNumSMs * 64) blocks = kNumSMs * 64;
  roi_align_nchw_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
      input, C, H, W, rois, total, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned != 0, output);
  return check_launch("roi_align_nchw_kernel");
}

extern "C" int lvcb200_assign_boxes_to_levels(const float* boxes, int64_t R, int min_level, int max_level,
                                              int canonical_box_size, int canonical_level, int64_t* levels, void* stream) {
  if (R == 0) return 0;
  LVC_REQUIRE(boxes && levels && ((uintptr_t)boxes % 16 == 0), "assign_levels: NULL or unaligned boxes");
  assign_levels_kernel<<<(unsigned)ceil_div64(R, 256), 256, 0, (cudaStream_t)stream>>>(
      boxes, R, min_level, max_level, canonical_box_size, canonical_level, levels);
  return check_launch("assign_levels_kernel");
}

extern "C" int lvcb200_roi_pool_fpn(const lvcb200_fmap* levels, int n_levels, int in_dtype, int C, const float* rois,
                                    int64_t R, int pooled, int sampling_ratio, int canonical_box_size,
                                    int canonical_level, int min_level, void* out, int out_dtype, int out_layout,
                                    int64_t out_pitch, int64_t* levels_out, void* stream) {
  LVC_REQUIRE(n_levels >= 1 && n_levels <= 4 && levels, "roi_pool_fpn: 1..4 levels");
  if (R == 0) return 0;
  int V = in_dtype == LVCB200_BF16 ? 8 : 4;
  LVC_REQUIRE(C % V == 0, "roi_pool_fpn: C must be a multiple of the 16-byte vector width");
  LVC_REQUIRE(out_pitch >= (int64_t)C * pooled * pooled, "roi_pool_fpn: out_pitch too small");
  PoolLevels L;
  L.n_levels = n_levels;
  for (int i = 0; i < n_levels; i++) {
    L.lv[i] = levels[i];
    LVC_REQUIRE(((uintptr_t)levels[i].base % 16) == 0 && levels[i].c_stride % V == 0, "roi_pool_fpn: level not 16B aligned");
  }
  int threads = 256;
  int64_t warps = R * pooled;   // one warp per (RoI, bin row)
  int64_t blocks = ceil_div64(warps, threads / 32);
  if (blocks > (int64_t)kNumSMs * 256) blocks = (int64_t)kNumSMs * 256;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(TI, TO)                                                                                              \
  roi_pool_fpn_kernel<TI, TO><<<(unsigned)blocks, threads, 0, s>>>(L, C, rois, R, pooled, sampling_ratio,           \
                                                                   canonical_box_size, canonical_level, min_level, \
                                                                   (TO*)out, out_layout, out_pitch, levels_out)
  if (in_dtype == LVCB200_BF16 && out_dtype == LVCB200_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == LVCB200_BF16 && out_dtype == LVCB200_F32) LAUNCH(__nv_bfloat16, float);
  else if (in_dtype == LVCB200_F32 && out_dtype == LVCB200_F32) LAUNCH(float, float);
  else if (in_dtype == LVCB200_F32 && out_dtype == LVCB200_BF16) LAUNCH(float, __nv_bfloat16);
  else return set_error(LVCB200_EINVAL, "roi_pool_fpn: bad dtype");
#undef LAUNCH
  return check_launch("roi_pool_fpn_kernel");
}
