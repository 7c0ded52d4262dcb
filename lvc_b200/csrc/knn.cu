// kNN label verification (tools/run_nearest_neighbours.py:142-162, 214-227).
//   knn_prepare: bank mean (crop_mean, :144) and the centred, L2-normalised bank rows (F.cosine_similarity's
//                clamp of the norm at eps = 1e-8), done once after the support bank is gathered.
//   knn_verify : CTA per 32 queries; centred query tile and bank tile staged in shared memory, 4x4 register
//                tiles, fp32 FMA; running top-k per query in shared memory; votes / mode / keep fused at the end.
// Round-1 kernel is fp32 SIMT (exact-arithmetic friendly); the tcgen05 path for the contraction is planned.
#include "common.cuh"

namespace lvcb200 {

constexpr int QT = 32, BT = 128, KT = 32, KNN_MAXK = 16;

__global__ void knn_mean_kernel(const float* __restrict__ bank, int S, int D, float* __restrict__ mean) {
  int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double a = 0.0;
  for (int s = 0; s < S; s++) a += (double)bank[(size_t)s * D + d];
  mean[d] = (float)(a / (double)S);
}

__global__ void __launch_bounds__(256)
knn_normalize_kernel(const float* __restrict__ bank, int D, const float* __restrict__ mean, float* __restrict__ bhat) {
  const int s = blockIdx.x;
  __shared__ double red[8];
  __shared__ float s_inv;
  double a = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = __fsub_rn(bank[(size_t)s * D + d], mean[d]);
    a += (double)v * (double)v;
  }
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0; for (int i = 0; i < 8; i++) t += red[i];
    float n = (float)sqrt(t);
    s_inv = n > 1e-8f ? n : 1e-8f;
  }
  __syncthreads();
  const float nrm = s_inv;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    bhat[(size_t)s * D + d] = __fdiv_rn(__fsub_rn(bank[(size_t)s * D + d], mean[d]), nrm);
}

__global__ void __launch_bounds__(256)
knn_verify_kernel(const float* __restrict__ mean, const float* __restrict__ bhat, const int64_t* __restrict__ bank_cls, int S, int D,
                  const float* __restrict__ queries, const int64_t* __restrict__ query_cls, int64_t Q, int topk, int knn,
                  int64_t* __restrict__ top_idx, float* __restrict__ top_sim, int64_t* __restrict__ votes, uint8_t* __restrict__ keep) {
  __shared__ __align__(16) float As[KT][QT];
  __shared__ __align__(16) float Bs[KT][BT];
  __shared__ float Ss[QT][BT + 1];
  __shared__ float qn[QT];
  __shared__ float tk_sim[QT][KNN_MAXK];
  __shared__ int tk_idx[QT][KNN_MAXK];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * QT;
  if (tid < QT) for (int k = 0; k < KNN_MAXK; k++) { tk_sim[tid][k] = -INFINITY; tk_idx[tid][k] = -1; }
  const int lq = tid >> 3, lk = (tid & 7) * 4;  // loader mapping for the query tile
  float qss = 0.f;                              // partial sum of squares of the centred query (first bank tile only)
  for (int b0 = 0; b0 < S; b0 += BT) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += KT) {
      {  // query tile, centred
        float4 v = make_float4(0, 0, 0, 0);
        if (q0 + lq < Q && k0 + lk < D) {
          v = __ldg(reinterpret_cast<const float4*>(queries + (size_t)(q0 + lq) * D + k0 + lk));
          float4 m = __ldg(reinterpret_cast<const float4*>(mean + k0 + lk));
          v.x = __fsub_rn(v.x, m.x); v.y = __fsub_rn(v.y, m.y); v.z = __fsub_rn(v.z, m.z); v.w = __fsub_rn(v.w, m.w);
        }
        As[lk][lq] = v.x; As[lk + 1][lq] = v.y; As[lk + 2][lq] = v.z; As[lk + 3][lq] = v.w;
        if (b0 == 0) qss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {  // bank tile
        int r = (tid >> 3) + 32 * j;
        float4 v = make_float4(0, 0, 0, 0);
        if (b0 + r < S && k0 + lk < D) v = __ldg(reinterpret_cast<const float4*>(bhat + (size_t)(b0 + r) * D + k0 + lk));
        Bs[lk][r] = v.x; Bs[lk + 1][r] = v.y; Bs[lk + 2][r] = v.z; Bs[lk + 3][r] = v.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KT; k++) {
        float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
        acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
        acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
        acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
      }
      __syncthreads();
    }
    if (b0 == 0) {  // finish the query norms: 8 consecutive lanes share a query
      float s = qss;
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      if ((tid & 7) == 0) { float n = sqrtf(s); qn[lq] = n > 1e-8f ? n : 1e-8f; }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int r = b0 + tx * 4 + j;
        Ss[ty * 4 + i][tx * 4 + j] = (r < S) ? __fdiv_rn(acc[i][j], qn[ty * 4 + i]) : -INFINITY;
      }
    __syncthreads();
    if (tid < QT) {  // running top-k, ascending bank index so that ties keep the lower index
      float worst = tk_sim[tid][topk - 1];
      for (int j = 0; j < BT; j++) {
        float s = Ss[tid][j];
        if (s > worst) {
          int p = topk - 1;
          while (p > 0 && tk_sim[tid][p - 1] < s) { tk_sim[tid][p] = tk_sim[tid][p - 1]; tk_idx[tid][p] = tk_idx[tid][p - 1]; p--; }
          tk_sim[tid][p] = s; tk_idx[tid][p] = b0 + j;
          worst = tk_sim[tid][topk - 1];
        }
      }
    }
    __syncthreads();
  }
  if (tid < QT && q0 + tid < Q) {
    const int64_t q = q0 + tid;
    int64_t v[KNN_MAXK];
    for (int k = 0; k < topk; k++) {
      int idx = tk_idx[tid][k];
      v[k] = idx >= 0 ? bank_cls[idx] : -1;
      top_idx[q * topk + k] = idx;
      votes[q * topk + k] = v[k];
      if (top_sim) top_sim[q * topk + k] = tk_sim[tid][k];
    }
    int64_t best_v = 0; int best_c = 0;   // torch.mode: most frequent, smallest value on ties
    const int kk = knn < topk ? knn : topk;
    for (int a = 0; a < kk; a++) {
      int c = 0;
      for (int b = 0; b < kk; b++) c += (v[b] == v[a]);
      if (c > best_c || (c == best_c && v[a] < best_v)) { best_c = c; best_v = v[a]; }
    }
    keep[q] = (query_cls[q] == best_v) ? 1 : 0;
  }
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" size_t lvcb200_knn_prepared_bytes(int S, int D) { return sizeof(float) * ((size_t)D + (size_t)S * D); }

extern "C" int lvcb200_knn_prepare(const float* bank, int S, int D, void* bank_prepared, void* stream) {
  LVC_REQUIRE(S >= 1 && D >= 4 && D % 4 == 0, "knn_prepare: need S >= 1 and D a positive multiple of 4");
  LVC_REQUIRE(bank && bank_prepared, "knn_prepare: NULL pointer");
  cudaStream_t s = (cudaStream_t)stream;
  float* mean = (float*)bank_prepared;
  float* bhat = mean + D;
  knn_mean_kernel<<<(D + 255) / 256, 256, 0, s>>>(bank, S, D, mean);
  int rc = check_launch("knn_mean_kernel");
  if (rc) return rc;
  knn_normalize_kernel<<<S, 256, 0, s>>>(bank, D, mean, bhat);
  return check_launch("knn_normalize_kernel");
}

extern "C" int lvcb200_knn_verify(const void* bank_prepared, const int64_t* bank_cls, int S, int D, const float* queries,
                                  const int64_t* query_cls, int64_t Q, int topk, int knn, int64_t* top_idx, float* top_sim,
                                  int64_t* votes, uint8_t* keep, void* stream) {
  LVC_REQUIRE(S >= 1 && S <= 4096 && D >= 4 && D % 4 == 0, "knn_verify: need 1 <= S <= 4096 and D a multiple of 4");
  LVC_REQUIRE(topk >= 1 && topk <= KNN_MAXK && topk <= S && knn >= 1, "knn_verify: need 1 <= topk <= min(16, S), knn >= 1");
  if (Q == 0) return 0;
  LVC_REQUIRE(bank_prepared && bank_cls && queries && query_cls && top_idx && votes && keep, "knn_verify: NULL pointer");
  LVC_REQUIRE(((uintptr_t)queries % 16) == 0, "knn_verify: queries must be 16-byte aligned");
  const float* mean = (const float*)bank_prepared;
  knn_verify_kernel<<<(unsigned)ceil_div64(Q, QT), 256, 0, (cudaStream_t)stream>>>(
      mean, mean + D, bank_cls, S, D, queries, query_cls, Q, topk, knn, top_idx, top_sim, votes, keep);
  return check_launch("knn_verify_kernel");
}
