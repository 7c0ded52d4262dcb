// subsample_labels (detectron2/modeling/sampling.py:9-54; lvc/modeling/sampling.py is the same function) and its RPN wrapper
// RPN._subsample_labels (proposal_generator/rpn.py:249-266), training-side users of the mining path's label vectors.
//
// The reference draws torch.randperm(n_pos)[:num_pos] and randperm(n_neg)[:num_neg]: two host synchronisations (numel()) and a
// generator stream that cannot be reproduced outside torch.  Here the randomness is an INPUT: one uint32 key per element (any
// generator).  A class's sample is its `take` elements with the smallest (key, index) pairs, emitted in that order -- exactly
// `cls_idx[argsort(keys[cls_idx], stable)[:take]]`, i.e. the reference with randperm(n) := argsort of n i.i.d. keys (a uniform random
// permutation).  The counts follow the reference: num_pos = min(n_pos, int(num_samples * positive_fraction)),
// num_neg = min(n_neg, num_samples - num_pos).
//
// One CTA per label vector, no host round trip: three histogram rounds (11 + 11 + 10 key bits) find, per class, the threshold key
// and how many elements equal to it are still needed; one pass collects the selection into shared memory; a bitonic sort orders it.
#include "common.cuh"

namespace lvcb200 {

constexpr int kSampMax = 1024;    // num_samples cap (the reference uses 256 for the RPN, 512 for the RoI heads)
constexpr int kSampBins = 2048;

template <typename L>
__device__ __forceinline__ int sample_class(const L* labels, int64_t i, int64_t bg) {   // 0 positive, 1 negative, -1 ignored
  const int64_t l = (int64_t)labels[i];
  return l == bg ? 1 : (l != -1 ? 0 : -1);
}

// f(class, key, index) over the labelled elements of one vector, eight independent loads in flight per thread (one CTA walks a
// whole vector: a dependent load per iteration would leave it latency bound).
template <typename L, typename F>
__device__ __forceinline__ void for_each_labelled(const L* __restrict__ labels, const uint32_t* __restrict__ keys, int64_t N, int64_t bg, F&& f) {
  constexpr int U = 8;
  for (int64_t base = 0; base < N; base += (int64_t)blockDim.x * U) {
    int c[U];
    uint32_t k[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
      c[u] = -1;
      k[u] = 0u;
      if (i < N) { c[u] = sample_class(labels, i, bg); k[u] = keys[i]; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (c[u] >= 0) f(c[u], k[u], base + (int64_t)u * blockDim.x + threadIdx.x);
  }
}

// Warp-wide search of the bin where the running count reaches `need` (1 <= need <= total of the row): every lane sums 64 contiguous
// bins, a shuffle scan finds the lane whose range holds the crossing, that lane walks its 64 bins.  Returns (bin, count below it).
__device__ __forceinline__ void find_bin(const int* __restrict__ h, int need, int& bin, int& below) {
  const int lane = threadIdx.x & 31;
  int sum = 0;
  for (int b = 0; b < kSampBins / 32; b++) sum += h[lane * (kSampBins / 32) + b];
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned m = __ballot_sync(0xffffffffu, incl >= need);
  const int owner = __ffs(m) - 1;
  int b = 0, cum = 0;
  if (lane == owner) {
    cum = incl - sum;
    b = lane * (kSampBins / 32);
    while (cum + h[b] < need) { cum += h[b]; b++; }
  }
  bin = __shfl_sync(0xffffffffu, b, owner);
  below = __shfl_sync(0xffffffffu, cum, owner);
}

template <typename L>
__global__ void __launch_bounds__(1024)
subsample_labels_kernel(const L* __restrict__ labels_all, const uint32_t* __restrict__ keys_all, int64_t N, int num_samples, int num_pos_cap,
                        int64_t bg, int64_t* __restrict__ pos_idx, int64_t* __restrict__ neg_idx, int32_t* __restrict__ counts,
                        int8_t* __restrict__ out_labels) {
  __shared__ int hist[2][kSampBins];
  __shared__ unsigned long long sel[2][kSampMax];
  __shared__ int s_tot[2], s_take[2], s_need[2], s_nsel[2], s_eq[2];
  __shared__ uint32_t s_prefix[2];
  const int tid = threadIdx.x;
  const L* labels = labels_all + (size_t)blockIdx.x * N;
  const uint32_t* keys = keys_all + (size_t)blockIdx.x * N;

  for (int round = 0; round < 3; round++) {
    const int shift = round == 0 ? 21 : (round == 1 ? 10 : 0);
    const int bits = round == 2 ? 10 : 11;
    for (int b = tid; b < 2 * kSampBins; b += blockDim.x) (&hist[0][0])[b] = 0;
    __syncthreads();
    const uint32_t p0 = round ? s_prefix[0] : 0u, p1 = round ? s_prefix[1] : 0u;
    const bool live0 = round == 0 || s_need[0] > 0, live1 = round == 0 || s_need[1] > 0;
    for_each_labelled(labels, keys, N, bg, [&](int c, uint32_t k, int64_t) {
      if (!(c ? live1 : live0)) return;
      if (round == 0 || (k >> (shift + bits)) == (c ? p1 : p0)) atomicAdd(&hist[c][(k >> shift) & ((1u << bits) - 1u)], 1);
    });
    __syncthreads();
    if (round == 0 && tid < 64) {           // totals per class: warp c sums row c
      const int c = tid >> 5;
      int sum = 0;
      for (int b = tid & 31; b < kSampBins; b += 32) sum += hist[c][b];
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if ((tid & 31) == 0) s_tot[c] = sum;
    }
    __syncthreads();
    if (round == 0 && tid == 0) {
      const int tp = min(s_tot[0], num_pos_cap);
      const int tn = min(s_tot[1], num_samples - tp);
      s_take[0] = tp; s_take[1] = tn; s_need[0] = tp; s_need[1] = tn;
      s_prefix[0] = s_prefix[1] = 0u;
      s_nsel[0] = s_nsel[1] = 0;
    }
    __syncthreads();
    if (tid < 64) {                         // warp c advances class c's prefix by this round's digit
      const int c = tid >> 5;
      const int need = s_need[c];
      if (need > 0) {                       // (uniform per warp) need <= matching elements, so the crossing exists
        int b, cum;
        find_bin(hist[c], need, b, cum);
        if ((tid & 31) == 0) {
          s_need[c] = need - cum;
          s_prefix[c] = (s_prefix[c] << bits) | (uint32_t)b;
          if (round == 2) s_eq[c] = hist[c][b];
        }
      } else if (round == 2 && (tid & 31) == 0) {
        s_eq[c] = 0;
      }
    }
    __syncthreads();
  }
  // s_prefix[c] = threshold key T_c; s_need[c] = elements equal to T_c still to take (>= 1 when take > 0); s_eq[c] = how many there are
  const bool tie0 = s_take[0] > 0 && s_eq[0] != s_need[0], tie1 = s_take[1] > 0 && s_eq[1] != s_need[1];
  for_each_labelled(labels, keys, N, bg, [&](int c, uint32_t k, int64_t i) {
    if (s_take[c] == 0) return;
    const uint32_t T = s_prefix[c];
    if (k < T || (k == T && !(c ? tie1 : tie0))) {
      const int pos = atomicAdd(&s_nsel[c], 1);
      sel[c][pos] = ((unsigned long long)k << 32) | (unsigned long long)(uint32_t)i;
    }
  });
  __syncthreads();
  if ((tie0 || tie1) && tid < 32) {       // several elements share the threshold key: the smallest indices win (stable argsort); rare
    for (int c = 0; c < 2; c++) {
      if (!(c ? tie1 : tie0)) continue;
      int need = s_need[c];
      const uint32_t T = s_prefix[c];
      for (int64_t i0 = 0; i0 < N && need > 0; i0 += 32) {
        const int64_t i = i0 + tid;
        const bool eq = i < N && sample_class(labels, i, bg) == c && keys[i] == T;
        const unsigned m = __ballot_sync(0xffffffffu, eq);
        const int r = __popc(m & ((1u << tid) - 1u));
        if (eq && r < need) sel[c][s_nsel[c] + r] = ((unsigned long long)T << 32) | (unsigned long long)(uint32_t)i;
        const int used = min(__popc(m), need);
        __syncwarp();
        if (tid == 0) s_nsel[c] += used;
        need -= used;
        __syncwarp();
      }
    }
  }
  __syncthreads();
  for (int c = 0; c < 2; c++) {           // bitonic sort of the selection by (key, index)
    const int n = s_take[c];
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int i = n + tid; i < p2; i += blockDim.x) sel[c][i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < p2; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const unsigned long long a = sel[c][i], b = sel[c][l];
            if (((i & k) == 0) == (a > b)) { sel[c][i] = b; sel[c][l] = a; }
          }
        }
        __syncthreads();
      }
  }
  int64_t* po = pos_idx + (size_t)blockIdx.x * num_samples;
  int64_t* no = neg_idx + (size_t)blockIdx.x * num_samples;
  for (int i = tid; i < num_samples; i += blockDim.x) {
    po[i] = i < s_take[0] ? (int64_t)(uint32_t)sel[0][i] : -1;
    no[i] = i < s_take[1] ? (int64_t)(uint32_t)sel[1][i] : -1;
  }
  if (tid == 0) { counts[blockIdx.x * 2] = s_take[0]; counts[blockIdx.x * 2 + 1] = s_take[1]; }
  if (out_labels != nullptr) {            // rpn.py:262-265: fill -1, scatter 1 at the positives, 0 at the negatives
    int8_t* ol = out_labels + (size_t)blockIdx.x * N;
    for (int64_t i = tid; i < N; i += blockDim.x) ol[i] = -1;
    __syncthreads();
    for (int i = tid; i < s_take[0]; i += blockDim.x) ol[(uint32_t)sel[0][i]] = 1;
    for (int i = tid; i < s_take[1]; i += blockDim.x) ol[(uint32_t)sel[1][i]] = 0;
  }
}

}  // namespace lvcb200

using namespace lvcb200;

extern "C" int lvcb200_subsample_labels(const void* labels, int labels_are_int8, const uint32_t* keys, int n_vectors, int64_t N,
                                        int num_samples, double positive_fraction, int64_t bg_label, int64_t* pos_idx, int64_t* neg_idx,
                                        int32_t* counts, int8_t* out_labels, void* stream) {
  LVC_REQUIRE(num_samples >= 1 && num_samples <= kSampMax, "subsample_labels: num_samples must be in [1, 1024]");
  LVC_REQUIRE(positive_fraction >= 0.0 && positive_fraction <= 1.0, "subsample_labels: positive_fraction must be in [0, 1]");
  LVC_REQUIRE(N >= 0 && N < (1ll << 31), "subsample_labels: at most 2^31 - 1 labels per vector");
  if (n_vectors == 0) return 0;
  LVC_REQUIRE(pos_idx && neg_idx && counts && (N == 0 || (labels && keys)), "subsample_labels: NULL pointer");
  const int num_pos_cap = (int)((double)num_samples * positive_fraction);       // int(num_samples * positive_fraction), sampling.py:41
  cudaStream_t s = (cudaStream_t)stream;
  if (labels_are_int8)
    subsample_labels_kernel<int8_t><<<n_vectors, 1024, 0, s>>>((const int8_t*)labels, keys, N, num_samples, num_pos_cap, bg_label, pos_idx,
                                                              neg_idx, counts, out_labels);
  else
    subsample_labels_kernel<int64_t><<<n_vectors, 1024, 0, s>>>((const int64_t*)labels, keys, N, num_samples, num_pos_cap, bg_label, pos_idx,
                                                               neg_idx, counts, out_labels);
  return check_launch("subsample_labels_kernel");
}
