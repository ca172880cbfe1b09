// audio8_b200 — device-side random draws of the pre-training step (second half): the span mask of
// `create_mask` (reference wav2vec2.py:189-216) and the negative indices of `Sampler.negatives` (:955-976) drawn by
// counter-based Philox on the GPU, so that a training step needs no host RNG work, no index upload and can replay as
// CUDA graphs whose draws change with a seed word in device memory.
//
// This is an OPT-IN mode (`audio8_b200.wav2vec2.set_device_draws`): the default path keeps numpy's global generator in
// the reference's call order, which is what makes masks and negatives bit-identical to the reference.  The device mode
// draws from the same DISTRIBUTIONS (uniform k-subsets of span starts, every row cut down to the batch-minimum count by
// a uniform subset, negatives uniform over the other masked steps of the same utterance) with a different generator.
// Everything below is specified exactly (Philox4x32-10 counters, Floyd's subset sampling, multiply-high range
// reduction) so that `tests/emu.py` reproduces the kernels bit for bit.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

// Philox4x32-10 (Salmon et al.), counter (c0..c3), key (k0, k1)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// one 32-bit word of the stream (tag, row, index): counter = (index, row, tag, 0)
__device__ __forceinline__ uint32_t draw32(unsigned long long seed, uint32_t tag, uint32_t row, uint32_t i) {
  return philox4x32_10(i, row, tag, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)).x;
}
enum : uint32_t { TAG_NUM = 1, TAG_START = 2, TAG_DROP = 3, TAG_NEG = 4 };

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Floyd's algorithm: a uniformly distributed k-subset of {0..n-1}, marked in `marks` (bytes, zero on entry).
// One thread; k dependent steps (k ~ 0.065 T span starts per row, or the few frames a row has above the batch minimum).
__device__ __forceinline__ void floyd_subset(uint8_t* marks, int n, int k, unsigned long long seed, uint32_t tag,
                                             uint32_t row) {
  for (int j = n - k; j < n; ++j) {
    const int t = (int)__umulhi(draw32(seed, tag, row, (uint32_t)j), (uint32_t)(j + 1));  // uniform in [0, j]
    if (marks[t]) marks[j] = 1;
    else marks[t] = 1;
  }
}

struct SpanArgs {
  unsigned long long seed;
  const unsigned long long* seed_dev;
  double p_start;
  int B, T, mask_length, R_max;
  int32_t* rows;   // [R_max + 1]: flat row indices b*T + t in row-major order, -1 padding, rows[R_max] = count
  uint8_t* mask;   // [B, T] 0 / 1
};

// One CTA, a warp per utterance (looped).  Shared memory: m[B][T] mask bytes, s[B][T] subset marks, lens[B].
__global__ void __launch_bounds__(1024) span_mask_kernel(const SpanArgs a) {
  extern __shared__ uint8_t sm[];
  const int B = a.B, T = a.T, L = a.mask_length;
  const int BT = B * T;
  uint8_t* m = sm;
  uint8_t* s = sm + BT;
  int* lens = reinterpret_cast<int*>(sm + ((2 * BT + 15) & ~15));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const unsigned long long seed = a.seed + seed_base_ld(a.seed_dev);
  for (int i = tid; i < 2 * BT; i += blockDim.x) sm[i] = 0;
  // reference :192  num_mask = int(p_start * input_length / float(mask_length) + np.random.rand())
  const double u0 = (double)(draw32(seed, TAG_NUM, 0u, 0u) >> 8) * (1.0 / 16777216.0);
  int num_mask = (int)(a.p_start * (double)T / (double)L + u0);
  int span = L;                                   // reference :199-201: only the RANGE of the starts shrinks
  if (T - span <= num_mask) span = T - num_mask - 1;
  const int n_start = T - span;
  if (num_mask > n_start) num_mask = n_start;     // the host wrapper rejects shapes where this could happen
  __syncthreads();
  for (int b = warp; b < B; b += nwarps) {
    uint8_t* mb = m + b * T;
    uint8_t* sb = s + b * T;
    if (lane == 0) floyd_subset(sb, n_start, num_mask, seed, TAG_START, (uint32_t)b);  // :203 choice(sz - min_len, num_mask)
    __syncwarp();
    for (int t = lane; t < n_start; t += 32)
      if (sb[t])
        for (int o = 0; o < L && t + o < T; ++o) mb[t + o] = 1;                        // :205-207 spans, clipped at sz
    __syncwarp();
    int cnt = 0;
    for (int t = lane; t < T; t += 32) {
      cnt += mb[t];
      sb[t] = 0;
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) lens[b] = cnt;
  }
  __syncthreads();
  int keep = lens[0];
  for (int b = 1; b < B; ++b) keep = min(keep, lens[b]);                               // :209-210
  for (int b = warp; b < B; b += nwarps) {
    uint8_t* mb = m + b * T;
    uint8_t* sb = s + b * T;
    const int len = lens[b], drop = len - keep;
    // :212-213 keeps a uniform `keep`-subset of the row's frames == drops a uniform (len - keep)-subset (by rank)
    if (drop > 0 && lane == 0) floyd_subset(sb, len, drop, seed, TAG_DROP, (uint32_t)b);
    __syncwarp();
    int rank0 = 0, out0 = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const bool bit = t < T && mb[t] != 0;
      const uint32_t ball = __ballot_sync(0xffffffffu, bit);
      const uint32_t lt = (1u << lane) - 1u;
      const int rank = rank0 + __popc(ball & lt);
      const bool kept = bit && !(drop > 0 && sb[rank] != 0);
      const uint32_t kball = __ballot_sync(0xffffffffu, kept);
      if (kept) a.rows[b * keep + out0 + __popc(kball & lt)] = b * T + t;
      if (t < T) a.mask[b * T + t] = kept ? 1 : 0;
      rank0 += __popc(ball);
      out0 += __popc(kball);
    }
  }
  const int total = B * keep;
  for (int i = total + tid; i < a.R_max; i += blockDim.x) a.rows[i] = -1;
  if (tid == 0) a.rows[a.R_max] = total;
}

// negatives: element e = r*K + k of the [R_max, K] index table; 4 elements per Philox call
__global__ void __launch_bounds__(256) negatives_kernel(unsigned long long seed0, const unsigned long long* seed_dev,
                                                        const int32_t* n_valid_ptr, int B, int K, long long total,
                                                        int32_t* out) {
  const unsigned long long seed = seed0 + seed_base_ld(seed_dev);
  const int n_valid = __ldg(n_valid_ptr);
  const int Tm = n_valid / B;  // masked steps per utterance (equal for every row of the batch)
  const long long groups = (total + 3) >> 2;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const uint4 w4 = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), TAG_NEG, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = 4 * g + i;
      const int r = (int)(e / K);
      v[i] = 0;
      if (r < n_valid && Tm > 1) {
        const int b = r / Tm, t = r - b * Tm;
        int n = (int)__umulhi(w[i], (uint32_t)(Tm - 1));  // reference :967 randint(0, T - 1)
        n += (n >= t);                                     // :969 never the positive
        v[i] = n + b * Tm;                                 // :970-974 offset into the flattened [B*T] latents
      } else if (r < n_valid) {
        v[i] = r;  // a single masked step per utterance: the reference's randint(0, 0) raises; here the step itself
      }
    }
    if (4 * g + 3 < total) {
      *reinterpret_cast<int4*>(out + 4 * g) = make_int4(v[0], v[1], v[2], v[3]);
    } else {
      for (int i = 0; i < 4 && 4 * g + i < total; ++i) out[4 * g + i] = v[i];
    }
  }
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_span_mask_draw(uint64_t seed, const void* seed_dev, int32_t B, int32_t T, double p_start,
                                 int32_t mask_length, int32_t R_max, int32_t* rows, uint8_t* mask, void* stream_v) {
  A8_REQUIRE(B > 0 && T > 0 && mask_length > 0 && R_max >= 0 && p_start >= 0.0, "span_mask_draw: bad arguments");
  // worst case of the draw: num_mask <= int(p T / L + 1); the reference raises inside np.random.choice when the starts do
  // not fit without replacement, and so does this wrapper (the kernel cannot)
  const int nm_max = (int)(p_start * (double)T / (double)mask_length + 1.0);
  int span = mask_length;
  if (T - span <= nm_max) span = T - nm_max - 1;
  A8_REQUIRE(span >= 0 && nm_max <= T - span, "span_mask_draw: %d spans do not fit %d frames", nm_max, T);
  const long long keep_max = (long long)nm_max * mask_length < T ? (long long)nm_max * mask_length : T;
  A8_REQUIRE((long long)B * keep_max <= R_max, "span_mask_draw: R_max %d below the worst case %lld", R_max,
             (long long)B * keep_max);
  const size_t smem = (size_t)((2ll * B * T + 15) & ~15ll) + sizeof(int) * (size_t)B;
  A8_REQUIRE(smem <= 200 * 1024, "span_mask_draw: B*T = %lld frames exceed the kernel's shared-memory plan",
             (long long)B * T);
  static bool configured = false;
  if (!configured) {
    A8_CUDA(cudaFuncSetAttribute(span_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  SpanArgs a{seed, static_cast<const unsigned long long*>(seed_dev), p_start, B, T, mask_length, R_max, rows, mask};
  const int threads = B >= 32 ? 1024 : 32 * (B < 4 ? 4 : B);
  span_mask_kernel<<<1, threads, smem, static_cast<cudaStream_t>(stream_v)>>>(a);
  return check_launch("span_mask_kernel");
}

extern "C" int a8_negatives_draw(uint64_t seed, const void* seed_dev, const int32_t* n_valid, int32_t B, int32_t K,
                                 int32_t R_max, int32_t* out, void* stream_v) {
  A8_REQUIRE(B > 0 && K > 0 && R_max > 0 && n_valid != nullptr, "negatives_draw: bad arguments");
  A8_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "negatives_draw: output must be 16-byte aligned");
  const long long total = (long long)R_max * K;
  const long long groups = (total + 3) / 4;
  const int grid = (int)(groups + 255) / 256 < 148 * 8 ? (int)((groups + 255) / 256) : 148 * 8;
  negatives_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      seed, static_cast<const unsigned long long*>(seed_dev), n_valid, B, K, total, out);
  return check_launch("negatives_kernel");
}
