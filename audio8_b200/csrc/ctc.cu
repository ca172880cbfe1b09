// audio8_b200 — CTC helpers for sm_100a: target preparation and greedy best-path decode.  The loss itself (alpha sweep,
// beta sweep + gradient, optional fused log-softmax) lives in ctc_loss.cu.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

// ------------------------------------------------------------------------------------------------
// Greedy best-path decode (reference: ctc.py:161-162, `lp.argmax(-1).unique_consecutive()` then drop blank): the only
// alignment the reference produces (eval, CER/WER inputs).  One CTA per utterance: frame arg-max (first index wins
// ties, as torch.argmax), collapse of repeats, blank removal, order-preserving compaction by a block-wide scan.
// Integer output, bit-exact against the reference's ops on the same log-probs.
// ------------------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 256;
__global__ void __launch_bounds__(DEC_THREADS) ctc_greedy_kernel(const float* lp, long long sb, long long st, long long sv,
                                                                 int T, int V, const int* in_len, int blank, int* out,
                                                                 int* out_len) {
  extern __shared__ int s_arg[];  // [T] frame arg-max
  __shared__ int s_scan[DEC_THREADS];
  const int b = blockIdx.x;
  const int Tb = min(in_len ? in_len[b] : T, T);
  const float* base = lp + (long long)b * sb;
  for (int t = threadIdx.x; t < Tb; t += blockDim.x) {
    const float* row = base + (long long)t * st;
    float best = row[0];
    int bi = 0;
    for (int v = 1; v < V; ++v) {
      const float x = row[(long long)v * sv];
      if (x > best || (x != x && !(best != best))) {  // first maximum; NaN ranks highest like torch.argmax
        best = x;
        bi = v;
      }
    }
    s_arg[t] = bi;
  }
  __syncthreads();
  // each thread owns a contiguous run of frames so that the compaction keeps time order
  const int per = (Tb + blockDim.x - 1) / blockDim.x;
  const int t0 = threadIdx.x * per, t1 = min(Tb, t0 + per);
  int cnt = 0;
  for (int t = t0; t < t1; ++t) {
    const int a = s_arg[t];
    cnt += (a != blank && (t == 0 || s_arg[t - 1] != a)) ? 1 : 0;
  }
  s_scan[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = 1; off < DEC_THREADS; off <<= 1) {  // inclusive Hillis-Steele scan
    const int v = (threadIdx.x >= off) ? s_scan[threadIdx.x - off] : 0;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = s_scan[threadIdx.x] - cnt;
  int* ob = out + (long long)b * T;
  for (int t = t0; t < t1; ++t) {
    const int a = s_arg[t];
    if (a != blank && (t == 0 || s_arg[t - 1] != a)) ob[pos++] = a;
  }
  const int total = s_scan[DEC_THREADS - 1];
  for (int i = total + threadIdx.x; i < T; i += blockDim.x) ob[i] = -1;
  if (threadIdx.x == 0) out_len[b] = total;
}

// Target preparation (ctc.py:193-194): flat row-major compaction of the entries that are neither PAD nor
// EOS, utterance b = flat[cumsum(target_lengths)[b-1] ...], exactly as F.ctc_loss consumes them.  One CTA.
__global__ void __launch_bounds__(1024) ctc_prep_kernel(const long long* targets, long long stride_b,
                                                        long long stride_s, int B, int S, int pad, int eos,
                                                        const long long* tgt_len64, const long long* in_len64,
                                                        int* flat, int* row_start, int* tgt_off, int* tgt_len,
                                                        int* in_len) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = w; b < B; b += nw) {
    int cnt = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      long long v = (s < S) ? targets[b * stride_b + s * stride_s] : pad;
      cnt += __popc(__ballot_sync(0xffffffffu, v != pad && v != eos));
    }
    if (lane == 0) row_start[b] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, off = 0;
    for (int b = 0; b < B; ++b) {
      const int c = row_start[b];
      row_start[b] = run;
      run += c;
      tgt_off[b] = off;
      tgt_len[b] = (int)tgt_len64[b];
      off += (int)tgt_len64[b];
      in_len[b] = (int)in_len64[b];
    }
    in_len[B] = 0;  // completion ticket of the alpha sweep's loss reduction (ctc_loss.cu)
  }
  __syncthreads();
  for (int b = w; b < B; b += nw) {
    int pos = row_start[b];
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      long long v = (s < S) ? targets[b * stride_b + s * stride_s] : pad;
      const bool keep = v != pad && v != eos;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) flat[pos + __popc(m & ((1u << lane) - 1u))] = (int)v;
      pos += __popc(m);
    }
  }
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_ctc_prep(const int64_t* targets, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S,
                           int32_t pad, int32_t eos, const int64_t* target_lengths, const int64_t* input_lengths,
                           int32_t* flat, int32_t* row_start, int32_t* tgt_offsets, int32_t* tgt_lengths,
                           int32_t* in_lengths, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(B > 0 && S >= 0, "ctc_prep: bad shape B=%d S=%d", B, S);
  ctc_prep_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const long long*>(targets), stride_b, stride_s, B, S,
                                          pad, eos, reinterpret_cast<const long long*>(target_lengths),
                                          reinterpret_cast<const long long*>(input_lengths), flat, row_start,
                                          tgt_offsets, tgt_lengths, in_lengths);
  return check_launch("ctc_prep_kernel");
}

extern "C" int a8_ctc_greedy(const float* lp, int64_t stride_b, int64_t stride_t, int64_t stride_v, int32_t B, int32_t T,
                             int32_t V, const int32_t* in_len, int32_t blank, int32_t* out, int32_t* out_len,
                             void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(B > 0 && T > 0 && V > 0 && (size_t)T * sizeof(int) <= 200 * 1024, "ctc_greedy: unsupported shape B=%d T=%d V=%d",
             B, T, V);
  const size_t smem = (size_t)T * sizeof(int);
  if (smem > 48 * 1024)
    A8_CUDA(cudaFuncSetAttribute(ctc_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_greedy_kernel<<<B, DEC_THREADS, smem, st>>>(lp, stride_b, stride_t, stride_v, T, V, in_len, blank, out, out_len);
  return check_launch("ctc_greedy_kernel");
}
