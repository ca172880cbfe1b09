// audio8_b200 — CTC loss for sm_100a: warp-per-utterance log2-space alpha/beta recursions.
//
// Replaces torch.nn.functional.ctc_loss as called by the reference (audio8/ctc.py:197-205; ATen's
// ctc_loss_log_alpha / log_beta / collect kernels, one thread per extended-label state with a
// __syncthreads() per time step).  Here one CTA of two warps owns an utterance: warp 0 sweeps alpha forward
// while warp 1 sweeps beta backward, concurrently.  Each lane keeps NS consecutive extended-label states in
// registers, neighbours are exchanged with two shuffles per time step, there is no block barrier inside the
// time loop, and the log-prob rows stream through a cp.async ring in shared memory.  All arithmetic is in
// the log2 domain (one ex2 per term, one lg2 per state).  The gradient kernel is fully parallel over (b, t).
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int RING = 8;  // log-prob rows in flight per warp (7 time steps of look-ahead cover an L2/HBM round trip)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b + 2^c) with -inf handling (never forms inf - inf)
__device__ __forceinline__ float lse3(float a, float b, float c) {
  // branch-free (the recursion evaluates NS of these per lane and time step; a divergent early return serialises
  // them): with every input -inf the shift is 0, ex2(-inf) = 0 and lg2(0) = -inf is the result
  const float m = fmaxf(a, fmaxf(b, c));
  const float ms = (m == -INFINITY) ? 0.f : m;
  return ms + lg2(ex2(a - ms) + ex2(b - ms) + ex2(c - ms));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct CtcArgs {
  const float* lp;
  long long st, sb, sv;
  int T, B, V;
  const int* targets;
  const int* tgt_off;
  const int* tgt_len;
  const int* in_len;
  int blank;
  int epad;  // 32 * NS
  float* alpha;
  float* beta;
  float* nll;
};

// one CTA (2 warps) per utterance; warp 0 = alpha sweep, warp 1 = beta sweep
template <int NS>
__global__ void __launch_bounds__(64) ctc_recursion_kernel(const CtcArgs a) {
  extern __shared__ float smem[];  // [2 warps][RING][V]
  const int b = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  const int* lab = a.targets + a.tgt_off[b];
  const int V = a.V;
  float* ring = smem + (size_t)w * RING * V;
  const uint32_t ring_u32 = smem_u32(ring);
  const float* lpb = a.lp + (long long)b * a.sb;
  float* out = (w == 0 ? a.alpha : a.beta) + (long long)b * a.T * a.epad;

  if (Tb <= 0 || E > 32 * NS) {  // degenerate: no frames (host guarantees E fits)
    if (threadIdx.x == 0) a.nll[b] = (S == 0 && Tb <= 0) ? 0.f : INFINITY;
    return;
  }

  // per-state class and skip flags (registers)
  int cls[NS];
  bool skp[NS];  // alpha: may come from s-2;  beta: may go to s+2
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int s = lane * NS + i;
    int c = a.blank;
    bool k = false;
    if (s < E && (s & 1)) {
      c = lab[s >> 1];
      if (w == 0) k = (s >= 3) && (lab[(s >> 1) - 1] != c);
      else k = (s + 2 < E) && (lab[(s >> 1) + 1] != c);
    }
    cls[i] = c;
    skp[i] = k;
  }

  // time order of this warp: alpha t = 0..Tb-1, beta t = Tb-1..0
  auto t_of = [&](int step) { return w == 0 ? step : Tb - 1 - step; };
  auto issue_row = [&](int step) {
    if (step < Tb) {
      const float* src = lpb + (long long)t_of(step) * a.st;
      const uint32_t dst = ring_u32 + (uint32_t)((step % RING) * V) * 4u;
      for (int v = lane; v < V; v += 32) cp_async4(dst + 4u * v, src + (long long)v * a.sv);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < RING - 1; ++i) issue_row(i);

  float cur[NS];
  for (int step = 0; step < Tb; ++step) {
    issue_row(step + RING - 1);
    cp_async_wait<RING - 1>();
    __syncwarp();
    const float* row = ring + (step % RING) * V;
    float nxt[NS];
    if (step == 0) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int s = lane * NS + i;
        const bool init = (w == 0) ? (s <= 1 && s < E) : (s >= E - 2 && s < E);
        nxt[i] = init ? row[cls[i]] * LOG2E : -INFINITY;
      }
    } else {
      float n1, n2;  // neighbour states across the lane boundary
      if (w == 0) {
        n1 = __shfl_up_sync(0xffffffffu, cur[NS - 1], 1);
        n2 = __shfl_up_sync(0xffffffffu, cur[NS - 2], 1);
        if (lane == 0) n1 = n2 = -INFINITY;
      } else {
        n1 = __shfl_down_sync(0xffffffffu, cur[0], 1);
        n2 = __shfl_down_sync(0xffffffffu, cur[1], 1);
        if (lane == 31) n1 = n2 = -INFINITY;
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int s = lane * NS + i;
        float x1, x2;
        if (w == 0) {
          x1 = (i >= 1) ? cur[i - 1] : n1;
          x2 = (i >= 2) ? cur[i - 2] : (i == 1 ? n1 : n2);
        } else {
          x1 = (i + 1 < NS) ? cur[i + 1] : n1;
          x2 = (i + 2 < NS) ? cur[i + 2] : (i + 1 < NS ? n1 : n2);
        }
        const float acc = lse3(cur[i], x1, skp[i] ? x2 : -INFINITY);
        nxt[i] = (s < E) ? row[cls[i]] * LOG2E + acc : -INFINITY;
      }
    }
    // vectorised store of this lane's NS states (16-byte aligned: epad = 32*NS, NS % 4 == 0)
    float4* dst = reinterpret_cast<float4*>(out + (long long)t_of(step) * a.epad + lane * NS);
#pragma unroll
    for (int i = 0; i < NS / 4; ++i) dst[i] = make_float4(nxt[4 * i], nxt[4 * i + 1], nxt[4 * i + 2], nxt[4 * i + 3]);
#pragma unroll
    for (int i = 0; i < NS; ++i) cur[i] = nxt[i];
    __syncwarp();  // row buffer (step % RING) is re-filled by issue_row(step + RING) next iteration
  }

  if (w == 0) {
    // ll = log2-sum of alpha[Tb-1, E-1] and alpha[Tb-1, E-2]
    float last = -INFINITY, prev = -INFINITY;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int s = lane * NS + i;
      if (s == E - 1) last = cur[i];
      if (s == E - 2) prev = cur[i];
    }
    last = __shfl_sync(0xffffffffu, last, (E - 1) / NS);
    prev = (E >= 2) ? __shfl_sync(0xffffffffu, prev, (E - 2) / NS) : -INFINITY;
    if (lane == 0) {
      const float ll2 = lse3(last, prev, -INFINITY);
      a.nll[b] = (ll2 == -INFINITY) ? INFINITY : -ll2 * LN2;
    }
  }
}

// loss = reduce_b nll[b] (deterministic single-warp sum)
__global__ void ctc_reduce_kernel(const float* nll, const int* tgt_len, int B, int mean, int zero_inf, float* loss) {
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    float v = nll[b];
    if (zero_inf && isinf(v)) v = 0.f;
    if (mean) v = v / (float)max(tgt_len[b], 1);
    acc += v;
  }
  acc = warp_sum(acc);
  if (threadIdx.x == 0) *loss = mean ? acc / (float)B : acc;
}

struct CtcGradArgs {
  CtcArgs c;
  const float* grad_out;
  long long go_stride;
  int mean, zero_inf;
  float* grad;  // [T,B,V] contiguous
};

constexpr int GRAD_WARPS = 8;
// grid (ceil(T / (GRAD_WARPS*TPW)), B): each warp handles TPW consecutive time steps of utterance b.
// The state posteriors are normalised per time step by their own sum (in exact arithmetic that sum equals the
// utterance likelihood for every t): this cancels the common-mode rounding drift that fp32 log-space alpha/beta
// of magnitude ~|nll| accumulate over T steps, so each gradient row sums to zero to fp32 precision.
template <int NS>
__global__ void __launch_bounds__(GRAD_WARPS * 32) ctc_grad_kernel(const CtcGradArgs g, int tpw) {
  extern __shared__ float smem[];  // [GRAD_WARPS][2][V] : lp row, bins ; then int ext[epad]
  const CtcArgs& a = g.c;
  const int b = blockIdx.y;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = a.V;
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  int* ext = reinterpret_cast<int*>(smem + (size_t)GRAD_WARPS * 2 * V);
  const int* lab = a.targets + a.tgt_off[b];
  for (int s = threadIdx.x; s < E; s += blockDim.x) ext[s] = (s & 1) ? lab[s >> 1] : a.blank;
  __syncthreads();
  float* row = smem + (size_t)w * 2 * V;
  float* bins = row + V;
  const float nll = a.nll[b];
  const bool dead = isinf(nll) || E > 32 * NS;
  float scale = g.grad_out[(long long)b * g.go_stride];
  if (g.mean) scale /= (float)(max(S, 1) * a.B);
  const int t0 = (blockIdx.x * GRAD_WARPS + w) * tpw;
  for (int t = t0; t < min(t0 + tpw, a.T); ++t) {
    float* gout = g.grad + ((long long)t * a.B + b) * V;
    if (t >= Tb || dead) {
      // PyTorch: zero for t >= input_length; an infeasible row (loss +inf) is zeroed by zero_infinity — without
      // zero_infinity the reference's gradient is NaN garbage, we return 0 there as well
      for (int c = lane; c < V; c += 32) gout[c] = 0.f;
      continue;
    }
    const float* lpr = a.lp + (long long)b * a.sb + (long long)t * a.st;
    for (int c = lane; c < V; c += 32) {
      row[c] = lpr[(long long)c * a.sv];
      bins[c] = 0.f;
    }
    __syncwarp();
    const float* al = a.alpha + ((long long)b * a.T + t) * a.epad;
    const float* be = a.beta + ((long long)b * a.T + t) * a.epad;
    float wv[NS];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int s = lane + 32 * i;
      wv[i] = -INFINITY;
      if (s < E) {
        const float v = al[s] + be[s];
        if (v > -INFINITY) wv[i] = v - row[ext[s]] * LOG2E;  // alpha and beta both include the emission at t
      }
      mx = fmaxf(mx, wv[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      wv[i] = (wv[i] > -INFINITY) ? ex2(wv[i] - mx) : 0.f;
      sum += wv[i];
    }
    sum = warp_sum(sum);
    const float inv = (sum > 0.f) ? 1.f / sum : 0.f;
    float blank_sum = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int s = lane + 32 * i;
      if (s < E && wv[i] > 0.f) {
        const float occ = wv[i] * inv;
        if (s & 1) atomicAdd(&bins[ext[s]], occ);
        else blank_sum += occ;
      }
    }
    blank_sum = warp_sum(blank_sum);
    __syncwarp();
    if (lane == 0) bins[a.blank] += blank_sum;
    __syncwarp();
    for (int c = lane; c < V; c += 32) gout[c] = (__expf(row[c]) - bins[c]) * scale;
    __syncwarp();
  }
}

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

int ns_for(int max_S) {
  const int E = 2 * max_S + 1;
  int ns = 4;
  while (32 * ns < E) ns += 4;
  return ns;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" size_t a8_ctc_scratch_floats(int32_t T, int32_t B, int32_t max_S) {
  return (size_t)T * (size_t)B * (size_t)(32 * ns_for(max_S));
}

extern "C" int a8_ctc_forward(const float* log_probs, int64_t stride_t, int64_t stride_b, int64_t stride_v,
                              int32_t T, int32_t B, int32_t V, const int32_t* targets,
                              const int32_t* tgt_offsets, const int32_t* tgt_lengths,
                              const int32_t* in_lengths, int32_t max_S, int32_t blank, int32_t reduction_mean,
                              int32_t zero_infinity, float* alpha, float* beta, float* nll, float* loss,
                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem T=%d B=%d V=%d", T, B, V);
  A8_REQUIRE(blank >= 0 && blank < V, "ctc: blank %d outside [0,%d)", blank, V);
  A8_REQUIRE(max_S >= 0 && max_S <= 511, "ctc: target length %d unsupported (max 511)", max_S);
  const int ns = ns_for(max_S);
  const size_t smem = (size_t)2 * RING * V * sizeof(float);
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large for the row ring", V);
  CtcArgs a{log_probs, stride_t, stride_b, stride_v, T, B, V, targets, tgt_offsets, tgt_lengths, in_lengths,
            blank, 32 * ns, alpha, beta, nll};
#define A8_CTC_CASE(NS)                                                                              \
  case NS: {                                                                                         \
    if (smem > 48 * 1024)                                                                            \
      A8_CUDA(cudaFuncSetAttribute(ctc_recursion_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)smem));                                                      \
    ctc_recursion_kernel<NS><<<B, 64, smem, stream>>>(a);                                            \
  } break;
  switch (ns) {
    A8_CTC_CASE(4) A8_CTC_CASE(8) A8_CTC_CASE(12) A8_CTC_CASE(16) A8_CTC_CASE(20) A8_CTC_CASE(24)
    A8_CTC_CASE(28) A8_CTC_CASE(32)
    default: set_error("ctc: bad NS %d", ns); return -1;
  }
#undef A8_CTC_CASE
  int rc = check_launch("ctc_recursion_kernel");
  if (rc) return rc;
  if (loss != nullptr) {
    ctc_reduce_kernel<<<1, 32, 0, stream>>>(nll, tgt_lengths, B, reduction_mean, zero_infinity, loss);
    rc = check_launch("ctc_reduce_kernel");
  }
  return rc;
}

extern "C" int a8_ctc_backward(const float* log_probs, int64_t stride_t, int64_t stride_b, int64_t stride_v,
                               int32_t T, int32_t B, int32_t V, const int32_t* targets,
                               const int32_t* tgt_offsets, const int32_t* tgt_lengths,
                               const int32_t* in_lengths, int32_t max_S, int32_t blank, const float* alpha,
                               const float* beta, const float* nll, const float* grad_out,
                               int64_t grad_out_stride, int32_t reduction_mean, int32_t zero_infinity,
                               float* grad, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem");
  const int ns = ns_for(max_S);
  CtcGradArgs g{{log_probs, stride_t, stride_b, stride_v, T, B, V, targets, tgt_offsets, tgt_lengths,
                 in_lengths, blank, 32 * ns, const_cast<float*>(alpha), const_cast<float*>(beta),
                 const_cast<float*>(nll)},
                grad_out, grad_out_stride, reduction_mean, zero_infinity, grad};
  const size_t smem = (size_t)GRAD_WARPS * 2 * V * sizeof(float) + (size_t)(32 * ns) * sizeof(int);
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large", V);
  // enough CTAs to fill 148 SMs a few times over, at least 1 step per warp
  int tpw = 1;
  while ((long long)cdiv(T, GRAD_WARPS * tpw) * B > 148 * 16 && tpw < 16) tpw *= 2;
  dim3 grid(cdiv(T, GRAD_WARPS * tpw), B);
#define A8_CTCG_CASE(NS)                                                                               \
  case NS: {                                                                                           \
    if (smem > 48 * 1024)                                                                              \
      A8_CUDA(cudaFuncSetAttribute(ctc_grad_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                   (int)smem));                                                        \
    ctc_grad_kernel<NS><<<grid, GRAD_WARPS * 32, smem, stream>>>(g, tpw);                              \
  } break;
  switch (ns) {
    A8_CTCG_CASE(4) A8_CTCG_CASE(8) A8_CTCG_CASE(12) A8_CTCG_CASE(16) A8_CTCG_CASE(20) A8_CTCG_CASE(24)
    A8_CTCG_CASE(28) A8_CTCG_CASE(32)
    default: set_error("ctc: bad NS %d", ns); return -1;
  }
#undef A8_CTCG_CASE
  return check_launch("ctc_grad_kernel");
}

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
