// audio8_b200 — CTC loss for sm_100a, second generation: warp-specialised log2-space alpha and beta sweeps running
// CONCURRENTLY (one CTA per utterance and direction), then a gradient kernel that is parallel over (utterance, time).
//
// Replaces torch.nn.functional.ctc_loss as called by the reference (audio8/ctc.py:197-205; ATen's
// ctc_loss_log_alpha / log_beta / collect kernels) and, when the input is the classifier's LOGITS, the log_softmax in
// front of it as well (audio8/wav2vec2.py:770): rows are normalised on the fly and the backward pass returns
// d loss / d logits = softmax - occupancy, the composition of both backward formulas.
//
// Sweep kernel, one CTA per (utterance, direction) (alpha: t ascending, beta: t descending), each with
//   W recursion warps: lane g (= 32*warp + lane) keeps NS consecutive extended-label states in registers; neighbours
//                  through two shuffles, across warps through a double-buffered smem slot and ONE named barrier per time
//                  step (only when W > 1).  log2 domain; the largest term of every log-sum-exp is factored out, so a
//                  blank state costs 1 ex2 + 1 lg2 and a label state 2 ex2 + 1 lg2.
//   2 producer warps: stream the log-prob / logit rows through a cp.async ring, convert them to log2 units (logits:
//                  minus the row's log-sum-exp) and publish each row on an mbarrier; they run ahead of the recursion by up
//                  to RING rows, so global-memory latency never sits on the serial chain.
// alpha (emission included) and beta~ (the sum over successors, emission excluded) go to fp32 scratch [B, T, Epad].
// nll reduction: the last CTA to finish (atomic ticket) sums nll[0..B) in a fixed order.
// Gradient kernel: a warp per time step; occupancy(s) = 2^(alpha + beta~ - ll), binned per class in shared memory.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int RING = 8;     // emission rows in flight (producers run at most this far ahead)
constexpr int AHEAD = 5;    // cp.async depth of each producer (rows) and of the alpha prefetch (time steps)
constexpr int NPROD = 2;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b): the larger term is factored out (1 ex2 + 1 lg2); all -inf stays -inf without forming inf - inf
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b), n = fminf(a, b);
  const float ms = (m == -INFINITY) ? 0.f : m;
  return m + lg2f(1.f + ex2f(n - ms));
}
// log2(2^a + 2^b + 2^c): 2 ex2 + 1 lg2
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  const float ms = (m == -INFINITY) ? 0.f : m;
  // the two terms that are not the maximum (ties: any consistent choice is exact)
  const float x = (a == m) ? b : a;
  const float y = (a == m || b == m) ? c : b;
  return m + lg2f(1.f + ex2f(x - ms) + ex2f(y - ms));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct CtcArgs {
  const float* x;        // log-probs or logits, element strides (st, sb, sv) for (t, b, v)
  long long st, sb, sv;
  int T, B, V;
  int from_logits;
  const int* targets;
  const int* tgt_off;
  const int* tgt_len;
  const int* in_len;     // [B] followed by one int: the alpha sweep's completion ticket (zero on entry, zero on exit)
  int blank;
  int W;                 // recursion warps per CTA
  int epad;              // 32 * W * NS: row pitch of alpha
  float* alpha;          // [B, T, epad] log2 domain, emission included
  float* beta;           // [B, T, epad] log2 domain, emission excluded
  float* nll;            // [B]
  // alpha sweep: reduced loss
  float* loss;
  int mean, zero_inf;
  // beta sweep: gradient
  const float* grad_out;
  long long go_stride;
  float* grad;           // element strides (gt, gb, 1) for (t, b, v)
  long long gt, gb;
};

// shared-memory carve-up of one sweep direction (floats): ring[RING][Vp] | xch[2][W][2] | mbarriers
struct Smem {
  float* ring;
  float* xch;
  uint32_t bars;  // shared-space address of the mbarrier block
};
constexpr int BAR_FULL = 0;               // [RING] row converted
constexpr int BAR_FREE = RING;            // [RING] row consumed by every recursion warp
constexpr int NBARS = 2 * RING;

__device__ __forceinline__ uint32_t bar_addr(const Smem& s, int i) { return s.bars + 8u * i; }

__host__ __device__ inline int vpad(int V) { return (V + 31) & ~31; }
__host__ __device__ inline size_t ctc_smem_bytes(int V, int W, int NS) {  // one direction
  size_t f = (size_t)RING * vpad(V) + 2 * W * 2;
  (void)NS;
  return f * sizeof(float) + NBARS * 8 + 16;
}

__device__ __forceinline__ Smem carve(float* base, int V, int W, int NS) {
  Smem s;
  (void)NS;
  s.ring = base;
  s.xch = s.ring + (size_t)RING * vpad(V);
  float* end = s.xch + 2 * W * 2;
  s.bars = (smem_u32(end) + 7u) & ~7u;
  return s;
}

// ---- producers: row r of the sweep (time t_of(r)) -> ring[r % RING] in log2 units
template <bool BACKWARD>
__device__ __forceinline__ void producer_loop(const CtcArgs& a, const Smem& s, int b, int Tb, int pid, int lane) {
  const int V = a.V, Vp = vpad(V);
  const float* xb = a.x + (long long)b * a.sb;
  auto t_of = [&](int r) { return BACKWARD ? Tb - 1 - r : r; };
  auto issue = [&](int k) {  // k-th row of THIS producer: r = pid + NPROD * k
    const int r = pid + NPROD * k;
    if (r < Tb) {
      if (r >= RING) mbar_wait(bar_addr(s, BAR_FREE + r % RING), ((r / RING) - 1) & 1);
      const float* src = xb + (long long)t_of(r) * a.st;
      const uint32_t dst = smem_u32(s.ring + (size_t)(r % RING) * Vp);
      for (int v = lane; v < V; v += 32) cp_async4(dst + 4u * v, src + (long long)v * a.sv);
    }
    cp_async_commit();
  };
  constexpr int DEPTH = AHEAD / NPROD + 1;  // rows of this producer in flight
#pragma unroll 1
  for (int k = 0; k < DEPTH; ++k) issue(k);
#pragma unroll 1
  for (int k = 0; pid + NPROD * k < Tb; ++k) {
    cp_async_wait<DEPTH - 1>();
    __syncwarp();
    const int r = pid + NPROD * k;
    float* row = s.ring + (size_t)(r % RING) * Vp;
    if (a.from_logits) {
      float mx = -INFINITY;
      for (int v = lane; v < V; v += 32) mx = fmaxf(mx, row[v]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int v = lane; v < V; v += 32) sum += ex2f((row[v] - mx) * LOG2E);
      sum = warp_sum(sum);
      const float shift = mx * LOG2E + lg2f(sum);
      for (int v = lane; v < V; v += 32) row[v] = fmaf(row[v], LOG2E, -shift);
    } else {
      for (int v = lane; v < V; v += 32) row[v] *= LOG2E;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_addr(s, BAR_FULL + r % RING));
    issue(k + DEPTH);
  }
  cp_async_wait<0>();
}

// per-state class and skip flag of the lane's NS states
template <int NS, bool BACKWARD>
__device__ __forceinline__ void state_setup(const CtcArgs& a, int b, int g, int E, int (&cls)[NS], bool (&skp)[NS]) {
  const int* lab = a.targets + a.tgt_off[b];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int st = g * NS + i;
    int c = a.blank;
    bool k = false;
    if (st < E && (st & 1)) {
      c = lab[st >> 1];
      if (!BACKWARD) k = (st >= 3) && (lab[(st >> 1) - 1] != c);      // alpha: may come from s-2
      else k = (st + 2 < E) && (lab[(st >> 1) + 1] != c);              // beta: may go to s+2
    }
    cls[i] = c;
    skp[i] = k;
  }
}

// one step of the recursion for the lane's NS states: out[i] = LSE over the allowed predecessors (no emission added)
template <int NS, bool BACKWARD>
__device__ __forceinline__ void recur(const float (&cur)[NS], float n1, float n2, const bool (&skp)[NS], float (&acc)[NS]) {
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    float x1, x2;
    if (!BACKWARD) {
      x1 = (i >= 1) ? cur[i - 1] : n1;
      x2 = (i >= 2) ? cur[i - 2] : (i == 1 ? n1 : n2);
    } else {
      x1 = (i + 1 < NS) ? cur[i + 1] : n1;
      x2 = (i + 2 < NS) ? cur[i + 2] : (i + 1 < NS ? n1 : n2);
    }
    // NS is even, so the parity of a state is the parity of i: even = blank (never skips)
    acc[i] = (i & 1) ? lse3(cur[i], x1, skp[i] ? x2 : -INFINITY) : lse2(cur[i], x1);
  }
}

// neighbour states across the lane / warp boundary
template <int NS, bool BACKWARD>
__device__ __forceinline__ void neighbours(const float (&cur)[NS], const Smem& s, int w, int W, int lane, int par, float& n1,
                                           float& n2) {
  if (!BACKWARD) {
    n1 = __shfl_up_sync(0xffffffffu, cur[NS - 1], 1);
    n2 = __shfl_up_sync(0xffffffffu, cur[NS - 2], 1);
    if (lane == 0) {
      if (w == 0) n1 = n2 = -INFINITY;
      else { n1 = s.xch[(par * W + (w - 1)) * 2]; n2 = s.xch[(par * W + (w - 1)) * 2 + 1]; }
    }
  } else {
    n1 = __shfl_down_sync(0xffffffffu, cur[0], 1);
    n2 = __shfl_down_sync(0xffffffffu, cur[1], 1);
    if (lane == 31) {
      if (w == W - 1) n1 = n2 = -INFINITY;
      else { n1 = s.xch[(par * W + (w + 1)) * 2]; n2 = s.xch[(par * W + (w + 1)) * 2 + 1]; }
    }
  }
}
// ================================================================================================ alpha || beta sweeps
// one direction's recursion warp: BACKWARD = false writes alpha (emission included), true writes beta~ (emission
// excluded: exactly the factor the occupancy needs next to alpha)
template <int NS, bool BACKWARD>
__device__ __forceinline__ void recursion_loop(const CtcArgs& a, const Smem& s, int b, int Tb, int E, int w, int lane, int bar_id,
                                               float* __restrict__ scratch, float (&cur)[NS]) {
  const int W = a.W;
  const int g = w * 32 + lane;
  int cls[NS];
  bool skp[NS];
  state_setup<NS, BACKWARD>(a, b, g, E, cls, skp);
  const int Vp = vpad(a.V);
  float* out = scratch + (long long)b * a.T * a.epad + (long long)g * NS;
#pragma unroll 1
  for (int r = 0; r < Tb; ++r) {
    float acc[NS];
    if (r == 0) {
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int st = g * NS + i;
        acc[i] = (BACKWARD ? (st < E && st >= E - 2) : (st <= 1)) ? 0.f : -INFINITY;
      }
    } else {
      float n1, n2;
      neighbours<NS, BACKWARD>(cur, s, w, W, lane, (r - 1) & 1, n1, n2);
      recur<NS, BACKWARD>(cur, n1, n2, skp, acc);
    }
    mbar_wait(bar_addr(s, BAR_FULL + r % RING), (r / RING) & 1);
    const float* row = s.ring + (size_t)(r % RING) * Vp;
#pragma unroll
    for (int i = 0; i < NS; ++i) cur[i] = (g * NS + i < E) ? row[cls[i]] + acc[i] : -INFINITY;
    // hand the boundary states to the neighbouring warps first: everything below is off the serial chain
    if (W > 1) {
      if (!BACKWARD) {
        if (lane == 31) { s.xch[((r & 1) * W + w) * 2] = cur[NS - 1]; s.xch[((r & 1) * W + w) * 2 + 1] = cur[NS - 2]; }
      } else {
        if (lane == 0) { s.xch[((r & 1) * W + w) * 2] = cur[0]; s.xch[((r & 1) * W + w) * 2 + 1] = cur[1]; }
      }
      named_bar(bar_id, 32 * W);
    } else {
      __syncwarp();
    }
    if (lane == 0) mbar_arrive(bar_addr(s, BAR_FREE + r % RING));  // every lane's row reads precede the barrier / this point
    const int t = BACKWARD ? Tb - 1 - r : r;
    float4* dst = reinterpret_cast<float4*>(out + (long long)t * a.epad);
#pragma unroll
    for (int i = 0; i < NS / 4; ++i) {
      if (BACKWARD) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
      else dst[i] = make_float4(cur[4 * i], cur[4 * i + 1], cur[4 * i + 2], cur[4 * i + 3]);
    }
  }
}

template <int NS>
__global__ void __launch_bounds__(32 * 8) ctc_sweep_kernel(const CtcArgs a) {
  extern __shared__ float smem_f[];
  // the two directions of an utterance are independent until the gradient: separate CTAs (neighbouring block ids), so a
  // small batch spreads over twice as many SMs and neither sweep shares its SM's MUFU / issue slots with the other
  const int b = blockIdx.x >> 1;
  const int dir = blockIdx.x & 1;        // 0: alpha, 1: beta
  const int W = a.W;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = w;
  const Smem s = carve(smem_f, a.V, W, NS);
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  const bool degenerate = (Tb <= 0) || (E > 32 * W * NS);
  if (w == 0 && lane == 0) {
    for (int i = 0; i < RING; ++i) {
      mbar_init(bar_addr(s, BAR_FULL + i), 1);
      mbar_init(bar_addr(s, BAR_FREE + i), W);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (!degenerate) {
    if (w >= W) {
      if (dir == 0) producer_loop<false>(a, s, b, Tb, w - W, lane);
      else producer_loop<true>(a, s, b, Tb, w - W, lane);
    } else if (dir == 1) {
      float cur[NS];
      recursion_loop<NS, true>(a, s, b, Tb, E, w, lane, 2, a.beta, cur);
    } else {
      float cur[NS];
      recursion_loop<NS, false>(a, s, b, Tb, E, w, lane, 1, a.alpha, cur);
      // ll = log2-sum of alpha[Tb-1, E-1] and alpha[Tb-1, E-2]: the owners drop them into the exchange slots
      const int g = w * 32 + lane;
      if (W > 1) named_bar(1, 32 * W);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int st = g * NS + i;
        if (st == E - 1) s.xch[0] = cur[i];
        if (st == E - 2) s.xch[1] = cur[i];
      }
      if (E < 2 && g == 0) s.xch[1] = -INFINITY;
      if (W > 1) named_bar(1, 32 * W);
      else __syncwarp();
      if (g == 0) {
        const float ll2 = lse2(s.xch[0], s.xch[1]);
        a.nll[b] = (ll2 == -INFINITY) ? INFINITY : -ll2 * LN2;
      }
    }
  } else if (threadIdx.x == 0 && dir == 0) {
    a.nll[b] = (S == 0 && Tb <= 0) ? 0.f : INFINITY;
  }
  // ---- reduced loss: the last alpha CTA to finish sums nll in a fixed order (deterministic), and re-arms the ticket
  if (a.loss != nullptr && dir == 0) {
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      __threadfence();
      int* ticket = const_cast<int*>(a.in_len) + a.B;
      s_last = (atomicAdd(ticket, 1) == a.B - 1);
      if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last && wg == 0) {
      __threadfence();
      float acc = 0.f;
      for (int i = lane; i < a.B; i += 32) {
        float v = __ldcg(a.nll + i);
        if (a.zero_inf && isinf(v)) v = 0.f;
        if (a.mean) v = v / (float)max(a.tgt_len[i], 1);
        acc += v;
      }
      acc = warp_sum(acc);
      if (lane == 0) *a.loss = a.mean ? acc / (float)a.B : acc;
    }
  }
}

// ================================================================================================ gradient
constexpr int GRAD_WARPS = 8;
// grid (ceil(T / (GRAD_WARPS*tpw)), B): each warp handles tpw consecutive time steps of utterance b.
// The state posteriors are normalised per time step by their own sum (in exact arithmetic that sum equals the utterance
// likelihood for every t): this cancels the common-mode rounding drift that fp32 log-space alpha / beta of magnitude
// ~|nll| accumulate over T steps, so each gradient row sums to zero to fp32 precision.
__global__ void __launch_bounds__(GRAD_WARPS * 32) ctc_grad_kernel(const CtcArgs a, int tpw) {
  extern __shared__ float smem_f[];  // [GRAD_WARPS][2][Vp] : log2-prob row, bins ; then int ext[epad]
  const int b = blockIdx.y;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = a.V, Vp = vpad(V);
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  int* ext = reinterpret_cast<int*>(smem_f + (size_t)GRAD_WARPS * 2 * Vp);
  const int* lab = a.targets + a.tgt_off[b];
  for (int st = threadIdx.x; st < E && st < a.epad; st += blockDim.x) ext[st] = (st & 1) ? lab[st >> 1] : a.blank;
  __syncthreads();
  float* row = smem_f + (size_t)w * 2 * Vp;
  float* bins = row + Vp;
  const float nll = a.nll[b];
  const bool dead = isinf(nll) || isnan(nll) || E > a.epad;
  float scale = a.grad_out[(long long)b * a.go_stride];
  if (a.mean) scale /= (float)(max(S, 1) * a.B);
  const float ll2 = -nll * LOG2E;
  const int t0 = (blockIdx.x * GRAD_WARPS + w) * tpw;
  for (int t = t0; t < min(t0 + tpw, a.T); ++t) {
    float* gout = a.grad + (long long)t * a.gt + (long long)b * a.gb;
    if (t >= Tb || dead) {
      // PyTorch: zero for t >= input_length; an infeasible row (loss +inf) is zeroed by zero_infinity — without
      // zero_infinity the reference's gradient is NaN garbage, zero is returned there as well
      for (int c = lane; c < V; c += 32) gout[c] = 0.f;
      continue;
    }
    const float* xr = a.x + (long long)b * a.sb + (long long)t * a.st;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) {
      const float v = xr[(long long)c * a.sv];
      row[c] = v;
      bins[c] = 0.f;
      mx = fmaxf(mx, v);
    }
    float shift = 0.f;
    if (a.from_logits) {  // log2-softmax of the row
      mx = warp_max(mx);
      float sum = 0.f;
      for (int c = lane; c < V; c += 32) sum += ex2f((row[c] - mx) * LOG2E);
      sum = warp_sum(sum);
      shift = mx * LOG2E + lg2f(sum);
    }
    __syncwarp();
    const float* al = a.alpha + ((long long)b * a.T + t) * a.epad;
    const float* be = a.beta + ((long long)b * a.T + t) * a.epad;
    // occupancy of state s: alpha (emission included) * beta~ (emission excluded) / likelihood, renormalised per step
    float tot = 0.f, blank_sum = 0.f;
    for (int st = lane; st < E; st += 32) {
      const float o = al[st] + be[st] - ll2;
      const float occ = (o > -INFINITY) ? ex2f(o) : 0.f;
      tot += occ;
      if (st & 1) {
        if (occ > 0.f) atomicAdd(&bins[ext[st]], occ);
      } else {
        blank_sum += occ;
      }
    }
    tot = warp_sum(tot);
    blank_sum = warp_sum(blank_sum);
    __syncwarp();
    if (lane == 0) bins[a.blank] += blank_sum;
    __syncwarp();
    const float inv = (tot > 0.f) ? 1.f / tot : 0.f;
    for (int c = lane; c < V; c += 32) gout[c] = (ex2f(fmaf(row[c], LOG2E, -shift)) - bins[c] * inv) * scale;
    __syncwarp();
  }
}

int pick_layout(int max_S, int* ns_out) {  // -> W (recursion warps), NS (states per lane)
  const int E = 2 * max_S + 1;
  int ns = 4;
  int W = (E + 32 * ns - 1) / (32 * ns);
  if (W > 6) {
    ns = 8;
    W = (E + 32 * ns - 1) / (32 * ns);
  }
  *ns_out = ns;
  return W < 1 ? 1 : W;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" size_t a8_ctc_scratch_floats(int32_t T, int32_t B, int32_t max_S) {
  int ns;
  const int W = pick_layout(max_S, &ns);
  return (size_t)T * (size_t)B * (size_t)(32 * W * ns);
}

extern "C" int a8_ctc_forward(const float* x, int64_t stride_t, int64_t stride_b, int64_t stride_v, int32_t T, int32_t B,
                              int32_t V, int32_t from_logits, const int32_t* targets, const int32_t* tgt_offsets,
                              const int32_t* tgt_lengths, const int32_t* in_lengths, int32_t max_S, int32_t blank,
                              int32_t reduction_mean, int32_t zero_infinity, float* alpha, float* beta, float* nll,
                              float* loss, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem T=%d B=%d V=%d", T, B, V);
  A8_REQUIRE(blank >= 0 && blank < V, "ctc: blank %d outside [0,%d)", blank, V);
  A8_REQUIRE(max_S >= 0 && max_S <= 511, "ctc: target length %d unsupported (max 511)", max_S);
  int ns;
  const int W = pick_layout(max_S, &ns);
  const size_t smem = (ctc_smem_bytes(V, W, ns) + 15) / 16 * 16;
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large for the row rings", V);
  CtcArgs a{};
  a.x = x; a.st = stride_t; a.sb = stride_b; a.sv = stride_v; a.T = T; a.B = B; a.V = V; a.from_logits = from_logits;
  a.targets = targets; a.tgt_off = tgt_offsets; a.tgt_len = tgt_lengths; a.in_len = in_lengths; a.blank = blank;
  a.W = W; a.epad = 32 * W * ns; a.alpha = alpha; a.beta = beta; a.nll = nll; a.loss = loss; a.mean = reduction_mean;
  a.zero_inf = zero_infinity;
  const int threads = 32 * (W + NPROD);
  if (ns == 4) {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_sweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_sweep_kernel<4><<<2 * B, threads, smem, stream>>>(a);
  } else {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_sweep_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_sweep_kernel<8><<<2 * B, threads, smem, stream>>>(a);
  }
  return check_launch("ctc_sweep_kernel");
}

extern "C" int a8_ctc_backward(const float* x, int64_t stride_t, int64_t stride_b, int64_t stride_v, int32_t T, int32_t B,
                               int32_t V, int32_t from_logits, const int32_t* targets, const int32_t* tgt_offsets,
                               const int32_t* tgt_lengths, const int32_t* in_lengths, int32_t max_S, int32_t blank,
                               const float* alpha, const float* beta, const float* nll, const float* grad_out,
                               int64_t grad_out_stride, int32_t reduction_mean, int32_t zero_infinity, float* grad,
                               int64_t grad_stride_t, int64_t grad_stride_b, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem");
  int ns;
  const int W = pick_layout(max_S, &ns);
  CtcArgs a{};
  a.x = x; a.st = stride_t; a.sb = stride_b; a.sv = stride_v; a.T = T; a.B = B; a.V = V; a.from_logits = from_logits;
  a.targets = targets; a.tgt_off = tgt_offsets; a.tgt_len = tgt_lengths; a.in_len = in_lengths; a.blank = blank;
  a.W = W; a.epad = 32 * W * ns; a.alpha = const_cast<float*>(alpha); a.beta = const_cast<float*>(beta);
  a.nll = const_cast<float*>(nll);
  a.mean = reduction_mean; a.zero_inf = zero_infinity; a.grad_out = grad_out; a.go_stride = grad_out_stride;
  a.grad = grad; a.gt = grad_stride_t; a.gb = grad_stride_b;
  const size_t smem = (size_t)GRAD_WARPS * 2 * vpad(V) * sizeof(float) + (size_t)a.epad * sizeof(int);
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large", V);
  // enough CTAs to fill 148 SMs a few times over, at least 1 step per warp
  int tpw = 1;
  while ((long long)cdiv(T, GRAD_WARPS * tpw) * B > 148 * 16 && tpw < 16) tpw *= 2;
  dim3 grid(cdiv(T, GRAD_WARPS * tpw), B);
  if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_grad_kernel<<<grid, GRAD_WARPS * 32, smem, stream>>>(a, tpw);
  return check_launch("ctc_grad_kernel");
}
