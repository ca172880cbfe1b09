// audio8_b200 — CTC loss for sm_100a, second generation: warp-specialised log2-space alpha sweep and a beta sweep that
// emits the gradient directly (no beta scratch, no separate gradient pass).
//
// Replaces torch.nn.functional.ctc_loss as called by the reference (audio8/ctc.py:197-205; ATen's
// ctc_loss_log_alpha / log_beta / collect kernels) and, when the input is the classifier's LOGITS, the log_softmax in
// front of it as well (audio8/wav2vec2.py:770): the row normalisation is done by the producer warps on the fly and the
// backward pass returns d loss / d logits = softmax - occupancy, the composition of both backward formulas.
//
// One CTA per utterance:
//   warps 0..W-1   recursion: lane g (= 32*warp + lane) keeps NS consecutive extended-label states in registers;
//                  neighbours through two shuffles, across warps through a double-buffered smem slot and ONE named barrier
//                  per time step (only when W > 1).  log2 domain; the largest of the three terms of every log-sum-exp is
//                  factored out, so a blank state costs 1 ex2 + 1 lg2 and a label state 2 ex2 + 1 lg2.
//   warps W, W+1   producers: stream the log-prob / logit rows through a cp.async ring, convert them to log2 units
//                  (logits: minus the row's log-sum-exp) and publish each row on an mbarrier; they run ahead of the
//                  recursion by up to RING rows, so global-memory latency never sits on the serial chain.
//   warp  W+2      (beta sweep only) emitter: per time step turns the per-class occupancy bins the recursion warps
//                  accumulate in shared memory into one gradient row, (softmax - occupancy / sum) * scale.
// Scratch: alpha only, fp32 [B, T, Epad]; the beta sweep prefetches its own states of alpha with per-thread cp.async.
// nll reduction: the last CTA of the alpha sweep to finish (atomic ticket) sums nll[0..B) in a fixed order.
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
  float* alpha;          // [B, T, epad] log2 domain
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

// shared-memory carve-up (floats): ring[RING][Vp] | xch[2][W][2] | bins[2][Vp] | alpha ring (beta sweep) | mbarriers
struct Smem {
  float* ring;
  float* xch;
  float* bins;
  float* aring;
  uint32_t bars;  // shared-space address of the mbarrier block
};
constexpr int BAR_FULL = 0;               // [RING] row converted
constexpr int BAR_FREE = RING;            // [RING] row consumed by every reader
constexpr int BAR_BINFULL = 2 * RING;     // [2]
constexpr int BAR_BINFREE = 2 * RING + 2; // [2]
constexpr int NBARS = 2 * RING + 4;

__device__ __forceinline__ uint32_t bar_addr(const Smem& s, int i) { return s.bars + 8u * i; }

__host__ __device__ inline int vpad(int V) { return (V + 31) & ~31; }
__host__ __device__ inline size_t ctc_smem_bytes(int V, int W, int NS, bool beta) {
  size_t f = (size_t)RING * vpad(V) + 2 * W * 2 + (beta ? 2 * vpad(V) : 0) + (beta ? (size_t)(AHEAD + 1) * 32 * W * NS : 0);
  return f * sizeof(float) + NBARS * 8 + 16;
}

__device__ __forceinline__ Smem carve(float* base, int V, int W, int NS, bool beta) {
  Smem s;
  s.ring = base;
  s.xch = s.ring + (size_t)RING * vpad(V);
  s.bins = s.xch + 2 * W * 2;
  s.aring = s.bins + (beta ? 2 * vpad(V) : 0);
  float* end = s.aring + (beta ? (size_t)(AHEAD + 1) * 32 * W * NS : 0);
  s.bars = (smem_u32(end) + 7u) & ~7u;
  return s;
}

// ---- producers: row r of the sweep (time t_of(r)) -> ring[r % RING] in log2 units
template <bool BACKWARD>
__device__ __forceinline__ void producer_loop(const CtcArgs& a, const Smem& s, int b, int Tb, int pid, int lane, int readers) {
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
  (void)readers;
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
template <int NS, bool BACKWARD>
__device__ __forceinline__ void publish(const float (&cur)[NS], const Smem& s, int w, int W, int lane, int par) {
  if (W == 1) return;
  if (!BACKWARD) {
    if (lane == 31) { s.xch[(par * W + w) * 2] = cur[NS - 1]; s.xch[(par * W + w) * 2 + 1] = cur[NS - 2]; }
  } else {
    if (lane == 0) { s.xch[(par * W + w) * 2] = cur[0]; s.xch[(par * W + w) * 2 + 1] = cur[1]; }
  }
  named_bar(1, 32 * W);
}

__device__ __forceinline__ void init_barriers(const CtcArgs& a, const Smem& s, bool beta) {
  if (threadIdx.x == 0) {
    const int readers = a.W + (beta ? 1 : 0);
    for (int i = 0; i < RING; ++i) {
      mbar_init(bar_addr(s, BAR_FULL + i), 1);
      mbar_init(bar_addr(s, BAR_FREE + i), readers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_addr(s, BAR_BINFULL + i), a.W);
      mbar_init(bar_addr(s, BAR_BINFREE + i), 1);
    }
    mbar_fence_init();
  }
}

// ================================================================================================ alpha sweep
template <int NS>
__global__ void __launch_bounds__(32 * 8) ctc_alpha_kernel(const CtcArgs a) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = a.W;
  const Smem s = carve(smem_f, a.V, W, NS, false);
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  const bool degenerate = (Tb <= 0) || (E > 32 * W * NS);
  init_barriers(a, s, false);
  __syncthreads();
  if (!degenerate) {
    if (w >= W) {
      producer_loop<false>(a, s, b, Tb, w - W, lane, W);
    } else {
      const int g = w * 32 + lane;
      int cls[NS];
      bool skp[NS];
      state_setup<NS, false>(a, b, g, E, cls, skp);
      const int Vp = vpad(a.V);
      float cur[NS];
      float* out = a.alpha + (long long)b * a.T * a.epad + (long long)g * NS;
#pragma unroll 1
      for (int r = 0; r < Tb; ++r) {
        float acc[NS];
        if (r == 0) {
#pragma unroll
          for (int i = 0; i < NS; ++i) acc[i] = (g * NS + i <= 1) ? 0.f : -INFINITY;
        } else {
          float n1, n2;
          neighbours<NS, false>(cur, s, w, W, lane, (r - 1) & 1, n1, n2);
          recur<NS, false>(cur, n1, n2, skp, acc);
        }
        mbar_wait(bar_addr(s, BAR_FULL + r % RING), (r / RING) & 1);
        const float* row = s.ring + (size_t)(r % RING) * Vp;
#pragma unroll
        for (int i = 0; i < NS; ++i) cur[i] = (g * NS + i < E) ? row[cls[i]] + acc[i] : -INFINITY;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_addr(s, BAR_FREE + r % RING));
        float4* dst = reinterpret_cast<float4*>(out + (long long)r * a.epad);
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) dst[i] = make_float4(cur[4 * i], cur[4 * i + 1], cur[4 * i + 2], cur[4 * i + 3]);
        publish<NS, false>(cur, s, w, W, lane, r & 1);
      }
      // ll = log2-sum of alpha[Tb-1, E-1] and alpha[Tb-1, E-2]: the owners drop them into the exchange slots
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
  } else if (threadIdx.x == 0) {
    a.nll[b] = (S == 0 && Tb <= 0) ? 0.f : INFINITY;
  }
  // ---- reduced loss: the last CTA to finish sums nll in a fixed order (deterministic), and re-arms the ticket
  if (a.loss != nullptr) {
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      __threadfence();
      int* ticket = const_cast<int*>(a.in_len) + a.B;
      s_last = (atomicAdd(ticket, 1) == a.B - 1);
      if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last && w == 0) {
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

// ================================================================================================ beta sweep + gradient
template <int NS>
__global__ void __launch_bounds__(32 * 9) ctc_beta_grad_kernel(const CtcArgs a) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = a.W;
  const int V = a.V, Vp = vpad(V);
  const Smem s = carve(smem_f, V, W, NS, true);
  const int Tb = min(a.in_len[b], a.T);
  const int S = a.tgt_len[b];
  const int E = 2 * S + 1;
  const float nll = a.nll[b];
  const bool dead = (Tb <= 0) || (E > 32 * W * NS) || isinf(nll) || isnan(nll);
  float* gb = a.grad + (long long)b * a.gb;
  // rows past the utterance (and whole infeasible utterances: zero_infinity semantics; without zero_infinity the
  // reference's gradient is NaN garbage, zero is returned there as well) are zero
  {
    const int t0 = dead ? 0 : Tb;
    const long long n = (long long)(a.T - t0) * V;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) gb[(t0 + i / V) * a.gt + (i % V)] = 0.f;
  }
  if (dead) return;
  init_barriers(a, s, true);
  for (int i = threadIdx.x; i < 2 * Vp; i += blockDim.x) s.bins[i] = 0.f;
  __syncthreads();
  float scale = a.grad_out[(long long)b * a.go_stride];
  if (a.mean) scale /= (float)(max(S, 1) * a.B);
  const float ll2 = -nll * LOG2E;  // log2 likelihood

  if (w >= W + NPROD) {
    // ------------------------------------------------------------------------------------ emitter
#pragma unroll 1
    for (int r = 0; r < Tb; ++r) {
      const int par = r & 1;
      mbar_wait(bar_addr(s, BAR_BINFULL + par), (r >> 1) & 1);
      const float* row = s.ring + (size_t)(r % RING) * Vp;  // still held: the emitter is one of the row's readers
      float* bins = s.bins + par * Vp;
      float tot = 0.f;
      for (int v = lane; v < Vp; v += 32) tot += bins[v];
      tot = warp_sum(tot);
      const float inv = (tot > 0.f) ? 1.f / tot : 0.f;
      float* gout = gb + (long long)(Tb - 1 - r) * a.gt;
      for (int v = lane; v < V; v += 32) {
        gout[v] = (ex2f(row[v]) - bins[v] * inv) * scale;
        bins[v] = 0.f;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_addr(s, BAR_BINFREE + par));
        mbar_arrive(bar_addr(s, BAR_FREE + r % RING));
      }
    }
  } else if (w >= W) {
    producer_loop<true>(a, s, b, Tb, w - W, lane, W + 1);
  } else {
    // ------------------------------------------------------------------------------------ recursion
    const int g = w * 32 + lane;
    int cls[NS];
    bool skp[NS];
    state_setup<NS, true>(a, b, g, E, cls, skp);
    // this lane's NS states of alpha, AHEAD time steps in flight (per-thread cp.async: no cross-thread hand-over)
    const float* asrc = a.alpha + (long long)b * a.T * a.epad + (long long)g * NS;
    const uint32_t adst = smem_u32(s.aring + (size_t)g * NS);
    const uint32_t apitch = (uint32_t)(32 * W * NS) * 4u;
    auto issue_alpha = [&](int r) {
      if (r < Tb) {
        const float* src = asrc + (long long)(Tb - 1 - r) * a.epad;
        const uint32_t dst = adst + (uint32_t)(r % (AHEAD + 1)) * apitch;
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) cp_async16(dst + 16u * i, src + 4 * i);
      }
      cp_async_commit();
    };
#pragma unroll 1
    for (int r = 0; r < AHEAD; ++r) issue_alpha(r);
    float cur[NS];
#pragma unroll 1
    for (int r = 0; r < Tb; ++r) {
      const int par = r & 1;
      float inner[NS];
      if (r == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const int st = g * NS + i;
          inner[i] = (st < E && st >= E - 2) ? 0.f : -INFINITY;
        }
      } else {
        float n1, n2;
        neighbours<NS, true>(cur, s, w, W, lane, (r - 1) & 1, n1, n2);
        recur<NS, true>(cur, n1, n2, skp, inner);
      }
      issue_alpha(r + AHEAD);
      cp_async_wait<AHEAD>();
      const float* al = s.aring + (size_t)(r % (AHEAD + 1)) * (32 * W * NS) + (size_t)g * NS;
      mbar_wait(bar_addr(s, BAR_FULL + r % RING), (r / RING) & 1);
      const float* row = s.ring + (size_t)(r % RING) * Vp;
      if (r >= 2) mbar_wait(bar_addr(s, BAR_BINFREE + par), ((r >> 1) - 1) & 1);
      float* bins = s.bins + par * Vp;
      float blank_sum = 0.f;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int st = g * NS + i;
        const bool live = st < E;
        cur[i] = live ? row[cls[i]] + inner[i] : -INFINITY;
        // occupancy of state st at this time step: alpha (emission included) * beta without the emission / likelihood
        const float o = al[i] + inner[i] - ll2;
        const float occ = (live && o > -INFINITY) ? ex2f(o) : 0.f;
        if (i & 1) {
          if (occ > 0.f) atomicAdd(&bins[cls[i]], occ);
        } else {
          blank_sum += occ;
        }
      }
      blank_sum = warp_sum(blank_sum);
      if (lane == 0 && blank_sum > 0.f) atomicAdd(&bins[a.blank], blank_sum);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_addr(s, BAR_BINFULL + par));
        mbar_arrive(bar_addr(s, BAR_FREE + r % RING));
      }
      publish<NS, true>(cur, s, w, W, lane, par);
    }
    cp_async_wait<0>();
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
                              int32_t reduction_mean, int32_t zero_infinity, float* alpha, float* nll, float* loss,
                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem T=%d B=%d V=%d", T, B, V);
  A8_REQUIRE(blank >= 0 && blank < V, "ctc: blank %d outside [0,%d)", blank, V);
  A8_REQUIRE(max_S >= 0 && max_S <= 511, "ctc: target length %d unsupported (max 511)", max_S);
  int ns;
  const int W = pick_layout(max_S, &ns);
  const size_t smem = ctc_smem_bytes(V, W, ns, false);
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large for the row ring", V);
  CtcArgs a{};
  a.x = x; a.st = stride_t; a.sb = stride_b; a.sv = stride_v; a.T = T; a.B = B; a.V = V; a.from_logits = from_logits;
  a.targets = targets; a.tgt_off = tgt_offsets; a.tgt_len = tgt_lengths; a.in_len = in_lengths; a.blank = blank;
  a.W = W; a.epad = 32 * W * ns; a.alpha = alpha; a.nll = nll; a.loss = loss; a.mean = reduction_mean;
  a.zero_inf = zero_infinity;
  const int threads = 32 * (W + NPROD);
  if (ns == 4) {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_alpha_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_alpha_kernel<4><<<B, threads, smem, stream>>>(a);
  } else {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_alpha_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_alpha_kernel<8><<<B, threads, smem, stream>>>(a);
  }
  return check_launch("ctc_alpha_kernel");
}

extern "C" int a8_ctc_backward(const float* x, int64_t stride_t, int64_t stride_b, int64_t stride_v, int32_t T, int32_t B,
                               int32_t V, int32_t from_logits, const int32_t* targets, const int32_t* tgt_offsets,
                               const int32_t* tgt_lengths, const int32_t* in_lengths, int32_t max_S, int32_t blank,
                               const float* alpha, const float* nll, const float* grad_out, int64_t grad_out_stride,
                               int32_t reduction_mean, int32_t zero_infinity, float* grad, int64_t grad_stride_t,
                               int64_t grad_stride_b, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(T > 0 && B > 0 && V > 0, "ctc: empty problem");
  int ns;
  const int W = pick_layout(max_S, &ns);
  const size_t smem = ctc_smem_bytes(V, W, ns, true);
  A8_REQUIRE(smem <= 200 * 1024, "ctc: vocabulary %d too large", V);
  CtcArgs a{};
  a.x = x; a.st = stride_t; a.sb = stride_b; a.sv = stride_v; a.T = T; a.B = B; a.V = V; a.from_logits = from_logits;
  a.targets = targets; a.tgt_off = tgt_offsets; a.tgt_len = tgt_lengths; a.in_len = in_lengths; a.blank = blank;
  a.W = W; a.epad = 32 * W * ns; a.alpha = const_cast<float*>(alpha); a.nll = const_cast<float*>(nll);
  a.mean = reduction_mean; a.zero_inf = zero_infinity; a.grad_out = grad_out; a.go_stride = grad_out_stride;
  a.grad = grad; a.gt = grad_stride_t; a.gb = grad_stride_b;
  const int threads = 32 * (W + NPROD + 1);
  if (ns == 4) {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_beta_grad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_beta_grad_kernel<4><<<B, threads, smem, stream>>>(a);
  } else {
    if (smem > 48 * 1024) A8_CUDA(cudaFuncSetAttribute(ctc_beta_grad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_beta_grad_kernel<8><<<B, threads, smem, stream>>>(a);
  }
  return check_launch("ctc_beta_grad_kernel");
}
