// audio8_b200 — Gumbel vector quantizer row kernels (reference: wav2vec2.py:547-576).
//
// The reference runs ~12 ATen kernels and materialises a [rows, G*V, var_dim] broadcast product (680 MB at the
// base config) to select one codeword per (row, group).  Here one warp owns one (row, group): it reads the V
// logits (+ Gumbel noise) once, accumulates the pooled softmax for the perplexity, takes the arg-max and
// copies the selected codeword (a 512-byte gather).  Backward is the closed form of D.2; the only dense
// contraction it needs (dq . vars^T) is done by the tcgen05 GEMM beforehand.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr int VPL = 16;  // logits per lane: V <= 512

struct VqArgs {
  const float* z;      // [R, G*V]
  const float* noise;  // [R*G, V] or null (eval)
  float tau;
  const float* vars;   // [G*V, vd]
  int R, G, V, vd;
  const int* n_valid;  // device scalar: rows [0, *n_valid) of the R rows exist, the rest is padding; null = all
  float* q;            // [R, G*vd]
  __nv_bfloat16* q_bf16;  // nullable
  int* kidx;           // [R*G]
  float* avg_sums;     // [V]
  // backward
  const float* a;      // [R, G*V]  = dq . vars^T per group
  const float* dq;     // [R, G*vd]
  const float* ppl;
  const float* dppl;
  __nv_bfloat16* dz;   // [R, G*V]
  float* dvars;        // [G*V, vd]
};

// All loads of the row are issued before the first dependent instruction: clamped addresses instead of per-lane
// branches (a divergent `if (v < V)` around load + arithmetic serialised the 2 x VPL loads of a warp, which is what
// these L2-resident kernels spend their time on).
__device__ __forceinline__ void load_row(const VqArgs& a, int n, int lane, float (&zv)[VPL], float (&uv)[VPL]) {
  const int r = n / a.G, g = n - r * a.G;
  const float* zr = a.z + ((long long)r * a.G + g) * a.V;
  const float* nr = a.noise ? a.noise + (long long)n * a.V : zr;
  const bool has_noise = a.noise != nullptr;
  const float it = 1.f / a.tau;
  const int nvl = (a.V + 31) >> 5;  // warp-uniform trip count
  float zraw[VPL], nraw[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (i < nvl) {
      const int vc = min(lane + 32 * i, a.V - 1);
      zraw[i] = __ldg(zr + vc);
      nraw[i] = __ldg(nr + vc);
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int v = lane + 32 * i;
    if (i < nvl && v < a.V) {
      zv[i] = zraw[i];
      uv[i] = has_noise ? (zraw[i] + nraw[i]) * it : zraw[i];
    } else {
      zv[i] = -INFINITY;
      uv[i] = -INFINITY;
    }
  }
}

// softmax of the lane-distributed row in place
__device__ __forceinline__ void softmax_row(float (&x)[VPL]) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPL; ++i) mx = fmaxf(mx, x[i]);
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    x[i] = expf(x[i] - mx);  // exp(-inf) = 0 for the padding slots
    sum += x[i];
  }
  const float inv = 1.f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < VPL; ++i) x[i] *= inv;
}

__device__ __forceinline__ int vq_rows(const VqArgs& a) { return a.n_valid ? min(__ldg(a.n_valid), a.R) : a.R; }

__global__ void __launch_bounds__(256) vq_fwd_kernel(const VqArgs a) {
  extern __shared__ float s_avg[];  // [V]
  for (int v = threadIdx.x; v < a.V; v += blockDim.x) s_avg[v] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int N = a.R * a.G;
  const int Nv = vq_rows(a) * a.G;  // padding rows still get a (meaningless) code, but stay out of the statistics
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    float zv[VPL], uv[VPL];
    load_row(a, n, lane, zv, uv);
    // arg-max of u, first index wins ties (torch.max / argmax semantics)
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < a.V && (uv[i] > best || bi == 0x7fffffff)) {
        best = uv[i];
        bi = v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    softmax_row(zv);  // s = softmax(z): pooled perplexity statistics (wav2vec2.py:554)
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < a.V && n < Nv) atomicAdd(&s_avg[v], zv[i]);
    }
    const int r = n / a.G, g = n - r * a.G;
    if (lane == 0) a.kidx[n] = bi;
    const float* cw = a.vars + ((long long)g * a.V + bi) * a.vd;
    const long long qo = ((long long)r * a.G + g) * a.vd;
    for (int d = lane; d < a.vd; d += 32) {
      const float c = cw[d];
      a.q[qo + d] = c;
      if (a.q_bf16) a.q_bf16[qo + d] = __float2bfloat16(c);
    }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < a.V; v += blockDim.x) atomicAdd(a.avg_sums + v, s_avg[v]);
}

// ppl = exp(-sum_v q_v log(q_v + 1e-7)),  q = avg_sums / N   (wav2vec2.py:565)
__global__ void vq_ppl_kernel(const float* avg_sums, int V, int N, const int* n_valid, int G, float* ppl) {
  if (n_valid != nullptr) N = min(N, *n_valid * G);
  float acc = 0.f;
  for (int v = threadIdx.x; v < V; v += 32) {
    const float q = avg_sums[v] / (float)N;
    acc += q * logf(q + 1e-7f);
  }
  acc = warp_sum(acc);
  if (threadIdx.x == 0) *ppl = expf(-acc);
}

__global__ void __launch_bounds__(256) vq_bwd_kernel(const VqArgs a) {
  extern __shared__ float s_dqb[];  // [V] : d loss / d q_bar
  const int Nall = a.R * a.G;
  const int N = vq_rows(a) * a.G;
  {
    const float c = (*a.dppl) * (*a.ppl) / (float)N;
    for (int v = threadIdx.x; v < a.V; v += blockDim.x) {
      const float q = a.avg_sums[v] / (float)N;
      s_dqb[v] = -c * (logf(q + 1e-7f) + q / (q + 1e-7f));
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const float it = 1.f / a.tau;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < Nall; n += warps) {
    const int r = n / a.G, g = n - r * a.G;
    if (n >= N) {  // padding row: no gradient
      __nv_bfloat16* dzp = a.dz + ((long long)r * a.G + g) * a.V;
      for (int v = lane; v < a.V; v += 32) dzp[v] = __float2bfloat16(0.f);
      continue;
    }
    float zv[VPL], uv[VPL];
    load_row(a, n, lane, zv, uv);
    softmax_row(zv);  // s
    float dot_s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < a.V) dot_s += zv[i] * s_dqb[v];
    }
    dot_s = warp_sum(dot_s);
    float dzv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      dzv[i] = (v < a.V) ? zv[i] * (s_dqb[v] - dot_s) : 0.f;
    }
    if (a.noise != nullptr) {  // straight-through Gumbel-softmax: gradient flows through softmax(u)
      softmax_row(uv);         // p
      const float* ar = a.a + ((long long)r * a.G + g) * a.V;
      float av[VPL];
      float dot_p = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        av[i] = (v < a.V) ? ar[v] : 0.f;
        dot_p += uv[i] * av[i];
      }
      dot_p = warp_sum(dot_p);
#pragma unroll
      for (int i = 0; i < VPL; ++i) dzv[i] += uv[i] * (av[i] - dot_p) * it;
    }
    __nv_bfloat16* dzr = a.dz + ((long long)r * a.G + g) * a.V;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + 32 * i;
      if (v < a.V) dzr[v] = __float2bfloat16(dzv[i]);
    }
    // codebook gradient: the forward value is exactly the selected codeword
    const int k = a.kidx[n];
    float* dv = a.dvars + ((long long)g * a.V + k) * a.vd;
    const float* dqr = a.dq + ((long long)r * a.G + g) * a.vd;
    for (int d = lane; d < a.vd; d += 32) atomicAdd(dv + d, dqr[d]);
  }
}

int vq_grid(int N) {  // one (row, group) per warp when they fit in one wave of 8-warp CTAs
  int g = cdiv(N, 8);
  return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g);
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_vq_fwd(const float* z, const float* noise, float tau, const float* vars, int32_t R, int32_t G,
                         int32_t V, int32_t vd, const int32_t* n_valid, float* q, void* q_bf16, int32_t* kidx, float* avg_sums, float* ppl,
                         void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && G > 0 && V > 0 && V <= 32 * VPL && vd > 0, "vq: unsupported shape R=%d G=%d V=%d vd=%d", R, G, V, vd);
  A8_REQUIRE(tau > 0.f, "vq: temperature must be positive");
  VqArgs a{z, noise, tau, vars, R, G, V, vd, n_valid, q, (__nv_bfloat16*)q_bf16, kidx, avg_sums,
           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  A8_CUDA(cudaMemsetAsync(avg_sums, 0, sizeof(float) * V, st));
  vq_fwd_kernel<<<vq_grid(R * G), 256, V * sizeof(float), st>>>(a);
  int rc = check_launch("vq_fwd_kernel");
  if (rc) return rc;
  vq_ppl_kernel<<<1, 32, 0, st>>>(avg_sums, V, R * G, n_valid, G, ppl);
  return check_launch("vq_ppl_kernel");
}

extern "C" int a8_vq_bwd(const float* z, const float* noise, float tau, int32_t R, int32_t G, int32_t V, int32_t vd,
                         const int32_t* n_valid, const float* a_dot, const float* dq, const int32_t* kidx, const float* avg_sums,
                         const float* ppl, const float* dppl, void* dz, float* dvars, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && G > 0 && V > 0 && V <= 32 * VPL && vd > 0, "vq_bwd: unsupported shape");
  VqArgs a{z, noise, tau, nullptr, R, G, V, vd, n_valid, nullptr, nullptr, const_cast<int*>(kidx),
           const_cast<float*>(avg_sums), a_dot, dq, ppl, dppl, (__nv_bfloat16*)dz, dvars};
  vq_bwd_kernel<<<vq_grid(R * G), 256, V * sizeof(float), st>>>(a);
  return check_launch("vq_bwd_kernel");
}
