// audio8_b200 — optimizer side of the training step: gradient-norm clipping + AdamW as two multi-tensor
// launches over every parameter of the model, replacing `torch.nn.utils.clip_grad_norm_` + `torch.optim.AdamW.step`
// (+ eight_mile's `scale_grads`) at /root/reference/audio8/pretrain.py:182-184 and train.py:323-325.
//
// HBM-bound: per parameter element the update reads p, g, m, v and writes p, m, v (28 B; +2 B when the bf16 operand
// copy of the weight is refreshed in the same pass), the norm pass reads g once (4 B).  Work is cut into fixed-size
// chunks (host-built chunk -> tensor map, static per model), one CTA per chunk, 128-bit accesses.
// The squared norm is reduced deterministically: one partial per chunk, and every CTA of the update kernel folds the
// partials (a few KB, L2-resident) in double precision, so no atomics, no zero-fill, no host synchronisation.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

struct OptTensor {      // one row of the device table (6 x int64)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  __nv_bfloat16* pb;    // optional bf16 copy of the updated parameter (GEMM operand), or null
};

constexpr int OPT_THREADS = 256;

__device__ __forceinline__ float block_sum(float x, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = x;
  __syncthreads();
  float t = (threadIdx.x < OPT_THREADS / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(OPT_THREADS)
optim_sqnorm_kernel(const OptTensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                    const long long* __restrict__ chunk_off, int n_chunks, int chunk, float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const OptTensor t = tab[chunk_tensor[c]];
    const long long off = chunk_off[c];
    const int len = (int)min((long long)chunk, t.n - off);
    const float* g = t.g + off;
    float s = 0.f;
    if (t.g != nullptr) {
      if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(g);
        const int n4 = len >> 2;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
          const float4 a = __ldg(g4 + i);
          s0 = fmaf(a.x, a.x, s0); s1 = fmaf(a.y, a.y, s1); s2 = fmaf(a.z, a.z, s2); s3 = fmaf(a.w, a.w, s3);
        }
        s = (s0 + s1) + (s2 + s3);
        for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) s = fmaf(g[i], g[i], s);
      } else {
        for (int i = threadIdx.x; i < len; i += OPT_THREADS) s = fmaf(g[i], g[i], s);
      }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partials[c] = s;
  }
}

// every constant below is derived on the host in double precision and rounded once, exactly like the Python scalars
// torch.optim.AdamW hands to its kernels (1.f - 0.999f is 4.7e-5 away from float(1 - 0.999))
struct AdamArgs {
  float decay;      // 1 - lr * weight_decay
  float omb1;       // 1 - beta1
  float beta2, omb2;
  float eps, bc2_sqrt, step_size;  // step_size = lr / bias_correction1
  float max_norm, grad_scale;
  int scale_grads_only;  // 1: clip_grad_norm_ / scale_grads semantics (scale the gradients in place, touch nothing else)
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a, float gs) {
  g *= gs;
  p *= a.decay;                                     // torch: param.mul_(1 - lr * weight_decay)
  m = m + (g - m) * a.omb1;                         // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + a.omb2 * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= a.step_size * (m / denom);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(OPT_THREADS)
optim_adamw_kernel(const OptTensor* __restrict__ tab, const int* __restrict__ chunk_tensor,
                   const long long* __restrict__ chunk_off, int n_chunks, int chunk, const float* __restrict__ partials,
                   int n_partials, AdamArgs a, float* __restrict__ total_norm_out) {
  __shared__ double dred[OPT_THREADS / 32];
  __shared__ float s_gs;
  // ---- global gradient norm from the per-chunk partials (deterministic, double accumulation) -> clip coefficient
  float gs = a.grad_scale;
  if (partials != nullptr) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) s += (double)partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < OPT_THREADS / 32; ++i) t += dred[i];
      const float total = (float)sqrt(t) * fabsf(a.grad_scale);  // norm of the gradients AFTER scale_grads
      if (blockIdx.x == 0 && total_norm_out != nullptr) *total_norm_out = total;
      float coef = 1.f;
      if (a.max_norm > 0.f) coef = fminf(a.max_norm / (total + 1e-6f), 1.f);  // torch.nn.utils.clip_grad_norm_
      s_gs = a.grad_scale * coef;
    }
    __syncthreads();
    gs = s_gs;
  }
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const OptTensor t = tab[chunk_tensor[c]];
    if (t.g == nullptr) continue;  // parameter without a gradient this step: untouched, like torch's optimizers
    const long long off = chunk_off[c];
    const int len = (int)min((long long)chunk, t.n - off);
    float* gw = const_cast<float*>(t.g) + off;
    if (a.scale_grads_only) {
      for (int i = threadIdx.x; i < len; i += OPT_THREADS) gw[i] *= gs;
      continue;
    }
    float* p = t.p + off;
    const float* g = t.g + off;
    float* m = t.m + off;
    float* v = t.v + off;
    __nv_bfloat16* pb = t.pb ? t.pb + off : nullptr;
    const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(pb) & 7u) == 0);
    int done = 0;
    if (vec) {
      const int n4 = len >> 2;
      float4* p4 = reinterpret_cast<float4*>(p);
      const float4* g4 = reinterpret_cast<const float4*>(g);
      float4* m4 = reinterpret_cast<float4*>(m);
      float4* v4 = reinterpret_cast<float4*>(v);
      for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = __ldg(g4 + i);
        adam_elem(pp.x, gg.x, mm.x, vv.x, a, gs);
        adam_elem(pp.y, gg.y, mm.y, vv.y, a, gs);
        adam_elem(pp.z, gg.z, mm.z, vv.z, a, gs);
        adam_elem(pp.w, gg.w, mm.w, vv.w, a, gs);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (pb != nullptr) {
          uint2 o;
          o.x = pack_bf16(pp.x, pp.y);
          o.y = pack_bf16(pp.z, pp.w);
          reinterpret_cast<uint2*>(pb)[i] = o;
        }
      }
      done = n4 << 2;
    }
    for (int i = done + threadIdx.x; i < len; i += OPT_THREADS) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_elem(pp, g[i], mm, vv, a, gs);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (pb != nullptr) pb[i] = __float2bfloat16(pp);
    }
  }
}

int opt_grid(int n_chunks) {
  const int cap = 148 * 8;
  return n_chunks < cap ? n_chunks : cap;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_optim_grad_sqnorm(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                                    int32_t n_chunks, int32_t chunk, float* partials, void* stream_v) {
  A8_REQUIRE(n_chunks > 0 && chunk > 0 && chunk % 4 == 0, "optim_grad_sqnorm: bad chunking");
  optim_sqnorm_kernel<<<opt_grid(n_chunks), OPT_THREADS, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<const OptTensor*>(table), chunk_tensor, reinterpret_cast<const long long*>(chunk_off), n_chunks, chunk,
      partials);
  return check_launch("optim_sqnorm_kernel");
}

extern "C" int a8_optim_adamw(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_off, int32_t n_chunks,
                              int32_t chunk, const float* partials, float max_norm, float grad_scale, double lr,
                              double beta1, double beta2, double eps, double weight_decay, double bias_correction1,
                              double bias_correction2_sqrt, int32_t scale_grads_only, float* total_norm_out,
                              void* stream_v) {
  A8_REQUIRE(n_chunks > 0 && chunk > 0 && chunk % 4 == 0, "optim_adamw: bad chunking");
  A8_REQUIRE(max_norm <= 0.f || partials != nullptr, "optim_adamw: clipping needs the partial squared norms");
  AdamArgs a{(float)(1.0 - lr * weight_decay), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
             (float)bias_correction2_sqrt, (float)(bias_correction1 != 0.0 ? lr / bias_correction1 : 0.0), max_norm,
             grad_scale, scale_grads_only};
  optim_adamw_kernel<<<opt_grid(n_chunks), OPT_THREADS, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<const OptTensor*>(table), chunk_tensor, reinterpret_cast<const long long*>(chunk_off), n_chunks, chunk,
      partials, n_chunks, a, total_norm_out);
  return check_launch("optim_adamw_kernel");
}
