// audio8_b200 — HBM-bound row kernels: LayerNorm (+residual, +dropout) fwd/bwd, column sums (bias gradients),
// element-wise dropout, GELU backward, log-softmax fwd/bwd.  (Attention's softmax lives inside the fused kernels, attn.cu.)
//
// Replaces the ATen kernels the reference dispatches for nn.LayerNorm (wav2vec2.py:623,639,904,930 and the
// ln1/ln2 of every eight_mile TransformerEncoder layer), the residual adds and nn.Dropout around them,
// softmax / masked_fill / dropout inside eight_mile's SeqScaledDotProductAttention, and F.log_softmax
// (wav2vec2.py:770).  One warp owns one row; 128-bit accesses; statistics in fp32; no shared-memory staging
// of rows (each element is read exactly once).
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
const unsigned long long* seed_source();  // a8_api.cu
namespace {

// ------------------------------------------------------------------------------------------------
// counter-based dropout mask: Philox-4x32 (7 rounds) -> 8 x 16-bit uniforms per call, one call per group of
// 8 consecutive elements.  Forward and backward regenerate the same mask from (seed, element group).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox8(unsigned long long seed, unsigned long long group) {
  uint32_t c0 = (uint32_t)group, c1 = (uint32_t)(group >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
struct DropMask8 {
  float m[8];
};
// keep-scale per element: 0 or 1/(1-p)
// Seeds: every launch passes a per-call-site constant; kernels add the 64-bit word at `seed_source()` (device memory,
// nullable) so that a captured CUDA graph draws fresh masks on every replay (a8_set_seed_source).
__device__ __forceinline__ unsigned long long seed_base(const unsigned long long* src) { return seed_base_ld(src); }

__device__ __forceinline__ DropMask8 drop_mask8(float p, unsigned long long seed, unsigned long long group) {
  DropMask8 d;
  if (p <= 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d.m[i] = 1.f;
    return d;
  }
  const uint4 r = philox8(seed, group);
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const float s = 1.f / (1.f - p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    d.m[2 * i] = ((w[i] & 0xFFFFu) >= thr) ? s : 0.f;
    d.m[2 * i + 1] = ((w[i] >> 16) >= thr) ? s : 0.f;
  }
  return d;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward:  s = x + drop_h(h);  y = drop_y(LN(s) * gamma + beta)
// ------------------------------------------------------------------------------------------------
struct LnFwdArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* h;  // nullable
  float p_h;
  unsigned long long seed_h;
  __nv_bfloat16* s_out;  // nullable (required when h != null and backward is wanted)
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* y;
  float* y_f32;  // nullable
  float p_y;
  unsigned long long seed_y;
  float* mean;
  float* rstd;
  int R, C;
  const unsigned long long* seed_src;
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 t;
  t = unpack_bf16(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// NCH chunks of 8 elements per lane (C <= NCH*256); FULL: C == NCH*256, no column guards; HAS_H: residual input.
// Every global load of a row is issued before the first dependent instruction (the earlier version interleaved one
// load, its Philox mask and its arithmetic per chunk behind branches: nine serial memory round trips per row).
// gamma / beta are read where they are used (a warp normally owns ONE row, so holding them in 2 x 8 x NCH registers bought
// nothing and capped the kernel at 2 CTAs per SM: the 4494-row launches ran as two waves): with <= 64 registers the
// whole launch is resident at once.
template <int NCH, bool FULL, bool HAS_H>
__global__ void __launch_bounds__(256, NCH <= 3 ? 4 : 2) ln_fwd_kernel(const LnFwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long sbase = seed_base(a.seed_src);
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int C = a.C;
  const float invC = 1.f / (float)C;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < a.R; row += warps) {
    const long long ro = (long long)row * C;
    uint4 xr[NCH], hr[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      const bool ok = FULL || c < C;
      xr[i] = ok ? ldg128(a.x + ro + c) : make_uint4(0u, 0u, 0u, 0u);
      if (HAS_H) hr[i] = ok ? ldg128(a.h + ro + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    float v[NCH][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      const bool ok = FULL || c < C;
      unpack8(xr[i], v[i]);
      if (HAS_H) {
        float hv[8];
        unpack8(hr[i], hv);
        const DropMask8 d = drop_mask8(a.p_h, a.seed_h + sbase, (unsigned long long)(ro + c) >> 3);
        // statistics are taken on the bf16-rounded sum, the value backward re-reads
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(__float2bfloat16(v[i][j] + hv[j] * d.m[j]));
        if (ok && a.s_out != nullptr) store8(a.s_out + ro + c, v[i]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];  // guarded-off chunks hold zeros
    }
    const float mu = warp_sum(sum) * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (FULL || c < C) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mu;
          sq += d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(sq) * invC + a.eps);
    if (lane == 0) {
      a.mean[row] = mu;
      a.rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (FULL || c < C) {
        float o[8], g[8], bt[8];
        load8f(a.gamma + c, g);
        load8f(a.beta + c, bt);
        const DropMask8 d = drop_mask8(a.p_y, a.seed_y + sbase, (unsigned long long)(ro + c) >> 3);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ((v[i][j] - mu) * rs * g[j] + bt[j]) * d.m[j];
        store8(a.y + ro + c, o);
        if (a.y_f32 != nullptr) store8f(a.y_f32 + ro + c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  g = drop_y(dy (+ dy_f32));  ds = rstd * (g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
// (+ dres: extra gradient arriving directly at s); dh = drop_h(ds); dgamma += g*xhat; dbeta += g; dbias_h += dh
// ------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const __nv_bfloat16* dy;
  const float* dy_f32;  // nullable, added to dy
  float p_y;
  unsigned long long seed_y;
  const __nv_bfloat16* s;
  const float* mean;
  const float* rstd;
  const float* gamma;
  __nv_bfloat16* ds;
  __nv_bfloat16* dh;  // nullable
  float p_h;
  unsigned long long seed_h;
  float* dgamma;   // [C] accumulated (atomics)
  float* dbeta;    // [C]
  float* dbias_h;  // nullable [C]
  int R, C;
  const unsigned long long* seed_src;
};

// One warp per row, two rows in flight per warp (the next row's dy / s tiles are loaded before the current row's
// arithmetic starts); column partials (dgamma, dbeta, dbias) are reduced across the CTA's warps in shared memory and
// leave as 16-byte vector reductions (one red.v4 per 4 columns per CTA).
template <int NCH, bool FULL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned long long sbase = seed_base(a.seed_src);
  __shared__ __align__(16) float red[8][NCH * 256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int C = a.C;
  const float invC = 1.f / (float)C;
  float acc_g[NCH][8], acc_b[NCH][8], acc_h[NCH][8], gm[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = (lane + 32 * i) * 8;
    load8f(a.gamma + ((FULL || c < C) ? c : 0), gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_g[i][j] = acc_b[i][j] = acc_h[i][j] = 0.f;
  }
  const bool has_f32 = a.dy_f32 != nullptr;
  uint4 dyr[NCH], sr[NCH];
  float mu = 0.f, rs = 0.f;
  auto load_row = [&](int row) {
    const long long ro = (long long)row * C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      const bool ok = FULL || c < C;
      dyr[i] = ok ? ldg128(a.dy + ro + c) : make_uint4(0u, 0u, 0u, 0u);
      sr[i] = ok ? ldg128(a.s + ro + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    mu = __ldg(a.mean + row);
    rs = __ldg(a.rstd + row);
  };
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row < a.R) load_row(row);
  for (; row < a.R; row += warps) {
    const long long ro = (long long)row * C;
    float g[NCH][8], xh[NCH][8];
    float s1 = 0.f, s2 = 0.f;
    const float mu_c = mu, rs_c = rs;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      const bool ok = FULL || c < C;
      float dy[8], sv[8];
      unpack8(dyr[i], dy);
      unpack8(sr[i], sv);
      if (has_f32 && ok) {
        float e[8];
        load8f(a.dy_f32 + ro + c, e);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[j] += e[j];
      }
      const DropMask8 d = drop_mask8(a.p_y, a.seed_y + sbase, (unsigned long long)(ro + c) >> 3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = ok ? dy[j] * d.m[j] : 0.f;
        xh[i][j] = ok ? (sv[j] - mu_c) * rs_c : 0.f;
        acc_g[i][j] += gg * xh[i][j];
        acc_b[i][j] += gg;
        g[i][j] = gg * gm[i][j];
        s1 += g[i][j];
        s2 += g[i][j] * xh[i][j];
      }
    }
    if (row + warps < a.R) load_row(row + warps);  // next row's tiles are in flight during this row's reductions
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      if (FULL || c < C) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs_c * (g[i][j] - s1 - xh[i][j] * s2);
        store8(a.ds + ro + c, o);
        if (a.dh != nullptr) {
          const DropMask8 d = drop_mask8(a.p_h, a.seed_h + sbase, (unsigned long long)(ro + c) >> 3);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] *= d.m[j];
            acc_h[i][j] += o[j];
          }
          store8(a.dh + ro + c, o);
        } else if (a.dbias_h != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_h[i][j] += o[j];
        }
      }
    }
  }
  // block reduction of the column partials, then one 16-byte vector reduction per 4 columns per CTA
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? a.dgamma : (pass == 1 ? a.dbeta : a.dbias_h);
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = (lane + 32 * i) * 8;
      float* rp = &red[w][c];
      if (pass == 0) { store8f(rp, acc_g[i]); } else if (pass == 1) { store8f(rp, acc_b[i]); } else { store8f(rp, acc_h[i]); }
    }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 u = *reinterpret_cast<const float4*>(&red[k][c]);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                   : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sum: out[c] += sum_r x[r][c]   (bf16 [R, ld] -> fp32 atomics; out zeroed by the caller)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* x, long long ld, int R, int C,
                                                     int rows_per_block, float* out) {
  pdl_launch_dependents();
  pdl_wait();
  // 64 column-octets x 4 row phases per CTA (512 columns, blockIdx.y selects the slab); each thread streams its
  // rows with 4 independent 16-byte loads in flight, phases are reduced in smem, one atomic per column per CTA
  __shared__ float red[4][64 * 8];
  const int oct = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c = (blockIdx.y * 64 + oct) * 8;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < C) {
    int r = r0 + ph;
    for (; r + 12 < r1; r += 16) {
      float v0[8], v1[8], v2[8], v3[8];
      load8(x + (long long)r * ld + c, v0);
      load8(x + (long long)(r + 4) * ld + c, v1);
      load8(x + (long long)(r + 8) * ld + c, v2);
      load8(x + (long long)(r + 12) * ld + c, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < r1; r += 4) {
      float v[8];
      load8(x + (long long)r * ld + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ph][oct * 8 + j] = acc[j];
  __syncthreads();
  for (int e = threadIdx.x; e < 512; e += 256) {
    const int cc = blockIdx.y * 512 + e;
    if (cc < C) atomicAdd(out + cc, red[0][e] + red[1][e] + red[2][e] + red[3][e]);
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise dropout on bf16 (n % 8 == 0): out = x * mask/(1-p); the same call is its own backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dropout_kernel(const __nv_bfloat16* x, __nv_bfloat16* out, long long n8,
                                                      float p, unsigned long long seed,
                                                      const unsigned long long* seed_src) {
  seed += seed_base(seed_src);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    load8(x + i * 8, v);
    const DropMask8 d = drop_mask8(p, seed, (unsigned long long)i);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= d.m[j];
    store8(out + i * 8, v);
  }
}

__global__ void __launch_bounds__(256) dropout_f32_kernel(const float* x, float* out, long long n8, float p,
                                                          unsigned long long seed,
                                                          const unsigned long long* seed_src) {
  seed += seed_base(seed_src);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    load8f(x + i * 8, v);
    const DropMask8 d = drop_mask8(p, seed, (unsigned long long)i);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= d.m[j];
    store8f(out + i * 8, v);
  }
}

// dz = dy * gelu'(z)   (bf16, n % 8 == 0)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* z,
                                                       __nv_bfloat16* dz, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], zz[8];
    load8(dy + i * 8, g);
    load8(z + i * 8, zz);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= gelu_erf_grad(zz[j]);
    store8(dz + i * 8, g);
  }
}

__global__ void __launch_bounds__(256) mul_dgelu_kernel(const __nv_bfloat16* a, const __half* b, __nv_bfloat16* out,
                                                        long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[8];
    load8(a + i * 8, x);
    const uint4 raw = *reinterpret_cast<const uint4*>(b + i * 8);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = unpack_f16(w[j]);
      x[2 * j] *= t.x;
      x[2 * j + 1] *= t.y;
    }
    store8(out + i * 8, x);
  }
}

// ------------------------------------------------------------------------------------------------
// log-softmax over a small class dimension: x fp32 [R,V] -> y fp32 [R,V];  backward dx = dy - exp(y)*sum(dy)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_softmax_fwd_kernel(const float* x, float* y, int R, int V) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < R; row += warps) {
    const float* xr = x + (long long)row * V;
    float mx = -INFINITY;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < V; c += 32) sum += expf(xr[c] - mx);
    const float lse = mx + logf(warp_sum(sum));
    for (int c = lane; c < V; c += 32) y[(long long)row * V + c] = xr[c] - lse;
  }
}
__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* dy, long long s_r, long long s_v,
                                                              const float* y, __nv_bfloat16* dx, int R, int V,
                                                              int rows_inner, long long s_outer) {
  // dy is addressed as dy[(row / rows_inner) * s_outer + (row % rows_inner) * s_r + c * s_v] so that the
  // [T,B,V] gradient CTC returns can be consumed without a transpose copy
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < R; row += warps) {
    const float* dr = dy + (long long)(row / rows_inner) * s_outer + (long long)(row % rows_inner) * s_r;
    float sum = 0.f;
    for (int c = lane; c < V; c += 32) sum += dr[(long long)c * s_v];
    sum = warp_sum(sum);
    for (int c = lane; c < V; c += 32)
      dx[(long long)row * V + c] =
          __float2bfloat16(dr[(long long)c * s_v] - expf(y[(long long)row * V + c]) * sum);
  }
}

int row_grid(int rows) {
  const int blocks = cdiv(rows, 8);
  return blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_layernorm_fwd(const void* x, const void* h, float p_h, uint64_t seed_h, void* s_out,
                                const float* gamma, const float* beta, float eps, void* y, float* y_f32, float p_y,
                                uint64_t seed_y, float* mean, float* rstd, int32_t R, int32_t C, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && C > 0 && C % 8 == 0 && C <= 1024, "layernorm: unsupported shape R=%d C=%d", R, C);
  LnFwdArgs a{(const __nv_bfloat16*)x, (const __nv_bfloat16*)h, p_h, seed_h, (__nv_bfloat16*)s_out, gamma, beta,
              eps, (__nv_bfloat16*)y, y_f32, p_y, seed_y, mean, rstd, R, C, seed_source()};
  const int nch = cdiv(C, 256);
  const int grid = row_grid(R);
  const bool full = (C == nch * 256);
#define A8_LN_FWD(N, F, HH) A8_CUDA(launch_pdl(ln_fwd_kernel<N, F, HH>, dim3(grid), dim3(256), 0, stream, 1, a))
#define A8_LN_FWD_N(N)                                   \
  do {                                                   \
    if (full && h != nullptr) A8_LN_FWD(N, true, true);  \
    else if (full) A8_LN_FWD(N, true, false);            \
    else if (h != nullptr) A8_LN_FWD(N, false, true);    \
    else A8_LN_FWD(N, false, false);                     \
  } while (0)
  switch (nch) {
    case 1: A8_LN_FWD_N(1); break;
    case 2: A8_LN_FWD_N(2); break;
    case 3: A8_LN_FWD_N(3); break;
    default: A8_LN_FWD_N(4); break;
  }
#undef A8_LN_FWD_N
#undef A8_LN_FWD
  return check_launch("ln_fwd_kernel");
}

extern "C" int a8_layernorm_bwd(const void* dy, const float* dy_f32, float p_y, uint64_t seed_y, const void* s,
                                const float* mean, const float* rstd, const float* gamma, void* ds, void* dh,
                                float p_h, uint64_t seed_h, float* dgamma, float* dbeta, float* dbias_h, int32_t R,
                                int32_t C, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && C > 0 && C % 8 == 0 && C <= 1024, "layernorm_bwd: unsupported shape R=%d C=%d", R, C);
  LnBwdArgs a{(const __nv_bfloat16*)dy, dy_f32, p_y, seed_y, (const __nv_bfloat16*)s, mean, rstd, gamma,
              (__nv_bfloat16*)ds, (__nv_bfloat16*)dh, p_h, seed_h, dgamma, dbeta, dbias_h, R, C, seed_source()};
  A8_REQUIRE(C % 4 == 0 && (reinterpret_cast<uintptr_t>(dgamma) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dbeta) & 15u) == 0 &&
                 (reinterpret_cast<uintptr_t>(dbias_h) & 15u) == 0,
             "layernorm_bwd: accumulators must be 16-byte aligned (vector reductions)");
  const int nch = cdiv(C, 256);
  int grid = cdiv(R, 8 * 2);  // two rows per warp: the column partials amortise their reductions
  const int cap = 148 * (nch >= 3 ? 1 : 2);  // the 3- and 4-chunk variants hold ~250 registers: one CTA per SM
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  const bool full = (C == nch * 256);
#define A8_LN_BWD(N)                                                                                        \
  do {                                                                                                      \
    if (full) A8_CUDA(launch_pdl(ln_bwd_kernel<N, true>, dim3(grid), dim3(256), 0, stream, 1, a));          \
    else A8_CUDA(launch_pdl(ln_bwd_kernel<N, false>, dim3(grid), dim3(256), 0, stream, 1, a));              \
  } while (0)
  switch (nch) {
    case 1: A8_LN_BWD(1); break;
    case 2: A8_LN_BWD(2); break;
    case 3: A8_LN_BWD(3); break;
    default: A8_LN_BWD(4); break;
  }
#undef A8_LN_BWD
  return check_launch("ln_bwd_kernel");
}

extern "C" int a8_colsum(const void* x, int64_t ld, int32_t R, int32_t C, float* out, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0, "colsum: bad shape R=%d C=%d ld=%lld", R, C, (long long)ld);
  const int slabs = cdiv(C, 512);
  int chunks = cdiv(148 * 4, slabs);  // ~4 CTAs per SM over the whole grid
  int rpb = cdiv(R, chunks);
  rpb = rpb < 32 ? 32 : rpb;
  dim3 grid(cdiv(R, rpb), slabs);
  A8_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, 1, (const __nv_bfloat16*)x, (long long)ld, (int)R, (int)C, rpb, out));
  return check_launch("colsum_kernel");
}

extern "C" int a8_dropout(const void* x, void* out, int32_t dtype, int64_t n, float p, uint64_t seed,
                          void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(n > 0 && n % 8 == 0, "dropout: n=%lld must be a positive multiple of 8", (long long)n);
  const long long n8 = n / 8;
  const int grid = (int)(n8 / 256 + 1 > 148 * 8 ? 148 * 8 : n8 / 256 + 1);
  if (dtype == 0) dropout_f32_kernel<<<grid, 256, 0, stream>>>((const float*)x, (float*)out, n8, p, seed, seed_source());
  else dropout_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, n8, p, seed, seed_source());
  return check_launch("dropout_kernel");
}

extern "C" int a8_gelu_bwd(const void* dy, const void* z, void* dz, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(n > 0 && n % 8 == 0, "gelu_bwd: n=%lld must be a positive multiple of 8", (long long)n);
  const long long n8 = n / 8;
  const int grid = (int)(n8 / 256 + 1 > 148 * 8 ? 148 * 8 : n8 / 256 + 1);
  A8_CUDA(launch_pdl(gelu_bwd_kernel, dim3(grid), dim3(256), 0, stream, 1, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)z,
                     (__nv_bfloat16*)dz, n8));
  return check_launch("gelu_bwd_kernel");
}

extern "C" int a8_mul_dgelu(const void* a, const void* b, void* out, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(n > 0 && n % 8 == 0, "mul_dgelu: n must be a positive multiple of 8");
  const long long n8 = n / 8;
  long long gl = (n8 + 255) / 256;
  const int grid = (int)(gl > 148 * 8 ? 148 * 8 : gl);
  A8_CUDA(launch_pdl(mul_dgelu_kernel, dim3(grid), dim3(256), 0, stream, 1, (const __nv_bfloat16*)a, (const __half*)b,
                     (__nv_bfloat16*)out, n8));
  return check_launch("mul_dgelu_kernel");
}

extern "C" int a8_log_softmax_fwd(const float* x, float* y, int32_t R, int32_t V, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && V > 0, "log_softmax: bad shape");
  log_softmax_fwd_kernel<<<row_grid(R), 256, 0, stream>>>(x, y, R, V);
  return check_launch("log_softmax_fwd_kernel");
}

extern "C" int a8_log_softmax_bwd(const float* dy, int64_t stride_outer, int64_t stride_row, int64_t stride_v,
                                  int32_t rows_inner, const float* y, void* dx, int32_t R, int32_t V,
                                  void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && V > 0 && rows_inner > 0, "log_softmax_bwd: bad shape");
  log_softmax_bwd_kernel<<<row_grid(R), 256, 0, stream>>>(dy, stride_row, stride_v, y, (__nv_bfloat16*)dx, R, V,
                                                          rows_inner, stride_outer);
  return check_launch("log_softmax_bwd_kernel");
}
