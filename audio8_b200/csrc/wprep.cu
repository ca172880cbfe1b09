// audio8_b200 — parameter re-layout kernels: fp32 master weights -> bf16 GEMM operands, once per step.
//
// The reference keeps fp32 parameters in PyTorch layouts (Linear [out,in], Conv1d [out,in,k], weight-normed
// pos-conv g [1,1,k] / v [out,in/groups,k]: wav2vec2.py:419,426,600-609).  The tcgen05 GEMM wants bf16, K-major
// operands (conv taps outermost, pos-conv groups padded to 64 channels).  These kernels do the cast, the permutes,
// the weight-norm and its backward in a handful of launches instead of ~100 small PyTorch ops per step.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

// ---------------------------------------------------------------------------------------------- multi-tensor cast
// table[i] = {src (fp32*), dst (bf16* or fp32*), numel, dst_is_f32}; grid (chunks, n)
struct CastEntry {
  const float* src;
  void* dst;
  long long n;
  long long f32;
};
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastEntry* table) {
  const CastEntry e = table[blockIdx.y];
  const long long n4 = e.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (e.f32) {
    float* d = reinterpret_cast<float*>(e.dst);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e.n; i += stride) d[i] = e.src[i];
    return;
  }
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(e.dst);
  const bool vec = ((reinterpret_cast<uintptr_t>(e.src) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d) & 7u) == 0);
  if (vec) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(e.src)[i];
      uint2 o;
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
      reinterpret_cast<uint2*>(d)[i] = o;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e.n; i += stride)
      d[i] = __float2bfloat16(e.src[i]);
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e.n; i += stride)
      d[i] = __float2bfloat16(e.src[i]);
  }
}

// ---------------------------------------------------------------------------------------------- conv weight pack
// w [Cout,Cin,k] fp32 -> wk [Cout, k*Cin] (forward / wgrad layout) and, per stride phase p,
// wt_p [Cin, ntaps_p*Cout] with column i*Cout+co = w[co, ci, taps_p[i]]  (data-gradient layout)
__global__ void __launch_bounds__(256) conv_pack_kernel(const float* w, int Cout, int Cin, int k, int s,
                                                        __nv_bfloat16* wk, __nv_bfloat16* wt0, __nv_bfloat16* wt1) {
  const long long n = (long long)Cout * Cin * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the OUTPUT wk (coalesced writes): i = (co*k + j)*Cin + ci
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int j = (int)(r % k), co = (int)(r / k);
    const __nv_bfloat16 v = __float2bfloat16(w[((long long)co * Cin + ci) * k + j]);
    wk[i] = v;
    if (wt0 != nullptr) {
      const int p = j % s, ord = j / s;  // taps of phase p are p, p+s, ...: ordinal = j / s
      const int ntaps = (k - p + s - 1) / s;
      __nv_bfloat16* wt = p == 0 ? wt0 : wt1;
      wt[(long long)ci * ntaps * Cout + (long long)ord * Cout + co] = v;
    }
  }
}

// dwk [Cout, k*Cin] fp32 -> dw [Cout,Cin,k] fp32
__global__ void __launch_bounds__(256) conv_unpack_kernel(const float* dwk, int Cout, int Cin, int k, float* dw) {
  const long long n = (long long)Cout * Cin * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int j = (int)(r % k), co = (int)(r / k);
    dw[((long long)co * Cin + ci) * k + j] = dwk[i];
  }
}

// ---------------------------------------------------------------------------------------------- pos-conv weight norm
// norm2[j] = sum_{o,i} v[o,i,j]^2   (v [D,cg,k], k contiguous).  grid (chunks), block 128+: thread = tap.
// Deterministic (the packed bf16 weights must not change between two calls on the same parameters): every block
// stores its partial sums, and the block that finishes last adds them in block order.  scratch = [gridDim.x][k]
// partials followed by one zeroed counter word.
__global__ void posconv_norm_kernel(const float* v, int rows, int k, int rows_per_block, float* norm2, float* scratch) {
  const int j = threadIdx.x;
  __shared__ int s_last;
  if (j < k) {
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float x = v[(long long)r * k + j];
      acc = fmaf(x, x, acc);
    }
    scratch[(long long)blockIdx.x * k + j] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (long long)gridDim.x * k);
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && j < k) {
    __threadfence();
    float tot = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) tot += __ldcg(scratch + (long long)b * k + j);
    norm2[j] = tot;
  }
}

// w = g[j] * v / sqrt(norm2[j]); packed bf16 [D, k*64]: wp row = output channel (g*cg+co), col j*64+ci;
// wpt row = input channel (g*cg+ci), col j*64+co; channels >= cg zero.  One thread per (row, j, c64).
__global__ void __launch_bounds__(256) posconv_pack_kernel(const float* g, const float* v, const float* norm2, int D,
                                                           int cg, int k, __nv_bfloat16* wp, __nv_bfloat16* wpt) {
  const long long n = (long long)D * k * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 63);
    const long long r = i >> 6;
    const int j = (int)(r % k), row = (int)(r / k);
    const int grp = row / cg, rr = row - grp * cg;
    float a = 0.f, b = 0.f;
    if (c < cg) {
      const float sc = g[j] * rsqrtf(norm2[j]);
      a = v[((long long)row * cg + c) * k + j] * sc;                        // wp:  (co=rr, ci=c)
      b = v[((long long)(grp * cg + c) * cg + rr) * k + j] * sc;            // wpt: (ci=rr, co=c)
    }
    wp[i] = __float2bfloat16(a);
    if (wpt != nullptr) wpt[i] = __float2bfloat16(b);
  }
}

// backward of the weight norm, from the GEMM's packed gradient dwp fp32 [G, k*64, 64] (dwp[g][j*64+ci][co]):
//   dW[o,i,j] = dwp[g][j*64+i][o - g*cg];  t[j] = sum dW*v;  dg[j] = t[j]/||v_j||;
//   dv = g/||v|| * (dW - v * t[j]/||v||^2)
__global__ void posconv_wn_bwd_sums_kernel(const float* dwp, const float* v, int D, int cg, int k, int rows_per_block,
                                           float* t) {
  const int j = threadIdx.x;
  if (j >= k) return;
  const int rows = D * cg;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int o = r / cg, i = r - o * cg;
    const int grp = o / cg, oo = o - grp * cg;
    const float dW = dwp[((long long)grp * k * 64 + (long long)j * 64 + i) * 64 + oo];
    acc = fmaf(dW, v[(long long)r * k + j], acc);
  }
  atomicAdd(t + j, acc);
}
__global__ void __launch_bounds__(256) posconv_wn_bwd_kernel(const float* dwp, const float* g, const float* v,
                                                             const float* norm2, const float* t, int D, int cg, int k,
                                                             float* dv, float* dg) {
  const long long n = (long long)D * cg * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % k);
    const long long r = idx / k;
    const int o = (int)(r / cg), i = (int)(r - (long long)o * cg);
    const int grp = o / cg, oo = o - grp * cg;
    const float dW = dwp[((long long)grp * k * 64 + (long long)j * 64 + i) * 64 + oo];
    const float inv = rsqrtf(norm2[j]);
    dv[idx] = g[j] * inv * (dW - v[idx] * t[j] * inv * inv);
    if (r == 0) dg[j] = t[j] * inv;
  }
}

int egrid(long long n) {
  long long g = n / 1024 + 1;
  return (int)(g > 148 * 8 ? 148 * 8 : g);
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_cast_multi(const void* table, int32_t n_entries, void* stream_v) {
  A8_REQUIRE(n_entries > 0, "cast_multi: empty table");
  dim3 grid(64, n_entries);
  cast_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(reinterpret_cast<const CastEntry*>(table));
  return check_launch("cast_multi_kernel");
}

extern "C" int a8_conv_pack(const float* w, int32_t Cout, int32_t Cin, int32_t k, int32_t s, void* wk, void* wt0,
                            void* wt1, void* stream_v) {
  A8_REQUIRE(Cout > 0 && Cin > 0 && k > 0 && s >= 1 && s <= 2, "conv_pack: unsupported shape (stride must be 1 or 2)");
  conv_pack_kernel<<<egrid((long long)Cout * Cin * k), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      w, Cout, Cin, k, s, (__nv_bfloat16*)wk, (__nv_bfloat16*)wt0, (__nv_bfloat16*)wt1);
  return check_launch("conv_pack_kernel");
}

extern "C" int a8_conv_unpack(const float* dwk, int32_t Cout, int32_t Cin, int32_t k, float* dw, void* stream_v) {
  conv_unpack_kernel<<<egrid((long long)Cout * Cin * k), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(dwk, Cout, Cin,
                                                                                                     k, dw);
  return check_launch("conv_unpack_kernel");
}

extern "C" int a8_posconv_pack(const float* g, const float* v, int32_t D, int32_t cg, int32_t k, float* norm2,
                               void* wp, void* wpt, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(k <= 1024 && cg <= 64 && D % cg == 0, "posconv_pack: unsupported shape D=%d cg=%d k=%d", D, cg, k);
  // norm2: k results followed by scratch of a8_posconv_norm_scratch_floats(D, cg, k) floats (partials + counter)
  const int rows = D * cg, rpb = 128;
  const int nblk = cdiv(rows, rpb);
  float* scratch = norm2 + k;
  A8_CUDA(cudaMemsetAsync(scratch + (long long)nblk * k, 0, sizeof(float), st));
  posconv_norm_kernel<<<nblk, ((k + 31) / 32) * 32, 0, st>>>(v, rows, k, rpb, norm2, scratch);
  int rc = check_launch("posconv_norm_kernel");
  if (rc) return rc;
  posconv_pack_kernel<<<egrid((long long)D * k * 64), 256, 0, st>>>(g, v, norm2, D, cg, k, (__nv_bfloat16*)wp,
                                                                    (__nv_bfloat16*)wpt);
  return check_launch("posconv_pack_kernel");
}

extern "C" int64_t a8_posconv_norm_scratch_floats(int32_t D, int32_t cg, int32_t k) {
  return (int64_t)cdiv((long long)D * cg, 128) * k + 1;
}

extern "C" int a8_posconv_wn_bwd(const float* dwp, const float* g, const float* v, const float* norm2, int32_t D,
                                 int32_t cg, int32_t k, float* t_scratch, float* dv, float* dg, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(k <= 1024 && cg <= 64, "posconv_wn_bwd: unsupported shape");
  A8_CUDA(cudaMemsetAsync(t_scratch, 0, sizeof(float) * k, st));
  const int rows = D * cg, rpb = 128;
  posconv_wn_bwd_sums_kernel<<<cdiv(rows, rpb), ((k + 31) / 32) * 32, 0, st>>>(dwp, v, D, cg, k, rpb, t_scratch);
  int rc = check_launch("posconv_wn_bwd_sums_kernel");
  if (rc) return rc;
  posconv_wn_bwd_kernel<<<egrid((long long)D * cg * k), 256, 0, st>>>(dwp, g, v, norm2, t_scratch, D, cg, k, dv, dg);
  return check_launch("posconv_wn_bwd_kernel");
}
