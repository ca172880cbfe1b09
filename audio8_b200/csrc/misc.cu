// audio8_b200 — index / mask / cast kernels around the time-masking of wav2vec2.
//
// Replaces the boolean-mask index_put / gathers of the reference (wav2vec2.py:939 `features[time_mask] =
// mask_emb`, :946 `unmasked_features[time_mask]`, :381 `outputs[time_mask]`, :632 `x[~pad_mask] = 0`,
// :717,721 fine-tuning masks), each of which costs a `nonzero` + device->host sync under eager PyTorch.
// Here the host uploads the (numpy-bit-exact) row indices once per step and everything is index driven.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// gather: dst[i,:] = src[idx[i],:]   scatter: dst[idx[i],:] = src[i,:]   (one warp per row).  A negative index marks a
// padding entry (index lists are padded to their worst-case length so that shapes are static): gather writes a zero
// row, scatter skips it.
template <typename TS, typename TD, bool SCATTER>
__global__ void __launch_bounds__(256) rows_copy_kernel(const TS* src, TD* dst, const int* idx, int n, int C) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const long long r = idx[i];
    if (r < 0) {
      if (!SCATTER)
        for (int c = lane; c < C; c += 32) dst[(long long)i * C + c] = from_f<TD>(0.f);
      continue;
    }
    const TS* s = src + (SCATTER ? (long long)i : r) * C;
    TD* d = dst + (SCATTER ? r : (long long)i) * C;
    for (int c = lane; c < C; c += 32) d[c] = from_f<TD>(to_f<TS>(s[c]));
  }
}

// x[idx[i], :] = vec  (bf16 rows, fp32 vector)
__global__ void __launch_bounds__(256) rows_set_kernel(__nv_bfloat16* x, const int* idx, int n, int C, const float* vec) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    if (idx[i] < 0) continue;  // padding entry
    __nv_bfloat16* d = x + (long long)idx[i] * C;
    for (int c = lane; c < C; c += 32) d[c] = __float2bfloat16(vec[c]);
  }
}

// backward of rows_set: dvec += sum_i dx[idx[i], :], then dx[idx[i], :] = 0.  A CTA owns RSB rows: their indices are
// staged in shared memory and a thread's RSB loads of one channel pair are independent (the first version chased
// index -> row -> add serially, 48 dependent round trips per thread).
constexpr int RSB = 16;
__global__ void __launch_bounds__(256) rows_set_bwd_kernel(__nv_bfloat16* dx, const int* idx, int n, int C,
                                                           int rows_per_block, float* dvec) {
  __shared__ int s_idx[RSB];
  const int i0 = blockIdx.x * RSB;
  if (threadIdx.x < RSB) s_idx[threadIdx.x] = (i0 + threadIdx.x < n) ? idx[i0 + threadIdx.x] : -1;
  __syncthreads();
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {  // C is even (bf16x2 accesses)
    uint32_t w[RSB];
#pragma unroll
    for (int i = 0; i < RSB; ++i) {
      const int r = s_idx[i];
      w[i] = (r >= 0) ? *reinterpret_cast<const uint32_t*>(dx + (long long)r * C + c) : 0u;
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < RSB; ++i) {
      const float2 t = unpack_bf16(w[i]);
      a0 += t.x;
      a1 += t.y;
      const int r = s_idx[i];
      if (r >= 0) *reinterpret_cast<uint32_t*>(dx + (long long)r * C + c) = 0u;
    }
    atomicAdd(dvec + c, a0);
    atomicAdd(dvec + c + 1, a1);
  }
}

// in place: zero rows with row_keep == 0 and channels with chan_zero != 0 (x bf16 [B,T,C])
__global__ void __launch_bounds__(256) mask_apply_kernel(__nv_bfloat16* x, const unsigned char* row_keep,
                                                         const unsigned char* chan_zero, int B, int T, int C) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int rows = B * T;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    __nv_bfloat16* d = x + (long long)r * C;
    const bool dead = row_keep != nullptr && row_keep[r] == 0;
    const unsigned char* cz = chan_zero ? chan_zero + (long long)(r / T) * C : nullptr;
    if (!dead && cz == nullptr) continue;
    for (int c = lane; c < C; c += 32)
      if (dead || cz[c]) d[c] = __float2bfloat16(0.f);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) cast_kernel(const TS* src, TD* dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<TD>(to_f<TS>(src[i]));
}

// bf16x3 split for near-fp32 tensor-core products: dst[r] = [hi(x) | hi(x) | lo(x)]  (a_side) or
// [hi(w) | lo(w) | hi(w)] (b_side), so that one K=3C bf16 GEMM yields hi*hi + hi*lo + lo*hi.
__global__ void __launch_bounds__(256) split3_kernel(const float* src, __nv_bfloat16* dst, int R, int C, int b_side) {
  const long long n = (long long)R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float v = src[i];
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    __nv_bfloat16* d = dst + r * 3 * C + c;
    d[0] = hi;
    d[C] = b_side ? lo : hi;
    d[2 * C] = b_side ? hi : lo;
  }
}

int grid_rows(int n) {
  int g = cdiv(n, 8);
  return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g);
}
int grid_elems(long long n) {
  long long g = n / 1024 + 1;
  return (int)(g > 148 * 8 ? 148 * 8 : g);
}

}  // namespace
}  // namespace a8

using namespace a8;

// dtype codes: 0 = fp32, 1 = bf16
extern "C" int a8_rows_copy(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, const int32_t* idx,
                            int32_t n, int32_t C, int32_t scatter, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(n > 0 && C > 0, "rows_copy: empty");
  const int g = grid_rows(n);
#define A8_RC(TS, TD, SC) rows_copy_kernel<TS, TD, SC><<<g, 256, 0, st>>>((const TS*)src, (TD*)dst, idx, n, C)
  if (src_dtype == 0 && dst_dtype == 0) { if (scatter) A8_RC(float, float, true); else A8_RC(float, float, false); }
  else if (src_dtype == 0 && dst_dtype == 1) { if (scatter) A8_RC(float, __nv_bfloat16, true); else A8_RC(float, __nv_bfloat16, false); }
  else if (src_dtype == 1 && dst_dtype == 0) { if (scatter) A8_RC(__nv_bfloat16, float, true); else A8_RC(__nv_bfloat16, float, false); }
  else { if (scatter) A8_RC(__nv_bfloat16, __nv_bfloat16, true); else A8_RC(__nv_bfloat16, __nv_bfloat16, false); }
#undef A8_RC
  return check_launch("rows_copy_kernel");
}

extern "C" int a8_rows_set(void* x, const int32_t* idx, int32_t n, int32_t C, const float* vec, void* stream_v) {
  A8_REQUIRE(n > 0 && C > 0, "rows_set: empty");
  rows_set_kernel<<<grid_rows(n), 256, 0, static_cast<cudaStream_t>(stream_v)>>>((__nv_bfloat16*)x, idx, n, C, vec);
  return check_launch("rows_set_kernel");
}

extern "C" int a8_rows_set_bwd(void* dx, const int32_t* idx, int32_t n, int32_t C, float* dvec, void* stream_v) {
  A8_REQUIRE(n > 0 && C > 0 && C % 2 == 0, "rows_set_bwd: empty or odd channel count");
  const int rpb = RSB;
  rows_set_bwd_kernel<<<cdiv(n, rpb), 256, 0, static_cast<cudaStream_t>(stream_v)>>>((__nv_bfloat16*)dx, idx, n, C,
                                                                                    rpb, dvec);
  return check_launch("rows_set_bwd_kernel");
}

extern "C" int a8_mask_apply(void* x, const uint8_t* row_keep, const uint8_t* chan_zero, int32_t B, int32_t T,
                             int32_t C, void* stream_v) {
  A8_REQUIRE(B > 0 && T > 0 && C > 0, "mask_apply: empty");
  mask_apply_kernel<<<grid_rows(B * T), 256, 0, static_cast<cudaStream_t>(stream_v)>>>((__nv_bfloat16*)x, row_keep,
                                                                                      chan_zero, B, T, C);
  return check_launch("mask_apply_kernel");
}

extern "C" int a8_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(n > 0 && src_dtype != dst_dtype, "cast: bad arguments");
  if (src_dtype == 0) cast_kernel<float, __nv_bfloat16><<<grid_elems(n), 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else cast_kernel<__nv_bfloat16, float><<<grid_elems(n), 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  return check_launch("cast_kernel");
}

extern "C" int a8_split3(const float* src, void* dst, int32_t R, int32_t C, int32_t b_side, void* stream_v) {
  A8_REQUIRE(R > 0 && C > 0, "split3: empty");
  split3_kernel<<<grid_elems((long long)R * C), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      src, (__nv_bfloat16*)dst, R, C, b_side);
  return check_launch("split3_kernel");
}
