// audio8_b200 — contrastive loss of wav2vec2 pre-training, fused (reference: wav2vec2.py:377-392).
//
// The reference gathers K negatives per masked step into a [K,B,Tm,C] tensor (213 MB at the base config),
// concatenates, runs cosine_similarity as several kernels and then cross_entropy.  Here one warp owns one
// masked step: it keeps x_i in registers, walks its K+1 candidate rows of y (L2 resident, 1 KB coalesced
// reads addressed by the host-generated indices), reduces dot products with shuffles and finishes the
// log-softmax / cross-entropy in registers.  Nothing but the [R,K+1] probabilities is written.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr int CPL = 16;           // channels per lane: C <= 512
constexpr float COS_EPS = 1e-8f;  // torch.cosine_similarity eps

__global__ void __launch_bounds__(256) row_norm_kernel(const float* x, const float* y, int R, int C, float* xn,
                                                       float* yn) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < 2 * R; i += warps) {
    const float* p = (i < R ? x + (long long)i * C : y + (long long)(i - R) * C);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += p[c] * p[c];
    s = warp_sum(s);
    if (lane == 0) {
      const float n = fmaxf(sqrtf(s), COS_EPS);
      if (i < R) xn[i] = n;
      else yn[i - R] = n;
    }
  }
}

struct ConArgs;
__device__ __forceinline__ int rows_valid(const ConArgs& a);

struct ConArgs {
  const float* x;   // [R,C]
  const float* y;   // [R,C]
  const int* idx;   // [R,K] rows of y (negatives)
  const float* xn;  // [R] clamped norms
  const float* yn;
  int R, C, K;
  const int* n_valid;  // device scalar: only rows [0, *n_valid) exist (the rest of the R rows is padding); null = all R
  float* cosv;      // [R,K+1]
  float* prob;      // [R,K+1] softmax over candidates
  float* row_loss;  // [R]
  // backward
  const float* dce;  // scalar: d loss / d CE
  float* dx;         // [R,C]
  float* dy;         // [R,C] zeroed, atomics
};

__device__ __forceinline__ int rows_valid(const ConArgs& a) { return a.n_valid ? min(__ldg(a.n_valid), a.R) : a.R; }
// backward: the padding rows of dx (dy is zero-filled by the host side) get zeros
__device__ __forceinline__ void zero_padding_rows(const ConArgs& a, int Rv) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = Rv + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < a.R; i += warps)
    for (int c = lane; c < a.C; c += 32) a.dx[(long long)i * a.C + c] = 0.f;
}

__global__ void __launch_bounds__(256) contrastive_fwd_kernel(const ConArgs a) {
  const int Rv = rows_valid(a);
  extern __shared__ float s_cos[];  // [8 warps][K+1]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float* cs = s_cos + w * (a.K + 1);
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < Rv; i += warps) {
    float xv[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      xv[q] = c < a.C ? a.x[(long long)i * a.C + c] : 0.f;
    }
    const float inx = 1.f / a.xn[i];
    for (int j = 0; j <= a.K; ++j) {
      const int cand = (j == 0) ? i : a.idx[(long long)i * a.K + j - 1];
      const float* yr = a.y + (long long)cand * a.C;
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < a.C) dot = fmaf(xv[q], yr[c], dot);
      }
      dot = warp_sum(dot);
      if (lane == 0) cs[j] = dot * inx / a.yn[cand];
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j <= a.K; j += 32) mx = fmaxf(mx, cs[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j <= a.K; j += 32) sum += expf(cs[j] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    for (int j = lane; j <= a.K; j += 32) {
      a.cosv[(long long)i * (a.K + 1) + j] = cs[j];
      a.prob[(long long)i * (a.K + 1) + j] = expf(cs[j] - lse);
    }
    if (lane == 0) a.row_loss[i] = lse - cs[0];
    __syncwarp();
  }
}

// ce = mean_i row_loss;  loss = xe_w * ce + div_w * (n_vars - ppl) / n_vars   (single CTA, deterministic)
__global__ void __launch_bounds__(1024) contrastive_finalize_kernel(const float* row_loss, int R, const int* n_valid,
                                                                    const float* ppl, float n_vars, float xe_w,
                                                                    float div_w, float* ce, float* loss) {
  __shared__ float red[32];
  if (n_valid != nullptr) R = min(R, *n_valid);
  float acc = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) acc += row_loss[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float c = acc / (float)R;
    *ce = c;
    *loss = xe_w * c + (ppl ? div_w * (n_vars - *ppl) / n_vars : 0.f);
  }
}

__global__ void __launch_bounds__(256) contrastive_bwd_kernel(const ConArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int Rv = rows_valid(a);
  zero_padding_rows(a, Rv);
  const float scale = (*a.dce) / (float)Rv;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < Rv; i += warps) {
    float xh[CPL], dxv[CPL];
    const float inx = 1.f / a.xn[i];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      xh[q] = c < a.C ? a.x[(long long)i * a.C + c] * inx : 0.f;
      dxv[q] = 0.f;
    }
    for (int j = 0; j <= a.K; ++j) {
      const int cand = (j == 0) ? i : a.idx[(long long)i * a.K + j - 1];
      const float cosj = a.cosv[(long long)i * (a.K + 1) + j];
      const float dcos = (a.prob[(long long)i * (a.K + 1) + j] - (j == 0 ? 1.f : 0.f)) * scale;
      const float iny = 1.f / a.yn[cand];
      const float* yr = a.y + (long long)cand * a.C;
      float* dyr = a.dy + (long long)cand * a.C;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < a.C) {
          const float yh = yr[c] * iny;
          dxv[q] = fmaf(dcos, yh - cosj * xh[q], dxv[q]);
          atomicAdd(dyr + c, dcos * (xh[q] - cosj * yh) * iny);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const int c = lane + 32 * q;
      if (c < a.C) a.dx[(long long)i * a.C + c] = dxv[q] * inx;
    }
  }
}

int con_grid(int R) {
  int g = cdiv(R, 8);
  return g < 1 ? 1 : (g > 148 * 4 ? 148 * 4 : g);
}

// ---- fast path (C a multiple of 32 with C/32 in {2,4,8,16,24}): a warp still owns one masked step, but works on FOUR
// candidates at a time: 8 lanes per candidate, each lane NQ 16-byte pieces of the candidate row (128 contiguous bytes
// per 8 lanes and piece), 3 shuffles per dot product.  Candidate indices / saved cosines are staged in the warp's slice
// of shared memory first, so the row loads of consecutive iterations are independent of each other (the first version
// chased index -> row -> reduce serially, ~800 clk per candidate).
template <int NQ>
__global__ void __launch_bounds__(256) contrastive_fwd_fast_kernel(const ConArgs a) {
  const int Rv = rows_valid(a);
  extern __shared__ float s_dyn[];  // per warp: [K+1] cos values, [K+1] candidate ids
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int K1 = a.K + 1, K1p = (K1 + 3) & ~3;
  float* cs = s_dyn + w * 2 * K1p;
  int* cid = reinterpret_cast<int*>(cs + K1p);
  const int sub = lane & 7, grp = lane >> 3;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < Rv; i += warps) {
    for (int j = lane; j < K1p; j += 32) cid[j] = (j == 0) ? i : (j < K1 ? __ldg(a.idx + (long long)i * a.K + j - 1) : i);
    float4 xv[NQ];
    const float4* xr = reinterpret_cast<const float4*>(a.x + (long long)i * a.C);
#pragma unroll
    for (int q = 0; q < NQ; ++q) xv[q] = __ldg(xr + sub + 8 * q);
    const float inx = 1.f / a.xn[i];
    __syncwarp();
#pragma unroll 2
    for (int j0 = 0; j0 < K1p; j0 += 4) {
      const int cand = cid[j0 + grp];
      const float4* yr = reinterpret_cast<const float4*>(a.y + (long long)cand * a.C);
      const float ynv = __ldg(a.yn + cand);
      float dot = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 yv = __ldg(yr + sub + 8 * q);
        dot = fmaf(xv[q].x, yv.x, dot); dot = fmaf(xv[q].y, yv.y, dot);
        dot = fmaf(xv[q].z, yv.z, dot); dot = fmaf(xv[q].w, yv.w, dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (sub == 0 && j0 + grp < K1) cs[j0 + grp] = dot * inx / ynv;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < K1; j += 32) mx = fmaxf(mx, cs[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < K1; j += 32) sum += expf(cs[j] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    for (int j = lane; j < K1; j += 32) {
      a.cosv[(long long)i * K1 + j] = cs[j];
      a.prob[(long long)i * K1 + j] = expf(cs[j] - lse);
    }
    if (lane == 0) a.row_loss[i] = lse - cs[0];
    __syncwarp();
  }
}

template <int NQ>
__global__ void __launch_bounds__(256) contrastive_bwd_fast_kernel(const ConArgs a) {
  extern __shared__ float s_dyn[];  // per warp: [K+1] cos, [K+1] dcos, [K+1] candidate ids
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int K1 = a.K + 1, K1p = (K1 + 3) & ~3;
  float* cs = s_dyn + w * 3 * K1p;
  float* dc = cs + K1p;
  int* cid = reinterpret_cast<int*>(dc + K1p);
  const int sub = lane & 7, grp = lane >> 3;
  const int Rv = rows_valid(a);
  zero_padding_rows(a, Rv);
  const float scale = (*a.dce) / (float)Rv;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < Rv; i += warps) {
    for (int j = lane; j < K1p; j += 32) {
      const bool ok = j < K1;
      cid[j] = (j == 0 || !ok) ? i : __ldg(a.idx + (long long)i * a.K + j - 1);
      cs[j] = ok ? a.cosv[(long long)i * K1 + j] : 0.f;
      dc[j] = ok ? (a.prob[(long long)i * K1 + j] - (j == 0 ? 1.f : 0.f)) * scale : 0.f;  // padding contributes nothing
    }
    const float inx = 1.f / a.xn[i];
    float4 xh[NQ], dxv[NQ];
    const float4* xr = reinterpret_cast<const float4*>(a.x + (long long)i * a.C);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4 t = __ldg(xr + sub + 8 * q);
      xh[q] = make_float4(t.x * inx, t.y * inx, t.z * inx, t.w * inx);
      dxv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
#pragma unroll 2
    for (int j0 = 0; j0 < K1p; j0 += 4) {
      const int j = j0 + grp;
      const int cand = cid[j];
      const float cosj = cs[j], dcos = dc[j];
      const float iny = 1.f / __ldg(a.yn + cand);
      const float4* yr = reinterpret_cast<const float4*>(a.y + (long long)cand * a.C);
      float* dyr = a.dy + (long long)cand * a.C;
      const float k1 = dcos * iny, k2 = dcos * cosj * iny;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 t = __ldg(yr + sub + 8 * q);
        const float4 yh = make_float4(t.x * iny, t.y * iny, t.z * iny, t.w * iny);
        dxv[q].x = fmaf(dcos, yh.x - cosj * xh[q].x, dxv[q].x);
        dxv[q].y = fmaf(dcos, yh.y - cosj * xh[q].y, dxv[q].y);
        dxv[q].z = fmaf(dcos, yh.z - cosj * xh[q].z, dxv[q].z);
        dxv[q].w = fmaf(dcos, yh.w - cosj * xh[q].w, dxv[q].w);
        if (dcos != 0.f) {  // dy[cand] += dcos (xh - cos yh) / |y|: one 16-byte vector reduction per piece
          const float4 g = make_float4(k1 * xh[q].x - k2 * yh.x, k1 * xh[q].y - k2 * yh.y, k1 * xh[q].z - k2 * yh.z,
                                       k1 * xh[q].w - k2 * yh.w);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dyr + 4 * (sub + 8 * q)), "f"(g.x), "f"(g.y),
                       "f"(g.z), "f"(g.w)
                       : "memory");
        }
      }
    }
    // the four candidate groups hold partial dx for the same channels: fold them, group 0 stores
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float4 t = dxv[q];
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
      }
      if (grp == 0)
        reinterpret_cast<float4*>(a.dx + (long long)i * a.C)[sub + 8 * q] = make_float4(t.x * inx, t.y * inx, t.z * inx, t.w * inx);
    }
    __syncwarp();
  }
}

template <bool BWD>
int launch_contrastive_fast(const ConArgs& a, cudaStream_t st) {
  const int nq = a.C / 32;
  const int K1p = (a.K + 4) & ~3;
  const size_t smem = (size_t)8 * (BWD ? 3 : 2) * K1p * sizeof(float);
  const int grid = con_grid(a.R);
#define A8_CON(N)                                                                  \
  case N:                                                                          \
    if (BWD) contrastive_bwd_fast_kernel<N><<<grid, 256, smem, st>>>(a);           \
    else contrastive_fwd_fast_kernel<N><<<grid, 256, smem, st>>>(a);               \
    return 1;
  if (a.C % 32 != 0 || smem > 48 * 1024 || (reinterpret_cast<uintptr_t>(a.x) & 15u) || (reinterpret_cast<uintptr_t>(a.y) & 15u) ||
      (BWD && ((reinterpret_cast<uintptr_t>(a.dx) & 15u) || (reinterpret_cast<uintptr_t>(a.dy) & 15u))))
    return 0;
  switch (nq) {
    A8_CON(2) A8_CON(4) A8_CON(8) A8_CON(16) A8_CON(24)
  }
#undef A8_CON
  return 0;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_contrastive_fwd(const float* x, const float* y, const int32_t* idx, int32_t R, int32_t C, int32_t K,
                                  const int32_t* n_valid, const float* ppl, float n_vars, float xe_w, float div_w, float* xn, float* yn,
                                  float* cosv, float* prob, float* row_loss, float* ce, float* loss, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && C > 0 && (C <= 32 * CPL || C == 768) && K >= 0 && K <= 4095,
             "contrastive: unsupported shape R=%d C=%d K=%d", R, C, K);
  row_norm_kernel<<<con_grid(2 * R), 256, 0, st>>>(x, y, R, C, xn, yn);
  int rc = check_launch("row_norm_kernel");
  if (rc) return rc;
  ConArgs a{x, y, idx, xn, yn, R, C, K, n_valid, cosv, prob, row_loss, nullptr, nullptr, nullptr};
  if (!launch_contrastive_fast<false>(a, st)) {
    // the generic kernel holds 32 * CPL channels per row: wider rows exist only on the fast path
    A8_REQUIRE(C <= 32 * CPL, "contrastive: C=%d needs the vectorised path (16-byte aligned x/y, K <= ~760)", C);
    contrastive_fwd_kernel<<<con_grid(R), 256, 8 * (K + 1) * sizeof(float), st>>>(a);
  }
  rc = check_launch("contrastive_fwd_kernel");
  if (rc) return rc;
  contrastive_finalize_kernel<<<1, 1024, 0, st>>>(row_loss, R, n_valid, ppl, n_vars, xe_w, div_w, ce, loss);
  return check_launch("contrastive_finalize_kernel");
}

extern "C" int a8_contrastive_bwd(const float* x, const float* y, const int32_t* idx, int32_t R, int32_t C, int32_t K,
                                  const int32_t* n_valid, const float* xn, const float* yn, const float* cosv, const float* prob,
                                  const float* dce, float* dx, float* dy, void* stream_v) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(R > 0 && C > 0 && (C <= 32 * CPL || C == 768) && K >= 0, "contrastive_bwd: unsupported shape");
  A8_CUDA(cudaMemsetAsync(dy, 0, sizeof(float) * (size_t)R * C, st));
  ConArgs a{x, y, idx, xn, yn, R, C, K, n_valid, const_cast<float*>(cosv), const_cast<float*>(prob), nullptr, dce, dx, dy};
  if (!launch_contrastive_fast<true>(a, st)) {
    A8_REQUIRE(C <= 32 * CPL, "contrastive_bwd: C=%d needs the vectorised path (16-byte aligned tensors, K <= ~500)", C);
    contrastive_bwd_kernel<<<con_grid(R), 256, 0, st>>>(a);
  }
  return check_launch("contrastive_bwd_kernel");
}
