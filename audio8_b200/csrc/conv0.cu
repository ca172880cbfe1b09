// audio8_b200 — feature-encoder layer 0, fully fused: Conv1d(1->C, k, stride) + GroupNorm(C,C) + GELU.
//
// Replaces cuDNN conv + ATen group_norm (RowwiseMoments + elementwise) + ATen GELU, three round trips over a
// [B,C,L0] fp32 tensor in the reference (wav2vec2.py:419-422), by kernels that only ever read the waveform
// (fp32 [B,L], L2 resident) and write/read the bf16 channels-last activation once:
//   * statistics: because z[t,c] = sum_j w[c,j] x[s*t+j] is linear in x, the per-(b,c) mean and variance over
//     time follow from the k first moments and k(k+1)/2 second moments of the strided windows of x:
//     mean_c = w_c . m,  E[z^2]_c = w_c^T R w_c.  One pass over x in fp64, no pass over z at all.
//   * forward: recompute z from x, normalise, affine, GELU, store bf16 [B,L0,C]  (HBM: one 2-byte write/elem)
//   * backward: two passes over the incoming gradient (sums for the norm backward, then dz -> dW), z recomputed.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr int KMAX = 10;
constexpr int NMOM = KMAX + KMAX * (KMAX + 1) / 2;  // 65
constexpr int TILE_T = 64;

// ---------------------------------------------------------------------------------------------- moments
// mom[b][0..k) = sum_t x[s t + j];  mom[b][k + idx(j,j')] = sum_t x[s t + j] x[s t + j'] (j <= j')
__global__ void __launch_bounds__(256) conv0_moments_kernel(const float* x, long long L, int L0, int k, int s,
                                                            double* mom) {
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * L;
  float acc[NMOM];
#pragma unroll
  for (int i = 0; i < NMOM; ++i) acc[i] = 0.f;
  const int t0 = blockIdx.x * (256 * 16);
  for (int it = 0; it < 16; ++it) {
    const int t = t0 + it * 256 + threadIdx.x;
    if (t < L0) {
      float w[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) w[j] = (j < k) ? xb[(long long)t * s + j] : 0.f;
      int idx = KMAX;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        acc[j] += w[j];
#pragma unroll
        for (int jj = j; jj < KMAX; ++jj) acc[idx++] += w[j] * w[jj];
      }
    }
  }
  __shared__ double red[8][NMOM];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NMOM; ++i) {
    double v = (double)acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NMOM) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    atomicAdd(mom + (long long)b * NMOM + threadIdx.x, v);
  }
}

// mean[b,c], rstd[b,c] from the moments
__global__ void conv0_stats_kernel(const double* mom, const float* w, int C, int k, int L0, float eps, float* mean,
                                   float* rstd) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* m = mom + (long long)b * NMOM;
  double wj[KMAX];
  for (int j = 0; j < KMAX; ++j) wj[j] = j < k ? (double)w[c * k + j] : 0.0;
  double mu = 0.0, e2 = 0.0;
  int idx = KMAX;
  for (int j = 0; j < KMAX; ++j) {
    mu += wj[j] * m[j];
    for (int jj = j; jj < KMAX; ++jj) e2 += (jj == j ? 1.0 : 2.0) * wj[j] * wj[jj] * m[idx++];
  }
  mu /= L0;
  e2 /= L0;
  const double var = fmax(e2 - mu * mu, 0.0);
  mean[b * C + c] = (float)mu;
  rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---------------------------------------------------------------------------------------------- main kernels
// thread mapping shared by forward and both backward passes: 256 threads = 64 channel-octets x 4 row phases;
// a CTA walks its [t_begin, t_end) range in tiles of TILE_T rows with the waveform window staged in smem
struct Conv0Args {
  const float* x;
  long long L;
  int L0, k, s, C, rows_per_cta;
  const float* w;      // [C,k]
  const float* gamma;  // [C]
  const float* beta;
  const float* mean;   // [B,C]
  const float* rstd;
  __nv_bfloat16* y;          // fwd out [B,L0,C]
  const __nv_bfloat16* da;   // bwd in  [B,L0,C]
  float* sums;               // [B,C,2]: sum dy, sum dy*xhat
  float* dw;                 // [C,k]
  float* dgamma;
  float* dbeta;
};

enum { MODE_FWD = 0, MODE_BWD_SUMS = 1, MODE_BWD_W = 2 };

template <int MODE>
__global__ void __launch_bounds__(256) conv0_kernel(const Conv0Args a) {
  __shared__ float xs[TILE_T * 8 + KMAX + 8];
  __shared__ float red[(MODE == MODE_BWD_W) ? 4 * 64 * 8 * 3 : (MODE == MODE_BWD_SUMS ? 4 * 64 * 16 : 1)];
  const int b = blockIdx.y;
  const int oct = threadIdx.x & 63, ph = threadIdx.x >> 6;
  const int c0 = oct * 8;
  const bool active = c0 < a.C;
  const int k = a.k, s = a.s;
  float w[8][KMAX], sc[8], sh[8], mu[8], rs[8], gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = min(c0 + i, a.C - 1);
#pragma unroll
    for (int j = 0; j < KMAX; ++j) w[i][j] = (j < k) ? a.w[c * k + j] : 0.f;
    mu[i] = a.mean[b * a.C + c];
    rs[i] = a.rstd[b * a.C + c];
    gm[i] = a.gamma[c];
    bt[i] = a.beta[c];
    sc[i] = rs[i] * gm[i];
    sh[i] = bt[i] - mu[i] * sc[i];
  }
  float s1m[8], s2m[8];  // BWD_W: mean(dy), mean(dy*xhat) per channel
  if (MODE == MODE_BWD_W) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = min(c0 + i, a.C - 1);
      s1m[i] = a.sums[(b * a.C + c) * 2] / (float)a.L0;
      s2m[i] = a.sums[(b * a.C + c) * 2 + 1] / (float)a.L0;
    }
  }
  float acc1[8], acc2[8], accw[8][KMAX];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc1[i] = acc2[i] = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) accw[i][j] = 0.f;
  }

  const float* xb = a.x + (long long)b * a.L;
  const int t_begin = blockIdx.x * a.rows_per_cta;
  const int t_end = min(a.L0, t_begin + a.rows_per_cta);
  for (int tt = t_begin; tt < t_end; tt += TILE_T) {
    const int nrows = min(TILE_T, t_end - tt);
    const int nx = (nrows - 1) * s + k;
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += 256) xs[i] = xb[(long long)tt * s + i];
    __syncthreads();
    if (!active) continue;
    for (int r = ph; r < nrows; r += 4) {
      float xv[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) xv[j] = (j < k) ? xs[r * s + j] : 0.f;
      float z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) acc = fmaf(w[i][j], xv[j], acc);
        z[i] = acc;
      }
      const long long off = ((long long)b * a.L0 + tt + r) * a.C + c0;
      if (MODE == MODE_FWD) {
        uint4 o;
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = gelu_fast(fmaf(z[i], sc[i], sh[i]));
        o.x = pack_bf16(g[0], g[1]); o.y = pack_bf16(g[2], g[3]);
        o.z = pack_bf16(g[4], g[5]); o.w = pack_bf16(g[6], g[7]);
        *reinterpret_cast<uint4*>(a.y + off) = o;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(a.da + off);
        float d[8];
        float2 t2;
        t2 = unpack_bf16(u.x); d[0] = t2.x; d[1] = t2.y;
        t2 = unpack_bf16(u.y); d[2] = t2.x; d[3] = t2.y;
        t2 = unpack_bf16(u.z); d[4] = t2.x; d[5] = t2.y;
        t2 = unpack_bf16(u.w); d[6] = t2.x; d[7] = t2.y;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (z[i] - mu[i]) * rs[i];
          const float dy = d[i] * gelu_grad_fast(fmaf(xh, gm[i], bt[i]));
          if (MODE == MODE_BWD_SUMS) {
            acc1[i] += dy;
            acc2[i] += dy * xh;
          } else {
            const float dz = sc[i] * (dy - s1m[i] - xh * s2m[i]);
#pragma unroll
            for (int j = 0; j < KMAX; ++j) accw[i][j] = fmaf(dz, xv[j], accw[i][j]);
          }
        }
      }
    }
  }
  if (MODE == MODE_BWD_SUMS) {
    // reduce the 4 row phases in smem, then one atomic per (channel, stat) per CTA
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[(ph * 64 + oct) * 16 + i] = acc1[i];
      red[(ph * 64 + oct) * 16 + 8 + i] = acc2[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int o = e >> 4, q = e & 15;
      const int c = o * 8 + (q & 7);
      if (c < a.C) {
        const float v = red[e] + red[64 * 16 + e] + red[2 * 64 * 16 + e] + red[3 * 64 * 16 + e];
        atomicAdd(a.sums + ((long long)b * a.C + c) * 2 + (q >> 3), v);
      }
    }
  } else if (MODE == MODE_BWD_W) {
    // taps in 4 passes of 3 (smem budget): reduce phases, atomics into dw[C,k]
#pragma unroll
    for (int j0 = 0; j0 < 12; j0 += 3) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj)
          red[((ph * 64 + oct) * 8 + i) * 3 + jj] = (j0 + jj < KMAX) ? accw[i][(j0 + jj < KMAX) ? j0 + jj : 0] : 0.f;
      __syncthreads();
      for (int e = threadIdx.x; e < 64 * 8 * 3; e += 256) {
        const int c = e / 3, jj = e % 3;
        if (c < a.C && j0 + jj < k) {
          const float v = red[e] + red[64 * 24 + e] + red[2 * 64 * 24 + e] + red[3 * 64 * 24 + e];
          atomicAdd(a.dw + c * k + j0 + jj, v);
        }
      }
    }
    if (blockIdx.x == 0) {  // dgamma / dbeta from the pass-1 sums (one CTA per batch item adds its share)
      for (int c = threadIdx.x; c < a.C; c += 256) {
        atomicAdd(a.dbeta + c, a.sums[((long long)b * a.C + c) * 2]);
        atomicAdd(a.dgamma + c, a.sums[((long long)b * a.C + c) * 2 + 1]);
      }
    }
  }
}

int rows_per_cta(int L0, int B) {
  int per_batch = (148 * 2) / (B > 0 ? B : 1);
  if (per_batch < 1) per_batch = 1;
  int rows = cdiv(L0, per_batch);
  rows = cdiv(rows, TILE_T) * TILE_T;
  return rows;
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_conv0_stats(const float* x, int32_t B, int64_t L, const float* w, int32_t C, int32_t k, int32_t s,
                              float eps, double* moments, float* mean, float* rstd, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(k >= 1 && k <= KMAX && s >= 1 && s <= 8, "conv0: kernel %d / stride %d unsupported", k, s);
  A8_REQUIRE(L >= k, "conv0: input shorter than the kernel");
  const int L0 = (int)((L - k) / s + 1);
  A8_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * NMOM * B, stream));
  dim3 g1(cdiv(L0, 256 * 16), B);
  conv0_moments_kernel<<<g1, 256, 0, stream>>>(x, L, L0, k, s, moments);
  int rc = check_launch("conv0_moments_kernel");
  if (rc) return rc;
  dim3 g2(cdiv(C, 128), B);
  conv0_stats_kernel<<<g2, 128, 0, stream>>>(moments, w, C, k, L0, eps, mean, rstd);
  return check_launch("conv0_stats_kernel");
}

static int conv0_check(int32_t C, int32_t k, int32_t s) {
  A8_REQUIRE(C % 8 == 0 && C <= 512, "conv0: C=%d must be a multiple of 8, <= 512", C);
  A8_REQUIRE(k >= 1 && k <= KMAX && s >= 1 && s <= 8, "conv0: kernel %d / stride %d unsupported", k, s);
  return 0;
}

extern "C" int a8_conv0_fwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma,
                            const float* beta, const float* mean, const float* rstd, int32_t C, int32_t k,
                            int32_t s, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (conv0_check(C, k, s)) return -1;
  const int L0 = (int)((L - k) / s + 1);
  Conv0Args a{x, L, L0, k, s, C, rows_per_cta(L0, B), w, gamma, beta, mean, rstd, (__nv_bfloat16*)y,
              nullptr, nullptr, nullptr, nullptr, nullptr};
  dim3 grid(cdiv(L0, a.rows_per_cta), B);
  conv0_kernel<MODE_FWD><<<grid, 256, 0, stream>>>(a);
  return check_launch("conv0_kernel<fwd>");
}

extern "C" int a8_conv0_bwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma,
                            const float* beta, const float* mean, const float* rstd, int32_t C, int32_t k,
                            int32_t s, const void* da, float* sums, float* dw, float* dgamma, float* dbeta,
                            void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (conv0_check(C, k, s)) return -1;
  const int L0 = (int)((L - k) / s + 1);
  Conv0Args a{x, L, L0, k, s, C, rows_per_cta(L0, B), w, gamma, beta, mean, rstd, nullptr,
              (const __nv_bfloat16*)da, sums, dw, dgamma, dbeta};
  dim3 grid(cdiv(L0, a.rows_per_cta), B);
  A8_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * B * C, stream));
  conv0_kernel<MODE_BWD_SUMS><<<grid, 256, 0, stream>>>(a);
  int rc = check_launch("conv0_kernel<bwd_sums>");
  if (rc) return rc;
  conv0_kernel<MODE_BWD_W><<<grid, 256, 0, stream>>>(a);
  return check_launch("conv0_kernel<bwd_w>");
}
