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
// Thread mapping shared by forward and backward: one thread = 2 adjacent channels (one bf16x2 word), a CTA of C/2
// threads walks its [t_begin, t_end) rows in tiles of TILE_T with the waveform window staged in smem (every thread
// of the CTA reads the same window element: a broadcast).  ~70 registers per thread keep 3 CTAs per SM resident.
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
  float* acc;                // bwd: [B,C,12] = sum_t dy*x_j (j<10), sum_t dy, sum_t dy*xhat
};

template <bool BWD>
__global__ void __launch_bounds__(256) conv0_kernel(const Conv0Args a) {
  __shared__ float xs[TILE_T * 8 + KMAX + 8];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  const bool active = c0 < a.C;
  const int k = a.k, s = a.s;
  float w[2][KMAX], sc[2], sh[2], mu[2], rs[2], gm[2], bt[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
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
  float acc1[2] = {0.f, 0.f}, acc2[2] = {0.f, 0.f}, accx[2][KMAX];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < KMAX; ++j) accx[i][j] = 0.f;

  const float* xb = a.x + (long long)b * a.L;
  const int t_begin = blockIdx.x * a.rows_per_cta;
  const int t_end = min(a.L0, t_begin + a.rows_per_cta);
  for (int tt = t_begin; tt < t_end; tt += TILE_T) {
    const int nrows = min(TILE_T, t_end - tt);
    const int nx = (nrows - 1) * s + k;
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = xb[(long long)tt * s + i];
    __syncthreads();
    if (!active) continue;
    const long long off0 = ((long long)b * a.L0 + tt) * a.C + c0;
#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
      float xv[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) xv[j] = (j < k) ? xs[r * s + j] : 0.f;
      float z[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) t = fmaf(w[i][j], xv[j], t);
        z[i] = t;
      }
      const long long off = off0 + (long long)r * a.C;
      if (!BWD) {
        *reinterpret_cast<uint32_t*>(a.y + off) =
            pack_bf16(gelu_fast(fmaf(z[0], sc[0], sh[0])), gelu_fast(fmaf(z[1], sc[1], sh[1])));
      } else {
        const float2 d = unpack_bf16(__ldg(reinterpret_cast<const uint32_t*>(a.da + off)));
        const float dd[2] = {d.x, d.y};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float xh = (z[i] - mu[i]) * rs[i];
          const float dy = dd[i] * gelu_grad_fast(fmaf(xh, gm[i], bt[i]));
          acc1[i] += dy;
          acc2[i] = fmaf(dy, xh, acc2[i]);
#pragma unroll
          for (int j = 0; j < KMAX; ++j) accx[i][j] = fmaf(dy, xv[j], accx[i][j]);
        }
      }
    }
  }
  if (BWD && active) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (c0 + i < a.C) {
        float* dst = a.acc + ((long long)b * a.C + c0 + i) * 12;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) atomicAdd(dst + j, accx[i][j]);
        atomicAdd(dst + 10, acc1[i]);
        atomicAdd(dst + 11, acc2[i]);
      }
    }
  }
}

// The model's layer 0 (k = 10, stride 5), four rows per step: the 25 window samples of rows r..r+3 are contiguous and
// 16-byte aligned in the staged waveform (row r starts at float 5 r, r a multiple of 4), so they arrive as 7 LDS.128
// instead of 40 scalar shared loads (the scalar version kept the LSU as busy as the FMA pipe: 5 broadcast loads per
// output), and the normalisation is folded into the taps (forward: w * rstd * gamma, start value = the shift;
// backward: w * rstd, start value -mean * rstd, which yields xhat directly).
template <bool BWD>
__global__ void __launch_bounds__(256, BWD ? 3 : 4) conv0_k10s5_kernel(const Conv0Args a) {
  constexpr int K = 10, S = 5;
  __shared__ __align__(16) float xs[TILE_T * S + 16];
  const int b = blockIdx.y;
  const int c0 = threadIdx.x * 2;
  const bool active = c0 < a.C;
  // the thread's two channels ride in the two halves of packed fp32 pairs (FFMA2: one issue slot per tap for both)
  f32x2 W[K], INIT, GM, BT;
  {
    float w[2][K], init[2], gm[2], bt[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = min(c0 + i, a.C - 1);
      const float mu = a.mean[b * a.C + c], rs = a.rstd[b * a.C + c];
      gm[i] = a.gamma[c];
      bt[i] = a.beta[c];
      const float sc = BWD ? rs : rs * gm[i];
      init[i] = BWD ? -mu * rs : bt[i] - mu * sc;
#pragma unroll
      for (int j = 0; j < K; ++j) w[i][j] = a.w[c * K + j] * sc;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) W[j] = pk2(w[0][j], w[1][j]);
    INIT = pk2(init[0], init[1]);
    GM = pk2(gm[0], gm[1]);
    BT = pk2(bt[0], bt[1]);
  }
  f32x2 ACC1 = bc2(0.f), ACC2 = bc2(0.f), ACCX[K];
#pragma unroll
  for (int j = 0; j < K; ++j) ACCX[j] = bc2(0.f);

  const float* xb = a.x + (long long)b * a.L;
  const int t_begin = blockIdx.x * a.rows_per_cta;
  const int t_end = min(a.L0, t_begin + a.rows_per_cta);
  for (int tt = t_begin; tt < t_end; tt += TILE_T) {
    const int nrows = min(TILE_T, t_end - tt);
    const int nx = (nrows - 1) * S + K;
    __syncthreads();
    for (int i = threadIdx.x; i < TILE_T * S + 16; i += blockDim.x) xs[i] = i < nx ? xb[(long long)tt * S + i] : 0.f;
    __syncthreads();
    if (!active) continue;
    const long long off0 = ((long long)b * a.L0 + tt) * a.C + c0;
#pragma unroll 1
    for (int r = 0; r < nrows; r += 4) {
      float xv[28];
      const float4* src = reinterpret_cast<const float4*>(xs + r * S);
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float4 v = src[q];
        xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
      }
      uint32_t dr[4];
      if (BWD) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          dr[rr] = (r + rr < nrows) ? __ldg(reinterpret_cast<const uint32_t*>(a.da + off0 + (long long)(r + rr) * a.C)) : 0u;
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        f32x2 Z = INIT;
#pragma unroll
        for (int j = 0; j < K; ++j) Z = fma2(W[j], bc2(xv[rr * S + j]), Z);
        if (!BWD) {
          if (r + rr < nrows) {
            float y0, y1;
            unpk2(gelu_fast2(Z), y0, y1);
            *reinterpret_cast<uint32_t*>(a.y + off0 + (long long)(r + rr) * a.C) = pack_bf16(y0, y1);
          }
        } else {
          const float2 d = unpack_bf16(dr[rr]);  // zero beyond the tile's last row: contributes nothing
          const f32x2 DY = mul2(pk2(d.x, d.y), gelu_grad_fast2(fma2(Z, GM, BT)));  // Z = xhat here
          ACC1 = add2(ACC1, DY);
          ACC2 = fma2(DY, Z, ACC2);
#pragma unroll
          for (int j = 0; j < K; ++j) ACCX[j] = fma2(DY, bc2(xv[rr * S + j]), ACCX[j]);
        }
      }
    }
  }
  if (BWD && active) {
    float acc1[2], acc2[2], accx[2][K];
    unpk2(ACC1, acc1[0], acc1[1]);
    unpk2(ACC2, acc2[0], acc2[1]);
#pragma unroll
    for (int j = 0; j < K; ++j) unpk2(ACCX[j], accx[0][j], accx[1][j]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (c0 + i < a.C) {
        float* dst = a.acc + ((long long)b * a.C + c0 + i) * 12;
#pragma unroll
        for (int j = 0; j < K; ++j) atomicAdd(dst + j, accx[i][j]);
        atomicAdd(dst + 10, acc1[i]);
        atomicAdd(dst + 11, acc2[i]);
      }
    }
  }
}

// dW, dgamma, dbeta from the single backward pass and the forward's window moments.  With dz = sc (dy - mean(dy) -
// xhat mean(dy xhat)) and xhat linear in the window,  sum_t dz x_j  needs only  sum_t dy x_j  and the moments:
//   sum_t xhat x_j = rstd ( w . R[:,j] - mean m_j )
__global__ void conv0_bwd_finalize_kernel(const float* acc, const double* mom, const float* w, const float* gamma,
                                          const float* mean, const float* rstd, int B, int C, int k, int L0,
                                          float* dw, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dwc[KMAX];
  for (int j = 0; j < KMAX; ++j) dwc[j] = 0.0;
  double dg = 0.0, db = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* ac = acc + ((long long)b * C + c) * 12;
    const double* m = mom + (long long)b * NMOM;
    const double rs = rstd[b * C + c], mu = mean[b * C + c], sc = rs * gamma[c];
    const double s1 = ac[10], s2 = ac[11];
    const double s1m = s1 / L0, s2m = s2 / L0;
    db += s1;
    dg += s2;
    for (int j = 0; j < k; ++j) {
      double wr = 0.0;  // sum_j' w[c][j'] R[j'][j]
      for (int jj = 0; jj < k; ++jj) {
        const int lo = jj < j ? jj : j, hi = jj < j ? j : jj;
        // packed upper triangle over KMAX taps: row lo starts at KMAX + lo*KMAX - lo*(lo-1)/2
        const int idx = KMAX + lo * KMAX - (lo * (lo - 1)) / 2 + (hi - lo);
        wr += (double)w[c * k + jj] * m[idx];
      }
      const double sxh = rs * (wr - mu * m[j]);
      dwc[j] += sc * ((double)ac[j] - s1m * m[j] - s2m * sxh);
    }
  }
  for (int j = 0; j < k; ++j) dw[c * k + j] = (float)dwc[j];
  dgamma[c] = (float)dg;
  dbeta[c] = (float)db;
}

// one wave: the grid is sized to the number of CTAs that are resident at once (registers decide: 3 per SM for the
// forward kernels, 2 for the k10/s5 backward; a grid sized for 3 with only 2 resident ran two rounds, +33 %)
template <typename K>
int resident_ctas(K kern, int threads) {
  int per_sm = 0, dev = 0, sms = 148;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * (sms > 0 ? sms : 148);
}

int rows_per_cta(int L0, int B, int ctas = 148 * 3) {
  int per_batch = ctas / (B > 0 ? B : 1);
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
  A8_REQUIRE(C % 64 == 0 && C <= 512, "conv0: C=%d must be a multiple of 64, <= 512", C);
  A8_REQUIRE(k >= 1 && k <= KMAX && s >= 1 && s <= 8, "conv0: kernel %d / stride %d unsupported", k, s);
  return 0;
}

extern "C" int a8_conv0_fwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma,
                            const float* beta, const float* mean, const float* rstd, int32_t C, int32_t k,
                            int32_t s, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (conv0_check(C, k, s)) return -1;
  const int L0 = (int)((L - k) / s + 1);
  const bool fast = (k == 10 && s == 5);
  static int ctas[2] = {0, 0}, ctas_c[2] = {0, 0};
  if (ctas_c[fast] != C) {
    ctas[fast] = fast ? resident_ctas(conv0_k10s5_kernel<false>, C / 2) : resident_ctas(conv0_kernel<false>, C / 2);
    ctas_c[fast] = C;
  }
  Conv0Args a{x, L, L0, k, s, C, rows_per_cta(L0, B, ctas[fast]), w, gamma, beta, mean, rstd, (__nv_bfloat16*)y, nullptr, nullptr};
  dim3 grid(cdiv(L0, a.rows_per_cta), B);
  if (k == 10 && s == 5) conv0_k10s5_kernel<false><<<grid, C / 2, 0, stream>>>(a);
  else conv0_kernel<false><<<grid, C / 2, 0, stream>>>(a);
  return check_launch("conv0_kernel<fwd>");
}

extern "C" int a8_conv0_bwd(const float* x, int32_t B, int64_t L, const float* w, const float* gamma,
                            const float* beta, const float* mean, const float* rstd, const double* moments,
                            int32_t C, int32_t k, int32_t s, const void* da, float* acc, float* dw, float* dgamma,
                            float* dbeta, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (conv0_check(C, k, s)) return -1;
  const int L0 = (int)((L - k) / s + 1);
  const bool fast = (k == 10 && s == 5);
  static int ctas[2] = {0, 0}, ctas_c[2] = {0, 0};
  if (ctas_c[fast] != C) {
    ctas[fast] = fast ? resident_ctas(conv0_k10s5_kernel<true>, C / 2) : resident_ctas(conv0_kernel<true>, C / 2);
    ctas_c[fast] = C;
  }
  Conv0Args a{x, L, L0, k, s, C, rows_per_cta(L0, B, ctas[fast]), w, gamma, beta, mean, rstd, nullptr,
              (const __nv_bfloat16*)da, acc};
  dim3 grid(cdiv(L0, a.rows_per_cta), B);
  A8_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 12 * B * C, stream));
  if (k == 10 && s == 5) conv0_k10s5_kernel<true><<<grid, C / 2, 0, stream>>>(a);
  else conv0_kernel<true><<<grid, C / 2, 0, stream>>>(a);
  int rc = check_launch("conv0_kernel<bwd>");
  if (rc) return rc;
  conv0_bwd_finalize_kernel<<<cdiv(C, 128), 128, 0, stream>>>(acc, moments, w, gamma, mean, rstd, B, C, k, L0, dw,
                                                                dgamma, dbeta);
  return check_launch("conv0_bwd_finalize_kernel");
}
