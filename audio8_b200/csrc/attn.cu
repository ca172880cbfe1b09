// audio8_b200 — fused scaled-dot-product attention for sm_100a (d_k = 64): forward, dQ and dK/dV kernels.
//
// Replaces eight_mile's SeqScaledDotProductAttention as called through `wav2vec2.py:644` (QK^T * d_k^-1/2,
// masked_fill(pad, -1e9), softmax, dropout, PV) and its autograd.  The [B,H,T,T] score / probability tensors are
// never written to HBM (the reference materialises 161 MB per layer at base / 15 s): scores live in TMEM, the
// probabilities go back to TMEM as the bf16 A-operand of the next tcgen05.mma.
//
// All three kernels share one shape: a CTA owns 128 rows (queries, or keys in the dK/dV kernel) and walks over
// 128-wide blocks of the other sequence axis;
//   warp 0  TMA producer (Q/K/V/dO tiles of the fused [B,T,3D] projection buffer, 128B swizzle, 2-stage ring)
//   warp 1  tcgen05.mma issuer: "score" MMAs (both operands from smem) into TMEM, then "accumulate" MMAs whose
//           A operand is the packed bf16 tile the math warps stored into TMEM (B = smem tile read MN-major)
//   warp 2  TMEM allocator
//   warps 4..  one thread per TMEM lane (= row): tcgen05.ld scores, exp2 / dropout / dS math, tcgen05.st
// Softmax statistics: forward keeps a running max that is only raised when a block exceeds it by 2^8 (then the
// O accumulator is rescaled through registers), so the common case is a single pass per block; log2-sum-exp is
// saved per row and the backward kernels recompute probabilities from it.
// Dropout: keep decisions are a pure function of (seed, b, h, q, k) (2 multiplies per 2x2 patch of the score
// matrix, usable from both orientations), regenerated in backward; nothing is stored.
#include "a8_common.cuh"
#include "a8_tmap.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {

const unsigned long long* seed_source();  // a8_api.cu

namespace {

// A8_ATTN_DIAG (experiments, scripts/attn_diag.py; results are WRONG with any bit set): 1 = forward softmax math replaced
// by a register move, 2 = no tcgen05.ld of the scores, 4 = no tcgen05.mma issued (commits only)
#ifndef A8_ATTN_DIAG
#define A8_ATTN_DIAG 0
#endif
#ifndef A8_ATTN_TRACE
#define A8_ATTN_TRACE 0
#endif
constexpr uint32_t TILE = 128 * 64 * 2;  // one [128 x 64] bf16 tile, 128B rows
constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);  // f32 acc, bf16 A/B, M=128
constexpr uint32_t IDESC_S = IDESC_BASE | ((uint32_t)(128 >> 3) << 17);                 // N=128, A and B K-major
constexpr uint32_t IDESC_S64 = IDESC_BASE | ((uint32_t)(64 >> 3) << 17);                 // N=64, A and B K-major
constexpr uint32_t IDESC_ACC = IDESC_BASE | ((uint32_t)(64 >> 3) << 17) | (1u << 16);   // N=64, B MN-major
constexpr uint32_t HALF_TILE = 64 * 64 * 2;  // 64 rows of a [128 x 64] tile (a multiple of the 1024-byte swizzle atom)

struct AttnArgs {
  int B, H, T, nblk;
  const unsigned char* key_keep;  // [B,T] or null
  float scale, scale_log2, keep_scale;
  uint32_t thr;                   // keep iff r(q,k) >= thr  (thr = p * 2^32)
  unsigned long long seed;
  const unsigned long long* seed_src;
  __nv_bfloat16* ctx;             // fwd out [B,T,D]
  const __nv_bfloat16* ctx_in;    // bwd in
  const __nv_bfloat16* dctx;      // bwd in
  float* lse;                     // [B,H,T]  log2-sum-exp2 of the scaled scores
  float* delta;                   // [B,H,T]  rowsum(dO * O)
  __nv_bfloat16* dqkv;            // bwd out [B,T,3D]
  float* dbias;                   // bwd, nullable [3D]: += column sums of dqkv as stored (the QKV bias gradient)
  long long* trace;               // A8_ATTN_TRACE builds: [5 CTAs][3 roles][64] clock64 stamps (scripts/attn_trace.py)
};

#if A8_ATTN_TRACE
__device__ __forceinline__ int trace_cta() {
  const int b = blockIdx.x;
  return b == 0 ? 0 : (b == 1 ? 1 : (b == 150 ? 2 : (b == 300 ? 3 : (b == 431 ? 4 : -1))));
}
__device__ __forceinline__ void atrace(const AttnArgs& a, int role, int slot) {
  const int c = trace_cta();
  if (a.trace != nullptr && c >= 0 && slot < 64) a.trace[(c * 3 + role) * 64 + slot] = clock64();
}
#define ATRACE(role, slot) atrace(a, role, slot)
#else
#define ATRACE(role, slot)
#endif

// ---------------------------------------------------------------------------------------------- dropout hash
// keep(q,k) <=> r(q,k) >= thr32, with  x = xorshift(((q/2 << 16) + k/2 + key) * M1)  shared by a 2x2 patch of the score
// matrix and r = x * M[(q&1)*2 + (k&1)].  Multiplies run on the FMA pipe; per element only the compare and the select
// (plus the patch's shift/xor) hit the half-rate ALU pipe, which is what bounds these kernels.  Usable from both orientations (query-row threads
// in forward / dQ, key-row threads in dK/dV).
constexpr uint32_t HM1 = 0x9E3779B1u;
__device__ __forceinline__ uint32_t hmix(uint32_t x) {
  x *= 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(unsigned long long seed, int bh) {
  return hmix((uint32_t)seed ^ hmix((uint32_t)(seed >> 32) + 0x9E3779B1u * (uint32_t)(bh + 1)));
}
__device__ __forceinline__ uint32_t drop_mul(int qodd, int kodd) {
  return qodd ? (kodd ? 0x165667B1u : 0x27D4EB2Fu) : (kodd ? 0xC2B2AE3Du : 0x85EBCA77u);
}
// linear part (additive in qp and kp, so loops advance it with one multiply-add) ...
__device__ __forceinline__ uint32_t drop_w(uint32_t key, int qp, int kp) {
  return (((uint32_t)qp << 16) + (uint32_t)kp + key) * HM1;
}
// ... one xorshift to break the lattice, then an element-specific multiply; the compare reads the high bits
__device__ __forceinline__ uint32_t drop_x(uint32_t w) { return w ^ (w >> 15); }
__device__ __forceinline__ uint32_t drop_r(uint32_t x, uint32_t mul) { return x * mul; }

__device__ __forceinline__ uint32_t valid_word(const unsigned char* keep, int T, int k0) {
  uint32_t w = 0;
  for (int i = 0; i < 32; ++i) {
    const int k = k0 + i;
    if (k < T && (keep == nullptr || keep[k] != 0)) w |= (1u << i);
  }
  return w;
}

// Column sums of a 32 (lanes = rows) x 32 (registers = columns) tile: 5 exchange steps, each halves the number of live
// columns per lane (31 shuffles in all); on return lane l holds the sum of column l in v[0].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
// the two bf16 values of a packed pair, as the fp32 numbers the consumer of the stored tensor will read
__device__ __forceinline__ void bf16_pair_values(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xFFFF0000u);
}

// shared prologue: barrier ids are kernel specific; TMEM allocation by warp 2
__device__ __forceinline__ uint32_t bar_at(uint32_t bars, int i) { return bars + 8u * i; }

// ---------------------------------------------------------------------------------------------- per-chunk math
// 32 score columns of one row -> 16 packed bf16 probability pairs; returns the row-sum contribution.  MASKED is the
// rare case (a chunk that contains padded or masked-out keys: the last key block, ragged batches): the common chunk
// carries no per-element validity test at all (as `if (word != ~0u)` inside the loop the compiler predicated the three
// mask instructions per element instead of branching around them: 20 % of the kernel's issue slots).
template <bool DROP, bool MASKED>
__device__ __forceinline__ float softmax_chunk(const uint32_t (&r)[32], uint32_t (&pk)[16], float c, float m,
                                               uint32_t word, uint32_t wch, uint32_t mul_e, uint32_t mul_o,
                                               uint32_t thr) {
  float lsum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float t0 = fmaf(__uint_as_float(r[i]), c, -m), t1 = fmaf(__uint_as_float(r[i + 1]), c, -m);
    if (MASKED) {
      if (!((word >> i) & 1u)) t0 = -INFINITY;
      if (!((word >> (i + 1)) & 1u)) t1 = -INFINITY;
    }
    float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
    lsum += p0 + p1;
    if (DROP) {
      const uint32_t w = drop_x(wch + (uint32_t)(i >> 1) * HM1);
      p0 = (drop_r(w, mul_e) >= thr) ? p0 : 0.f;
      p1 = (drop_r(w, mul_o) >= thr) ? p1 : 0.f;
    }
    pk[i >> 1] = pack_bf16(p0, p1);
  }
  return lsum;
}

// 32 columns of S and dP of one query row -> 16 packed bf16 dS pairs (dQ kernel)
template <bool DROP, bool MASKED>
__device__ __forceinline__ void ds_chunk(const uint32_t (&rs)[32], const uint32_t (&rp)[32], uint32_t (&pk)[16],
                                         float c, float nlse, float dscale, float ndelta, uint32_t word, uint32_t wch,
                                         uint32_t mul_e, uint32_t mul_o, uint32_t thr) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float t0 = fmaf(__uint_as_float(rs[i]), c, nlse), t1 = fmaf(__uint_as_float(rs[i + 1]), c, nlse);
    if (MASKED) {
      if (!((word >> i) & 1u)) t0 = -INFINITY;
      if (!((word >> (i + 1)) & 1u)) t1 = -INFINITY;
    }
    const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
    // dS = p * (keep * dP / (1-p_drop) - delta) * scale, with the scale folded into the two constants
    float d0 = fmaf(__uint_as_float(rp[i]), dscale, ndelta), d1 = fmaf(__uint_as_float(rp[i + 1]), dscale, ndelta);
    if (DROP) {
      const uint32_t w = drop_x(wch + (uint32_t)(i >> 1) * HM1);
      d0 = (drop_r(w, mul_e) >= thr) ? d0 : ndelta;
      d1 = (drop_r(w, mul_o) >= thr) ? d1 : ndelta;
    }
    pk[i >> 1] = pack_bf16(p0 * d0, p1 * d1);
  }
}

// ================================================================================================ forward
// The key axis is walked in UNITS of 64 keys (half of a staged 128-key tile) with S and P double-buffered in TMEM: the
// score MMA of unit u+1 runs while the softmax warps work on unit u and the PV MMA of unit u-1 is in flight.
// TMEM columns: S stage s [64s, 64s+64)  P (packed bf16) stage s [128+32s, +32)  O [192,256)
template <bool DROP>
__global__ void __launch_bounds__(256, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + TILE, sV = base + 3 * TILE;
  const uint32_t bars = base + 5 * TILE;
  enum { Q_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, S_FULL = 5, P_READY = 7, PV_DONE = 9 };  // two of each but Q_FULL
  const uint32_t tmem_slot = bars + 96;
  uint32_t* sValid = reinterpret_cast<uint32_t*>(smem_raw + (bars + 128 - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  if (threadIdx.x == 128) ATRACE(2, 0);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 128) ATRACE(2, 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = a.nblk;
  const int U = 2 * nblk;
  const int qb = blockIdx.x % nblk;
  const int bh = blockIdx.x / nblk;
  const int h = bh % a.H, b = bh / a.H;
  const int D = a.H * 64;

  if (warp == 0) {
    if (elect_one()) tma_prefetch_desc(&map_qkv);
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_init(bar_at(bars, Q_FULL), 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_at(bars, KV_FULL + i), 1);
        mbar_init(bar_at(bars, KV_EMPTY + i), 1);
        mbar_init(bar_at(bars, S_FULL + i), 1);
        mbar_init(bar_at(bars, P_READY + i), 128);
        mbar_init(bar_at(bars, PV_DONE + i), 1);
      }
      mbar_fence_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
  } else if (warp >= 4) {
    const unsigned char* keep = a.key_keep ? a.key_keep + (long long)b * a.T : nullptr;
    for (int w = threadIdx.x - 128; w < nblk * 4; w += 128) sValid[w] = valid_word(keep, a.T, w * 32);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 128) ATRACE(2, 2);

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_at(bars, Q_FULL), TILE);
      tma_load_3d(&map_qkv, bar_at(bars, Q_FULL), sQ, h * 64, qb * 128, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(bar_at(bars, KV_EMPTY + st), ((j >> 1) & 1) ^ 1u);
        ATRACE(0, j);
        mbar_expect_tx(bar_at(bars, KV_FULL + st), 2 * TILE);
        tma_load_3d(&map_qkv, bar_at(bars, KV_FULL + st), sK + st * TILE, D + h * 64, j * 128, b);
        tma_load_3d(&map_qkv, bar_at(bars, KV_FULL + st), sV + st * TILE, 2 * D + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t tO = tmem_base + 192;
      // descriptor low words (16-byte units); a unit's operand = tile base + stage * TILE + half * HALF_TILE
      const uint32_t q_lo = umma_desc_lo(sQ, 0), k_lo0 = umma_desc_lo(sK, 0), v_lo0 = umma_desc_lo(sV, 8192);
      auto issue_pv = [&](int u) {  // O += P(u) [128 x 64 keys, TMEM] * V(u) [64 keys x 64, read MN-major]
        const uint32_t v_lo = v_lo0 + (((u >> 1) & 1) * TILE + (u & 1) * HALF_TILE) / 16;
        const uint32_t tP = tmem_base + 128 + (u & 1) * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(A8_ATTN_DIAG & 4))
            umma_bf16_ts(tO, tP + k * 8, umma_desc_sbo1024(v_lo + k * (2048 / 16)), IDESC_ACC, (u > 0 || k > 0) ? 1u : 0u);
        if (u & 1) umma_commit(bar_at(bars, KV_EMPTY + ((u >> 1) & 1)));
        umma_commit(bar_at(bars, PV_DONE + (u & 1)));
      };
      mbar_wait(bar_at(bars, Q_FULL), 0);
      ATRACE(1, 0);
      for (int u = 0; u < U; ++u) {
        const int j = u >> 1, s = u & 1;
        if (s == 0) mbar_wait(bar_at(bars, KV_FULL + (j & 1)), (j >> 1) & 1);
        ATRACE(1, 1 + 3 * u);
        // PV(u-2), the last reader of P stage s, must have COMPLETED before the scores of unit u are issued: the softmax
        // warps take S_FULL(u) as the licence to overwrite that P buffer.  Relying on the S_FULL commit merely being
        // issued after PV(u-2) gave timing-dependent errors on peaked score distributions (scripts/attn_accuracy.py).
        // Waiting here costs this one thread ~100 cycles it has to spare; in the softmax warps it would cost every unit.
        if (u >= 2) mbar_wait(bar_at(bars, PV_DONE + s), ((u - 2) >> 1) & 1);
        tc_fence_after();
        const uint32_t kk_lo = k_lo0 + ((j & 1) * TILE + s * HALF_TILE) / 16;
        const uint32_t tS = tmem_base + s * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(A8_ATTN_DIAG & 4))
            umma_bf16(tS, umma_desc_sbo1024(q_lo + k * 2), umma_desc_sbo1024(kk_lo + k * 2), IDESC_S64, k > 0 ? 1u : 0u);
        umma_commit(bar_at(bars, S_FULL + s));
        ATRACE(1, 2 + 3 * u);
        if (u > 0) {
          mbar_wait(bar_at(bars, P_READY + (s ^ 1)), ((u - 1) >> 1) & 1);
          tc_fence_after();
          ATRACE(1, 3 + 3 * u);
          issue_pv(u - 1);
        }
      }
      mbar_wait(bar_at(bars, P_READY + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
      tc_fence_after();
      issue_pv(U - 1);
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const int row = qb * 128 + quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem_row = tmem_base + lane_addr, tO = tmem_row + 192;
    const float c = a.scale_log2;
    uint32_t dkey = 0;
    if (DROP) dkey = drop_key(a.seed + seed_base_ld(a.seed_src), bh);
    const uint32_t mul_e = drop_mul(row & 1, 0), mul_o = drop_mul(row & 1, 1);
    const uint32_t wrow = drop_w(dkey, row >> 1, 0);
    float m = -INFINITY, l = 0.f;
#pragma unroll 1
    for (int u = 0; u < U; ++u) {
      const int s = u & 1;
      const uint32_t tS = tmem_row + s * 64, tP = tmem_row + 128 + s * 32;
      if (threadIdx.x == 128) ATRACE(2, 3 + 2 * u);
      mbar_wait(bar_at(bars, S_FULL + s), (u >> 1) & 1);  // S(u) is only issued once PV(u-2) has released this P stage
      tc_fence_after();
      if (threadIdx.x == 128) ATRACE(2, 4 + 2 * u);
      uint32_t r[32];
      if (A8_ATTN_DIAG & 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint((float)(lane + i + u) * 0.01f);
      }
      // The running max is only raised when a unit would overflow the row sum (any exp2 above 2^64): the common case
      // is ONE pass per unit with no per-element max.  Unit 0 and the rare overflow take the exact-max pass first.
      bool need_max = (u == 0);
      while (true) {
        if (need_max) {
          float mx = -INFINITY;
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            tmem_ld_32x32(tS + ch * 32, r);
            tmem_ld_wait();
            const uint32_t word = sValid[u * 2 + ch];
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((word >> i) & 1u) ? __uint_as_float(r[i]) : -INFINITY);
          }
          float m_new = fmaxf(m, mx * c);
          if (m_new == -INFINITY) m_new = 0.f;
          if (u > 0) {  // rescale the row sum and the O accumulator once every PV issued so far has completed
            mbar_wait(bar_at(bars, PV_DONE + (s ^ 1)), ((u - 1) >> 1) & 1);
            tc_fence_after();
            const float f = ex2_approx(m - m_new);
            l *= f;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
              tmem_ld_32x32(tO + ch * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st_32x32(tO + ch * 32, r);
            }
            tmem_st_wait();
          }
          m = m_new;
        }
        float lsum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          if (!(A8_ATTN_DIAG & 2)) {
            tmem_ld_32x32(tS + ch * 32, r);
            tmem_ld_wait();
          }
          const uint32_t word = sValid[u * 2 + ch];
          const uint32_t wch = wrow + (uint32_t)((u * 64 + ch * 32) >> 1) * HM1;
          uint32_t pk[16];
          if (A8_ATTN_DIAG & 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = r[2 * i] + wch;
            lsum = 1.f;
          } else if (word == 0xFFFFFFFFu) lsum += softmax_chunk<DROP, false>(r, pk, c, m, word, wch, mul_e, mul_o, a.thr);
          else lsum += softmax_chunk<DROP, true>(r, pk, c, m, word, wch, mul_e, mul_o, a.thr);
          tmem_st_32x16(tP + ch * 16, pk);
        }
        if (!need_max && __any_sync(0xffffffffu, !(lsum < 1.8e19f))) {
          need_max = true;
          tmem_st_wait();
          continue;
        }
        l += lsum;
        break;
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_at(bars, P_READY + s));
    }
    if (threadIdx.x == 128) ATRACE(2, 3 + 2 * U);
    mbar_wait(bar_at(bars, PV_DONE + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
    tc_fence_after();
    if (threadIdx.x == 128) ATRACE(2, 4 + 2 * U);
    const float inv = l > 0.f ? a.keep_scale / l : 0.f;
    uint32_t r[32];
    __nv_bfloat16* dst = a.ctx + ((long long)b * a.T + row) * D + h * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld_32x32(tO + ch * 32, r);
      tmem_ld_wait();
      if (row < a.T) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + ch * 32 + g * 8) = o;
        }
      }
    }
    if (row < a.T) a.lse[(long long)bh * a.T + row] = l > 0.f ? m + log2f(l) : INFINITY;
    if (threadIdx.x == 128) ATRACE(2, 5 + 2 * U);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 128) ATRACE(2, 6 + 2 * U);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ================================================================================================ backward: dQ
// thread = query row; 8 math warps (two per lane quarter).  The key axis is walked in UNITS of 64 keys (half of a
// staged 128-key tile); S / dP / dS are double-buffered in TMEM, so the score MMAs of unit u+1 run while the math warps
// work on unit u and the dQ accumulation of unit u-1 is in flight.
// TMEM columns: stage s: S [128s, 128s+64)  dP [128s+64, 128s+128);  dS packed [256+32s, +32);  dQ [320,384)
template <bool DROP>
__global__ void __launch_bounds__(384, 1)
attn_dq_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
               const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sdO = base + TILE, sK = base + 2 * TILE, sV = base + 4 * TILE;
  const uint32_t bars = base + 6 * TILE;
  enum { QDO_FULL = 0, KV_FULL = 1, KV_EMPTY = 3, S_FULL = 5, DS_READY = 7, DQ_DONE = 9 };
  const uint32_t tmem_slot = bars + 96;
  uint32_t* sValid = reinterpret_cast<uint32_t*>(smem_raw + (bars + 128 - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = a.nblk;
  const int U = 2 * nblk;
  const int qb = blockIdx.x % nblk;
  const int bh = blockIdx.x / nblk;
  const int h = bh % a.H, b = bh / a.H;
  const int D = a.H * 64;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_qkv);
      tma_prefetch_desc(&map_do);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_init(bar_at(bars, QDO_FULL), 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_at(bars, KV_FULL + i), 1);
        mbar_init(bar_at(bars, KV_EMPTY + i), 1);
        mbar_init(bar_at(bars, S_FULL + i), 1);
        mbar_init(bar_at(bars, DS_READY + i), 256);
        mbar_init(bar_at(bars, DQ_DONE + i), 1);
      }
      mbar_fence_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
  } else if (warp >= 4) {
    const unsigned char* keep = a.key_keep ? a.key_keep + (long long)b * a.T : nullptr;
    for (int w = threadIdx.x - 128; w < nblk * 4; w += 256) sValid[w] = valid_word(keep, a.T, w * 32);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_at(bars, QDO_FULL), 2 * TILE);
      tma_load_3d(&map_qkv, bar_at(bars, QDO_FULL), sQ, h * 64, qb * 128, b);
      tma_load_3d(&map_do, bar_at(bars, QDO_FULL), sdO, h * 64, qb * 128, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(bar_at(bars, KV_EMPTY + st), ((j >> 1) & 1) ^ 1u);
        mbar_expect_tx(bar_at(bars, KV_FULL + st), 2 * TILE);
        tma_load_3d(&map_qkv, bar_at(bars, KV_FULL + st), sK + st * TILE, D + h * 64, j * 128, b);
        tma_load_3d(&map_qkv, bar_at(bars, KV_FULL + st), sV + st * TILE, 2 * D + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t tdQ = tmem_base + 320;
      const uint32_t q_lo = umma_desc_lo(sQ, 0), do_lo = umma_desc_lo(sdO, 0), k_lo0 = umma_desc_lo(sK, 0), v_lo0 = umma_desc_lo(sV, 0);
      const uint32_t kmn_lo0 = umma_desc_lo(sK, 8192);  // K read MN-major (the dQ accumulation)
      auto issue_dq = [&](int u) {  // dQ += dS(u) [128 x 64 keys, TMEM] * K(u) [64 keys x 64, read MN-major]
        const uint32_t kk_lo = kmn_lo0 + (((u >> 1) & 1) * TILE + (u & 1) * HALF_TILE) / 16;
        const uint32_t tdS = tmem_base + 256 + (u & 1) * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tdQ, tdS + k * 8, umma_desc_sbo1024(kk_lo + k * (2048 / 16)), IDESC_ACC, (u > 0 || k > 0) ? 1u : 0u);
        if (u & 1) umma_commit(bar_at(bars, KV_EMPTY + ((u >> 1) & 1)));
        umma_commit(bar_at(bars, DQ_DONE + (u & 1)));
      };
      mbar_wait(bar_at(bars, QDO_FULL), 0);
      for (int u = 0; u < U; ++u) {
        const int j = u >> 1, s = u & 1;
        if (s == 0) mbar_wait(bar_at(bars, KV_FULL + (j & 1)), (j >> 1) & 1);
        // Stage s: the math warps' DS_READY of unit u-2 was awaited before dQ(u-2) was issued, and dQ(u-2) itself must
        // have COMPLETED before the scores of unit u are issued: the math warps take S_FULL(u) as the licence to overwrite
        // this stage's dS buffer.  (Relying on the S_FULL commit being issued after dQ(u-2) was not enough in the forward
        // variant of this pipeline: timing-dependent errors, scripts/attn_accuracy.py.)  The wait costs this one thread
        // ~100 cycles it has to spare; in the math warps it would cost every unit.
        if (u >= 2) mbar_wait(bar_at(bars, DQ_DONE + s), ((u - 2) >> 1) & 1);
        tc_fence_after();
        const uint32_t off = ((j & 1) * TILE + s * HALF_TILE) / 16;
        const uint32_t kk_lo = k_lo0 + off, vv_lo = v_lo0 + off;
        const uint32_t tS = tmem_base + s * 128, tdP = tS + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tS, umma_desc_sbo1024(q_lo + k * 2), umma_desc_sbo1024(kk_lo + k * 2), IDESC_S64, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdP, umma_desc_sbo1024(do_lo + k * 2), umma_desc_sbo1024(vv_lo + k * 2), IDESC_S64, k > 0 ? 1u : 0u);
        umma_commit(bar_at(bars, S_FULL + s));
        if (u > 0) {
          mbar_wait(bar_at(bars, DS_READY + (s ^ 1)), ((u - 1) >> 1) & 1);
          tc_fence_after();
          issue_dq(u - 1);
        }
      }
      mbar_wait(bar_at(bars, DS_READY + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
      tc_fence_after();
      issue_dq(U - 1);
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int row = qb * 128 + quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem_row = tmem_base + lane_addr, tdQ = tmem_row + 320;
    const float c = a.scale_log2;
    uint32_t dkey = 0;
    if (DROP) dkey = drop_key(a.seed + seed_base_ld(a.seed_src), bh);
    const uint32_t mul_e = drop_mul(row & 1, 0), mul_o = drop_mul(row & 1, 1);
    const uint32_t wrow = drop_w(dkey, row >> 1, 0);
    // delta = rowsum(dO * O); lse of this row
    float delta = 0.f, nlse = -INFINITY;
    if (row < a.T) {
      const long long ro = ((long long)b * a.T + row) * D + h * 64;
      const uint4* po = reinterpret_cast<const uint4*>(a.ctx_in + ro);
      const uint4* pd = reinterpret_cast<const uint4*>(a.dctx + ro);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 o = __ldg(po + g), d = __ldg(pd + g);
        const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(ow[t]), y = unpack_bf16(dw[t]);
          delta = fmaf(x.x, y.x, delta);
          delta = fmaf(x.y, y.y, delta);
        }
      }
      nlse = -a.lse[(long long)bh * a.T + row];
      if (half == 0) a.delta[(long long)bh * a.T + row] = delta;
    }
    const float dscale = a.keep_scale * a.scale, ndelta = -delta * a.scale;
#pragma unroll 1
    for (int u = 0; u < U; ++u) {
      const int s = u & 1;
      mbar_wait(bar_at(bars, S_FULL + s), (u >> 1) & 1);  // S(u) is only issued once dQ(u-2) has released this dS stage
      tc_fence_after();
      const int col0 = u * 64 + half * 32;  // first key of this thread's 32 columns
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tmem_row + s * 128 + half * 32, rs);
      tmem_ld_32x32(tmem_row + s * 128 + 64 + half * 32, rp);
      tmem_ld_wait();
      const uint32_t word = sValid[col0 >> 5];
      const uint32_t wch = wrow + (uint32_t)(col0 >> 1) * HM1;
      uint32_t pk[16];
      if (word == 0xFFFFFFFFu) ds_chunk<DROP, false>(rs, rp, pk, c, nlse, dscale, ndelta, word, wch, mul_e, mul_o, a.thr);
      else ds_chunk<DROP, true>(rs, rp, pk, c, nlse, dscale, ndelta, word, wch, mul_e, mul_o, a.thr);
      tmem_st_32x16(tmem_row + 256 + s * 32 + half * 16, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_at(bars, DS_READY + s));
    }
    mbar_wait(bar_at(bars, DQ_DONE + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tdQ + half * 32, r);
    tmem_ld_wait();
    float cs[32];
    {
      __nv_bfloat16* dst = a.dqkv + ((long long)b * a.T + row) * (3 * D) + h * 64 + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
        o.y = pack_bf16(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
        o.z = pack_bf16(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
        o.w = pack_bf16(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
        if (row < a.T) *reinterpret_cast<uint4*>(dst + g * 8) = o;
        if (row >= a.T) o = make_uint4(0u, 0u, 0u, 0u);
        bf16_pair_values(o.x, cs[8 * g + 0], cs[8 * g + 1]);
        bf16_pair_values(o.y, cs[8 * g + 2], cs[8 * g + 3]);
        bf16_pair_values(o.z, cs[8 * g + 4], cs[8 * g + 5]);
        bf16_pair_values(o.w, cs[8 * g + 6], cs[8 * g + 7]);
      }
    }
    if (a.dbias != nullptr) {  // QKV bias gradient: column sums of this CTA's rows of dQ (no separate pass over dqkv)
      const float t = warp_colsum32(cs, lane);
      atomicAdd(a.dbias + h * 64 + half * 32 + lane, t);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ backward: dK, dV
// thread = key row, walks over the queries in UNITS of 64 (half of a staged 128-query tile), S^T / dP^T / P^T / dS^T
// double-buffered in TMEM like the dQ kernel.  TMEM columns:
//   stage s: S^T [128s, +64)  dP^T [128s+64, +64);  P^T packed [256+32s, +32)  dS^T packed [320+32s, +32);  dV [384,448)  dK [448,512)
template <bool DROP>
__global__ void __launch_bounds__(384, 1)
attn_dkv_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sK = base, sV = base + TILE, sQ = base + 2 * TILE, sdO = base + 4 * TILE;
  const uint32_t bars = base + 6 * TILE;
  enum { KV_FULL = 0, Q_FULL = 1, Q_EMPTY = 3, S_FULL = 5, PS_READY = 7, ACC_DONE = 9 };
  const uint32_t tmem_slot = bars + 96;
  float* sStat = reinterpret_cast<float*>(smem_raw + (bars + 128 - raw));  // [2 stages][nlse 128 | delta 128]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = a.nblk;
  const int U = 2 * nblk;
  const int kb = blockIdx.x % nblk;
  const int bh = blockIdx.x / nblk;
  const int h = bh % a.H, b = bh / a.H;
  const int D = a.H * 64;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_qkv);
      tma_prefetch_desc(&map_do);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_init(bar_at(bars, KV_FULL), 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_at(bars, Q_FULL + i), 1);
        mbar_init(bar_at(bars, Q_EMPTY + i), 1);
        mbar_init(bar_at(bars, S_FULL + i), 1);
        mbar_init(bar_at(bars, PS_READY + i), 256);
        mbar_init(bar_at(bars, ACC_DONE + i), 1);
      }
      mbar_fence_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // whole warp: lanes stage the per-query statistics of each block, lane 0 drives TMA and the barriers
    if (lane == 0) {
      mbar_expect_tx(bar_at(bars, KV_FULL), 2 * TILE);
      tma_load_3d(&map_qkv, bar_at(bars, KV_FULL), sK, D + h * 64, kb * 128, b);
      tma_load_3d(&map_qkv, bar_at(bars, KV_FULL), sV, 2 * D + h * 64, kb * 128, b);
    }
    for (int i = 0; i < nblk; ++i) {
      const int st = i & 1;
      mbar_wait(bar_at(bars, Q_EMPTY + st), ((i >> 1) & 1) ^ 1u);
      float* dst = sStat + st * 256;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int q = i * 128 + t * 32 + lane;
        const bool ok = q < a.T;
        dst[t * 32 + lane] = ok ? -a.lse[(long long)bh * a.T + q] : -INFINITY;  // -inf: p == 0 for q >= T
        dst[128 + t * 32 + lane] = ok ? -a.delta[(long long)bh * a.T + q] * a.scale : 0.f;  // -delta * scale
      }
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(bar_at(bars, Q_FULL + st), 2 * TILE);
        tma_load_3d(&map_qkv, bar_at(bars, Q_FULL + st), sQ + st * TILE, h * 64, i * 128, b);
        tma_load_3d(&map_do, bar_at(bars, Q_FULL + st), sdO + st * TILE, h * 64, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t tdV = tmem_base + 384, tdK = tmem_base + 448;
      const uint32_t k_lo = umma_desc_lo(sK, 0), v_lo = umma_desc_lo(sV, 0), q_lo0 = umma_desc_lo(sQ, 0), do_lo0 = umma_desc_lo(sdO, 0);
      const uint32_t qmn_lo0 = umma_desc_lo(sQ, 8192), domn_lo0 = umma_desc_lo(sdO, 8192);  // read MN-major (accumulations)
      auto issue_acc = [&](int u) {  // dV += P^T(u) dO(u),  dK += dS^T(u) Q(u): 64 queries of K extent each
        const uint32_t off = (((u >> 1) & 1) * TILE + (u & 1) * HALF_TILE) / 16;
        const uint32_t qq_lo = qmn_lo0 + off, dd_lo = domn_lo0 + off;
        const uint32_t tPT = tmem_base + 256 + (u & 1) * 32, tdST = tmem_base + 320 + (u & 1) * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tdV, tPT + k * 8, umma_desc_sbo1024(dd_lo + k * (2048 / 16)), IDESC_ACC, (u > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tdK, tdST + k * 8, umma_desc_sbo1024(qq_lo + k * (2048 / 16)), IDESC_ACC, (u > 0 || k > 0) ? 1u : 0u);
        if (u & 1) umma_commit(bar_at(bars, Q_EMPTY + ((u >> 1) & 1)));
        umma_commit(bar_at(bars, ACC_DONE + (u & 1)));
      };
      mbar_wait(bar_at(bars, KV_FULL), 0);
      for (int u = 0; u < U; ++u) {
        const int i = u >> 1, s = u & 1;
        if (s == 0) mbar_wait(bar_at(bars, Q_FULL + (i & 1)), (i >> 1) & 1);
        if (u >= 2) mbar_wait(bar_at(bars, ACC_DONE + s), ((u - 2) >> 1) & 1);  // as in the dQ kernel: acc(u-2) has completed
        tc_fence_after();
        const uint32_t off = ((i & 1) * TILE + s * HALF_TILE) / 16;
        const uint32_t qq_lo = q_lo0 + off, dd_lo = do_lo0 + off;
        const uint32_t tST = tmem_base + s * 128, tdPT = tST + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tST, umma_desc_sbo1024(k_lo + k * 2), umma_desc_sbo1024(qq_lo + k * 2), IDESC_S64, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tdPT, umma_desc_sbo1024(v_lo + k * 2), umma_desc_sbo1024(dd_lo + k * 2), IDESC_S64, k > 0 ? 1u : 0u);
        umma_commit(bar_at(bars, S_FULL + s));
        if (u > 0) {
          mbar_wait(bar_at(bars, PS_READY + (s ^ 1)), ((u - 1) >> 1) & 1);
          tc_fence_after();
          issue_acc(u - 1);
        }
      }
      mbar_wait(bar_at(bars, PS_READY + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
      tc_fence_after();
      issue_acc(U - 1);
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int key = kb * 128 + quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem_row = tmem_base + lane_addr, tdV = tmem_row + 384, tdK = tmem_row + 448;
    const bool valid_row = key < a.T && (a.key_keep == nullptr || a.key_keep[(long long)b * a.T + key] != 0);
    const float c = a.scale_log2;
    const float dscale = a.keep_scale * a.scale;
    uint32_t dkey = 0;
    if (DROP) dkey = drop_key(a.seed + seed_base_ld(a.seed_src), bh);
    const uint32_t mul_qe = drop_mul(0, key & 1), mul_qo = drop_mul(1, key & 1);
    const uint32_t wkey = drop_w(dkey, 0, key >> 1);
#pragma unroll 1
    for (int u = 0; u < U; ++u) {
      const int i = u >> 1, s = u & 1;
      if (s == 0) mbar_wait(bar_at(bars, Q_FULL + (i & 1)), (i >> 1) & 1);  // the block's statistics are staged (acquire)
      mbar_wait(bar_at(bars, S_FULL + s), (u >> 1) & 1);  // S(u) is only issued once acc(u-2) has released P^T / dS^T
      tc_fence_after();
      const int col0 = s * 64 + half * 32;  // first query (inside the 128-query tile) of this thread's 32 columns
      const float* stat = sStat + (i & 1) * 256 + col0;
      uint32_t rs[32], rp[32];
      tmem_ld_32x32(tmem_row + s * 128 + half * 32, rs);
      tmem_ld_32x32(tmem_row + s * 128 + 64 + half * 32, rp);
      tmem_ld_wait();
      const uint32_t wq0 = wkey + (uint32_t)((i * 128 + col0) >> 1) * (HM1 << 16);
      uint32_t pp[16], pd[16];
      if (valid_row) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 nl = *reinterpret_cast<const float2*>(stat + e);
          const float2 dl = *reinterpret_cast<const float2*>(stat + 128 + e);  // -delta * scale
          const float p0 = ex2_approx(fmaf(__uint_as_float(rs[e]), c, nl.x));
          const float p1 = ex2_approx(fmaf(__uint_as_float(rs[e + 1]), c, nl.y));
          // dS = p * (keep * dP / (1-p_drop) - delta) * scale, with the scale folded into dscale and the staged -delta
          float d0 = fmaf(__uint_as_float(rp[e]), dscale, dl.x), d1 = fmaf(__uint_as_float(rp[e + 1]), dscale, dl.y);
          float k0 = p0, k1 = p1;
          if (DROP) {
            const uint32_t w = drop_x(wq0 + (uint32_t)(e >> 1) * (HM1 << 16));  // patch ((q0 + e) / 2, key / 2)
            const bool keep0 = drop_r(w, mul_qe) >= a.thr, keep1 = drop_r(w, mul_qo) >= a.thr;
            k0 = keep0 ? p0 : 0.f;
            k1 = keep1 ? p1 : 0.f;
            d0 = keep0 ? d0 : dl.x;
            d1 = keep1 ? d1 : dl.y;
          }
          pp[e >> 1] = pack_bf16(k0, k1);
          pd[e >> 1] = pack_bf16(p0 * d0, p1 * d1);
        }
      } else {  // a key beyond T or masked out: its column of P and dS is zero
#pragma unroll
        for (int e = 0; e < 16; ++e) pp[e] = pd[e] = 0u;
      }
      tmem_st_32x16(tmem_row + 256 + s * 32 + half * 16, pp);
      tmem_st_32x16(tmem_row + 320 + s * 32 + half * 16, pd);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_at(bars, PS_READY + s));
    }
    mbar_wait(bar_at(bars, ACC_DONE + ((U - 1) & 1)), ((U - 1) >> 1) & 1);
    tc_fence_after();
    // half 0 writes dV (scaled by 1/(1-p)), half 1 writes dK
    const uint32_t src = half == 0 ? tdV : tdK;
    const float sc = half == 0 ? a.keep_scale : 1.f;
    __nv_bfloat16* dst = a.dqkv + ((long long)b * a.T + key) * (3 * D) + (half == 0 ? 2 * D : D) + h * 64;
    uint32_t r[32];
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld_32x32(src + ch * 32, r);
      tmem_ld_wait();
      float cs[32];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[8 * g + 0]) * sc, __uint_as_float(r[8 * g + 1]) * sc);
        o.y = pack_bf16(__uint_as_float(r[8 * g + 2]) * sc, __uint_as_float(r[8 * g + 3]) * sc);
        o.z = pack_bf16(__uint_as_float(r[8 * g + 4]) * sc, __uint_as_float(r[8 * g + 5]) * sc);
        o.w = pack_bf16(__uint_as_float(r[8 * g + 6]) * sc, __uint_as_float(r[8 * g + 7]) * sc);
        if (key < a.T) *reinterpret_cast<uint4*>(dst + ch * 32 + g * 8) = o;
        if (key >= a.T) o = make_uint4(0u, 0u, 0u, 0u);
        bf16_pair_values(o.x, cs[8 * g + 0], cs[8 * g + 1]);
        bf16_pair_values(o.y, cs[8 * g + 2], cs[8 * g + 3]);
        bf16_pair_values(o.z, cs[8 * g + 4], cs[8 * g + 5]);
        bf16_pair_values(o.w, cs[8 * g + 6], cs[8 * g + 7]);
      }
      if (a.dbias != nullptr) {  // QKV bias gradient, K (zero in exact arithmetic) and V thirds
        const float t = warp_colsum32(cs, lane);
        atomicAdd(a.dbias + (half == 0 ? 2 * D : D) + h * 64 + ch * 32 + lane, t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// keep decisions as bytes [B,H,T,T] (tests: the mask the fused kernels regenerate)
__global__ void attn_dropmask_kernel(unsigned char* out, int B, int H, int T, uint32_t thr, unsigned long long seed,
                                     const unsigned long long* seed_src) {
  const long long n = (long long)B * H * T * T;
  const unsigned long long s = seed + seed_base_ld(seed_src);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % T);
    const int q = (int)((i / T) % T);
    const int bh = (int)(i / ((long long)T * T));
    const uint32_t w = drop_x(drop_w(drop_key(s, bh), q >> 1, k >> 1));
    out[i] = drop_r(w, drop_mul(q & 1, k & 1)) >= thr ? 1 : 0;
  }
}

long long* g_attn_trace = nullptr;

int fill_args(AttnArgs& a, const uint8_t* key_keep, int B, int H, int T, float scale, float pdrop, uint64_t seed) {
  A8_REQUIRE(B > 0 && H > 0 && T > 0 && T <= 4096, "attention: unsupported shape B=%d H=%d T=%d (T <= 4096)", B, H, T);
  A8_REQUIRE(pdrop >= 0.f && pdrop < 1.f, "attention: dropout probability %f", pdrop);
  memset(&a, 0, sizeof(a));
  a.B = B; a.H = H; a.T = T; a.nblk = cdiv(T, 128);
  a.key_keep = key_keep;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.thr = (uint32_t)((double)pdrop * 4294967296.0);
  a.keep_scale = pdrop > 0.f ? (float)(4294967296.0 / (4294967296.0 - (double)a.thr)) : 1.f;  // 1 / realised keep rate
  a.seed = seed;
  a.seed_src = seed_source();
  a.trace = g_attn_trace;
  return 0;
}

constexpr int SMEM_FWD = 5 * TILE + 1024 + 128 + 512;
constexpr int SMEM_BWD = 6 * TILE + 1024 + 128 + 2048;

template <typename K>
int set_smem(K kern, int bytes) {
  A8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace
}  // namespace a8

using namespace a8;

#if A8_ATTN_TRACE
extern "C" void a8_attn_set_trace(void* buf) { g_attn_trace = static_cast<long long*>(buf); }  // debug builds only
#endif

extern "C" int a8_attn_fwd(const void* qkv, const uint8_t* key_keep, void* ctx, float* lse, int32_t B, int32_t H,
                           int32_t T, float scale, float pdrop, uint64_t seed, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  AttnArgs a;
  if (int rc = fill_args(a, key_keep, B, H, T, scale, pdrop, seed)) return rc;
  a.ctx = static_cast<__nv_bfloat16*>(ctx);
  a.lse = lse;
  const long long D3 = 3ll * H * 64;
  CUtensorMap mq;
  if (int rc = make_tmap_3d(&mq, qkv, D3, T, B, D3, (long long)T * D3, 64, 128, "attention qkv")) return rc;
  static bool cfg = false;
  if (!cfg) {
    if (int rc = set_smem(attn_fwd_kernel<false>, SMEM_FWD)) return rc;
    if (int rc = set_smem(attn_fwd_kernel<true>, SMEM_FWD)) return rc;
    cfg = true;
  }
  const int grid = a.nblk * B * H;
  if (pdrop > 0.f) A8_CUDA(launch_pdl(attn_fwd_kernel<true>, dim3(grid), dim3(256), SMEM_FWD, stream, 1, mq, a));
  else A8_CUDA(launch_pdl(attn_fwd_kernel<false>, dim3(grid), dim3(256), SMEM_FWD, stream, 1, mq, a));
  return check_launch("attn_fwd_kernel");
}

extern "C" int a8_attn_bwd(const void* qkv, const uint8_t* key_keep, const void* ctx, const void* dctx,
                           const float* lse, float* delta, void* dqkv, float* dbias, int32_t B, int32_t H, int32_t T,
                           float scale, float pdrop, uint64_t seed, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  AttnArgs a;
  if (int rc = fill_args(a, key_keep, B, H, T, scale, pdrop, seed)) return rc;
  a.ctx_in = static_cast<const __nv_bfloat16*>(ctx);
  a.dctx = static_cast<const __nv_bfloat16*>(dctx);
  a.lse = const_cast<float*>(lse);
  a.delta = delta;
  a.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  a.dbias = dbias;
  const long long D = 1ll * H * 64, D3 = 3 * D;
  CUtensorMap mq, md;
  if (int rc = make_tmap_3d(&mq, qkv, D3, T, B, D3, (long long)T * D3, 64, 128, "attention qkv")) return rc;
  if (int rc = make_tmap_3d(&md, dctx, D, T, B, D, (long long)T * D, 64, 128, "attention dctx")) return rc;
  static bool cfg = false;
  if (!cfg) {
    if (int rc = set_smem(attn_dq_kernel<false>, SMEM_BWD)) return rc;
    if (int rc = set_smem(attn_dq_kernel<true>, SMEM_BWD)) return rc;
    if (int rc = set_smem(attn_dkv_kernel<false>, SMEM_BWD)) return rc;
    if (int rc = set_smem(attn_dkv_kernel<true>, SMEM_BWD)) return rc;
    cfg = true;
  }
  const int grid = a.nblk * B * H;
  if (pdrop > 0.f) A8_CUDA(launch_pdl(attn_dq_kernel<true>, dim3(grid), dim3(384), SMEM_BWD, stream, 1, mq, md, a));
  else A8_CUDA(launch_pdl(attn_dq_kernel<false>, dim3(grid), dim3(384), SMEM_BWD, stream, 1, mq, md, a));
  if (int rc = check_launch("attn_dq_kernel")) return rc;
  if (pdrop > 0.f) A8_CUDA(launch_pdl(attn_dkv_kernel<true>, dim3(grid), dim3(384), SMEM_BWD, stream, 1, mq, md, a));
  else A8_CUDA(launch_pdl(attn_dkv_kernel<false>, dim3(grid), dim3(384), SMEM_BWD, stream, 1, mq, md, a));
  return check_launch("attn_dkv_kernel");
}

extern "C" int a8_attn_dropmask(uint8_t* keep_out, int32_t B, int32_t H, int32_t T, float pdrop, uint64_t seed,
                                void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  A8_REQUIRE(B > 0 && H > 0 && T > 0, "dropmask: bad shape");
  attn_dropmask_kernel<<<148 * 4, 256, 0, stream>>>(keep_out, B, H, T, (uint32_t)((double)pdrop * 4294967296.0), seed,
                                                     seed_source());
  return check_launch("attn_dropmask_kernel");
}
