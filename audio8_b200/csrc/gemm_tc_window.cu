// audio8_b200 — tcgen05 GEMM, tap-window variant (a8_gemm_t.reserved = A8_GEMM_TAP_WINDOW | k16 << 8).
//
// For the grouped positional convolution (wav2vec2.py:600-609,634: Conv1d(768, 768, k=128, groups=16)) every k-block of
// the implicit GEMM is one tap: its A tile is the SAME activation rows shifted by one time step.  The generic kernel
// fetches a 16 KB A tile and an 8 KB weight tile per 128 x 64 x 64 block of MMAs, i.e. 192 B per tensor-pipe clock
// against the ~43 B/clk an SM gets from L2 when every SM is loading: it ran at 0.18 of the tensor peak, L2-bound
// (profiles/r02_ncu_full.md #24: L2 throughput 47 %, tensor pipe 27 %).  Here
//   * a CTA owns a 256-row tile (two 128-row accumulators share every weight tile),
//   * taps are visited residue by residue (j = 8 q + r): the 256 + 8 (Q-1) activation rows that the Q taps of one
//     residue touch are staged ONCE (47 KB) and each tap's A operand is that window at a row offset of 8 q rows, a
//     multiple of the 1024-byte swizzle atom, so only the shared-memory descriptor's start address moves,
//   * the weight tile is fetched with exactly the rows (N rounded up to 16) and multiplied over exactly the 16-wide
//     k-steps (k16) that hold non-zero weights: 48 of the padded 64 on both axes for the 16 x 48-channel groups.
// Per 16 taps an SM now reads 47 + 16 x 6 KB for 16 x 2 x 3 MMAs of 128 x 48 x 16 (2304 clk): 62 B/clk.
// Measured at base / 15 s (B=6, T=749; scripts/kernel_table.py): 143 us (plain kernel) -> 88 us with an 8-deep weight ring.
// Roles, barriers and the epilogue are those of gemm_tc_kernel.cuh (same epilogue_chunk code).
#include "gemm_tc_kernel.cuh"
#include <stdlib.h>

namespace a8 {
namespace gemm {

int make_tmap(CUtensorMap* out, const a8_operand_t& v, int box0, int box1, const char* what);  // gemm_tc.cu

namespace {

constexpr int WIN_M = 256;                 // rows per tile (two accumulators)
constexpr int WIN_QMAX = 16;               // taps per residue class
constexpr int WIN_A_SLOTS = 2;
// The weight ring is deep: a tap's MMAs take ~150 clk, an L2 round trip under load ~2000, so 8 stages (the first version)
// left the tensor pipe waiting for weights (88 us at base / 15 s); 96 KB of ring = 16 stages of 48-row tiles.
constexpr int WIN_B_STAGES_MAX = 32;
constexpr uint32_t WIN_B_RING_BYTES = 96 * 1024;
constexpr uint32_t WIN_A_BYTES = (WIN_M + 8 * (WIN_QMAX - 1)) * 128;  // 48128 = 47 swizzle atoms
constexpr uint32_t WIN_BAR_BYTES = 1024;   // 2 x 2 + 2 x 32 + 4 barriers, the TMEM slot
constexpr uint32_t WIN_TMEM_COLS = 256;    // 2 stages x 2 halves x 64 columns
constexpr uint32_t WIN_SMEM_BYTES = WIN_A_SLOTS * WIN_A_BYTES + WIN_B_RING_BYTES + WIN_BAR_BYTES + 2 * 64 * 4 +
                                    EPI_WARPS * STG_BYTES + 1024;

struct WinParams {
  int Q;         // taps per residue (k_blocks / 8)
  int dir;       // +1: the A row coordinate grows with the tap index, -1: it shrinks (data gradient)
  int k16;       // 16-wide k-steps per tap that are multiplied (1..4)
  int n_mma;     // UMMA N = B rows fetched per tap (N rounded up to 16, <= 64)
  int m_tiles2;  // 256-row tiles per (hi, lo) block
  int group;     // taps per weight barrier: one wait / one commit of the MMA thread per `group` taps (a barrier test
                 // costs the issuing thread ~150 clk, as much as the 6 MMAs of a tap)
  int b_groups;  // depth of the weight ring in groups (<= WIN_B_STAGES_MAX); a tile = n_mma * 128 bytes
};

// debug timeline (a8_gemm_set_trace, scripts/win_trace.py): CTA 0, its first tile, per tap: [0..127] the MMA thread saw
// the weight tile, [128..255] it had issued the tap's MMAs and commit, [256..383] the producer saw the stage free
__device__ __forceinline__ void wtrace(const KParams& p, int slot) {
  if (p.trace != nullptr && blockIdx.x == 0 && slot < 384) p.trace[slot] = clock64();
}

// The MMA thread's loop.  What it costs per tap besides the MMAs themselves bounds this kernel (6 MMAs of 24 tensor-pipe
// clocks per tap): the first version rebuilt both 64-bit descriptors from byte addresses for every MMA (shift, mask, or:
// ~60 dependent uniform-datapath instructions and three constant-bank loads per tap, 630 clk per tap measured with
// scripts/win_trace.py against 144 clk of tensor work).  Here the descriptors' low words (address field, 16-byte units)
// are carried incrementally: a tap's operand is the previous one plus a constant, the k-steps add 2, the lower half-tile
// adds 1024; the high word is a constant.
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) {
  constexpr uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, 128B swizzle
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

template <int K16>
__device__ __forceinline__ void window_mma_loop(const KParams& p, const WinParams& w, uint32_t bars, uint32_t sA,
                                                uint32_t sB, uint32_t tmem_base) {
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(w.n_mma >> 3) << 17) |
                         ((uint32_t)(BLOCK_M >> 4) << 24);
  const int Q = opaque(w.Q), ngrp = opaque(w.b_groups), G = opaque(w.group), dir = opaque(w.dir), total = opaque(p.total_tiles);
  const uint32_t b_step = (uint32_t)opaque(w.n_mma) * 8u;           // one weight tile, 16-byte units
  const uint32_t a_first = dir > 0 ? 0u : (uint32_t)(Q - 1) * 64u;  // tap q = 0 of a residue: window offset, 16-byte units
  const int a_step = dir * 64;                                      // next tap: 8 rows = 1024 B further (or back)
  const uint32_t a_full0 = bars, a_empty0 = bars + 8u * WIN_A_SLOTS;
  const uint32_t b_full0 = bars + 8u * (2 * WIN_A_SLOTS), b_empty0 = bars + 8u * (2 * WIN_A_SLOTS + WIN_B_STAGES_MAX);
  const uint32_t tfull0 = bars + 8u * (2 * WIN_A_SLOTS + 2 * WIN_B_STAGES_MAX), tempty0 = tfull0 + 16u;
  const uint32_t b_lo0 = sB >> 4;
  int aslot = 0, grp = 0, iter = 0;
  uint32_t aphase = 0, bphase = 0;
  uint32_t b_lo = b_lo0, b_bar = 0;  // current weight group: descriptor word of its first tile, byte offset of its barriers
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++iter) {
    const int as = iter & 1;
    mbar_wait(tempty0 + 8u * as, ((iter >> 1) & 1u) ^ 1u);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + as * 128;
    for (int r = 0; r < 8; ++r) {
      mbar_wait(a_full0 + 8u * aslot, aphase);
      tc_fence_after();
      uint32_t a_lo = ((sA + aslot * WIN_A_BYTES) >> 4) + a_first;
      for (int q0 = 0; q0 < Q; q0 += G) {
        // TMA -> mbarrier -> tcgen05.mma needs no tcgen05 fence (both sides are the async proxy, the barrier orders them)
        mbar_wait(b_full0 + b_bar, bphase);
        for (int g = 0; g < G; ++g) {
          if (tracing && iter == 0) wtrace(p, 8 * (q0 + g) + r);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < K16; ++k)
              umma_bf16(d_tmem + h * 64, desc_from_lo(a_lo + h * (BLOCK_M * 128 / 16) + k * 2), desc_from_lo(b_lo + k * 2), idesc,
                        (r > 0 || q0 > 0 || g > 0 || k > 0) ? 1u : 0u);
          }
          if (tracing && iter == 0) wtrace(p, 128 + 8 * (q0 + g) + r);
          a_lo += a_step;
          b_lo += b_step;
        }
        umma_commit(b_empty0 + b_bar);
        b_bar += 8u;
        if (++grp == ngrp) {
          grp = 0;
          bphase ^= 1u;
          b_lo = b_lo0;
          b_bar = 0;
        }
      }
      umma_commit(a_empty0 + 8u * aslot);
      if (++aslot == WIN_A_SLOTS) {
        aslot = 0;
        aphase ^= 1u;
      }
    }
    umma_commit(tfull0 + 8u * as);
  }
}

template <int EK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_window_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b, const KParams p, const WinParams w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + WIN_A_SLOTS * WIN_A_BYTES;
  const uint32_t bars = sB + WIN_B_RING_BYTES;
  const int WIN_B_STAGES = w.b_groups;  // barrier pairs in use
  const uint32_t WIN_B_BYTES = (uint32_t)w.n_mma * 128u;  // a multiple of the 1024-byte swizzle atom (n_mma % 16 == 0)
  auto a_full = [&](int i) { return bars + 8u * i; };
  auto a_empty = [&](int i) { return bars + 8u * (WIN_A_SLOTS + i); };
  auto b_full = [&](int i) { return bars + 8u * (2 * WIN_A_SLOTS + i); };
  auto b_empty = [&](int i) { return bars + 8u * (2 * WIN_A_SLOTS + WIN_B_STAGES_MAX + i); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * WIN_A_SLOTS + 2 * WIN_B_STAGES_MAX + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * WIN_A_SLOTS + 2 * WIN_B_STAGES_MAX + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * WIN_A_SLOTS + 2 * WIN_B_STAGES_MAX + 4);
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + WIN_BAR_BYTES - raw_u32));  // [2 accumulator stages][64]
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + 2 * 64);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_a_hi);
      tma_prefetch_desc(&map_a_lo);
      tma_prefetch_desc(&map_b);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < WIN_A_SLOTS; ++i) {
        mbar_init(a_full(i), 1);
        mbar_init(a_empty(i), 1);
      }
      for (int i = 0; i < WIN_B_STAGES; ++i) {
        mbar_init(b_full(i), 1);
        mbar_init(b_empty(i), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(tfull_bar(i), 1);
        mbar_init(tempty_bar(i), EPI_WARPS);
      }
      mbar_fence_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, WIN_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  const int Q = w.Q;
  const uint32_t a_bytes = (uint32_t)(WIN_M + 8 * (Q - 1)) * 128u;
  const uint32_t b_bytes = (uint32_t)w.n_mma * 128u;
  auto decode = [&](int tile, int& mt2, int& lo, int& hi) {
    mt2 = tile % w.m_tiles2;
    const int r = tile / w.m_tiles2;
    lo = r % p.lo_count;
    hi = r / p.lo_count;
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int aslot = 0, stage = 0;
      uint32_t aphase = 0, bphase = 0;
      const bool tracing = p.trace != nullptr && blockIdx.x == 0;
      // weight-tile coordinates advance by a constant per tap: kept in registers instead of re-evaluating the affine
      // map (20 multiply-adds on constant-bank operands) for each of the 128 taps
      int cb8[4], cb1[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        cb1[d] = p.b.cb[d];
        cb8[d] = 8 * p.b.cb[d];
      }
      const uint32_t b_stage_bytes = WIN_B_BYTES;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int mt2, lo, hi;
        decode(tile, mt2, lo, hi);
        const int m0 = mt2 * WIN_M;
        int bt[4];
        op_coords(p.b, 0, 0, 0, lo, hi, bt);  // tap 0 of this (lo, hi)
        for (int r = 0; r < 8; ++r) {
          mbar_wait(a_empty(aslot), aphase ^ 1u);
          mbar_expect_tx(a_full(aslot), a_bytes);
          // the window starts at the smallest row any tap of this residue reads: tap r (dir > 0) or tap 8 (Q-1) + r
          int cc[4];
          op_coords(p.a, 0, w.dir > 0 ? r : 8 * (Q - 1) + r, m0, lo, hi, cc);
          const uint32_t a_dst = sA + aslot * WIN_A_BYTES;
          tma_load_4d(&map_a_hi, a_full(aslot), a_dst, cc[0], cc[1], cc[2], cc[3]);
          tma_load_4d(&map_a_lo, a_full(aslot), a_dst + WIN_M * 128, cc[0], cc[1] + WIN_M, cc[2], cc[3]);
          int bq[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) bq[d] = bt[d] + cb1[d] * r;
          for (int q0 = 0; q0 < Q; q0 += w.group) {
            mbar_wait(b_empty(stage), bphase ^ 1u);
            if (tracing && tile == 0) wtrace(p, 256 + 8 * q0 + r);
            mbar_expect_tx(b_full(stage), b_bytes * w.group);
            for (int g = 0; g < w.group; ++g) {
              tma_load_4d(&map_b, b_full(stage), sB + (stage * w.group + g) * b_stage_bytes, bq[0], bq[1], bq[2], bq[3]);
#pragma unroll
              for (int d = 0; d < 4; ++d) bq[d] += cb8[d];
            }
            if (++stage == WIN_B_STAGES) {
              stage = 0;
              bphase ^= 1u;
            }
          }
          if (++aslot == WIN_A_SLOTS) {
            aslot = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (elect_one()) {
      switch (w.k16) {
        case 1: window_mma_loop<1>(p, w, bars, sA, sB, tmem_base); break;
        case 2: window_mma_loop<2>(p, w, bars, sA, sB, tmem_base); break;
        case 3: window_mma_loop<3>(p, w, bars, sA, sB, tmem_base); break;
        default: window_mma_loop<4>(p, w, bars, sA, sB, tmem_base); break;
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int qd = warp & 3;
    const int part = (warp - 4) >> 2;
    constexpr int COLS = 64 / EPI_PARTS;
    constexpr int NCH = COLS / EPI_CW;
    const int tid_e = threadIdx.x - 128;
    int iter = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
      int mt2, lo, hi;
      decode(tile, mt2, lo, hi);
      const int as = iter & 1;
      float* sb = nullptr;
      if (p.bias != nullptr) {
        sb = s_bias + as * 64;
        const float* bsrc = p.bias + (long long)lo * p.bias_stride_lo;
        for (int i = tid_e; i < 64; i += 32 * EPI_WARPS) sb[i] = (i < p.N) ? __ldg(bsrc + i) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }
      mbar_wait(tfull_bar(as), (iter >> 1) & 1u);
      tc_fence_after();
      uint8_t* stg = s_stage + (warp - 4) * STG_BYTES;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        TileCoord t;
        t.nt = 0; t.mt = mt2 * 2 + h; t.lo = lo; t.hi = hi; t.kb_begin = 0; t.kb_end = p.k_blocks;
        t.M = p.M; t.N = p.N; t.c = p.c; t.ldc = p.ldc; t.g = 0;
        const int row0 = t.mt * BLOCK_M + qd * 32;
        if (row0 >= p.M) continue;  // warp-uniform: this quarter of the half tile lies beyond the last row
        const long long row_off0 = (long long)hi * p.c_stride_hi + (long long)lo * p.c_stride_lo + (long long)row0 * p.ldc;
        const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + as * 128 + h * 64 + part * COLS;
        const int nb0 = part * COLS;
        const float* sbw = sb ? sb + part * COLS : nullptr;
        uint32_t ra[EPI_CW];
        uint4 aux_pre[EPI_CW / 8];
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          if (nb0 + c * EPI_CW >= p.N) break;
          tmem_ld_chunk(t_addr + c * EPI_CW, ra);
          if (((EK >> 5) & 3) != AUX_NONE)  // the aux tile's global loads fly under the TMEM load
            aux_issue<EPI_CW / 8>(aux_pre, p.aux, row_off0, p.ldc, nb0 + c * EPI_CW, min(32, p.M - row0), (p.N + 7) & ~7, lane);
          tmem_ld_wait();
          epilogue_chunk<EK, EPI_CW>(p, t, ra, row_off0, row0, nb0 + c * EPI_CW, -1, aux_pre,
                                     sbw ? sbw + c * EPI_CW : nullptr, stg, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WIN_TMEM_COLS);
  }
}

template <int EK>
int launch_window_inst(const CUtensorMap& mah, const CUtensorMap& mal, const CUtensorMap& mb, const KParams& kp,
                       const WinParams& wp, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_tc_window_kernel<EK>;
  if (!configured) {
    A8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIN_SMEM_BYTES));
    configured = true;
  }
  const int grid = kp.total_tiles < num_sms() ? kp.total_tiles : num_sms();
  A8_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), WIN_SMEM_BYTES, stream, 1, mah, mal, mb, kp, wp));
  return check_launch("gemm_tc_window_kernel");
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[lo][j*64 + ci][co] = sum over (batch, t) of A[t + j][ci] * B[t][co]  (both operands MN-major: the contraction index t
// is the row index in memory).  The plain kernel computes a 128-row tile = 2 taps and fetches their two [64 t x 64 ci]
// atoms plus the B tile per k-block: 24 KB per 128 x 64 x 64 block of MMAs, L2-bound like the forward.  Here one CTA owns
// (group lo, residue r) = the Q taps j = 8 q + r: per k-block (64 values of t) it stages the 64 + 8 (Q-1) rows those taps
// read ONCE; tap q's atom is the window at a row offset of 8 q rows (1024 B), so the M = 128 tile of taps (2i, 2i+1) is
// one MN-major operand whose two atoms lie 1024 B apart (the descriptor's leading-dimension byte offset).  All Q/2
// accumulators (Q/2 x 64 TMEM columns) stay resident over the whole contraction; 31 KB per 32 MMAs of 128 x 64 x 16.
constexpr int WG_STAGES = 5;
constexpr uint32_t WG_A_BYTES = (64 + 8 * (WIN_QMAX - 1)) * 128;  // 23552 = 23 swizzle atoms
constexpr uint32_t WG_B_BYTES = 64 * 128;
constexpr uint32_t WG_SMEM_BYTES = WG_STAGES * (WG_A_BYTES + WG_B_BYTES) + 256 + EPI_WARPS * STG_BYTES + 1024;

struct WgParams {
  int Q;          // taps per residue (even)
  int tmem_cols;  // power of two >= 64 * Q / 2
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_wgrad_window_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const KParams p, const WgParams w) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + WG_STAGES * WG_A_BYTES;
  const uint32_t bars = sB + WG_STAGES * WG_B_BYTES;
  auto full_bar = [&](int i) { return bars + 8u * i; };
  auto empty_bar = [&](int i) { return bars + 8u * (WG_STAGES + i); };
  const uint32_t done_bar = bars + 8u * (2 * WG_STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * WG_STAGES + 1);
  uint8_t* s_stage = smem_raw + (bars + 256u - raw_u32);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < WG_STAGES; ++i) {
        mbar_init(full_bar(i), 1);
        mbar_init(empty_bar(i), 1);
      }
      mbar_init(done_bar, 1);
      mbar_fence_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)w.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  const int Q = w.Q;
  const int r = blockIdx.x & 7, lo = blockIdx.x >> 3;
  const uint32_t a_bytes = (uint32_t)(64 + 8 * (Q - 1)) * 128u;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int kin = 0, kbatch = 0;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), a_bytes + WG_B_BYTES);
        int cc[4];
        op_coords(p.a, kin, kbatch, r, lo, 0, cc);  // the atom of tap r (q = 0): the window's first row
        tma_load_4d(&map_a, full_bar(stage), sA + stage * WG_A_BYTES, cc[0], cc[1], cc[2], cc[3]);
        op_coords(p.b, kin, kbatch, 0, lo, 0, cc);
        tma_load_4d(&map_b, full_bar(stage), sB + stage * WG_B_BYTES, cc[0], cc[1], cc[2], cc[3]);
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
        if (++kin == p.k_inner) {
          kin = 0;
          ++kbatch;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                 ((uint32_t)(BLOCK_M >> 4) << 24);
      // descriptor low words carried incrementally (see window_mma_loop): address field in 16-byte units, the leading-
      // dimension offset (1024 B between the two taps' atoms for A, one 8 KB atom for B) in bits 16..29
      constexpr uint32_t A_LBO = (1024u >> 4) << 16, B_LBO = ((BLOCK_K * 128u) >> 4) << 16;
      const int halfQ = Q / 2;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_lo = (sA >> 4) | A_LBO, b_lo = (sB >> 4) | B_LBO, bar = 0;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(bars + bar, phase);
        tc_fence_after();
        uint32_t a_i = a_lo;
        for (int i = 0; i < halfQ; ++i) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)  // 16 contraction rows per MMA = 2048 B
            umma_bf16(tmem_base + i * 64, desc_from_lo(a_i + k * (UMMA_K * 128 / 16)), desc_from_lo(b_lo + k * (UMMA_K * 128 / 16)),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          a_i += 2048u >> 4;  // taps 2i+2, 2i+3: 16 rows further
        }
        umma_commit(bars + 8u * WG_STAGES + bar);
        a_lo += WG_A_BYTES >> 4;
        b_lo += WG_B_BYTES >> 4;
        bar += 8u;
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1u;
          a_lo = (sA >> 4) | A_LBO;
          b_lo = (sB >> 4) | B_LBO;
          bar = 0;
        }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int part = (warp - 4) >> 2;
    constexpr int COLS = 64 / EPI_PARTS;
    constexpr int NCH = COLS / EPI_CW;
    constexpr int EK = ek_make(OUT_F32, 0, 0, AUX_NONE);
    mbar_wait(done_bar, 0);
    tc_fence_after();
    uint8_t* stg = s_stage + (warp - 4) * STG_BYTES;
#pragma unroll 1
    for (int i = 0; i < Q / 2; ++i) {
      const int tap = 8 * (2 * i + (qd >> 1)) + r;       // TMEM lanes 0..63 of tile i: tap 2i, lanes 64..127: tap 2i+1
      const int row0 = tap * 64 + (qd & 1) * 32;
      TileCoord t;
      t.nt = 0; t.mt = 0; t.lo = lo; t.hi = 0; t.kb_begin = 0; t.kb_end = p.k_blocks;
      t.M = p.M; t.N = p.N; t.c = p.c; t.ldc = p.ldc; t.g = 0;
      const long long row_off0 = (long long)lo * p.c_stride_lo + (long long)row0 * p.ldc;
      const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + i * 64 + part * COLS;
      uint32_t ra[EPI_CW];
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        if (part * COLS + c * EPI_CW >= p.N) break;
        tmem_ld_chunk(t_addr + c * EPI_CW, ra);
        tmem_ld_wait();
        uint4 no_aux[EPI_CW / 8];
        epilogue_chunk<EK, EPI_CW>(p, t, ra, row_off0, row0, part * COLS + c * EPI_CW, -1, no_aux, nullptr, stg, lane);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)w.tmem_cols);
  }
}

}  // namespace

// called by a8_gemm (gemm_tc.cu) with kp filled as for the generic kernel
int launch_wgrad_window(const a8_gemm_t& g, KParams& kp, cudaStream_t stream) {
  A8_REQUIRE(g.M % 1024 == 0 && g.M >= 1024 && g.M <= 8 * WIN_QMAX * 64,
             "gemm wgrad window: M = %d must be taps x 64 with taps a multiple of 16 in [16, %d]", g.M, 8 * WIN_QMAX);
  A8_REQUIRE(g.N >= 1 && g.N <= 64 && g.N % 8 == 0, "gemm wgrad window: N = %d must be a multiple of 8, <= 64", g.N);
  A8_REQUIRE(g.c_dtype == OUT_F32 && g.split_k <= 1 && g.act == ACT_NONE && g.aux == nullptr && g.z_out == nullptr &&
                 g.bias == nullptr && g.hi_count <= 1 && g.alpha == 1.f,
             "gemm wgrad window: plain fp32 output only (no split-K, epilogue extras or hi batching)");
  // the 64-row atom index r of operand A is the tap: it moves the row coordinate (dim 1) by one and nothing else
  A8_REQUIRE(g.a.cr[1] == 1 && g.a.cr[0] == 0 && g.a.cr[2] == 0 && g.a.cr[3] == 0,
             "gemm wgrad window: operand A must advance one row of dim 1 per 64-row atom");
  WgParams wp;
  wp.Q = g.M / 64 / 8;
  int cols = 64 * wp.Q / 2;
  wp.tmem_cols = 32;
  while (wp.tmem_cols < cols) wp.tmem_cols *= 2;
  kp.total_tiles = 8 * kp.lo_count;
  CUtensorMap ma, mb;
  if (int rc = make_tmap(&ma, g.a, 64, 64 + 8 * (wp.Q - 1), "wgrad window A")) return rc;
  if (int rc = make_tmap(&mb, g.b, 64, BLOCK_K, "wgrad window B")) return rc;
  static bool configured = false;
  if (!configured) {
    A8_CUDA(cudaFuncSetAttribute(gemm_tc_wgrad_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM_BYTES));
    configured = true;
  }
  A8_CUDA(launch_pdl(gemm_tc_wgrad_window_kernel, dim3(kp.total_tiles), dim3(GEMM_THREADS), WG_SMEM_BYTES, stream, 1, ma, mb, kp, wp));
  return check_launch("gemm_tc_wgrad_window_kernel");
}

int launch_window(const a8_gemm_t& g, KParams& kp, int k16, cudaStream_t stream) {
  if (g.a.major == MAJOR_MN && g.b.major == MAJOR_MN) return launch_wgrad_window(g, kp, stream);
  A8_REQUIRE(g.a.major == MAJOR_K && g.b.major == MAJOR_K, "gemm window: both operands must be K-major (or both MN-major)");
  A8_REQUIRE(g.k_inner == 1 && (g.split_k <= 1), "gemm window: k_inner must be 1 and split_k 1 (one k-block per tap)");
  A8_REQUIRE(g.k_blocks % 8 == 0 && g.k_blocks >= 16 && g.k_blocks <= 8 * WIN_QMAX,
             "gemm window: k_blocks = %d must be a multiple of 8 in [16, %d]", g.k_blocks, 8 * WIN_QMAX);
  A8_REQUIRE(g.N >= 1 && g.N <= 64, "gemm window: N = %d must be <= 64", g.N);
  A8_REQUIRE(g.c_dtype == OUT_BF16, "gemm window: bf16 output only");
  // per tap the A tile moves by exactly one row (dim 1) and nothing else; rows of the tile are rows of dim 1
  A8_REQUIRE((g.a.cb[1] == 1 || g.a.cb[1] == -1) && g.a.cb[0] == 0 && g.a.cb[2] == 0 && g.a.cb[3] == 0 && g.a.cr[1] == 1 &&
                 g.a.cr[0] == 0 && g.a.cr[2] == 0 && g.a.cr[3] == 0,
             "gemm window: operand A must advance one row of dim 1 per k-block");
  A8_REQUIRE(k16 >= 0 && k16 <= 4, "gemm window: k16 = %d", k16);
  WinParams wp;
  wp.Q = g.k_blocks / 8;
  wp.dir = g.a.cb[1];
  wp.k16 = k16 == 0 ? 4 : k16;
  wp.n_mma = ((g.N + 15) / 16) * 16;
  wp.m_tiles2 = cdiv(g.M, WIN_M);
  wp.group = (wp.Q % 4 == 0) ? 4 : ((wp.Q % 2 == 0) ? 2 : 1);
  wp.b_groups = (int)(WIN_B_RING_BYTES / (wp.n_mma * 128u)) / wp.group;
  if (wp.b_groups > WIN_B_STAGES_MAX) wp.b_groups = WIN_B_STAGES_MAX;

  kp.m_tiles = wp.m_tiles2;
  kp.n_tiles = 1;
  const long long tiles = (long long)wp.m_tiles2 * kp.lo_count * kp.hi_count;
  A8_REQUIRE(tiles < (1ll << 30), "gemm window: too many tiles");
  kp.total_tiles = (int)tiles;
  CUtensorMap mah, mal, mb;
  if (int rc = make_tmap(&mah, g.a, BLOCK_K, WIN_M, "window A")) return rc;
  if (int rc = make_tmap(&mal, g.a, BLOCK_K, 8 * (wp.Q - 1), "window A tail")) return rc;
  if (int rc = make_tmap(&mb, g.b, BLOCK_K, wp.n_mma, "window B")) return rc;
  const int ek = ek_make(g.c_dtype, g.act, g.z_out != nullptr, g.aux_mode);
  switch (ek) {
    case ek_make(OUT_BF16, ACT_GELU_DZ, 1, AUX_ADD): return launch_window_inst<ek_make(OUT_BF16, ACT_GELU_DZ, 1, AUX_ADD)>(mah, mal, mb, kp, wp, stream);
    case ek_make(OUT_BF16, ACT_GELU, 0, AUX_ADD): return launch_window_inst<ek_make(OUT_BF16, ACT_GELU, 0, AUX_ADD)>(mah, mal, mb, kp, wp, stream);
    case ek_make(OUT_BF16, 0, 0, AUX_ADD): return launch_window_inst<ek_make(OUT_BF16, 0, 0, AUX_ADD)>(mah, mal, mb, kp, wp, stream);
    case ek_make(OUT_BF16, 0, 0, AUX_NONE): return launch_window_inst<ek_make(OUT_BF16, 0, 0, AUX_NONE)>(mah, mal, mb, kp, wp, stream);
  }
  return launch_window_inst<EK_GENERIC>(mah, mal, mb, kp, wp, stream);
}

}  // namespace gemm
}  // namespace a8
