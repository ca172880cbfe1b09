// audio8_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
// Roles inside one 384-thread CTA (one CTA per SM, persistent over output tiles):
//   warp 0 (one elected lane)  TMA producer: cp.async.bulk.tensor.4d -> 128B-swizzled smem ring
//   warp 1 (one elected lane)  MMA issuer:   tcgen05.mma (M=128, N=BN, K=16) -> TMEM accumulators
//   warp 2                     TMEM allocator / deallocator
//   warps 4..11                epilogue: tcgen05.ld -> registers -> bias / GELU (+ derivative) / residual or multiply
//                              (packed fp32 pairs) -> staged coalesced stores (+ fused column sums) -> HBM
// Pipelines: smem full/empty mbarrier ring (TMA <-> MMA) and a 2-deep TMEM accumulator ring
// (MMA <-> epilogue), so the epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle span
constexpr int UMMA_K = 16;
#ifndef A8_EPI16
#define A8_EPI16 0
#endif
// epilogue warps 4..: EPI_PARTS per TMEM lane quarter, each owns 1/EPI_PARTS of the tile's columns and drains them in
// chunks of EPI_CW columns.  A8_EPI16=1 (16 warps x 16-column chunks, one pipeline stage less) was measured on the
// step's shapes (scripts/gemm_bench.py, r01): 0-8 % slower than 8 warps x 32 columns on every GEMM, GELU kinds included.
constexpr int EPI_PARTS = A8_EPI16 ? 4 : 2;
constexpr int EPI_CW = A8_EPI16 ? 16 : 32;
constexpr int EPI_WARPS = 4 * EPI_PARTS;
constexpr int STG_PITCH = 80;                 // bytes per staged row: <= 64 payload + 16 pad (conflict-free 16B writes)
constexpr int STG_BYTES = 32 * STG_PITCH;     // per epilogue warp
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;
enum { MAJOR_K = A8_MAJOR_K, MAJOR_MN = A8_MAJOR_MN };
enum { OUT_BF16 = A8_OUT_BF16, OUT_F32 = A8_OUT_F32, OUT_F32_ATOMIC = A8_OUT_F32_ATOMIC };
enum { ACT_NONE = A8_ACT_NONE, ACT_GELU = A8_ACT_GELU, ACT_GELU_DZ = A8_ACT_GELU_DZ };
enum { AUX_NONE = A8_AUX_NONE, AUX_ADD = A8_AUX_ADD, AUX_MUL_GELU_GRAD = A8_AUX_MUL_GELU_GRAD, AUX_MUL = A8_AUX_MUL };

struct OpCoef {
  int base[4], ck[4], cb[4], cr[4], cl[4], ch[4];
};

// Epilogue kind, a template parameter of the kernel: EK_GENERIC reads every switch from KParams at run time (one
// large body: ~9000 SASS instructions, which thrashes the instruction cache of the 8 epilogue warps); any other
// value fixes (c_dtype | act << 2 | z_out << 4 | aux_mode << 5) at compile time so that the hot shapes run a
// straight-line epilogue a few hundred instructions long.
constexpr int EK_GENERIC = -1;
constexpr int ek_make(int c_dtype, int act, int z, int aux) { return c_dtype | (act << 2) | (z << 4) | (aux << 5); }

struct KParams {
  int M, N, m_tiles, n_tiles, lo_count, hi_count;  // m_tiles counts tile PAIRS when the kernel runs as 2-CTA clusters
  int k_blocks, k_inner, split_k;
  OpCoef a, b;
  void* c;
  int c_dtype;
  void* z_out;
  const void* aux;
  int aux_mode;
  const float* bias;
  int bias_stride_lo;
  int act;
  float alpha;
  long long ldc, c_stride_lo, c_stride_hi;
  int total_tiles;
  long long* trace;  // debug timeline (a8_gemm_set_trace), normally null
  float* colsum;     // nullable: += column sums of the stored bf16 output (a8_gemm_t::colsum)
};

// Grouped launch (a8_gemm_group): ONE persistent kernel walks the tiles of up to GROUP_MAX problems that share operand
// majors, tile shape, epilogue kind, coordinate maps and split-K factor but have their own operands (tensor maps),
// output, extents and contraction length.  Set-up, first-load latency and the exposed last epilogue are then paid once
// per group instead of once per problem (e.g. the 4 weight-gradient GEMMs of each of the 12 transformer layers).
constexpr int GROUP_MAX = 48;
struct GroupProb {
  void* c;
  long long ldc;
  int M, N, m_tiles, n_tiles, k_blocks, tile_end;  // tile_end: exclusive prefix over the group's tile counts
};
struct alignas(64) GroupParams {
  CUtensorMap map_a[GROUP_MAX];
  CUtensorMap map_b[GROUP_MAX];
  GroupProb prob[GROUP_MAX];
  int n_prob;
};

__device__ __forceinline__ void op_coords(const OpCoef& o, int kin, int kbatch, int r, int lo, int hi,
                                          int (&c)[4]) {
#pragma unroll
  for (int d = 0; d < 4; ++d)
    c[d] = o.base[d] + o.ck[d] * kin + o.cb[d] * kbatch + o.cr[d] * r + o.cl[d] * lo + o.ch[d] * hi;
}

// debug timeline: clock64() stamps of CTAs 0..3, roles {0 producer, 1 MMA issuer, 2 epilogue warp 4}, up to 8 tiles,
// 4 events each (scripts/gemm_trace.py).  One predicated store per event when the pointer is null-checked.
constexpr int TRACE_CTAS = 4, TRACE_TILES = 8, TRACE_EVENTS = 4;
__device__ __forceinline__ void trace_ev(const KParams& p, int role, int iter, int ev) {
  if (p.trace != nullptr && blockIdx.x < TRACE_CTAS && iter < TRACE_TILES)
    p.trace[((blockIdx.x * 3 + role) * TRACE_TILES + iter) * TRACE_EVENTS + ev] = clock64();
}

__device__ __forceinline__ void trace_wall(const KParams& p, int slot) {  // %globaltimer (ns) of CTAs 0..3, thread 0
  if (p.trace != nullptr && blockIdx.x < TRACE_CTAS && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[((blockIdx.x * 3 + 0) * TRACE_TILES + (TRACE_TILES - 1)) * TRACE_EVENTS + slot] = (long long)t;
  }
}

template <int BN, int CL = 1>
struct Cfg {
  // CL == 2: a CTA pair works on one 256 x BN tile with cta_group::2 MMAs; each CTA stages its own 128 rows of A and
  // its own HALF of the B tile (BN/2 rows), so a stage is smaller and the ring deeper.
  static constexpr int STAGES0 = (CL == 2) ? (BN == 256 ? 6 : (BN == 192 ? 7 : 8)) : ((BN == 256) ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8)));
  static constexpr int STAGES = STAGES0 - (EPI_WARPS > 8 ? 1 : 0);  // the 16-warp staging buffers take one stage's room
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = (BN / CL) * BLOCK_K * 2;
  static constexpr uint32_t TMEM_COLS = (BN == 192) ? 512 : 2 * BN;  // powers of two >= 32; 2 accumulator stages
  // ring | barriers | bias [2][BN] | column-sum partials [2][BN] | epilogue staging | alignment slack
  static constexpr uint32_t SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 + 4 * BN * 4 + EPI_WARPS * STG_BYTES + 1024;
};

// UMMA shared-memory matrix descriptor, 128B swizzle (layout type 2), descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

struct TileCoord {
  int nt, mt, lo, hi, kb_begin, kb_end;
  // the problem this tile belongs to (the launch's only one, or an entry of the group)
  int M, N;
  void* c;
  long long ldc;
  int g;
};
template <int CL>
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int tile, int rank) {
  TileCoord t;
  t.nt = tile % p.n_tiles;
  int r = tile / p.n_tiles;
  t.mt = (r % p.m_tiles) * CL + rank;
  r /= p.m_tiles;
  t.lo = r % p.lo_count;
  r /= p.lo_count;
  t.hi = r % p.hi_count;
  const int split = r / p.hi_count;
  t.kb_begin = (int)(((long long)split * p.k_blocks) / p.split_k);
  t.kb_end = (int)(((long long)(split + 1) * p.k_blocks) / p.split_k);
  t.M = p.M; t.N = p.N; t.c = p.c; t.ldc = p.ldc; t.g = 0;
  return t;
}
// grouped launch: `g` is the caller's cursor into the problem table (tiles are visited in increasing order)
template <int CL>
__device__ __forceinline__ TileCoord decode_tile_group(const KParams& p, const GroupProb* __restrict__ probs, int tile,
                                                       int rank, int& g) {
  while (tile >= probs[g].tile_end) ++g;
  const GroupProb& q = probs[g];
  int r = tile - (g > 0 ? probs[g - 1].tile_end : 0);
  TileCoord t;
  t.nt = r % q.n_tiles;
  r /= q.n_tiles;
  t.mt = (r % q.m_tiles) * CL + rank;
  const int split = r / q.m_tiles;
  t.lo = 0; t.hi = 0;
  t.kb_begin = (int)(((long long)split * q.k_blocks) / p.split_k);
  t.kb_end = (int)(((long long)(split + 1) * q.k_blocks) / p.split_k);
  t.M = q.M; t.N = q.N; t.c = q.c; t.ldc = q.ldc; t.g = g;
  return t;
}

// Coalesced copy between a warp's staging buffer (32 rows x PIECES*16 B) and global rows `row_off0 + r*ldc` (element
// offsets of element size ES): lane l moves 16 B of row (l/PIECES + (32/PIECES) i), piece (l % PIECES) — every
// instruction touches 32/PIECES rows x PIECES*16 contiguous bytes instead of 32 rows x 16 bytes.
enum { STG_STORE = 1, STG_RED = 2 };
template <int ES, int MODE, int PIECES>
__device__ __forceinline__ void stage_copy(uint8_t* stg, void* gbase, long long row_off0, long long ldc, int col0,
                                           int rows_valid, int cols_valid, int lane) {
  constexpr int EPP = 16 / ES;  // elements per 16-byte piece
  constexpr int RPI = 32 / PIECES;  // rows per instruction
  const int piece = lane % PIECES;
  const int col = col0 + piece * EPP;
  // all shared-memory reads first, then the global accesses: with a load -> store pair per piece the compiler reuses one
  // register quad and every store waits for its own LDS (ncu: 23 % of the epilogue's samples sat on those STGs)
  uint4 v[PIECES];
#pragma unroll
  for (int i = 0; i < PIECES; ++i)
    v[i] = *reinterpret_cast<const uint4*>(stg + (lane / PIECES + RPI * i) * STG_PITCH + piece * 16);
#pragma unroll
  for (int i = 0; i < PIECES; ++i) {
    const int r = lane / PIECES + RPI * i;
    if (r < rows_valid && col < cols_valid) {
      uint8_t* g = reinterpret_cast<uint8_t*>(gbase) + (row_off0 + (long long)r * ldc + col) * ES;
      if (MODE == STG_STORE) {
        *reinterpret_cast<uint4*>(g) = v[i];
      } else {  // split-K: one 16-byte vector reduction per lane
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(__uint_as_float(v[i].x)),
                     "f"(__uint_as_float(v[i].y)), "f"(__uint_as_float(v[i].z)), "f"(__uint_as_float(v[i].w))
                     : "memory");
      }
    }
  }
}
// The bf16 / fp16 aux tile of a chunk (residual to add, GELU' to multiply by) in the same coalesced lane pattern, in two
// steps: `aux_issue` puts the global loads in flight (for the NEXT chunk, one chunk of math ahead), `aux_commit` parks
// them in the warp's staging buffer from which every lane then reads its own row.  (Loading the tile where it is
// consumed left the full global-load latency exposed once per chunk: 60 % of the samples of the dgrad epilogues.)
template <int PIECES>
__device__ __forceinline__ void aux_issue(uint4 (&pre)[PIECES], const void* gbase, long long row_off0, long long ldc,
                                          int col0, int rows_valid, int cols_valid, int lane) {
  constexpr int RPI = 32 / PIECES;
  const int piece = lane % PIECES;
  const int col = col0 + piece * 8;
#pragma unroll
  for (int i = 0; i < PIECES; ++i) {
    const int r = lane / PIECES + RPI * i;
    pre[i] = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid && col < cols_valid)
      pre[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(gbase) +
                                                    (row_off0 + (long long)r * ldc + col) * 2));
  }
}
template <int PIECES>
__device__ __forceinline__ void aux_commit(uint8_t* stg, const uint4 (&pre)[PIECES], int lane) {
  constexpr int RPI = 32 / PIECES;
#pragma unroll
  for (int i = 0; i < PIECES; ++i)
    *reinterpret_cast<uint4*>(stg + (lane / PIECES + RPI * i) * STG_PITCH + (lane % PIECES) * 16) = pre[i];
}

// CW columns x 32 rows (one row per lane) of accumulators in registers -> epilogue math -> global memory.  Outputs (and
// the bf16 aux input) go through the warp's smem staging buffer so that global accesses are row-contiguous (direct
// 16-byte-per-row stores from registers were measured: slower, L2 sees partial sectors).
// `pre` holds this chunk's aux tile (aux_issue); it is refilled for the chunk at column nb_next (< 0: none) as soon as it
// has been parked.
template <int EK, int CW>
__device__ __forceinline__ void epilogue_chunk(const KParams& p, const TileCoord& tc, const uint32_t* r,
                                               long long row_off0, int row0, int nb, int nb_next, uint4 (&pre)[CW / 8],
                                               const float* sb, uint8_t* stg, int lane, float* cs_smem = nullptr) {
  constexpr bool GEN = (EK < 0);
  constexpr int NP = CW / 8;  // 16-byte bf16 pieces per row
  const int c_dtype = GEN ? p.c_dtype : (EK & 3);
  const int act = GEN ? p.act : ((EK >> 2) & 3);
  const bool do_gelu = act == ACT_GELU;
  const bool do_gelu_dz = act == ACT_GELU_DZ;  // out = gelu(v), z_out = gelu'(v)
  const bool do_z = GEN ? (p.z_out != nullptr) : (((EK >> 4) & 1) != 0);
  const int aux_mode = GEN ? p.aux_mode : ((EK >> 5) & 3);
  const int rows_valid = min(32, tc.M - row0);       // may be <= 0
  const int cols_valid = (tc.N + 7) & ~7;             // absolute column bound for the 16-byte pieces
  const long long ldc = tc.ldc;
  void* const cptr = tc.c;
  const bool has_aux = aux_mode != AUX_NONE;
  uint4* my = reinterpret_cast<uint4*>(stg + lane * STG_PITCH);
  uint4 a[NP];
  if (has_aux) {
    aux_commit<NP>(stg, pre, lane);
    __syncwarp();
#pragma unroll
    for (int g = 0; g < NP; ++g) a[g] = my[g];
    __syncwarp();
    if (nb_next >= 0) aux_issue<NP>(pre, p.aux, row_off0, ldc, nb_next, rows_valid, cols_valid, lane);
  }
  float v[CW];
  const float alpha = p.alpha;
  if (sb != nullptr) {  // broadcast LDS.128 (a scalar LDS per column costs a full shared-memory wavefront each)
    const float4* sb4 = reinterpret_cast<const float4*>(sb);
#pragma unroll
    for (int i = 0; i < CW / 4; ++i) {
      const float4 b4 = sb4[i];
      unpk2(fma2(pk2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), bc2(alpha), pk2(b4.x, b4.y)),
            v[4 * i], v[4 * i + 1]);
      unpk2(fma2(pk2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), bc2(alpha), pk2(b4.z, b4.w)),
            v[4 * i + 2], v[4 * i + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < CW; i += 2)
      unpk2(mul2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), bc2(alpha)), v[i], v[i + 1]);
  }
  if (do_gelu_dz) {  // the activation and its derivative from one erf / exp evaluation; the derivative goes to z_out as fp16
#pragma unroll
    for (int g = 0; g < NP; ++g) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) gelu_both_fast2(v[8 * g + j], v[8 * g + j + 1], d[j], d[j + 1]);
      if (do_z) my[g] = make_uint4(pack_f16(d[0], d[1]), pack_f16(d[2], d[3]), pack_f16(d[4], d[5]), pack_f16(d[6], d[7]));
    }
  } else if (do_z) {
#pragma unroll
    for (int g = 0; g < NP; ++g) {
      uint4 z;
      z.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]); z.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
      z.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); z.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
      my[g] = z;
    }
  }
  if (do_z) {
    __syncwarp();
    stage_copy<2, STG_STORE, NP>(stg, p.z_out, row_off0, ldc, nb, rows_valid, cols_valid, lane);
    __syncwarp();
  }
  if (do_gelu) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = gelu_fast(v[i]);
  }
  if (has_aux) {
#pragma unroll
    for (int g = 0; g < NP; ++g) {
      const uint32_t w[4] = {a[g].x, a[g].y, a[g].z, a[g].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = (aux_mode == AUX_MUL) ? unpack_f16(w[j]) : unpack_bf16(w[j]);
        if (aux_mode == AUX_ADD) {
          unpk2(add2(pk2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1]), pk2(t.x, t.y)), v[8 * g + 2 * j], v[8 * g + 2 * j + 1]);
        } else if (aux_mode == AUX_MUL) {
          unpk2(mul2(pk2(v[8 * g + 2 * j], v[8 * g + 2 * j + 1]), pk2(t.x, t.y)), v[8 * g + 2 * j], v[8 * g + 2 * j + 1]);
        } else {
          v[8 * g + 2 * j] *= gelu_grad_fast(t.x);
          v[8 * g + 2 * j + 1] *= gelu_grad_fast(t.y);
        }
      }
    }
  }
  if (c_dtype == OUT_BF16) {
#pragma unroll
    for (int g = 0; g < NP; ++g) {
      uint4 o;
      o.x = pack_bf16(v[8 * g + 0], v[8 * g + 1]); o.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
      o.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); o.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
      my[g] = o;
    }
    __syncwarp();
    if ((GEN || aux_mode == AUX_MUL) && cs_smem != nullptr) {
      // bias gradient of the producing layer, from the staged tile: lane l sums column l over the warp's 32 rows (rows of
      // the staging buffer are 80 B apart: a column's 32 reads touch 16 banks twice through the same words, no conflict)
      // and adds it to the CTA's per-tile partials in shared memory (cs_smem = this chunk's slice; the tile's total
      // goes to global memory once: per-warp global reductions on the same 12 KB serialised in the L2 atomic units)
      float cs = 0.f;
      if (lane < CW) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)
          if (rr < rows_valid)
            cs += __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(stg + rr * STG_PITCH + lane * 2)) << 16);
        atomicAdd(cs_smem + lane, cs);
      }
    }
    stage_copy<2, STG_STORE, NP>(stg, cptr, row_off0, ldc, nb, rows_valid, cols_valid, lane);
    __syncwarp();
  } else {  // fp32: plain stores, or vector reductions for split-K partial sums; 16 columns (64 B per row) at a time
#pragma unroll
    for (int h = 0; h < CW / 16; ++h) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        my[g] = make_uint4(__float_as_uint(v[16 * h + 4 * g]), __float_as_uint(v[16 * h + 4 * g + 1]),
                           __float_as_uint(v[16 * h + 4 * g + 2]), __float_as_uint(v[16 * h + 4 * g + 3]));
      __syncwarp();
      if (c_dtype == OUT_F32)
        stage_copy<4, STG_STORE, 4>(stg, cptr, row_off0, ldc, nb + 16 * h, rows_valid, cols_valid, lane);
      else
        stage_copy<4, STG_RED, 4>(stg, cptr, row_off0, ldc, nb + 16 * h, rows_valid, cols_valid, lane);
      __syncwarp();
    }
  }
}

// ---- 2-CTA mode (CL == 2, cta_group::2): the two CTAs of a cluster (one TPC) compute ONE 256 x BN tile.  CTA r
// stages A rows [128 r, 128 r + 128) and B rows [r BN/2, (r+1) BN/2) of the tile in its own shared memory; the leader
// (rank 0) issues tcgen05.mma.cta_group::2 with M = 256, which reads both CTAs' shared memory and writes each CTA's half
// of the accumulator into that CTA's own TMEM.  Per k-block every SM now reads 16 + 16 KB of operands from shared memory
// (instead of 16 + 32 KB single-CTA) and receives 32 KB from L2: the shared-memory port stops being the bound.
// Barriers: both producers' TMA loads complete on the LEADER's full barrier (the leader posts the expected bytes of
// both); the leader's tcgen05.commit is multicast to the empty / accumulator-full barriers of both CTAs; the peer's
// epilogue warps arrive remotely on the leader's accumulator-empty barrier.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x16(taddr, r); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Accumulator-empty arrivals are RELAXED: they only have to follow this warp's tcgen05.ld of the accumulator (ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync); a release arrive also waits until the warp's global stores of
// the tile are visible (ncu: ~10 % of the epilogue warps' samples sat on the ERRBAR that implements it).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA load whose completion bytes are posted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint32_t bar_cluster, uint32_t dst, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int MA, int MB, int BN, int CL, int EK, bool GROUP>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap* __restrict__ maps_a,
                                             const CUtensorMap* __restrict__ maps_b, const KParams& p,
                                             const GroupProb* __restrict__ probs, int n_prob) {
  using C = Cfg<BN, CL>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * C::A_BYTES;
  const uint32_t bars = sB + STAGES * C::B_BYTES;
  auto full_bar = [&](int i) { return bars + 8u * i; };
  auto empty_bar = [&](int i) { return bars + 8u * (STAGES + i); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * STAGES + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * STAGES + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 256u - raw_u32));  // [2 accumulator stages][BN]
  float* s_cs = s_bias + 2 * BN;                                                 // [2 tile parities][BN] column-sum partials
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_cs + 2 * BN);                  // [EPI_WARPS][32 rows][80 B]
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  trace_wall(p, 0);

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(maps_a);
      tma_prefetch_desc(maps_b);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(full_bar(i), 1);
        mbar_init(empty_bar(i), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(tfull_bar(i), 1);
        mbar_init(tempty_bar(i), CL * EPI_WARPS);  // lane 0 of every epilogue warp of every CTA of the pair
      }
      mbar_fence_init();
    }
  } else if (warp == 2) {
    if (CL == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    else tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);  // the same warp of BOTH CTAs issues the paired allocation
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();  // the next kernel's prologue may start; it waits for this grid before touching memory
  trace_wall(p, 1);
  pdl_wait();               // everything above overlapped the previous kernel's tail
  trace_wall(p, 2);
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int piter = 0;
      int gcur = 0;
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++piter) {
        const TileCoord t = GROUP ? decode_tile_group<CL>(p, probs, tile, rank, gcur) : decode_tile<CL>(p, tile, rank);
        const CUtensorMap& map_a = maps_a[t.g];
        const CUtensorMap& map_b = maps_b[t.g];
        const int m0 = t.mt * BLOCK_M, n0 = t.nt * BN;
        trace_ev(p, 0, piter, 0);
        int kin = t.kb_begin % p.k_inner, kbatch = t.kb_begin / p.k_inner;  // advanced incrementally: no division
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {                     // on the per-k-block issue path
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = sA + stage * C::A_BYTES;
          const uint32_t b_dst = sB + stage * C::B_BYTES;
          int cc[4];
          if (CL == 1) {
            mbar_expect_tx(full_bar(stage), C::A_BYTES + C::B_BYTES);
            if (MA == MAJOR_K) {
              op_coords(p.a, kin, kbatch, m0, t.lo, t.hi, cc);
              tma_load_4d(&map_a, full_bar(stage), a_dst, cc[0], cc[1], cc[2], cc[3]);
            } else {
#pragma unroll
              for (int at = 0; at < BLOCK_M / 64; ++at) {
                op_coords(p.a, kin, kbatch, m0 / 64 + at, t.lo, t.hi, cc);
                tma_load_4d(&map_a, full_bar(stage), a_dst + at * (BLOCK_K * 128), cc[0], cc[1], cc[2], cc[3]);
              }
            }
            if (MB == MAJOR_K) {
              op_coords(p.b, kin, kbatch, n0, t.lo, t.hi, cc);
              tma_load_4d(&map_b, full_bar(stage), b_dst, cc[0], cc[1], cc[2], cc[3]);
            } else {
#pragma unroll
              for (int at = 0; at < BN / 64; ++at) {
                op_coords(p.b, kin, kbatch, n0 / 64 + at, t.lo, t.hi, cc);
                tma_load_4d(&map_b, full_bar(stage), b_dst + at * (BLOCK_K * 128), cc[0], cc[1], cc[2], cc[3]);
              }
            }
          } else {
            // both CTAs' loads complete on the leader's barrier; the leader posts the byte count of the pair
            const uint32_t fb = mapa_rank(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (C::A_BYTES + C::B_BYTES));
            if (MA == MAJOR_K) {
              op_coords(p.a, kin, kbatch, m0, t.lo, t.hi, cc);
              tma_load_4d_2sm(&map_a, fb, a_dst, cc[0], cc[1], cc[2], cc[3]);
            } else {
#pragma unroll
              for (int at = 0; at < BLOCK_M / 64; ++at) {
                op_coords(p.a, kin, kbatch, m0 / 64 + at, t.lo, t.hi, cc);
                tma_load_4d_2sm(&map_a, fb, a_dst + at * (BLOCK_K * 128), cc[0], cc[1], cc[2], cc[3]);
              }
            }
            if (MB == MAJOR_K) {  // this CTA's half of the tile's B rows
              op_coords(p.b, kin, kbatch, n0 + rank * (BN / 2), t.lo, t.hi, cc);
              tma_load_4d_2sm(&map_b, fb, b_dst, cc[0], cc[1], cc[2], cc[3]);
            } else {
#pragma unroll
              for (int a2 = 0; a2 < BN / 128; ++a2) {
                op_coords(p.b, kin, kbatch, n0 / 64 + rank * (BN / 128) + a2, t.lo, t.hi, cc);
                tma_load_4d_2sm(&map_b, fb, b_dst + a2 * (BLOCK_K * 128), cc[0], cc[1], cc[2], cc[3]);
              }
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
          if (++kin == p.k_inner) {
            kin = 0;
            ++kbatch;
          }
        }
        trace_ev(p, 0, piter, 1);
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (rank == 0 && elect_one()) {  // pair mode: the leader CTA issues for both
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)MA << 15) |
                                 ((uint32_t)MB << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((BLOCK_M * CL) >> 4) << 24);
      // K-major: 8-row groups are 1024 B apart (SBO), one swizzle span along K (LBO unused).
      // MN-major: 64-element MN atoms are BLOCK_K*128 B apart (LBO), 8-row K groups 1024 B (SBO).
      constexpr uint32_t A_LBO = (MA == MAJOR_K) ? 0u : BLOCK_K * 128u;
      constexpr uint32_t B_LBO = (MB == MAJOR_K) ? 0u : BLOCK_K * 128u;
      constexpr uint32_t A_KADV = (MA == MAJOR_K) ? UMMA_K * 2u : UMMA_K * 128u;
      constexpr uint32_t B_KADV = (MB == MAJOR_K) ? UMMA_K * 2u : UMMA_K * 128u;
      // The per-k-block cost of this loop besides the MMAs bounds the short-K GEMMs (4 MMAs of 96-128 tensor-pipe clocks
      // per k-block): descriptors are not rebuilt from byte addresses for every MMA (shift, mask, or on the uniform
      // datapath) but carried as low words (address field in 16-byte units | LBO field) that advance by constants; the
      // high word (SBO 1024 B, version, 128B swizzle) is a constant.
      constexpr uint32_t DESC_HI = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      constexpr uint32_t A_LO0 = ((A_LBO >> 4) & 0x3FFFu) << 16, B_LO0 = ((B_LBO >> 4) & 0x3FFFu) << 16;
      auto desc = [](uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      int gcur = 0;
      uint32_t a_lo = (sA >> 4) | A_LO0, b_lo = (sB >> 4) | B_LO0, sbar = 0;  // of the current stage
      // (Probing the next stage's barrier with mbarrier.test_wait ahead of the MMAs was measured and is SLOWER: the thread
      // issues in order and stalls on the probe's result, ~145 clk, before the MMAs instead of after them.)
      for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++iter) {
        const TileCoord t = GROUP ? decode_tile_group<CL>(p, probs, tile, rank, gcur) : decode_tile<CL>(p, tile, rank);
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1u;
        trace_ev(p, 1, iter, 0);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        trace_ev(p, 1, iter, 1);
        const uint32_t d_tmem = tmem_base + as * BN;
        const int nkb = t.kb_end - t.kb_begin;
        for (int i = 0; i < nkb; ++i) {
          // TMA -> mbarrier -> tcgen05.mma needs no tcgen05 fence (both are the async proxy, the barrier orders them)
          mbar_wait(bars + sbar, phase);  // full_bar(stage)
          if (i == 0) trace_ev(p, 1, iter, 2);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (CL == 1) umma_bf16(d_tmem, desc(a_lo + k * (A_KADV >> 4)), desc(b_lo + k * (B_KADV >> 4)), idesc, (i > 0 || k > 0) ? 1u : 0u);
            else umma_bf16_2sm(d_tmem, desc(a_lo + k * (A_KADV >> 4)), desc(b_lo + k * (B_KADV >> 4)), idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (CL == 1) {
            umma_commit(bars + 8u * STAGES + sbar);  // empty_bar(stage)
            if (i == nkb - 1) umma_commit(tfull_bar(as));
          } else {  // the stage is free / the accumulator is ready in BOTH CTAs
            umma_commit_2sm(bars + 8u * STAGES + sbar, 3);
            if (i == nkb - 1) umma_commit_2sm(tfull_bar(as), 3);
          }
          if (i == nkb - 1) trace_ev(p, 1, iter, 3);
          a_lo += C::A_BYTES >> 4;
          b_lo += C::B_BYTES >> 4;
          sbar += 8u;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
            a_lo = (sA >> 4) | A_LO0;
            b_lo = (sB >> 4) | B_LO0;
            sbar = 0;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    // warp w may only touch TMEM lanes 32*(w%4)..+31; the two warps of a lane quarter split the tile's columns.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;    // which 1/EPI_PARTS of the tile's columns
    constexpr int COLS = BN / EPI_PARTS;  // columns per warp
    constexpr int NCH = COLS / EPI_CW;    // chunks per warp
    const int tid_e = threadIdx.x - 128;
    int iter = 0;
    int gcur = 0;
    // fused column sums (a8_gemm_t::colsum): compiled into the epilogues that can carry them only
    constexpr bool CS_KIND = !GROUP && ((EK < 0) || (((EK >> 5) & 3) == AUX_MUL));
    const bool do_cs = CS_KIND && p.colsum != nullptr;
    if (do_cs) {
      for (int i = tid_e; i < 2 * BN; i += 32 * EPI_WARPS) s_cs[i] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    }
    for (int tile = tile0; tile < p.total_tiles; tile += tile_step, ++iter) {
      const TileCoord t = GROUP ? decode_tile_group<CL>(p, probs, tile, rank, gcur) : decode_tile<CL>(p, tile, rank);
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1u;
      float* sb = nullptr;
      if (p.bias != nullptr) {
        // stage this tile's bias slice in shared memory while the main loop of the tile is still running
        sb = s_bias + as * BN;
        const float* bsrc = p.bias + (long long)t.lo * p.bias_stride_lo + t.nt * BN;
        for (int i = tid_e; i < BN; i += 32 * EPI_WARPS) sb[i] = (t.nt * BN + i < t.N) ? __ldg(bsrc + i) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }
      const int row0 = t.mt * BLOCK_M + q * 32;
      const long long row_off0 =
          (long long)t.hi * p.c_stride_hi + (long long)t.lo * p.c_stride_lo + (long long)row0 * t.ldc;
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * COLS;
      const int nb0 = t.nt * BN + half * COLS;
      // the first chunk's aux tile is requested before the wait for the accumulator
      uint4 aux_pre[EPI_CW / 8];
      if (((EK < 0) ? p.aux_mode : ((EK >> 5) & 3)) != AUX_NONE && nb0 < t.N)
        aux_issue<EPI_CW / 8>(aux_pre, p.aux, row_off0, t.ldc, nb0, min(32, t.M - row0), (t.N + 7) & ~7, lane);
      if (warp == 4 && lane == 0) trace_ev(p, 2, iter, 0);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (warp == 4 && lane == 0) trace_ev(p, 2, iter, 1);
      const float* sbw = sb ? sb + half * COLS : nullptr;
      uint8_t* stg = s_stage + (warp - 4) * STG_BYTES;
      // one chunk at a time, NOT unrolled: the chunk body is several hundred instructions and the epilogue warps must
      // stay inside the instruction cache (the other epilogue warps hide this warp's tcgen05.ld latency)
      uint32_t ra[EPI_CW];
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        if (nb0 + c * EPI_CW >= t.N) break;
        tmem_ld_chunk(t_addr + c * EPI_CW, ra);
        tmem_ld_wait();
        const int nb_next = (c + 1 < NCH && nb0 + (c + 1) * EPI_CW < t.N) ? nb0 + (c + 1) * EPI_CW : -1;
        epilogue_chunk<EK, EPI_CW>(p, t, ra, row_off0, row0, nb0 + c * EPI_CW, nb_next, aux_pre,
                                   sbw ? sbw + c * EPI_CW : nullptr, stg, lane,
                                   do_cs ? s_cs + as * BN + half * COLS + c * EPI_CW : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 4 && lane == 0) trace_ev(p, 2, iter, 2);
      if (lane == 0) {
        if (CL == 1) mbar_arrive_relaxed(tempty_bar(as));
        else mbar_arrive_cluster(mapa_rank(tempty_bar(as), 0));  // the leader's MMA thread waits for both CTAs
      }
      if (do_cs) {
        // the tile's column totals leave the CTA once all epilogue warps have added theirs; the buffer of this parity is
        // reused two tiles later, after the next tile's barrier
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        for (int i = tid_e; i < BN; i += 32 * EPI_WARPS) {
          const float v = s_cs[as * BN + i];
          s_cs[as * BN + i] = 0.f;
          if (t.nt * BN + i < t.N) atomicAdd(p.colsum + t.nt * BN + i, v);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // neither CTA exits while the peer may still write its smem or signal its barriers
  trace_wall(p, 3);
  if (warp == 2) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
    else tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

template <int MA, int MB, int BN, int CL, int EK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const KParams p) {
  gemm_tc_body<MA, MB, BN, CL, EK, false>(&map_a, &map_b, p, nullptr, 1);
}

template <int MA, int MB, int BN, int CL, int EK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_group_kernel(const __grid_constant__ GroupParams gp, const KParams p) {
  gemm_tc_body<MA, MB, BN, CL, EK, true>(gp.map_a, gp.map_b, p, gp.prob, gp.n_prob);
}

// ---------------------------------------------------------------------------------------------
// launch (instantiated per (MA, MB) in gemm_tc_inst_*.cu so that the translation units build in parallel)
// ---------------------------------------------------------------------------------------------
int num_sms();  // gemm_tc.cu

template <int MA, int MB, int BN, int CL, int EK>
int launch_inst(const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_tc_kernel<MA, MB, BN, CL, EK>;
  if (!configured) {
    A8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg<BN, CL>::SMEM_BYTES));
    configured = true;
  }
  const int slots = num_sms() / CL;
  const int grid = CL * (kp.total_tiles < slots ? kp.total_tiles : slots);
  A8_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg<BN, CL>::SMEM_BYTES, stream, CL, ma, mb, kp));
  return check_launch("gemm_tc_kernel");
}

template <int MA, int MB, int EK>
int launch_bn(int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp,
              cudaStream_t stream) {
  if (cl == 2) {
    switch (bn) {
      case 128: return launch_inst<MA, MB, 128, 2, EK>(ma, mb, kp, stream);
      case 256: return launch_inst<MA, MB, 256, 2, EK>(ma, mb, kp, stream);
      case 192:  // each CTA stages 96 B rows: only expressible for a K-major B (MN-major atoms are 64 wide)
        if constexpr (MB == MAJOR_K) return launch_inst<MA, MB, 192, 2, EK>(ma, mb, kp, stream);
        break;
    }
    set_error("gemm: pair mode needs block_n 128 or 256 (or 192 with a K-major B), got %d", bn);
    return -1;
  }
  switch (bn) {
    case 64: return launch_inst<MA, MB, 64, 1, EK>(ma, mb, kp, stream);
    case 128: return launch_inst<MA, MB, 128, 1, EK>(ma, mb, kp, stream);
    case 192: return launch_inst<MA, MB, 192, 1, EK>(ma, mb, kp, stream);
    case 256: return launch_inst<MA, MB, 256, 1, EK>(ma, mb, kp, stream);
  }
  set_error("gemm: unsupported block_n %d", bn);
  return -1;
}

template <int MA, int MB, int BN, int CL, int EK>
int launch_group_inst(const GroupParams& gp, const KParams& kp, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_tc_group_kernel<MA, MB, BN, CL, EK>;
  if (!configured) {
    A8_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN, CL>::SMEM_BYTES));
    configured = true;
  }
  const int slots = num_sms() / CL;
  const int grid = CL * (kp.total_tiles < slots ? kp.total_tiles : slots);
  A8_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg<BN, CL>::SMEM_BYTES, stream, CL, gp, kp));
  return check_launch("gemm_tc_group_kernel");
}
int launch_group_mnmn(int ek, int bn, int cl, const GroupParams& gp, const KParams& kp, cudaStream_t s);  // gemm_tc_inst_mnmn.cu

// one entry per (MA, MB): picks the specialised epilogue when (c_dtype, act, z_out, aux_mode) is on its list
int launch_kk(int ek, int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t s);
int launch_kmn(int ek, int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t s);
int launch_mnmn(int ek, int bn, int cl, const CUtensorMap& ma, const CUtensorMap& mb, const KParams& kp, cudaStream_t s);

}  // namespace gemm
}  // namespace a8
