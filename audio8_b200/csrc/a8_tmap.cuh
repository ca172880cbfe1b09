// audio8_b200 — host-side helper shared by the TMA-fed kernels: encode a bf16 tiled tensor map (128B swizzle).
#pragma once
#include "a8_common.cuh"

namespace a8 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tmap_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER entry point: it fails with CUDA_ERROR_INVALID_CONTEXT on a host thread that has
// not touched the runtime yet (autograd's worker thread on its first backward).  Bind the primary context once.
inline void tmap_ensure_context() {
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
    cudaFree(nullptr);
    bound = true;
  }
}

// 3-D bf16 view {d0 (contiguous), d1, d2} with element strides s1, s2 and a {box0, box1, 1} box.
inline int make_tmap_3d(CUtensorMap* out, const void* ptr, long long d0, long long d1, long long d2, long long s1,
                        long long s2, int box0, int box1, const char* what) {
  tmap_ensure_context();
  EncodeTiledFn fn = tmap_encode_fn();
  A8_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  A8_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "%s: pointer not 16B aligned", what);
  A8_REQUIRE(s1 % 8 == 0 && s2 % 8 == 0 && s1 > 0 && s2 > 0, "%s: strides must be positive multiples of 8", what);
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1 * 2ull, (cuuint64_t)s2 * 2ull};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  A8_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return 0;
}

#if defined(__CUDACC__)
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem, packed bf16 pairs per column] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 registers per thread -> 16 consecutive TMEM columns of the thread's lane
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// UMMA shared-memory matrix descriptor, 128B swizzle (layout type 2), descriptor version 1
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// The same descriptor split into its variable low word (address field, 16-byte units | leading-dimension field) and its
// constant high word (SBO 1024 B, version 1, 128B swizzle): MMA-issuing threads advance the low word by constants instead
// of rebuilding the descriptor from a byte address for every tcgen05.mma (shift, mask, or: three dependent uniform-
// datapath instructions per operand, which is what bounded the issue rate of small MMAs).
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t addr, uint32_t lbo_bytes) {
  return ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t umma_desc_sbo1024(uint32_t lo) {
  constexpr uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
#endif

}  // namespace a8
