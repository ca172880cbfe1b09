// audio8_b200 — shared device/host helpers for the sm_100a kernels.
// Everything here is hand-written PTX wrappers (mbarrier, TMA, tcgen05/TMEM) plus
// small math helpers shared by the memory-bound kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

namespace a8 {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local last error string, returned through a8_last_error())
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> 0 / negative code

#define A8_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      a8::set_error(__VA_ARGS__);             \
      return -1;                              \
    }                                         \
  } while (0)

#define A8_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      a8::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return -2;                                                             \
    }                                                                        \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining,
// so launch latency, the prologue (barrier init, TMEM allocation) and the predecessor's tail overlap.  Every kernel
// launched through launch_pdl() calls pdl_wait() before it touches global memory.  A8_PDL=0 turns the attribute off.
bool pdl_enabled();  // a8_api.cu
#if defined(__CUDACC__)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
// d/dx [0.5 x (1+erf(x/sqrt2))] = Phi(x) + x*phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): 1 rcp + 1 ex2 + 7 FMA.  Used where the result is stored
// in bf16 (8 mantissa bits): exact for the purpose and ~3x cheaper than erff().
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.f - p * t * ex2_approx(-1.4426950408889634f * ax * ax);
  return copysignf(e, x);
}
// exact-erf GELU with |error| < 1e-6 and ONE special-function op (A&S 7.1.28: 1-erf(u) = (1+a1 u+..+a6 u^6)^-16)
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(0.0000430638f, u, 0.0002765672f);
  p = fmaf(p, u, 0.0001520143f);
  p = fmaf(p, u, 0.0092705272f);
  p = fmaf(p, u, 0.0422820123f);
  p = fmaf(p, u, 0.0705230784f);
  p = fmaf(p, u, 1.f);
  p *= p; p *= p; p *= p; p *= p;              // overflows to +inf for |x| > ~25: rcp(inf) = 0, the right limit
  const float h = 0.5f * rcp_approx(p);        // (1 - erf(u)) / 2
  return x * (x >= 0.f ? 1.f - h : h);
}
// d/dx gelu(x) = Phi(x) + x phi(x), |error| < 1e-6: 1 rcp + 1 ex2, the exponential shared by erf (7.1.26) and phi
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, u, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-0.72134752044448170f * x * x);  // exp(-x^2/2)
  const float half_erf = 0.5f - 0.5f * p * t * e;
  const float cdf = 0.5f + copysignf(half_erf, x);
  return fmaf(x * 0.3989422804014327f, e, cdf);
}
// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 carry two values per instruction: half the issue slots, the same
// FP32 throughput).  A pair lives in a 64-bit register; ptxas keeps it in two adjacent 32-bit registers, so packing
// values that are produced next to each other costs nothing, and broadcast constants become immediates.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 bc2(float c) { return pk2(c, c); }
// gelu(x) and gelu'(x) of two values at once, with ONE special-function op per value.  The GEMM epilogue that applies
// GELU and stores its derivative runs at the rate of its narrowest pipe, and that is the 16-lane MUFU: the reciprocal of
// the rational form (7.1.26) is replaced by a polynomial.  With e = exp(-x^2/2):
//   Phi(-|x|) = e * w(|x|),  w(a) = erfcx(a / sqrt2) / 2   (smooth, 0.5 at 0, ~ 1 / (a sqrt(2 pi)) for large a)
// w is a degree-8 polynomial on [0, 8] fitted under the weight e (beyond 8, e < 2e-14 and the clamped argument keeps
// the product finite): |Phi error| < 1.2e-6, |gelu error| < 5.3e-6 up to |x| = 8, |gelu' error| < 1.3e-6 (fit and
// float32 check: DESIGN.md section 3).  The coefficients carry the minus sign of 0.5 - e w.
__device__ __forceinline__ void gelu_both_fast2(float& x0, float& x1, float& d0, float& d1) {
  const f32x2 X = pk2(x0, x1), A = pk2(fminf(fabsf(x0), 8.f), fminf(fabsf(x1), 8.f));
  f32x2 w = fma2(bc2(-3.910695158992894e-05f), A, bc2(0.0006307236035354435f));
  w = fma2(w, A, bc2(-0.004477267153561115f));
  w = fma2(w, A, bc2(0.01902402751147747f));
  w = fma2(w, A, bc2(-0.05642680823802948f));
  w = fma2(w, A, bc2(0.13017946481704712f));
  w = fma2(w, A, bc2(-0.24935509264469147f));
  w = fma2(w, A, bc2(0.3988867998123169f));
  w = fma2(w, A, bc2(-0.4999993145465851f));
  float a0, a1;
  unpk2(mul2(mul2(X, X), bc2(-0.72134752044448170f)), a0, a1);
  const f32x2 E = pk2(ex2_approx(a0), ex2_approx(a1));  // exp(-x^2/2)
  float h0, h1;
  unpk2(fma2(w, E, bc2(0.5f)), h0, h1);                 // 0.5 - Phi(-|x|) = erf(|x|/sqrt2) / 2
  const f32x2 CDF = add2(bc2(0.5f), pk2(copysignf(h0, x0), copysignf(h1, x1)));
  unpk2(fma2(mul2(X, bc2(0.3989422804014327f)), E, CDF), d0, d1);
  unpk2(mul2(X, CDF), x0, x1);
}
// gelu_fast / gelu_grad_fast on a pair (same formulas, same constants)
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 X) {
  float x0, x1;
  unpk2(X, x0, x1);
  const f32x2 U = pk2(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
  f32x2 p = fma2(bc2(0.0000430638f), U, bc2(0.0002765672f));
  p = fma2(p, U, bc2(0.0001520143f));
  p = fma2(p, U, bc2(0.0092705272f));
  p = fma2(p, U, bc2(0.0422820123f));
  p = fma2(p, U, bc2(0.0705230784f));
  p = fma2(p, U, bc2(1.f));
  p = mul2(p, p); p = mul2(p, p); p = mul2(p, p); p = mul2(p, p);
  float p0, p1;
  unpk2(p, p0, p1);
  float h0, h1;  // erf(|x|/sqrt2) / 2 = 0.5 - 0.5 / p
  unpk2(fma2(pk2(rcp_approx(p0), rcp_approx(p1)), bc2(-0.5f), bc2(0.5f)), h0, h1);
  return mul2(X, add2(bc2(0.5f), pk2(copysignf(h0, x0), copysignf(h1, x1))));
}
__device__ __forceinline__ f32x2 gelu_grad_fast2(f32x2 X) {
  float x0, x1;
  unpk2(X, x0, x1);
  const f32x2 T = pk2(rcp_approx(fmaf(0.23164188826636045f, fabsf(x0), 1.f)),
                      rcp_approx(fmaf(0.23164188826636045f, fabsf(x1), 1.f)));
  f32x2 p = fma2(bc2(-0.5307027145f), T, bc2(0.7265760135f));
  p = fma2(p, T, bc2(-0.7107068705f));
  p = fma2(p, T, bc2(0.142248368f));
  p = fma2(p, T, bc2(-0.127414796f));
  float a0, a1;
  unpk2(mul2(mul2(X, X), bc2(-0.72134752044448170f)), a0, a1);
  const f32x2 E = pk2(ex2_approx(a0), ex2_approx(a1));
  float h0, h1;
  unpk2(fma2(mul2(p, T), E, bc2(0.5f)), h0, h1);
  const f32x2 CDF = add2(bc2(0.5f), pk2(copysignf(h0, x0), copysignf(h1, x1)));
  return fma2(mul2(X, bc2(0.3989422804014327f)), E, CDF);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// per-step dropout seed word in device memory (a8_set_seed_source), nullable
__device__ __forceinline__ unsigned long long seed_base_ld(const unsigned long long* src) {
  return src != nullptr ? __ldg(src) : 0ull;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// IEEE fp16 pairs: the stored gelu'(z) factors (values in [-0.13, 1.13]: 11 mantissa bits instead of bf16's 8)
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// block-wide sum using one smem array of 32 floats; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// ---------------------------------------------------------------------------------------------
// PTX: shared-address helpers, mbarrier, TMA, tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (kernel aborts with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  bool timing = false;
  while (true) {
    // the suspend-time hint lets the hardware park the warp until the phase completes instead of returning after
    // its short default interval: waiting warps re-issued the try_wait / branch / clock sequence ~9 times per wait
    // (15 % of the issue slots of the attention kernels) next to the math warps of the same scheduler
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) break;
    if (!timing) {
      t0 = clock64();
      timing = true;
    } else if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz: never legitimate
      printf("a8: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// a kernel parameter copied into a register the compiler cannot rematerialise from the constant bank inside hot loops
__device__ __forceinline__ int opaque(int v) {
  int r;
  asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// TMA tiled load, 4-D tensor map, completes on an mbarrier of this CTA
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp executes
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace a8
