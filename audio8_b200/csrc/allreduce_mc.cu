// audio8_b200 — gradient all-reduce through the NVSwitch: the data-parallel exchange of
// /root/reference/audio8/pretrain.py:153 (DistributedDataParallel's bucketed NCCL all-reduce), re-done as ONE in-place
// pass over the gradient arena using the switch's multicast + in-fabric reduction (NVLS):
//
//   every rank owns 1/world of the range; for its part it issues multimem.ld_reduce on the MULTICAST address (the switch
//   fetches the 16 bytes from every GPU's copy, adds them, returns one value), scales by 1/world (DDP averages) and
//   multimem.st's the result back to the multicast address (the switch writes it into every GPU's copy).
//
// Per GPU and direction the NVLink traffic is 1 x the range (7/8 out + 1/8 in for the reduce, 1/8 out + 7/8 in for the
// broadcast) against 1.75 x for a ring at world = 8, no intermediate buffers, no reduction arithmetic on the SMs, and
// every rank ends with bit-identical values (one rank reduces each element, everyone receives that result).
// The cross-rank ordering (all gradients written before the first ld_reduce, all stores landed before anyone reads) is
// the caller's: parallel.py brackets the launch with the symmetric-memory barrier on the same stream.
// HBM/NVLink-bound: algorithmic bytes per element 4 B read x world (in the fabric) + 4 B written x world.
#include "a8_common.cuh"
#include "../../include/audio8_b200.h"

namespace a8 {
namespace {

constexpr int MC_THREADS = 512;
constexpr int MC_UNROLL = 4;

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void mc_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// vec4 indices [v0, v1) of the multicast window; MC_UNROLL independent 16-byte reductions in flight per thread
__global__ void __launch_bounds__(MC_THREADS)
allreduce_mc_kernel(float* __restrict__ mc, long long v0, long long v1, float scale) {
  const long long stride = (long long)gridDim.x * MC_THREADS;
  for (long long i = v0 + (long long)blockIdx.x * MC_THREADS + threadIdx.x; i < v1; i += stride * MC_UNROLL) {
    float4 v[MC_UNROLL];
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < v1) v[u] = mc_ld_reduce(mc + 4 * j);
    }
#pragma unroll
    for (int u = 0; u < MC_UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < v1) {
        v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
        mc_st(mc + 4 * j, v[u]);
      }
    }
  }
}

}  // namespace
}  // namespace a8

using namespace a8;

extern "C" int a8_allreduce_mc(void* multicast_base, int64_t begin, int64_t end, int32_t rank, int32_t world, float scale,
                               int32_t ctas, void* stream_v) {
  A8_REQUIRE(multicast_base != nullptr, "allreduce_mc: no multicast mapping (the fabric or the driver does not offer NVLS)");
  A8_REQUIRE(world >= 1 && rank >= 0 && rank < world, "allreduce_mc: rank %d of %d", rank, world);
  A8_REQUIRE(begin >= 0 && end >= begin && begin % 4 == 0 && end % 4 == 0 &&
             (reinterpret_cast<uintptr_t>(multicast_base) & 15u) == 0,
             "allreduce_mc: the range must be 16-byte aligned (begin %lld, end %lld)", (long long)begin, (long long)end);
  const long long nv = (end - begin) / 4;
  if (nv == 0) return 0;
  // contiguous vec4 slice of this rank
  const long long per = (nv + world - 1) / world;
  long long v0 = begin / 4 + per * rank, v1 = v0 + per;
  const long long vend = end / 4;
  if (v0 > vend) v0 = vend;
  if (v1 > vend) v1 = vend;
  if (v1 <= v0) return 0;
  if (ctas <= 0) ctas = 24;
  const long long need = (v1 - v0 + (long long)MC_THREADS * MC_UNROLL - 1) / ((long long)MC_THREADS * MC_UNROLL);
  const int grid = (int)(need < ctas ? need : ctas);
  allreduce_mc_kernel<<<grid, MC_THREADS, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<float*>(multicast_base), v0, v1, scale);
  return check_launch("allreduce_mc_kernel");
}
