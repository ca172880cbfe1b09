// Cycles of the synchronisation primitives the warp-specialised kernels use, measured with clock64 by one thread of an
// otherwise idle SM (nvcc -gencode arch=compute_100a,code=sm_100a -o sync_cost sync_cost.cu && ./sync_cost).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WHAT>
__global__ void k(long long* out) {
  __shared__ __align__(8) unsigned long long bar[4];
  __shared__ uint32_t tslot;
  const uint32_t b = smem_u32(&bar[0]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b + 8));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b));  // phase 0 of bar[0] complete
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      uint32_t done = 0;
      if (WHAT == 0) {  // empty loop: clock + loop overhead
      } else if (WHAT == 1) {
        asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
      } else if (WHAT == 2) {
        asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3; selp.u32 %0, 1, 0, P;}" : "=r"(done) : "r"(b), "r"(0u), "r"(0x989680u) : "memory");
      } else if (WHAT == 3) {
        asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
      } else if (WHAT == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      } else if (WHAT == 5) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else if (WHAT == 6) {  // commit with nothing outstanding, to a barrier nobody waits on
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b + 8) : "memory");
      } else if (WHAT == 7) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b + 8) : "memory");
      } else if (WHAT == 8) {  // test + dependent use after 40 independent integer ops
        asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P;}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
      } else if (WHAT == 9) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else if (WHAT == 10) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      acc += done;
    }
    long long t1 = clock64();
    out[0] = t1 - t0;
    out[1] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tslot));
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* d;
  cudaMalloc(&d, 16);
  const char* names[] = {"empty loop", "mbarrier.try_wait (complete phase)", "mbarrier.try_wait + suspend hint", "mbarrier.test_wait",
                         "tcgen05.fence::after_thread_sync", "tcgen05.fence::before_thread_sync", "tcgen05.commit (nothing outstanding)",
                         "mbarrier.arrive", "mbarrier.test_wait (same)", "tcgen05.wait::ld (nothing outstanding)", "tcgen05.wait::st (nothing outstanding)"};
  long long h[2];
#define RUN(W)                                              \
  k<W><<<1, 128>>>(d);                                      \
  cudaDeviceSynchronize();                                  \
  k<W><<<1, 128>>>(d);                                      \
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: error %s\n", names[W], cudaGetErrorString(cudaGetLastError())); return 1; } \
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);             \
  printf("%-42s %6.1f cycles per call\n", names[W], h[0] / 64.0);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(7) RUN(6)
  return 0;
}
