// micro-benchmark: issue rate of tcgen05.mma (cta_group::1, M = 128) from shared-memory operands for kind::tf32 / kind::f16,
// N = 32..256, swizzle-128B vs no-swizzle K-major layouts, same vs alternating operand buffers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ uint64_t desc_ns(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | (8ull << 16) | (64ull << 32) | (1ull << 46); }
__global__ void __launch_bounds__(128, 1) k(int N, int kind, int sw, int n_mma, int alt, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ unsigned long long bar; __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(su32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(su32(&tmem_base)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t fmt = kind == 0 ? 2u : 0u;     // tf32 : f16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long c0 = clock64();
    for (int i = 0; i < n_mma; i++) {
      const uint32_t a = su32(sm) + (alt ? (i % 3) * 16384u : 0u) + (uint32_t)(i & 3) * 32u * (sw ? 1u : 8u);
      const uint32_t b = su32(sm) + 65536u + (alt ? (i % 3) * 32768u : 0u) + (uint32_t)(i & 3) * 32u * (sw ? 1u : 8u);
      const uint64_t da = sw ? desc_sw(a) : desc_ns(a), db = sw ? desc_sw(b) : desc_ns(b);
      const uint32_t accum = i > 0;
      if (kind == 0) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" :: "r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" :: "r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
    }
    long long c1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(su32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(su32(&bar)) : "memory");
    long long c2 = clock64();
    out[blockIdx.x * 2] = (unsigned long long)(c2 - c0); out[blockIdx.x * 2 + 1] = (unsigned long long)(c1 - c0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(256u) : "memory"); }
}
int main() {
  unsigned long long* out; cudaMalloc(&out, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
  const int n_mma = 480;
  for (int grid : {1, 148})
    for (int kind : {0, 1})
      for (int sw : {1, 0})
        for (int alt : {0, 1})
          for (int N : {32, 64, 128, 256}) {
            k<<<grid, 128, 164 * 1024>>>(N, kind, sw, n_mma, alt, out);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("grid %3d kind %s layout %s operands %s N %3d: %7.1f cycles per MMA (issue %6.1f) %s\n", grid, kind == 0 ? "tf32" : "f16 ", sw ? "sw128" : "nosw ", alt ? "3 buffers" : "same     ", N,
                   (double)h[0] / n_mma, (double)h[1] / n_mma, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  return 0;
}
