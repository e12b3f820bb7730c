// micro-benchmark: throughput of cp.async.bulk (1-D, global -> shared, mbarrier complete_tx) per SM as a function of the copy size,
// the number of copies in flight and the number of CTAs pulling at once.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a bulkcopy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const unsigned char* src, size_t src_bytes, int copy_bytes, int n_stage, int iters, int pieces, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ unsigned long long bar[8];
  if (threadIdx.x == 0) { for (int i = 0; i < n_stage; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(su32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    size_t off = ((size_t)blockIdx.x * 7919u * 4096u) % (src_bytes - (size_t)copy_bytes * 2);
    off &= ~(size_t)1023;
    for (int it = 0; it < iters + n_stage; it++) {
      const int st = it % n_stage, use = it / n_stage;
      if (use > 0) {   // wait for the previous copy into this stage
        uint32_t ok = 0; const uint32_t par = (use - 1) & 1;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(su32(&bar[st])), "r"(par) : "memory");
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(su32(&bar[st])), "r"(copy_bytes) : "memory");
        const int pb = copy_bytes / pieces;
        for (int p = 0; p < pieces; p++)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(su32(sm + (size_t)st * copy_bytes + (size_t)p * pb)), "l"(src + off + (size_t)p * pb), "r"(pb), "r"(su32(&bar[st])) : "memory");
        off += copy_bytes; if (off + copy_bytes > src_bytes) off = 0;
      }
    }
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[blockIdx.x] = t1 - t0;
  }
}
int main() {
  const size_t SRC = 64u << 20;
  unsigned char* src; cudaMalloc(&src, SRC); cudaMemset(src, 1, SRC);
  unsigned long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 200;
  for (int grid : {1, 32, 148})
    for (int kb : {4, 16, 32, 64, 96})
      for (int stages : {2, 3})
        for (int pieces : {1, 4, 16}) {
          if (kb * stages > 196 || (kb * 1024 / pieces) % 16) continue;
          k<<<grid, 128, kb * stages * 1024>>>(src, SRC, kb * 1024, stages, iters, pieces, out);
          cudaError_t e = cudaDeviceSynchronize();
          unsigned long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
          unsigned long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
          printf("grid %3d copy %3d KB stages %d pieces %2d: %8.2f us per copy, %7.1f GB/s per SM, %8.1f GB/s total %s\n", grid, kb, stages, pieces, mx / 1000.0 / iters,
                 (double)kb * 1024 * iters / mx, (double)kb * 1024 * iters / mx * grid, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
