// micro-benchmarks of the primitives the persistent kernels are built from (B200): build with nvcc, run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int ld_acq(const unsigned int* p) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) { unsigned int v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// (a) grid barrier variants, 148 CTAs x 256 threads, N iterations
__global__ void k_bar(unsigned int* cnt, int iters, int mode, unsigned long long* out, float* scratch) {
  unsigned int epoch = 0;
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = 0;
  for (int i = 0; i < iters; i++) {
    if (i == 10 && blockIdx.x == 0 && threadIdx.x == 0) t0 = gtime();
    if (mode == 3) scratch[(size_t)blockIdx.x * 4096 + threadIdx.x + (i & 7) * 256] = (float)i;   // some stores before the barrier
    if (mode == 0) { grid.sync(); continue; }
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(cnt) : "memory");
      const unsigned int target = epoch * gridDim.x;
      if (mode == 2) { while (ld_relaxed(cnt) < target) { } __threadfence(); }
      else { while (ld_acq(cnt) < target) { } }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (gtime() - t0);
}
// (b) dependent load chain latency: pointer chase through a buffer of given size
__global__ void k_chase(const unsigned int* buf, int iters, unsigned long long* out) {
  unsigned int p = 0;
  unsigned long long t0 = gtime();
  for (int i = 0; i < iters; i++) p = buf[p];
  unsigned long long t1 = gtime();
  out[0] = t1 - t0; out[1] = p;
}
// (c) producer/consumer flag latency between two CTAs on different SMs (ping-pong)
__global__ void k_pingpong(unsigned int* flags, int iters, unsigned long long* out) {
  unsigned long long t0 = gtime();
  for (int i = 1; i <= iters; i++) {
    if (blockIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(flags) : "memory");
      while (ld_acq(flags + 32) < (unsigned)i) { }
    } else {
      while (ld_acq(flags) < (unsigned)i) { }
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(flags + 32) : "memory");
    }
  }
  if (blockIdx.x == 0) out[0] = gtime() - t0;
}
int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int nsm = p.multiProcessorCount;
  unsigned int* cnt; unsigned long long* out; float* scratch;
  cudaMalloc(&cnt, 4096); cudaMalloc(&out, 64); cudaMalloc(&scratch, (size_t)nsm * 4096 * 4);
  unsigned long long h[2];
  const char* names[] = {"cg::grid.sync", "release-add + acquire-poll", "release-add + relaxed-poll + fence", "as 1 with 256 stores/CTA before"};
  for (int mode = 0; mode < 4; mode++) {
    int iters = 2010;
    cudaMemset(cnt, 0, 4096);
    void* args[] = {&cnt, &iters, &mode, &out, &scratch};
    cudaLaunchCooperativeKernel((void*)k_bar, dim3(nsm), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("grid barrier [%s] %d CTAs: %.3f us each (%s)\n", names[mode], nsm, h[0] / 2000.0 / 1000.0, cudaGetErrorString(cudaGetLastError()));
  }
  for (int grid = 8; grid <= 64; grid *= 2) {
    int iters = 2010, mode = 1;
    cudaMemset(cnt, 0, 4096);
    void* args[] = {&cnt, &iters, &mode, &out, &scratch};
    cudaLaunchCooperativeKernel((void*)k_bar, dim3(grid), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("grid barrier [release/acquire] %d CTAs: %.3f us each\n", grid, h[0] / 2000.0 / 1000.0);
  }
  // pointer chase: 1 MB (L2 resident) and 1 GB stride pattern (DRAM)
  for (size_t bytes : {(size_t)1 << 20, (size_t)1 << 30}) {
    size_t n = bytes / 4;
    unsigned int* hb = (unsigned int*)malloc(bytes);
    size_t stride = (bytes == (size_t)1 << 20) ? 1031 : 4099 * 1021;
    for (size_t i = 0; i < n; i++) hb[i] = (unsigned int)((i + stride) % n);
    unsigned int* db; cudaMalloc(&db, bytes); cudaMemcpy(db, hb, bytes, cudaMemcpyHostToDevice);
    k_chase<<<1, 1>>>(db, 2000, out); k_chase<<<1, 1>>>(db, 2000, out);
    cudaDeviceSynchronize(); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("dependent load chain over %zu MB: %.1f ns per load\n", bytes >> 20, h[0] / 2000.0);
    cudaFree(db); free(hb);
  }
  cudaMemset(cnt, 0, 4096);
  k_pingpong<<<2, 32>>>(cnt, 2000, out);
  cudaDeviceSynchronize(); cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  printf("flag ping-pong between two CTAs: %.1f ns round trip (%s)\n", h[0] / 2000.0, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
