// g4r_persistent.cuh -- the whole window of mini-batches in ONE cooperative kernel: every CTA loops over the
// steps and over the phases of g4r_kernels.cuh, separated by a grid-wide barrier.  No host round trip, no kernel
// launch latency between the dependent phases of a step (the per-phase kernels spend most of their ~15 us each on
// launch + first-touch latency at the headline shape).  Included from g4r_lib.cu after the MD macro.
#pragma once


__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Grid barrier on a monotonically increasing arrival counter: the k-th barrier completes when the counter reaches
// k * nblocks.  One release-add and acquire-polls per CTA; no reset, no separate fences.
__device__ __forceinline__ void grid_barrier(GridBar* gb, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    const unsigned int target = epoch * nblocks;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(&gb->count) : "memory");
    while (ld_acquire_u32(&gb->count) < target) { }
  }
  __syncthreads();
}

constexpr int PK_THREADS = 256;

__global__ void __launch_bounds__(PK_THREADS, 1) k_persistent(int slot, int n_steps, GridBar* gb, unsigned long long* tstamp) {
  extern __shared__ __align__(16) float smem[];
  const ModelDev& md = MD;
  float* sA = smem;
  float* sB = smem + GK * (GB + 1);
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int B = md.B;
  unsigned int epoch = 0;
#define PK_STAMP(k) do { if (tstamp && cta == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); tstamp[(size_t)s * 16 + (k)] = t_; } } while (0)
  for (int s = 0; s < n_steps; s++) {
    PK_STAMP(0);
    if (md.mode != 0) { phase_gather_in(md, s, true, cta, ncta); grid_barrier(gb, ncta, epoch); }
    for (int li = 0; li < md.n_layers; li++) {
      const LayerDev& ly = md.layer[li];
      const int n1 = ((2 * ly.L + GB - 1) / GB) * ((B + GB - 1) / GB);
      for (int t = cta; t < n1; t += ncta) phase_f1(md, li, s, ly.H, t, sA, sB);
      grid_barrier(gb, ncta, epoch);
      if (li == 0) PK_STAMP(6);
      const int n2 = ((ly.L + GB - 1) / GB) * ((B + GB - 1) / GB);
      for (int t = cta; t < n2; t += ncta) phase_f2(md, li, s, ly.H, true, t, sA, sB);
      grid_barrier(gb, ncta, epoch);
    }
    PK_STAMP(1);
    for (int c = cta; c < md.NCH; c += ncta) phase_score(md, s, c, smem);
    grid_barrier(gb, ncta, epoch);
    PK_STAMP(2);
    phase_stats(md, s, cta, ncta, smem);
    grid_barrier(gb, ncta, epoch);
    PK_STAMP(3);
    for (int c = cta; c < md.NCH; c += ncta) phase_lossgrad(md, s, c, smem);
    grid_barrier(gb, ncta, epoch);
    PK_STAMP(4);
    for (int li = md.n_layers - 1; li >= 0; li--) {
      const LayerDev& ly = md.layer[li];
      phase_b1(md, li, s, cta, ncta);
      grid_barrier(gb, ncta, epoch);
      if (li == md.n_layers - 1) PK_STAMP(7);
      const int n2 = ((ly.L + GB - 1) / GB) * ((B + GB - 1) / GB);
      for (int t = cta; t < n2; t += ncta) phase_b2(md, li, s, t, sA, sB);
      grid_barrier(gb, ncta, epoch);
      if (li == md.n_layers - 1) PK_STAMP(8);
      if (ly.in_dim > 0) {
        const int n3 = ((ly.in_dim + GB - 1) / GB) * ((B + GB - 1) / GB);
        for (int t = cta; t < n3; t += ncta) phase_b3(md, li, s, t, sA, sB);
        grid_barrier(gb, ncta, epoch);
      }
      const DenseJobs dj = dense_jobs(ly.L, ly.in_dim);
      const int nj = dj.nWh + dj.nWrz + dj.nWx + dj.nBh;
      // dense jobs from the top of the grid, input-row updates from the bottom: disjoint arrays, same phase
      for (int j = ncta - 1 - cta; j < nj; j += ncta) phase_dense(md, li, s, j, sA, sB);
      if (li == 0) for (int b = cta; b < B; b += ncta) phase_sparse_in(md, s, b);
      grid_barrier(gb, ncta, epoch);
    }
    PK_STAMP(5);
  }
#undef PK_STAMP
}
