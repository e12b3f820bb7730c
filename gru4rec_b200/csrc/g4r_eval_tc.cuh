// g4r_eval_tc.cuh -- full-catalogue scoring of evaluate_gpu on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// What is computed (reference: yhat = h Wy^T + By, gru4rec.py:502; ranks = (others > targets).sum + 1, evaluation.py:57-64):
// for every lane b of the evaluation batch the number of catalogue items whose score beats / ties the score of the lane's
// target.  The [items x lanes] score matrix (37,483 x 512 at the RSC15 shape: 3.8 GFLOP per mini-batch, the largest dense
// contraction of the whole path) is never written: each 128-lane x 256-item tile is accumulated in TMEM by UMMA
// (tcgen05.mma kind::tf32, M = 128, N = 256, K = 8), read back with tcgen05.ld, and reduced to the two counters in registers.
//
// fp32 fidelity on TF32 tensor cores: 3xTF32 -- every fp32 operand x is split as hi = tf32(x), lo = tf32(x - hi) and the
// product is accumulated as lo*hi + hi*lo + hi*hi (the dropped lo*lo term and the roundings are ~2^-21 relative), i.e. ~1e-6
// relative on the scores, far inside the 1e-4 bar on Recall / MRR.  The target's own column is excluded explicitly (it is
// the one comparison that must be exact), so a rank can only move where two DIFFERENT items' scores differ by < 1e-6 relative.
//
// Structure (one CTA per SM, 320 threads, persistent over its item tiles):
//   pre-pass   k_tc_split writes the operands as hi / lo TF32 blocks in the K-major 128-byte-swizzle UMMA layout: the item table
//              once per evaluation, the hidden states once per mini-batch; one extra K column carries the item bias (1.0 on the
//              hidden-state side), so the accumulator is the complete pre-activation score
//   warp 8     TMA producer (one thread): two bulk copies (cp.async.bulk -> mbarrier complete_tx) per 32-wide K chunk fill a
//              96 KB stage [A hi | A lo | B hi | B lo]; two stages
//   warp 9     MMA issuer (one thread): 12 tcgen05.mma per chunk, tcgen05.commit hands the stage back / publishes the accumulator
//   warps 0-7  epilogue: wait for the accumulator (2 x 256 TMEM columns, double buffered), tcgen05.ld 32 columns at a time, two
//              compares per item against the lane's pre-activation thresholds (k_eval_tgt computes them once per lane) -- a
//              thread owns one evaluation lane (TMEM lane), so the two counters are thread-local; two warps per lane quarter
//              split the tile's columns
// All waits are mbarrier try_wait loops with a time-out that sets an error flag (a wrong phase must not hang the box).
#pragma once

constexpr int TC_M = 128;            // evaluation lanes per tile (UMMA M, TMEM lanes: one per epilogue thread)
constexpr int TC_N = 256;            // items per tile (UMMA N, TMEM columns per accumulator)
constexpr int TC_KC = 32;            // K chunk per pipeline stage (floats)
constexpr int TC_EPI_WARPS = 8;      // two warps per TMEM lane quarter (each takes half of the tile's columns)
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = TC_EPI_THREADS + 64;     // + TMA producer warp + MMA issuer warp
constexpr int TC_STAGES = 2;
constexpr uint32_t TC_A_BYTES = TC_M * TC_KC * 4;       // 16 KB per hi / lo array
constexpr uint32_t TC_B_BYTES = TC_N * TC_KC * 4;       // 32 KB
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 96 KB
constexpr unsigned long long TC_TIMEOUT_NS = 2000000000ull;

struct TcSmem {
  alignas(1024) unsigned char stage[TC_STAGES][TC_STAGE_BYTES];   // [A hi | A lo | B hi | B lo]
  alignas(8) unsigned long long stage_full[TC_STAGES];    // operand blocks of the stage have landed (TMA complete_tx)
  unsigned long long stage_free[TC_STAGES];    // MMAs that read the stage have completed (tcgen05.commit)
  unsigned long long acc_full[2];              // all MMAs of the tile have completed (tcgen05.commit)
  unsigned long long acc_free[2];              // epilogue has drained the accumulator (256 arrivals)
  uint32_t tmem_base;
  int err;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
}
// a wait that does not complete within TC_TIMEOUT_NS is a protocol bug: trap (the launch fails with an error) rather than hang
__device__ __forceinline__ bool tc_mbar_wait(unsigned long long* bar, unsigned int parity, int* err) {
  const uint32_t a = tc_smem_u32(bar);
  unsigned long long t0 = 0; unsigned int spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
    if ((++spins & 63u) == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      if (t - t0 > TC_TIMEOUT_NS) { *(volatile int*)err = 1; asm volatile("trap;"); }
    }
  }
}
__device__ __forceinline__ uint32_t tc_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
// K-major operand blocks in the 128-byte-swizzle layout: a row is 32 tf32 values = 128 bytes, eight rows form a 1024-byte atom in
// which the 16-byte piece c of row r sits at position c ^ (r % 8) (the tensor core reads a whole 128-byte row per access; the
// unswizzled 8 x 16-byte core-matrix layout measured ~8x slower operand fetch).  SBO (next 8 rows) = 1024 B, LBO unused (1);
// blocks are 1024-byte aligned, a K step of 8 values advances the start address by 32 bytes inside the atom.
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of the 16-byte piece (row r, k = 4 * kq .. 4 * kq + 3) inside a block of rows x 32 k-values
__device__ __forceinline__ uint32_t tc_block_off(int r, int kq) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((kq ^ (r & 7)) << 4)); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem_u32(bar)) : "memory");
}

// Pre-split operand blocks in global memory (written once per evaluation for the item table, once per mini-batch for the hidden
// states): block (rb, c) = rows [rb * RB, +RB) x k [32 c, +32) as [hi | lo], each in the K-major 128-byte-swizzle layout
// (tc_block_off).  The scoring kernel then feeds the tensor cores with plain bulk copies (TMA) -- no register staging on the
// critical path.
// The contraction runs over K + 1 values: column K holds `one` on the hidden-state side and the item bias on the table side
// (bias != nullptr), so the accumulator is the complete pre-activation score; table rows past the catalogue get a bias of -3e38,
// which no threshold ever reaches (the epilogue needs no per-column validity test).
constexpr float TC_PAD_BIAS = -3.0e38f;
template <int RB>
__global__ void __launch_bounds__(256) k_tc_split(const float* __restrict__ src, int nrows, int ld, int K, unsigned char* __restrict__ dst, int n_chunk,
                                                  const float* __restrict__ bias, float one) {
  const int rb = blockIdx.x, c = blockIdx.y;
  unsigned char* hi = dst + ((size_t)rb * n_chunk + c) * 2 * (RB * TC_KC * 4);
  unsigned char* lo = hi + RB * TC_KC * 4;
  const int k0 = c * TC_KC;
  for (int i = threadIdx.x; i < RB * 8; i += blockDim.x) {
    const int r = i % RB, cc = i / RB;            // rows fastest: consecutive threads write consecutive 16-byte pieces
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = rb * RB + r;
    if (row < nrows && k0 + cc * 4 < K) v = ld4(src + (size_t)row * ld + k0 + cc * 4);
    if (K - (k0 + cc * 4) >= 0 && K - (k0 + cc * 4) < 4) {       // the extra column (K % 4 == 0 is not required)
      const float x = bias ? (row < nrows ? bias[row] : TC_PAD_BIAS) : one;
      const int u = K - (k0 + cc * 4);
      if (u == 0) v = make_float4(x, 0.f, 0.f, 0.f); else if (u == 1) v.y = x, v.z = 0.f, v.w = 0.f; else if (u == 2) v.z = x, v.w = 0.f; else v.w = x;
    }
    const uint32_t off = tc_block_off(r, cc);
    uint4 h, l;
    h.x = tc_tf32(v.x); h.y = tc_tf32(v.y); h.z = tc_tf32(v.z); h.w = tc_tf32(v.w);
    l.x = tc_tf32(v.x - __uint_as_float(h.x)); l.y = tc_tf32(v.y - __uint_as_float(h.y));
    l.z = tc_tf32(v.z - __uint_as_float(h.z)); l.w = tc_tf32(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}
__device__ __forceinline__ void tc_bulk_copy(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(tc_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(tc_smem_u32(bar)) : "memory");
}

// Ranking compares act(x) with the target score t = act(x_t).  act is monotone non-decreasing, so for every lane there are two
// pre-activation thresholds with  act(x) > t  <=>  x > hi  and  act(x) == t  <=>  lo <= x <= hi : they are found once per lane by
// bisection over the ordered fp32 bit patterns with the very act_fwd the fp32 kernels use (64 evaluations per lane), and the
// per-item work drops to two compares.
__device__ __forceinline__ uint32_t tc_fkey(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float tc_fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
// x_t = the target's own pre-activation (act(x_t) == t): both thresholds are normally within a few ulps of it, so the search
// gallops away from key(x_t) (1, 2, 4, ... keys) and bisects the last bracket -- a handful of evaluations outside the flat
// regions of the activation, at most ~64 inside them.
__device__ __forceinline__ void tc_thresholds(const ActSpec a, bool elem_act, float t, float x_t, float& lo, float& hi) {
  if (!elem_act) { lo = hi = t; return; }
  const uint32_t kmin = tc_fkey(-INFINITY), kmax = tc_fkey(INFINITY), k0 = min(max(tc_fkey(x_t), kmin), kmax);
  // upper: first key above k0 whose activation exceeds t (kmax + 1 if none); invariant act(l - 1) <= t
  uint32_t l = k0 + 1u, r = kmax + 1u;
  for (uint32_t d = 1u; l < r; d <<= 1) {
    const uint32_t m = (kmax - l < d) ? kmax : l + d - 1u;          // probe
    if (act_fwd(a, tc_fkey_inv(m)) > t) { r = m; break; }
    l = m + 1u;
    if (d >= 0x80000000u) break;
  }
  while (l < r) { const uint32_t m = l + ((r - l) >> 1); if (act_fwd(a, tc_fkey_inv(m)) > t) r = m; else l = m + 1u; }
  hi = tc_fkey_inv(l - 1u);
  // lower: smallest key whose activation still reaches t; invariant act(r2) >= t
  uint32_t r2 = k0, l2 = kmin;
  for (uint32_t d = 1u; l2 < r2; d <<= 1) {
    const uint32_t m = (r2 - kmin < d) ? kmin : r2 - d;
    if (act_fwd(a, tc_fkey_inv(m)) >= t) { r2 = m; if (d >= 0x80000000u) break; } else { l2 = m + 1u; break; }
  }
  while (l2 < r2) { const uint32_t m = l2 + ((r2 - l2) >> 1); if (act_fwd(a, tc_fkey_inv(m)) >= t) r2 = m; else l2 = m + 1u; }
  lo = tc_fkey_inv(r2);
}

// cnt[b*2 + 0] += #items with score > target score of lane b; cnt[b*2 + 1] += #items with score == target (the target itself
// counts as one tie, exactly as in the fp32 kernel where its score equals the target score bit for bit).
// Tile = 128 evaluation lanes (UMMA M, TMEM lanes: one lane per epilogue thread, so the counting is thread-local) x 256 items
// (UMMA N, TMEM columns).  Asplit: hidden-state blocks of 128 lanes, Bsplit: item-table blocks of 256 items (k_tc_split).
__global__ void __launch_bounds__(TC_THREADS, 1) k_eval_tc(int slot, int s, const float* __restrict__ tgt, int tgt_stride, int* cnt,
                                                           const unsigned char* __restrict__ Asplit, const unsigned char* __restrict__ Bsplit) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(tc_raw);
  const ModelDev& md = MD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = md.wM[s], I = md.n_items, K = md.L + 1;        // + the bias column
  const int n_tiles = (I + TC_N - 1) / TC_N;         // item tiles
  const int n_lb = (M + TC_M - 1) / TC_M;            // lane blocks
  const int n_chunk = (K + TC_KC - 1) / TC_KC;
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; i++) { tc_mbar_init(&sm.stage_free[i], 1); tc_mbar_init(&sm.stage_full[i], 1); }
    for (int i = 0; i < 2; i++) { tc_mbar_init(&sm.acc_full[i], 1); tc_mbar_init(&sm.acc_free[i], TC_EPI_THREADS); }
    sm.err = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == TC_EPI_WARPS) {   // TMEM: 512 columns = two 128 x 256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem_u32(&sm.tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  if (warp == TC_EPI_WARPS) {
    // ================= TMA producer (one thread): operand blocks -> shared memory stages =================
    if (lane == 0) {
      unsigned int it = 0;
      for (int lb = 0; lb < n_lb; lb++)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
          for (int c = 0; c < n_chunk; c++, it++) {
            const uint32_t st = it & 1u, use = it >> 1;
            if (use > 0) tc_mbar_wait(&sm.stage_free[st], (use - 1) & 1u, &sm.err);     // the MMAs of the previous use are done
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tc_smem_u32(&sm.stage_full[st])), "r"(TC_STAGE_BYTES) : "memory");
            tc_bulk_copy(sm.stage[st], Asplit + ((size_t)lb * n_chunk + c) * 2 * TC_A_BYTES, 2 * TC_A_BYTES, &sm.stage_full[st]);
            tc_bulk_copy(sm.stage[st] + 2 * TC_A_BYTES, Bsplit + ((size_t)t * n_chunk + c) * 2 * TC_B_BYTES, 2 * TC_B_BYTES, &sm.stage_full[st]);
          }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      unsigned int it = 0, wi = 0;
      for (int lb = 0; lb < n_lb; lb++)
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, wi++) {
          const uint32_t acc = wi & 1u;
          if (wi >= 2) tc_mbar_wait(&sm.acc_free[acc], ((wi >> 1) - 1) & 1u, &sm.err);   // epilogue drained this accumulator
          for (int c = 0; c < n_chunk; c++, it++) {
            const uint32_t st = it & 1u, use = it >> 1;
            tc_mbar_wait(&sm.stage_full[st], use & 1u, &sm.err);                         // operand blocks have landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tc_smem_u32(sm.stage[st]), a_lo = a_hi + TC_A_BYTES, b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
            const uint32_t d = tmem + acc * TC_N;
            const int ksteps = (min(TC_KC, K - c * TC_KC) + 7) / 8;
            for (int j = 0; j < ksteps; j++) {
              const uint32_t o = (uint32_t)j * 32u;      // 8 values along K = 32 bytes inside the swizzle atom
              tc_mma_tf32(d, tc_desc(a_lo + o), tc_desc(b_hi + o), idesc, (c == 0 && j == 0) ? 0u : 1u);
              tc_mma_tf32(d, tc_desc(a_hi + o), tc_desc(b_lo + o), idesc, 1u);
              tc_mma_tf32(d, tc_desc(a_hi + o), tc_desc(b_hi + o), idesc, 1u);
            }
            tc_commit(&sm.stage_free[st]);
            if (c == n_chunk - 1) tc_commit(&sm.acc_full[acc]);
          }
        }
    }
  } else if (warp < TC_EPI_WARPS) {
    // ================= epilogue: TMEM -> registers -> thread-local counters =================
    // warp w reads TMEM lanes 32 * (w % 4) .. + 31 (its evaluation lanes) and the column half w / 4 of the tile
    unsigned int wi = 0;
    const bool elem_act = md.fact.kind <= G4R_ACT_SELU;
    const int q4 = warp & 3, half = warp >> 2;
    for (int lb = 0; lb < n_lb; lb++) {
      const int b = lb * TC_M + q4 * 32 + lane;
      const bool vrow = b < M;
      const int yit = vrow ? md.wY[(size_t)s * md.B + b] : -1;
      const float lo = vrow ? tgt[tgt_stride + b] : INFINITY, hi = vrow ? tgt[2 * tgt_stride + b] : INFINITY;   // k_eval_tgt
      int cgt = 0, cge = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, wi++) {
        const uint32_t acc = wi & 1u;
        const int i0 = t * TC_N;
        tc_mbar_wait(&sm.acc_full[acc], (wi >> 1) & 1u, &sm.err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int self_col = yit - i0;                       // the target's own column (if it falls into this tile)
#pragma unroll 1
        for (int q = 0; q < 4; q++) {
          const int c0 = half * 128 + q * 32;
          uint32_t r[32];
          const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + acc * TC_N + c0;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                       "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                         "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                         "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                         "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float x = __uint_as_float(r[j]);
            cgt += (x > hi) ? 1 : 0;
            cge += (x >= lo) ? 1 : 0;
          }
          if ((unsigned)(self_col - c0) < 32u) {             // rare: take the target's own column back out, it counts as exactly one tie
            float xs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j++) if (j == self_col - c0) xs = __uint_as_float(r[j]);
            cgt -= (xs > hi) ? 1 : 0;
            cge -= (xs >= lo) ? 1 : 0;
            cge += 1;                                        // == (self: not above) + one tie
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_mbar_arrive(&sm.acc_free[acc]);
      }
      const int ceq = cge - cgt;                            // lo <= x <= hi
      if (vrow) { if (cgt) atomicAdd(&cnt[b * 2], cgt); if (ceq) atomicAdd(&cnt[b * 2 + 1], ceq); }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
  }
}
