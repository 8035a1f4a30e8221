// umma_gemm.cu — (S1/S3) the two dense contractions of hybrid spreading on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA):
//     G = A^T K_u^-1 A      /root/reference/model/SpreadMethod/model.py:25   (np.dot, fp64)
//     F = A W               /root/reference/model/SpreadMethod/model.py:98   (np.dot, fp64)
// Both have a BINARY (0/1) left operand; the real right operand is carried as `planes`
// narrow matrices that are multiplied against the same A tile inside ONE MMA:
//   kind 0 (bf16 x bf16 -> fp32): planes = hi / mid / lo bf16 split of an fp32 matrix; every
//          product 0/1 x bf16 is exact, the three partial sums live in separate TMEM column
//          ranges and are added lo-first in the epilogue;
//   kind 1 (u8 x u8 -> s32): planes = base-256 digits of the fixed-point integer
//          q_u = round(2^s / k_u); all products and sums are exact integers, the digit sums
//          are recombined in float64 in the epilogue and rounded once to fp32.
//
// Accumulation numerics (measured on B200, tools/probe_umma_numerics.py): tcgen05 kind::f16
// adds the 16 products of one MMA and the fp32 accumulator with TRUNCATION (round toward
// zero, ~2 guard bits), so a long K loop is biased low by up to ~1 ulp per MMA step
// (1.5e-5 at K=4100 on heavy-tailed data).  The bf16 kind therefore drains the TMEM
// accumulator every `chunk_kb` K-blocks and carries the running sum in fp32 registers with
// round-to-nearest adds: the truncation error is then relative to a chunk's partial sum and
// bounded by (4*chunk_kb) * 2^-23, independent of K.  kind::i8 accumulates exactly in int32.
//
// Kernel shape (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer  : A tile 128 x 128B and `planes` B tiles NB x 128B per stage
//                               (SWIZZLE_128B, K-major), 4-stage mbarrier ring
//   warp 1      MMA issuer    : one thread issues 4 x tcgen05.mma (M=128, N=planes*NB, K=32B)
//                               per stage into a double-buffered TMEM accumulator (2 x 256 col)
//   warps 2..5  epilogue      : tcgen05.ld 32x32b -> combine planes -> scale -> fp32 store,
//                               overlapped with the next tile's main loop
#include "umma_common.cuh"
#include "select.cuh"

namespace lgc {
namespace umma {

struct GemmParams {
  int64_t M, N;      // output rows, output columns (per plane)
  int num_kb;        // K blocks of 128 bytes
  int planes, NB;    // planes and output columns per tile
  int tiles_m, tiles_n;
  float* C;
  int64_t ldc;
  const float* rs;
  const float* cs;
  double scale;
  int k_elems_per_kb;  // 64 (bf16) or 128 (u8)
  int chunk_kb;        // K-blocks accumulated in TMEM before a drain (>= num_kb: single chunk)
  // ---- tile schedule of the CTA-pair kernel ----
  //   0: every tile, column-block major (consecutive clusters share a B tile through L2)
  //   1: SYMMETRIC output (C == C^T bit for bit, M == N): only tiles that touch the upper triangle are computed,
  //      tiles strictly above the diagonal blocks also store their transpose
  //   2: ROW-OWNER segments (fused top-k): a work item is (256-row block, one of `segs` ranges of column tiles);
  //      the epilogue threads keep their row's selection state in registers across the item's tiles
  int mode;
  int segs;
  // ---- fused row-wise top-k epilogue (mode 2) ----
  const uint32_t* excl;          // bit-packed exclusion matrix (row r, col c -> bit (excl_row0 + r)*excl_stride + c) or null
  long long excl_stride, excl_row0;
  int k;
  unsigned long long* cand;      // [M * segs][kCandCap] candidate keys
  int* cand_cnt;                 // [M * segs]
  // ---- multi-GPU (schedule 1): this rank computes every n_ranks-th cluster slot of the symmetric schedule and stores
  //      each tile (and its mirror) into ALL replicas of C (peer pointers through CUDA IPC): a fused GEMM + all-gather ----
  float* Cpeer[8];
  int n_peers;                   // 0: store to C only
  int rank_slot, n_ranks;        // this rank's offset / the stride of the cluster slots (0 / 1 on one GPU)
};


struct TileSched {
  int m, n, seg, n_end;
  bool first, last, started;
  long long w;
};

// number of 256-row blocks of column tile n that touch the upper triangle (symmetric schedule)
__device__ __forceinline__ int sym_rows(const GemmParams& p, int n) {
  const long long c = ((long long)(n + 1) * p.NB - 1) / 256 + 1;
  return c < p.tiles_m ? (int)c : p.tiles_m;
}

// Every role of the CTA pair (both producers, the MMA issuer, all epilogue warps) walks the SAME sequence of tiles.
__device__ __forceinline__ bool sched_next(TileSched& s, const GemmParams& p, int cluster_id, int n_clusters) {
  if (p.mode == 0) {
    s.w = s.started ? s.w + n_clusters : cluster_id;
    s.started = true;
    if (s.w >= (long long)p.tiles_m * p.tiles_n) return false;
    s.m = (int)(s.w % p.tiles_m);
    s.n = (int)(s.w / p.tiles_m);
    return true;
  }
  if (p.mode == 1) {
    if (!s.started) { s.started = true; s.m = cluster_id; s.n = 0; }
    else s.m += n_clusters;
    while (s.n < p.tiles_n) {
      const int c = sym_rows(p, s.n);
      if (s.m < c) return true;
      s.m -= c;
      ++s.n;
    }
    return false;
  }
  if (s.started && s.n + 1 < s.n_end) {
    ++s.n;
    s.first = false;
    s.last = (s.n + 1 == s.n_end);
    return true;
  }
  s.w = s.started ? s.w + n_clusters : cluster_id;
  s.started = true;
  while (s.w < (long long)p.tiles_m * p.segs) {
    s.m = (int)(s.w / p.segs);
    s.seg = (int)(s.w % p.segs);
    const int n0 = (int)((long long)s.seg * p.tiles_n / p.segs), n1 = (int)((long long)(s.seg + 1) * p.tiles_n / p.segs);
    if (n1 > n0) {
      s.n = n0; s.n_end = n1; s.first = true; s.last = (n0 + 1 == n1);
      return true;
    }
    s.w += n_clusters;
  }
  return false;
}

template <int KIND, int PLANES>
__global__ void __launch_bounds__(kThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kStages * kAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* tfull = bars + 2 * kStages;     // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NB = PLANES == 3 ? 80 : PLANES == 4 ? 64 : 128;  // output columns per tile
  const int num_tiles = p.tiles_m * p.tiles_n;
  constexpr int n_mma = PLANES * NB;  // MMA N
  const int n_chunks = (p.num_kb + p.chunk_kb - 1) / p.chunk_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const uint32_t stage_tx = (uint32_t)(kAStage + n_mma * kKBytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], stage_tx);
          const int k0 = kb * p.k_elems_per_kb;
          tma_load_2d(&tmapA, &full[stage], smemA + stage * kAStage, k0, m_blk * kBlockM);
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl)
            tma_load_3d(&tmapB, &full[stage], smemB + stage * kBStage + pl * NB * kKBytes, k0, n_blk * NB, pl);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // instruction descriptor: D fmt | A fmt | B fmt | K-major both | N>>3 | M>>4
    uint32_t idesc = 0;
    if (KIND == 0) idesc |= (1u << 4) | (1u << 7) | (1u << 10);  // F32 accum, BF16 x BF16
    else idesc |= (2u << 4);                                     // S32 accum, U8 x U8 (fmt 0)
    idesc |= ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;  // accumulator-buffer use counter (one per chunk)
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int ch = 0; ch < n_chunks; ++ch, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        const int kb0 = ch * p.chunk_kb;
        const int kb1 = min(kb0 + p.chunk_kb, p.num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = make_smem_desc(smem_u32(smemA + stage * kAStage));
            const uint64_t bdesc = make_smem_desc(smem_u32(smemB + stage * kBStage));
#pragma unroll
            for (int k = 0; k < kKBytes / 32; ++k)  // 32 bytes of K per MMA: +2 in >>4 units
              mma_ss<KIND>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            tc_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
            if (kb == kb1 - 1) tc_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
      const int64_t row = (int64_t)m_blk * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float rscale = (row_ok && p.rs) ? __ldg(p.rs + row) : 1.0f;
      float* crow = p.C + (row_ok ? row : 0) * p.ldc;
      float run[NB];  // running fp32 (round-to-nearest) sum over chunks, bf16 kind only
      for (int ch = 0; ch < n_chunks; ++ch, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride;
        const bool first = ch == 0, last = ch == n_chunks - 1;
#pragma unroll
        for (int c0 = 0; c0 < NB; c0 += 16) {
          uint32_t r[PLANES][16];
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) tmem_ld16(t_row + pl * NB + c0, r[pl]);
          tmem_ld_wait();
          float out[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (KIND == 0) {
              // planes hold hi / mid / lo: add the small ones first
              float a = __uint_as_float(r[PLANES - 1][j]);
#pragma unroll
              for (int pl = PLANES - 2; pl >= 0; --pl) a += __uint_as_float(r[pl][j]);
              if (!first) a += run[c0 + j];
              run[c0 + j] = a;
              out[j] = a * (float)p.scale;
            } else {
              long long tot = 0;  // exact: digits recombined in 64-bit integers, one rounding
#pragma unroll
              for (int pl = PLANES - 1; pl >= 0; --pl) tot = (tot << 8) + (long long)(int)r[pl][j];
              out[j] = (float)((double)tot * p.scale);
            }
          }
          if (last) {
            const int64_t col0 = (int64_t)n_blk * NB + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int64_t col = col0 + j;
              const float cscale = (p.cs && col < p.N) ? __ldg(p.cs + col) : 1.0f;
              out[j] = out[j] * rscale * cscale;
            }
            if (row_ok) {
              if (col0 + 16 <= p.N && (p.ldc & 3) == 0 && ((uintptr_t)p.C & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (col0 + j < p.N) crow[col0 + j] = out[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}


// ================================================================================================
// CTA-pair variant (tcgen05 cta_group::2, cluster of two CTAs on one TPC, UMMA M = 256).
// The single-CTA kernel above is shared-memory-bandwidth bound: per K-block TMA writes 46 KB and
// the MMA reads 46 KB of smem in 480 cycles = 192 B/clk against the 128 B/clk port (tensor pipe
// 67-71 % measured, profiles/r1_ncu_gemm.txt).  Here each CTA stages its own 128 rows of A but only
// HALF of the B rows (N/2); the pair's tensor cores read both halves, so per-CTA smem traffic drops
// to 2 x 31 KB per 480 cycles = 129 B/clk.
//   * both CTAs run a TMA producer; all transaction bytes are signalled on the LEADER's (even CTA)
//     full barrier (barrier address with the peer bit cleared), count 2: leader arrive.expect_tx +
//     one remote arrive from the peer producer;
//   * only the leader's warp 1 issues MMAs; tcgen05.commit multicasts the "stage free" and
//     "accumulator full" arrivals to both CTAs;
//   * both CTAs run 4 epilogue warps on their own 128 TMEM lanes and arrive remotely on the leader's
//     "accumulator empty" barrier (count 8).
// ================================================================================================
// each CTA stages 16 KB of A + <= 16 KB of B per K-block, so six stages fit where the single-CTA kernel has four
constexpr int kPStages = 6;
constexpr int kPBStage = 128 * kKBytes;  // 16 KB
constexpr int kPStageBytes = kAStage + kPBStage;
static_assert((size_t)kPStages * kPStageBytes <= (size_t)kStages * kStageBytes, "pair pipeline must fit the same smem budget");

// EPI 0: store C (optionally mirrored, schedule 1);  EPI 1: fused row-wise top-k (schedule 2), C is never written.
// EPI 1 runs EIGHT epilogue warps (two per scheduler): two threads share an output row, each ranks one half of the
// tile's columns into its own candidate buffer, and they share the row's pruning threshold through shared memory.
// With four warps the epilogue was latency-bound (one warp per scheduler, IPC ~0.1, ncu round 2) and the buffer
// compactions landed on the critical path of a one-wave launch (0.45 ms against 0.21 ms for the plain GEMM at ML-1M).
constexpr int kEpiParts = 2;
constexpr int kThreadsTopk = 64 + 32 * 4 * kEpiParts;   // 320

template <int KIND, int PLANES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EPI == 1 ? kThreadsTopk : kThreads, 1)
umma_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                      const __grid_constant__ CUtensorMap tmapBh, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kPStages * kAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPStages * kPStageBytes);
  uint64_t* full = bars;                      // [kStages]  used on the leader
  uint64_t* empty = bars + kPStages;           // [kStages]  per CTA
  uint64_t* tfull = bars + 2 * kPStages;       // [2]        per CTA
  uint64_t* tempty = bars + 2 * kPStages + 2;  // [2]        used on the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPStages + 4);
  uint32_t* s_row_thr = tmem_slot + 4;   // [128] EPI 1: value key of the best known k-th score of each of the CTA's rows

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // cluster slot in the (possibly multi-GPU) tile schedule: rank r owns slots r, r + n_ranks, ... of every stride
  const int cluster_id = (int)(blockIdx.x >> 1) * (p.n_ranks > 0 ? p.n_ranks : 1) + p.rank_slot;
  const int n_clusters = (int)(gridDim.x >> 1) * (p.n_ranks > 0 ? p.n_ranks : 1);
  constexpr int NB = PLANES == 3 ? 80 : PLANES == 4 ? 64 : 128;  // output columns per tile
  constexpr int n_mma = PLANES * NB;                              // MMA N (both halves)
  constexpr int kEpiWarps = EPI == 1 ? 4 * kEpiParts : 4;
  constexpr int H = n_mma / 2;                                    // B rows staged by each CTA
  const int n_chunks = (p.num_kb + p.chunk_kb - 1) / p.chunk_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPStages; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapBh) : "memory");
  }
  cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast commit
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      const uint32_t cta_tx = (uint32_t)(kAStage + H * kKBytes);
      int stage = 0;
      uint32_t phase = 0;
      TileSched ts{};
      while (sched_next(ts, p, cluster_id, n_clusters)) {
        const int m_blk = ts.m, n_blk = ts.n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full[stage], 2 * cta_tx);
          else mbar_arrive_leader(&full[stage]);
          const int k0 = kb * p.k_elems_per_kb;
          tma2_load_2d(&tmapA, &full[stage], smemA + stage * kAStage, k0, m_blk * 256 + (int)rank * kBlockM);
          // this CTA's half of the stacked plane rows [rank*H, rank*H + H)
          uint8_t* dstB = smemB + stage * kPBStage;
          if (PLANES == 3) {
            if (leader) {
              tma2_load_3d(&tmapB, &full[stage], dstB, k0, n_blk * NB, 0);
              tma2_load_3d(&tmapBh, &full[stage], dstB + NB * kKBytes, k0, n_blk * NB, 1);
            } else {
              tma2_load_3d(&tmapBh, &full[stage], dstB, k0, n_blk * NB + NB / 2, 1);
              tma2_load_3d(&tmapB, &full[stage], dstB + (NB / 2) * kKBytes, k0, n_blk * NB, 2);
            }
          } else if (PLANES == 1) {
            tma2_load_3d(&tmapBh, &full[stage], dstB, k0, n_blk * NB + (int)rank * (NB / 2), 0);
          } else {
#pragma unroll
            for (int q = 0; q < PLANES / 2; ++q)
              tma2_load_3d(&tmapB, &full[stage], dstB + q * NB * kKBytes, k0, n_blk * NB, (int)rank * (PLANES / 2) + q);
          }
          if (++stage == kPStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader) {
      uint32_t idesc = 0;
      if (KIND == 0) idesc |= (1u << 4) | (1u << 7) | (1u << 10);
      else idesc |= (2u << 4);
      idesc |= ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TileSched ts{};
      while (sched_next(ts, p, cluster_id, n_clusters)) {
        for (int ch = 0; ch < n_chunks; ++ch, ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kAccStride;
          const int kb0 = ch * p.chunk_kb;
          const int kb1 = min(kb0 + p.chunk_kb, p.num_kb);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (lane == 0) {
              const uint64_t adesc = make_smem_desc(smem_u32(smemA + stage * kAStage));
              const uint64_t bdesc = make_smem_desc(smem_u32(smemB + stage * kPBStage));
#pragma unroll
              for (int k = 0; k < kKBytes / 32; ++k)
                mma2_ss<KIND>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
              tc_commit2_mc(&empty[stage]);
              if (kb == kb1 - 1) tc_commit2_mc(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == kPStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2.., both CTAs) ------------------------------
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int cpart = EPI == 1 ? (warp - 2) >> 2 : 0;   // EPI 1: which part of the tile's columns this warp ranks
    const int rloc = q * 32 + lane;                // row inside the CTA's 128
    constexpr int kGroups = NB / 16;
    constexpr int kG0 = EPI == 1 ? (kGroups + kEpiParts - 1) / kEpiParts : kGroups;   // groups of 16 columns of part 0
    const int g_begin = EPI == 1 ? cpart * kG0 : 0;
    const int g_end = EPI == 1 ? (cpart == kEpiParts - 1 ? kGroups : (cpart + 1) * kG0) : kGroups;
    constexpr int kPartCols = kG0 * 16;            // most columns one thread can offer per tile
    int it = 0;
    // fused top-k state of this thread's (row, column part) (EPI 1): own threshold key and fill of its candidate buffer
    unsigned long long sel_thr = 0ull;
    int sel_cnt = 0;
    unsigned long long* sel_buf = nullptr;
    TileSched ts{};
    while (sched_next(ts, p, cluster_id, n_clusters)) {
      const int m_blk = ts.m, n_blk = ts.n;
      const int64_t row = (int64_t)m_blk * 256 + (int64_t)rank * kBlockM + rloc;
      const bool row_ok = row < p.M;
      const float rscale = (row_ok && p.rs) ? __ldg(p.rs + row) : 1.0f;
      float* crow = (EPI == 0) ? p.C + (row_ok ? row : 0) * p.ldc : nullptr;
      // schedule 1: this tile lies strictly above the diagonal blocks -> its transpose is not computed anywhere else
      const bool mirror = EPI == 0 && p.mode == 1 && (long long)n_blk * NB >= (long long)(m_blk + 1) * 256;
      float sel_thr_f = -INFINITY;
      if (EPI == 1) {
        if (ts.first) {
          sel_thr = 0ull; sel_cnt = 0;
          sel_buf = p.cand + ((size_t)(row_ok ? row : 0) * (p.segs * kEpiParts) + ts.seg * kEpiParts + cpart) * kCandCap;
          // all epilogue warps of the CTA start a work item together: reset the shared row thresholds
          asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
          if (cpart == 0) s_row_thr[rloc] = 0u;
          asm volatile("bar.sync 1, %0;" ::"r"(32 * kEpiWarps) : "memory");
        }
        // pruning threshold of this tile: the row's best known k-th score over both column parts
        const uint32_t tk = s_row_thr[rloc];
        sel_thr_f = row_ok ? (tk ? key_float(tk) : -INFINITY) : INFINITY;
      }
      // EPI 1: the exclusion bits of this row for the tile's NB columns, fetched BEFORE waiting for the accumulator so
      // that their latency hides behind the tile's MMAs (a dependent global load per surviving value would serialise
      // the first tiles of an item, whose threshold is still -inf)
      constexpr int MW = (NB + 31) / 32;
      uint32_t mbits[MW];
#pragma unroll
      for (int i = 0; i < MW; ++i) mbits[i] = 0u;
      if (EPI == 1 && p.excl && row_ok) {
        const long long b0 = (p.excl_row0 + row) * p.excl_stride + (long long)n_blk * NB;
        const long long last_word = ((p.excl_row0 + row) * p.excl_stride + p.N - 1) >> 5;   // last word holding a bit of this row
        const long long w0 = b0 >> 5;
        const int sh = (int)(b0 & 31);
        uint32_t raw[MW + 1];
#pragma unroll
        for (int i = 0; i <= MW; ++i) raw[i] = (w0 + i <= last_word) ? __ldg(p.excl + w0 + i) : 0u;
#pragma unroll
        for (int i = 0; i < MW; ++i) mbits[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
      }
      float run[EPI == 1 ? 1 : NB];
      for (int ch = 0; ch < n_chunks; ++ch, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride;
        const bool first = ch == 0, last = ch == n_chunks - 1;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          if (EPI == 1 && (g < g_begin || g >= g_end)) continue;
          const int c0 = g * 16;
          uint32_t r[PLANES][16];
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) tmem_ld16(t_row + pl * NB + c0, r[pl]);
          tmem_ld_wait();
          float out[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (KIND == 0) {
              float a = __uint_as_float(r[PLANES - 1][j]);
#pragma unroll
              for (int pl = PLANES - 2; pl >= 0; --pl) a += __uint_as_float(r[pl][j]);
              if (EPI == 0) {
                if (!first) a += run[c0 + j];
                run[c0 + j] = a;
              }
              out[j] = a * (float)p.scale;
            } else {
              long long tot = 0;
#pragma unroll
              for (int pl = PLANES - 1; pl >= 0; --pl) tot = (tot << 8) + (long long)(int)r[pl][j];
              out[j] = (float)((double)tot * p.scale);
            }
          }
          if (last) {
            const int64_t col0 = (int64_t)n_blk * NB + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int64_t col = col0 + j;
              const float cscale = (p.cs && col < p.N) ? __ldg(p.cs + col) : 1.0f;
              out[j] = out[j] * rscale * cscale;
            }
            if (EPI == 0) {
              if (row_ok && p.n_peers == 0) {
                if (col0 + 16 <= p.N && (p.ldc & 3) == 0 && ((uintptr_t)p.C & 15) == 0) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (col0 + j < p.N) crow[col0 + j] = out[j];
                }
                if (mirror) {
                  // C[col, row] = C[row, col]: for a fixed j the 32 lanes of the warp own 32 consecutive rows, so
                  // each of these stores is one coalesced 128-byte segment of row (col0 + j) of C
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (col0 + j < p.N) p.C[(col0 + j) * p.ldc + row] = out[j];
                }
              } else if (row_ok) {
                // multi-GPU: the tile (and its mirror) goes into every rank's replica of C over NVLink
                const bool vec = col0 + 16 <= p.N && (p.ldc & 3) == 0;
                for (int pr = 0; pr < p.n_peers; ++pr) {
                  float* cb = p.Cpeer[pr];
                  float* cr = cb + row * p.ldc;
                  if (vec) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                      *reinterpret_cast<float4*>(cr + col0 + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                      if (col0 + j < p.N) cr[col0 + j] = out[j];
                  }
                  if (mirror) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                      if (col0 + j < p.N) cb[(col0 + j) * p.ldc + row] = out[j];
                  }
                }
              }
            } else {
              // ---- fused selection: ONE compare per 16 values (their maximum against the row's threshold); in a group
              //      with a survivor every value gets its exact 64-bit key (value, column), passes the exclusion bit
              //      test and is appended to the thread's private buffer (plain stores) ----
              float mx = fmaxf(fmaxf(fmaxf(out[0], out[1]), fmaxf(out[2], out[3])), fmaxf(fmaxf(out[4], out[5]), fmaxf(out[6], out[7])));
              mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(out[8], out[9]), fmaxf(out[10], out[11])),
                                   fmaxf(fmaxf(out[12], out[13]), fmaxf(out[14], out[15]))));
              if (mx >= sel_thr_f) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int64_t col = col0 + j;
                  if (out[j] >= sel_thr_f && col < p.N) {
                    const unsigned long long key = make_key(float_key(out[j]), (uint32_t)col);
                    const bool ex = (mbits[(c0 + j) >> 5] >> ((c0 + j) & 31)) & 1u;
                    if (key > sel_thr && !ex) sel_buf[sel_cnt++] = key;
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[acc]);   // the accumulator is free: the next tile's MMAs may start
      }
      if (EPI == 1) {
        // buffers that could overflow during the next tile are compacted now, one at a time, by the whole warp (radix
        // select in registers, select.cuh); the TMEM buffer has been released already, so this overlaps the tensor-core
        // work of the following tiles.  The new k-th best is published for the row's other column part.
        uint32_t need = __ballot_sync(0xffffffffu, sel_cnt + kPartCols > kCandCap);
        while (need) {
          const int rl = __ffs(need) - 1;
          need &= need - 1u;
          unsigned long long* b = reinterpret_cast<unsigned long long*>(
              __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(sel_buf), rl));
          const int c = __shfl_sync(0xffffffffu, sel_cnt, rl);
          unsigned long long t;
          const int kept = warp_select<kCandCap / 32>(b, c, p.k, lane, &t);
          if (lane == rl) {
            sel_cnt = kept;
            sel_thr = t;
            if (t) atomicMax(&s_row_thr[rloc], (uint32_t)(t >> 32));
          }
        }
        if (ts.last && row_ok) p.cand_cnt[(size_t)row * (p.segs * kEpiParts) + ts.seg * kEpiParts + cpart] = sel_cnt;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still touch its smem/TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

static int g_use_pair = 1;   // cta_group::2 kernel when the problem has at least one 256-row tile
static int g_chunk_kb = 8;  // K-blocks (of 64 bf16) per TMEM accumulation chunk, see header comment

// ---- host side: tensor maps through the driver entry point (no -lcuda link dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// naive on-device cross-check (double accumulate), one thread per output
template <int KIND>
__global__ void gemm_planes_simt_kernel(const void* A_, int64_t lda, const void* B_, int64_t ldb,
                                        int64_t plane_stride, int planes, int64_t M, int64_t N, int64_t K,
                                        float* C, int64_t ldc, const float* rs, const float* cs, double scale) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t m = blockIdx.y;
  if (n >= N || m >= M) return;
  double tot = 0.0;
  for (int pl = planes - 1; pl >= 0; --pl) {
    double s = 0.0;
    if (KIND == 0) {
      const uint16_t* A = (const uint16_t*)A_ + m * lda;
      const uint16_t* B = (const uint16_t*)B_ + pl * plane_stride + n * ldb;
      for (int64_t k = 0; k < K; ++k)
        s += (double)__uint_as_float((uint32_t)A[k] << 16) * (double)__uint_as_float((uint32_t)B[k] << 16);
      tot += s;
    } else {
      const uint8_t* A = (const uint8_t*)A_ + m * lda;
      const uint8_t* B = (const uint8_t*)B_ + pl * plane_stride + n * ldb;
      long long si = 0;
      for (int64_t k = 0; k < K; ++k) si += (long long)A[k] * (long long)B[k];
      tot += (double)si * (double)(1ll << (8 * pl));
    }
  }
  float v = (float)(tot * scale);
  if (KIND == 0) v = (float)tot * (float)scale;
  C[m * ldc + n] = v * (rs ? rs[m] : 1.f) * (cs ? cs[n] : 1.f);
}

static int check_gemm_args(int kind, const void* A, int64_t lda, const void* B, int64_t ldb,
                           int64_t plane_stride, int planes, int64_t M, int64_t N, int64_t K, float* C,
                           int64_t ldc) {
  LGC_REQUIRE(kind == 0 || kind == 1, "gemm: kind must be 0 (bf16) or 1 (u8)");
  LGC_REQUIRE(A && B && C, "gemm: null pointer");
  LGC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem");
  LGC_REQUIRE(lda >= K && ldb >= K && ldc >= N, "gemm: leading dimension smaller than extent");
  if (kind == 0) LGC_REQUIRE(planes >= 1 && planes <= 3, "gemm: bf16 kind takes 1..3 planes");
  else LGC_REQUIRE(planes >= 1 && planes <= 4, "gemm: u8 kind takes 1..4 planes");
  LGC_REQUIRE(planes == 1 || plane_stride >= N * ldb, "gemm: plane stride overlaps planes");
  return LGC_OK;
}

}  // namespace umma
}  // namespace lgc

using namespace lgc;
using namespace lgc::umma;

namespace lgc {
namespace umma {

struct TopkArgs {
  const uint32_t* excl;
  int64_t excl_stride, excl_row0;
  int k;
  int64_t* out_idx;
  float* out_val;
  void* scratch;
  size_t scratch_bytes;
};

// segments per 256-row block of the fused top-k schedule: fill the machine when there are fewer row blocks than
// CTA pairs (ML-1M: 24 row blocks x 3 segments = 72 work items on 74 pairs), one segment otherwise
static int topk_segments(int64_t M, int64_t N, int planes) {
  const int NB = planes == 3 ? 80 : planes == 4 ? 64 : 128;
  const int64_t tiles_m = ceil_div(M, 256), tiles_n = ceil_div(N, NB);
  const int clusters = num_sms() / 2;
  int64_t segs = tiles_m >= clusters ? 1 : clusters / tiles_m;
  if (segs > tiles_n) segs = tiles_n;
  return (int)(segs < 1 ? 1 : segs);
}

static size_t topk_scratch_bytes(int64_t M, int64_t N, int planes) {
  const size_t rows = (size_t)M * topk_segments(M, N, planes) * kEpiParts;
  return align_up(rows * kCandCap * sizeof(unsigned long long), 256) + align_up(rows * sizeof(int), 256);
}

// mode 0: plain; 1: symmetric output (kind 1, M == N, no row/column scales); 2: fused top-k (topk != null)
struct PeerArgs {
  float* const* tables;
  int n_peers;   // replicas to store into (1 with an NVSwitch multicast address)
  int rank, n_ranks;
};

static int gemm_planes_impl(int32_t kind, const void* A, int64_t lda, const void* B, int64_t ldb,
                            int64_t plane_stride, int32_t planes, int64_t M, int64_t N, int64_t K, float* C, int64_t ldc,
                            const float* rs, const float* cs, double scale, int mode, const TopkArgs* topk,
                            lgc_stream_t stream_, const PeerArgs* peers = nullptr) {
  const int esize = kind == 0 ? 2 : 1;
  LGC_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm: operands must be 16-byte aligned");
  LGC_REQUIRE((lda * esize) % 16 == 0 && (ldb * esize) % 16 == 0 && (plane_stride * esize) % 16 == 0,
              "gemm: leading dimensions / plane stride must be multiples of 16 bytes");
  LGC_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: extents exceed int32");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) LGC_FAIL(LGC_ERR_CUDA, "gemm: cuTensorMapEncodeTiled entry point not available");

  const int NB = planes == 3 ? 80 : planes == 4 ? 64 : 128;
  const int k_elems = kKBytes / esize;
  const CUtensorMapDataType dt = kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)lda * esize};
    cuuint32_t box[2] = {(cuuint32_t)k_elems, (cuuint32_t)kBlockM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmA, dt, 2, const_cast<void*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LGC_FAIL(LGC_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)ldb * esize, (cuuint64_t)(planes > 1 ? plane_stride : N * ldb) * esize};
    cuuint32_t box[3] = {(cuuint32_t)k_elems, (cuuint32_t)NB, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, dt, 3, const_cast<void*>(B), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LGC_FAIL(LGC_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }

  GemmParams p{};
  p.M = M; p.N = N;
  p.num_kb = (int)ceil_div(K, k_elems);
  p.planes = planes; p.NB = NB;
  p.tiles_m = (int)ceil_div(M, kBlockM);
  p.tiles_n = (int)ceil_div(N, NB);
  p.C = C; p.ldc = ldc; p.rs = rs; p.cs = cs; p.scale = scale;
  p.k_elems_per_kb = k_elems;
  p.chunk_kb = kind == 0 ? g_chunk_kb : p.num_kb;  // int32 accumulation is exact: one chunk
  p.mode = 0; p.segs = 1;
  p.n_peers = 0; p.rank_slot = 0; p.n_ranks = 1;
  if (peers) {
    p.n_peers = peers->n_peers; p.rank_slot = peers->rank; p.n_ranks = peers->n_ranks;
    for (int r = 0; r < peers->n_peers; ++r) p.Cpeer[r] = peers->tables[r];
  }
  cudaStream_t stream = (cudaStream_t)stream_;

  const bool pair = (g_use_pair || mode == 2 || peers) && M > kBlockM;
  if (peers && !pair) LGC_FAIL(LGC_ERR_UNSUPPORTED, "gemm bcast: needs more than 128 rows (CTA-pair kernel)");
  if (mode == 2 && !pair) LGC_FAIL(LGC_ERR_UNSUPPORTED, "resource_topk: needs more than 128 rows (CTA-pair kernel)");
  if (pair) {
    // CTA-pair kernel: 256-row tiles, each CTA stages half of the stacked plane rows of B
    CUtensorMap tmBh;
    {
      cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)planes};
      cuuint64_t strides[2] = {(cuuint64_t)ldb * esize, (cuuint64_t)(planes > 1 ? plane_stride : N * ldb) * esize};
      cuuint32_t box[3] = {(cuuint32_t)k_elems, (cuuint32_t)(NB / 2), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmBh, dt, 3, const_cast<void*>(B), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) LGC_FAIL(LGC_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B half) failed with %d", (int)r);
    }
    p.tiles_m = (int)ceil_div(M, 256);
    p.mode = mode;
    int64_t work = (int64_t)p.tiles_m * p.tiles_n;
    if (mode == 1) {
      work = 0;
      for (int n = 0; n < p.tiles_n; ++n) {
        const int64_t c = ((int64_t)(n + 1) * NB - 1) / 256 + 1;
        work += c < p.tiles_m ? c : p.tiles_m;
      }
    } else if (mode == 2) {
      p.segs = topk_segments(M, N, planes);
      work = (int64_t)p.tiles_m * p.segs;
      const size_t rows = (size_t)M * p.segs * kEpiParts;
      LGC_REQUIRE(topk->scratch && topk->scratch_bytes >= topk_scratch_bytes(M, N, planes), "resource_topk: scratch too small");
      LGC_REQUIRE(((uintptr_t)topk->scratch & 255) == 0, "resource_topk: scratch must be 256-byte aligned");
      p.cand = reinterpret_cast<unsigned long long*>(topk->scratch);
      p.cand_cnt = reinterpret_cast<int*>(reinterpret_cast<char*>(topk->scratch) +
                                          align_up(rows * kCandCap * sizeof(unsigned long long), 256));
      p.excl = topk->excl; p.excl_stride = topk->excl_stride; p.excl_row0 = topk->excl_row0; p.k = topk->k;
    }
    int clusters = num_sms() / 2;
    const int64_t my_work = ceil_div(work, (int64_t)p.n_ranks);
    if (my_work < clusters) clusters = (int)(my_work > 0 ? my_work : 1);
    const int grid = 2 * clusters;
    bool launched = false;
#define LGC_GEMM_PAIR_CASE(KD, PL, EP)                                                                          \
  if (kind == KD && planes == PL && (mode == 2) == (EP == 1)) {                                                 \
    static DeviceOnce attr;                                                                                     \
    if (attr.need()) {                                                                                          \
      LGC_CUDA(cudaFuncSetAttribute(umma_gemm_pair_kernel<KD, PL, EP>,                                          \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBytes + 512)));     \
      attr.mark();                                                                                              \
    }                                                                                                           \
    umma_gemm_pair_kernel<KD, PL, EP><<<grid, EP == 1 ? kThreadsTopk : kThreads, kSmemBytes + 512, stream>>>(   \
        tmA, tmB, tmBh, p);                                                                                     \
    launched = true;                                                                                            \
  }
    LGC_GEMM_PAIR_CASE(0, 1, 0) LGC_GEMM_PAIR_CASE(0, 2, 0) LGC_GEMM_PAIR_CASE(0, 3, 0)
    LGC_GEMM_PAIR_CASE(1, 1, 0) LGC_GEMM_PAIR_CASE(1, 2, 0) LGC_GEMM_PAIR_CASE(1, 3, 0) LGC_GEMM_PAIR_CASE(1, 4, 0)
    LGC_GEMM_PAIR_CASE(1, 3, 1) LGC_GEMM_PAIR_CASE(1, 4, 1)
#undef LGC_GEMM_PAIR_CASE
    if (!launched) LGC_FAIL(LGC_ERR_UNSUPPORTED, "gemm: kind %d / %d planes not available in this mode", kind, planes);
    LGC_LAUNCH_CHECK("umma_gemm_pair_kernel");
    if (mode == 2) {
      topk_merge_kernel<kCandCap><<<(unsigned)ceil_div(M, kMergeWarps), kMergeWarps * 32, 0, stream>>>(
          p.cand, p.cand_cnt, M, p.segs * kEpiParts, topk->k, topk->out_idx, topk->out_val);
      LGC_LAUNCH_CHECK("topk_merge_kernel");
    }
    return LGC_OK;
  }

  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
#define LGC_GEMM_CASE(KD, PL)                                                                             \
  if (kind == KD && planes == PL) {                                                                       \
    static DeviceOnce attr;                                                                               \
    if (attr.need()) {                                                                                    \
      LGC_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<KD, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)kSmemBytes));                                                    \
      attr.mark();                                                                                        \
    }                                                                                                     \
    umma_gemm_kernel<KD, PL><<<grid, kThreads, kSmemBytes, stream>>>(tmA, tmB, p);                        \
  }
  LGC_GEMM_CASE(0, 1) LGC_GEMM_CASE(0, 2) LGC_GEMM_CASE(0, 3)
  LGC_GEMM_CASE(1, 1) LGC_GEMM_CASE(1, 2) LGC_GEMM_CASE(1, 3) LGC_GEMM_CASE(1, 4)
#undef LGC_GEMM_CASE
  LGC_LAUNCH_CHECK("umma_gemm_kernel");
  return LGC_OK;
}

}  // namespace umma
}  // namespace lgc

extern "C" int hs_gemm_planes(int32_t kind, const void* A, int64_t lda, const void* B, int64_t ldb,
                              int64_t plane_stride, int32_t planes, int64_t M, int64_t N, int64_t K,
                              float* C, int64_t ldc, const float* rs, const float* cs, double scale,
                              lgc_stream_t stream_) {
  int rc = check_gemm_args(kind, A, lda, B, ldb, plane_stride, planes, M, N, K, C, ldc);
  if (rc) return rc;
  return gemm_planes_impl(kind, A, lda, B, ldb, plane_stride, planes, M, N, K, C, ldc, rs, cs, scale, 0, nullptr, stream_);
}

extern "C" int hs_gemm_planes_sym(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                                  int32_t planes, int64_t N, int64_t K, float* C, int64_t ldc, double scale,
                                  lgc_stream_t stream_) {
  int rc = check_gemm_args(1, A, lda, B, ldb, plane_stride, planes, N, N, K, C, ldc);
  if (rc) return rc;
  return gemm_planes_impl(1, A, lda, B, ldb, plane_stride, planes, N, N, K, C, ldc, nullptr, nullptr, scale, 1, nullptr,
                          stream_);
}

extern "C" int hs_gemm_planes_sym_bcast(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                                        int32_t planes, int64_t N, int64_t K, float* const* store_C_host, int32_t n_store,
                                        int32_t my_rank, int32_t n_ranks, int64_t ldc, double scale, lgc_stream_t stream_) {
  LGC_REQUIRE(store_C_host && n_store >= 1 && n_store <= 8, "gemm bcast: 1..8 store targets");
  LGC_REQUIRE(n_ranks >= 1 && n_ranks <= 64 && my_rank >= 0 && my_rank < n_ranks, "gemm bcast: bad rank");
  for (int r = 0; r < n_store; ++r)
    LGC_REQUIRE(store_C_host[r] && ((uintptr_t)store_C_host[r] & 15) == 0, "gemm bcast: bad store target");
  int rc = check_gemm_args(1, A, lda, B, ldb, plane_stride, planes, N, N, K, store_C_host[0], ldc);
  if (rc) return rc;
  PeerArgs pa{store_C_host, n_store, my_rank, n_ranks};
  return gemm_planes_impl(1, A, lda, B, ldb, plane_stride, planes, N, N, K, store_C_host[0], ldc, nullptr, nullptr, scale, 1,
                          nullptr, stream_, &pa);
}

extern "C" int64_t hs_resource_topk_scratch_bytes(int64_t M, int64_t N, int32_t planes) {
  if (M <= 0 || N <= 0) return 0;
  return (int64_t)topk_scratch_bytes(M, N, planes);
}

extern "C" int hs_resource_topk(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                                int32_t planes, int64_t M, int64_t N, int64_t K, const float* cs, double scale,
                                const uint32_t* excl_mask, int64_t mask_stride_bits, int64_t row_offset, int32_t k,
                                int64_t* out_idx, float* out_val, void* scratch, int64_t scratch_bytes,
                                lgc_stream_t stream_) {
  LGC_REQUIRE(A && B && out_idx && scratch, "resource_topk: null pointer");
  LGC_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldb >= K, "resource_topk: bad extents");
  LGC_REQUIRE(planes == 3 || planes == 4, "resource_topk: 3 or 4 uint8 digit planes");
  LGC_REQUIRE(plane_stride >= N * ldb, "resource_topk: plane stride overlaps planes");
  LGC_REQUIRE(k >= 1 && k <= 32 && k <= N, "resource_topk: k must be in [1, min(32, N)]");
  LGC_REQUIRE(!excl_mask || mask_stride_bits >= N, "resource_topk: mask stride smaller than the row");
  LGC_REQUIRE(((uintptr_t)excl_mask & 3) == 0, "resource_topk: mask must be 4-byte aligned");
  TopkArgs t{excl_mask, mask_stride_bits, row_offset, k, out_idx, out_val, scratch, (size_t)scratch_bytes};
  return gemm_planes_impl(1, A, lda, B, ldb, plane_stride, planes, M, N, K, nullptr, N, nullptr, cs, scale, 2, &t, stream_);
}

extern "C" int hs_gemm_use_cta_pair(int32_t on) {
  g_use_pair = on ? 1 : 0;
  return LGC_OK;
}

extern "C" int hs_gemm_config(int32_t chunk_kb) {
  LGC_REQUIRE(chunk_kb >= 1, "gemm config: chunk_kb must be >= 1");
  g_chunk_kb = chunk_kb;
  return LGC_OK;
}

extern "C" int hs_gemm_planes_simt(int32_t kind, const void* A, int64_t lda, const void* B, int64_t ldb,
                                   int64_t plane_stride, int32_t planes, int64_t M, int64_t N, int64_t K,
                                   float* C, int64_t ldc, const float* rs, const float* cs, double scale,
                                   lgc_stream_t stream) {
  int rc = check_gemm_args(kind, A, lda, B, ldb, plane_stride, planes, M, N, K, C, ldc);
  if (rc) return rc;
  LGC_REQUIRE(M < 65536, "simt gemm: M too large for the cross-check kernel");
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)M);
  if (kind == 0)
    gemm_planes_simt_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, plane_stride, planes, M, N, K,
                                                                        C, ldc, rs, cs, scale);
  else
    gemm_planes_simt_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, plane_stride, planes, M, N, K,
                                                                        C, ldc, rs, cs, scale);
  LGC_LAUNCH_CHECK("gemm_planes_simt_kernel");
  return LGC_OK;
}
