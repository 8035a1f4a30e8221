// score_topk_tc.cu — (P8) full-rank scoring + seen-pair rule + top-k on the 5th-gen tensor cores.
// Stands in for torch.matmul(users_emb, items_emb.T) + score[users, items] = -1024 + torch.topk at
// /root/reference/model/LightGCN/recommend.py:86-114 (== evaluation.py:34-52, SpreadLightGCN/model.py:77-102 and the
// Hadamard fusion F_new = G * F of SpreadLightGCN/model.py:151): a U x M x 64 contraction (617 GFLOP at the
// Amazon-Book shape) whose U x M result must never be written.  lgc_score_topk (score_topk.cu) runs it on the fp32
// FMA pipe at ~32 TFLOP/s; here it runs as a 3xTF32 split on tcgen05:
//
//     x = hi + lo,  hi = x rounded to TF32 (10-bit mantissa),  lo = x - hi (exact in fp32; the MMA reads its top 11 bits)
//     <a, b> ~= <a_hi, b_lo> + <a_lo, b_hi> + <a_hi, b_hi>          (24 tcgen05.mma kind::tf32 per 256 x 256 tile)
//
// Dropped: lo*lo (2^-22 of a term) and the truncation of lo (2^-21 of a term); the small cross terms are accumulated
// FIRST so that the tensor core's truncating fp32 accumulation (tools/probe_umma_numerics.py) acts on them while the
// accumulator is still tiny.  Measured against the fp32-FMA kernel / the fp64 oracle the scores agree to ~1e-6
// relative, inside the stated 1e-5; ids are identical except at float near-ties (north_star).
//
// Kernel shape: one CTA PAIR (cta_group::2, UMMA M = 256) per SM pair, persistent over work items
// (256-user block, segment of the 256-item tiles), 192 threads per CTA:
//   warp 0      TMA producer : the CTA's 128 user rows (hi + lo: 64 KB, ONCE per work item) and, per tile, its half of the
//                              tile's items (128 rows x (hi + lo) = 64 KB) into a 2-stage ring
//   warp 1      MMA issuer   : leader CTA only, 24 MMAs (M 256 x N 256 x K 8) per tile into a double-buffered TMEM
//                              accumulator (2 x 256 columns)
//   warps 2..17 epilogue     : 4 threads per user row (one per 64-column quarter of the tile): tcgen05.ld 16 scores at a
//                              time -> ONE compare of their maximum against the row's shared threshold -> (fusion: multiplier)
//                              -> seen-pair rule through a CURSOR into the row's sorted seen list (the tiles of an item
//                              advance monotonically through the item ids, so no search and no mask matrix is needed)
//                              -> one float compare against the row's current k-th best -> survivors appended to the
//                              row's private candidate buffer; full buffers are compacted by the warp (select.cuh)
// A final warp-per-row kernel merges the segments (topk_merge_kernel).
#include "select.cuh"
#include "umma_common.cuh"

namespace lgc {
namespace umma {

constexpr int kTcTileN = 256;                       // items per tile (MMA N)
constexpr int kTcKbBytes = 128 * 128;               // one K-block tile: 128 rows x 128 B (32 fp32), SWIZZLE_128B
constexpr int kTcStages = 2;
// Epilogue: 16 warps = 4 TMEM lane quarters x 4 COLUMN quarters of the tile.  TMEM reads return 64 B/clk and the MMAs
// of a tile take ~3 k cycles, so the selection must cost only a few instructions per score: with one warp per
// scheduler the 4-warp version was issue-bound at ~16 k cycles per tile.  Four threads share a user row, each ranks
// its own 64 columns of every tile into its own candidate buffer (merged at the end like segments) and they share
// the row's pruning threshold through shared memory.
constexpr int kTcColParts = 4;
constexpr int kTcEpiWarps = 4 * kTcColParts;
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int kTcColsPerPart = kTcTileN / kTcColParts;   // 64

struct ScoreTcParams {
  int64_t u0, u1;                                   // users [u0, u1) of the embedding table
  int n_items, tiles_m, tiles_n, segs, k;
  const int32_t* seen_ptr;                          // CSR over ALL users, ascending item ids per row (may be null)
  const int32_t* seen_idx;
  float fill;
  int exclude_seen;
  const float* mul;                                 // (u1-u0, ldmul) multiplier matrix or null
  int64_t ldmul;
  unsigned long long* cand;                         // [(u1-u0) * segs][kCandCap]
  int* cand_cnt;
};

// x -> (hi, lo): hi = round-to-nearest-even to 10 mantissa bits (TF32), lo = x - hi
__global__ void split_tf32_kernel(const float* __restrict__ X, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 x = *reinterpret_cast<const float4*>(X + i);
  const float xs[4] = {x.x, x.y, x.z, x.w};
  float h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t b = __float_as_uint(xs[j]);
    b += 0x00000fffu + ((b >> 13) & 1u);            // RNE at bit 13
    b &= 0xffffe000u;
    h[j] = __uint_as_float(b);
    l[j] = xs[j] - h[j];
  }
  *reinterpret_cast<float4*>(hi + i) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(lo + i) = make_float4(l[0], l[1], l[2], l[3]);
}

template <int KB /* K-blocks of 32 fp32: dim = 32 * KB */, bool MUL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
score_topk_tc_kernel(const __grid_constant__ CUtensorMap tmapU, const __grid_constant__ CUtensorMap tmapI,
                     const ScoreTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int kABytes = 2 * KB * kTcKbBytes;      // hi + lo of this CTA's 128 users
  constexpr int kBStageBytes = 2 * KB * kTcKbBytes; // hi + lo of this CTA's 128 items of the tile
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kABytes + kTcStages * kBStageBytes);
  uint64_t* a_full = bars;            // leader
  uint64_t* a_empty = bars + 1;       // per CTA
  uint64_t* b_full = bars + 2;        // [kTcStages] leader
  uint64_t* b_empty = bars + 2 + kTcStages;      // [kTcStages] per CTA
  uint64_t* tfull = bars + 2 + 2 * kTcStages;    // [2] per CTA
  uint64_t* tempty = bars + 4 + 2 * kTcStages;   // [2] leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kTcStages);
  // value key (monotone uint32 of the fp32 score) of the best known k-th score of each of the CTA's 128 user rows:
  // the maximum over the four column quarters' own thresholds — a score below it cannot be in the row's top-k
  uint32_t* s_row_thr = tmem_slot + 4;   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const long long n_work = (long long)p.tiles_m * p.segs;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 2);
    mbar_init(a_empty, 1);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&b_full[s], 2); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapU) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapI) : "memory");
  }
  cluster_sync_all();
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
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (long long w = cluster_id; w < n_work; w += n_clusters) {
        const int m_blk = (int)(w / p.segs), seg = (int)(w % p.segs);
        const int n0 = (int)((long long)seg * p.tiles_n / p.segs), n1 = (int)((long long)(seg + 1) * p.tiles_n / p.segs);
        if (n1 <= n0) continue;
        // this CTA's 128 user rows, hi and lo planes, once per work item
        mbar_wait(a_empty, a_phase ^ 1);
        if (leader) mbar_expect_tx(a_full, 2 * (uint32_t)kABytes);
        else mbar_arrive_leader(a_full);
        const int urow = (int)(p.u0 + (int64_t)m_blk * 256 + (int64_t)rank * 128);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma2_load_3d(&tmapU, a_full, smemA + (pl * KB + kb) * kTcKbBytes, kb * 32, urow, pl);
        a_phase ^= 1;
        for (int n = n0; n < n1; ++n) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&b_full[stage], 2 * (uint32_t)kBStageBytes);
          else mbar_arrive_leader(&b_full[stage]);
          uint8_t* dst = smemB + stage * kBStageBytes;
          const int irow = n * kTcTileN + (int)rank * 128;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
              tma2_load_3d(&tmapI, &b_full[stage], dst + (pl * KB + kb) * kTcKbBytes, kb * 32, irow, pl);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader) {
      // D fp32 | A tf32 | B tf32 | K-major both | N = 256 | M = 256
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcTileN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int it = 0;
      for (long long w = cluster_id; w < n_work; w += n_clusters) {
        const int seg = (int)(w % p.segs);
        const int n0 = (int)((long long)seg * p.tiles_n / p.segs), n1 = (int)((long long)(seg + 1) * p.tiles_n / p.segs);
        if (n1 <= n0) continue;
        mbar_wait(a_full, a_phase);
        a_phase ^= 1;
        tc_fence_after();
        for (int n = n0; n < n1; ++n, ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t d_tmem = tmem_base + acc * kAccStride;
            const uint32_t a0 = smem_u32(smemA), b0 = smem_u32(smemB + stage * kBStageBytes);
            uint32_t accum = 0;
            // cross terms first (hi x lo, lo x hi), the dominant hi x hi last
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t a_hi = make_smem_desc(a0 + (0 * KB + kb) * kTcKbBytes);
              const uint64_t a_lo = make_smem_desc(a0 + (1 * KB + kb) * kTcKbBytes);
              const uint64_t b_hi = make_smem_desc(b0 + (0 * KB + kb) * kTcKbBytes);
              const uint64_t b_lo = make_smem_desc(b0 + (1 * KB + kb) * kTcKbBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {          // 32 bytes (8 tf32) of K per MMA: +2 in >>4 units
                mma2_ss<2>(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, accum);
                accum = 1;
                mma2_ss<2>(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
              }
            }
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const uint64_t a_hi = make_smem_desc(a0 + (0 * KB + kb) * kTcKbBytes);
              const uint64_t b_hi = make_smem_desc(b0 + (0 * KB + kb) * kTcKbBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma2_ss<2>(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
            }
            tc_commit2_mc(&b_empty[stage]);
            tc_commit2_mc(&tfull[acc]);
            if (n == n1 - 1) tc_commit2_mc(a_empty);   // the user rows may be replaced once these MMAs retire
          }
          __syncwarp();
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..17, both CTAs) ------------------------------
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int cpart = (warp - 2) >> 2;            // column quarter of every tile this warp ranks
    const int rloc = q * 32 + lane;               // row inside the CTA's 128
    int it = 0;
    for (long long w = cluster_id; w < n_work; w += n_clusters) {
      const int m_blk = (int)(w / p.segs), seg = (int)(w % p.segs);
      const int n0 = (int)((long long)seg * p.tiles_n / p.segs), n1 = (int)((long long)(seg + 1) * p.tiles_n / p.segs);
      if (n1 <= n0) continue;
      const int64_t lrow = (int64_t)m_blk * 256 + (int64_t)rank * 128 + rloc;   // row inside [0, u1-u0)
      const int64_t urow = p.u0 + lrow;
      const bool row_ok = urow < p.u1;
      unsigned long long sel_thr = 0ull;
      int sel_cnt = 0;
      const int vseg = seg * kTcColParts + cpart;                               // (segment, column quarter) = merge unit
      unsigned long long* sel_buf = p.cand + ((size_t)(row_ok ? lrow : 0) * (p.segs * kTcColParts) + vseg) * kCandCap;
      // all epilogue warps of the CTA start the item together: reset the shared row thresholds
      asm volatile("bar.sync 1, %0;" ::"r"(32 * kTcEpiWarps) : "memory");
      if (cpart == 0) s_row_thr[rloc] = 0u;
      asm volatile("bar.sync 1, %0;" ::"r"(32 * kTcEpiWarps) : "memory");
      // cursor into the row's sorted seen list: first entry >= the segment's first item
      int s_cur = 0, s_end = 0;
      if (p.seen_ptr && row_ok) {
        s_cur = __ldg(p.seen_ptr + urow);
        s_end = __ldg(p.seen_ptr + urow + 1);
        if (n0 > 0) {
          const int first_item = n0 * kTcTileN;
          int lo = s_cur, hi = s_end;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(p.seen_idx + mid) < first_item) lo = mid + 1; else hi = mid;
          }
          s_cur = lo;
        }
      }
      int next_seen = s_cur < s_end ? __ldg(p.seen_idx + s_cur) : 0x7fffffff;
      const float* mrow = MUL ? p.mul + (row_ok ? lrow : 0) * p.ldmul : nullptr;

      for (int n = n0; n < n1; ++n, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        // pruning threshold of this tile: the row's best known k-th score over all four column quarters (rows past
        // the end never pass)
        const uint32_t tk = s_row_thr[rloc];
        const float thr_f = row_ok ? (tk ? key_float(tk) : -INFINITY) : INFINITY;
        const int cbase = n * kTcTileN + cpart * kTcColsPerPart;                // first column of this thread's slice
        // seen items of the other column quarters of the previous / this tile are skipped, not patched
        while (next_seen < cbase) {
          ++s_cur;
          next_seen = s_cur < s_end ? __ldg(p.seen_idx + s_cur) : 0x7fffffff;
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + cpart * kTcColsPerPart;
#pragma unroll
        for (int c1 = 0; c1 < kTcColsPerPart; c1 += 16) {
          const int col0 = cbase + c1;
          uint32_t r[16];
          tmem_ld16(t_row + c1, r);
          float mv[16];
          if (MUL) {
            if (row_ok && col0 + 16 <= p.n_items && (p.ldmul & 3) == 0 && ((uintptr_t)p.mul & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow + col0 + j));
                mv[j] = m4.x; mv[j + 1] = m4.y; mv[j + 2] = m4.z; mv[j + 3] = m4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) mv[j] = (row_ok && col0 + j < p.n_items) ? __ldg(mrow + col0 + j) : 0.f;
            }
          }
          tmem_ld_wait();
          float out[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) out[j] = MUL ? __uint_as_float(r[j]) * mv[j] : __uint_as_float(r[j]);
          // seen pairs inside these 16 columns (rare): fill value (x multiplier) or NaN = never selected
          while (next_seen < col0 + 16) {
            const int js = next_seen - col0;
            const float v = p.exclude_seen ? __int_as_float(0x7fc00000) : p.fill;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j == js) out[j] = (MUL && !p.exclude_seen) ? v * mv[j] : v;
            ++s_cur;
            next_seen = s_cur < s_end ? __ldg(p.seen_idx + s_cur) : 0x7fffffff;
          }
          // fast path: ONE compare per 16 scores (their maximum against the threshold; fmaxf ignores the NaN marks)
          float mx = fmaxf(fmaxf(fmaxf(out[0], out[1]), fmaxf(out[2], out[3])), fmaxf(fmaxf(out[4], out[5]), fmaxf(out[6], out[7])));
          mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(out[8], out[9]), fmaxf(out[10], out[11])),
                               fmaxf(fmaxf(out[12], out[13]), fmaxf(out[14], out[15]))));
          if (mx >= thr_f) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (out[j] >= thr_f && col0 + j < p.n_items) {
                const unsigned long long key = make_key(float_key(out[j]), (uint32_t)(col0 + j));
                if (key > sel_thr) sel_buf[sel_cnt++] = key;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[acc]);   // accumulator free: the next tile's MMAs may start
        // a thread offers at most 64 scores per tile: rows that could not take another 64 survivors are compacted now,
        // one row at a time, by the whole warp; the new k-th best is published for the row's other column quarters
        uint32_t need = __ballot_sync(0xffffffffu, sel_cnt + kTcColsPerPart > kCandCap);
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
      }
      if (row_ok) p.cand_cnt[(size_t)lrow * (p.segs * kTcColParts) + vseg] = sel_cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn2)ptr;
  }
  return fn;
}

static int tc_segments(int64_t rows, int64_t n_items) {
  const int64_t tiles_m = ceil_div(rows, 256), tiles_n = ceil_div(n_items, kTcTileN);
  const int clusters = num_sms() / 2;
  int64_t segs = tiles_m >= clusters ? 1 : clusters / tiles_m;
  if (segs > tiles_n) segs = tiles_n;
  return (int)(segs < 1 ? 1 : segs);
}

struct TcWs {
  size_t xu, xi, cand, cnt, total;
};

static TcWs tc_layout(int64_t n_users_total, int64_t n_items, int64_t rows, int dim) {
  TcWs w{};
  size_t off = 0;
  w.xu = off; off += align_up((size_t)2 * n_users_total * dim * sizeof(float), 256);
  w.xi = off; off += align_up((size_t)2 * n_items * dim * sizeof(float), 256);
  const size_t rs = (size_t)rows * tc_segments(rows, n_items) * kTcColParts;
  w.cand = off; off += align_up(rs * kCandCap * sizeof(unsigned long long), 256);
  w.cnt = off; off += align_up(rs * sizeof(int), 256);
  w.total = off;
  return w;
}

}  // namespace umma
}  // namespace lgc

using namespace lgc;
using namespace lgc::umma;

extern "C" int64_t lgc_score_topk_tc_workspace_bytes(int64_t n_users_total, int64_t n_items, int64_t rows, int32_t dim) {
  if (n_users_total <= 0 || n_items <= 0 || rows <= 0 || dim <= 0) return 0;
  return (int64_t)tc_layout(n_users_total, n_items, rows, dim).total;
}

extern "C" int lgc_score_topk_tc(const float* Xu, const float* Xi, int64_t n_users_total, int64_t u0, int64_t u1,
                                 int64_t n_items, int32_t dim, const int32_t* seen_ptr, const int32_t* seen_idx,
                                 float fill, int32_t exclude_seen, const float* mul, int64_t ldmul, int32_t k,
                                 int64_t* out_idx, float* out_val, void* workspace, int64_t workspace_bytes,
                                 lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(Xu && Xi && out_idx && workspace, "score_topk_tc: null pointer");
  LGC_REQUIRE(0 <= u0 && u0 < u1 && u1 <= n_users_total && n_items > 0 && n_items < (1ll << 31) - 512 &&
                  n_users_total < (1ll << 31) - 512,
              "score_topk_tc: bad extents");
  LGC_REQUIRE(dim == 32 || dim == 64, "score_topk_tc: embedding dim must be 32 or 64");
  LGC_REQUIRE(k >= 1 && k <= 32 && k <= n_items, "score_topk_tc: k must be in [1, min(32, n_items)]");
  LGC_REQUIRE(((uintptr_t)Xu & 15) == 0 && ((uintptr_t)Xi & 15) == 0 && ((uintptr_t)workspace & 255) == 0,
              "score_topk_tc: embeddings must be 16-byte aligned, the workspace 256-byte aligned");
  LGC_REQUIRE((seen_ptr == nullptr) == (seen_idx == nullptr), "score_topk_tc: seen_ptr / seen_idx mismatch");
  LGC_REQUIRE(!mul || ldmul >= n_items, "score_topk_tc: multiplier leading dimension smaller than the row");
  const int64_t rows = u1 - u0;
  const TcWs ws = tc_layout(n_users_total, n_items, rows, dim);
  LGC_REQUIRE((size_t)workspace_bytes >= ws.total, "score_topk_tc: workspace too small");
  EncodeTiledFn2 encode = encode_fn();
  if (!encode) LGC_FAIL(LGC_ERR_CUDA, "score_topk_tc: cuTensorMapEncodeTiled entry point not available");
  char* base = reinterpret_cast<char*>(workspace);
  float* xu = reinterpret_cast<float*>(base + ws.xu);
  float* xi = reinterpret_cast<float*>(base + ws.xi);

  // 1. hi / lo planes of both tables (plane-major: [hi rows | lo rows])
  {
    const int64_t nu = n_users_total * dim, ni = n_items * dim;
    split_tf32_kernel<<<(unsigned)ceil_div(nu / 4, 256), 256, 0, stream>>>(Xu, nu, xu, xu + nu);
    LGC_LAUNCH_CHECK("split_tf32_kernel");
    split_tf32_kernel<<<(unsigned)ceil_div(ni / 4, 256), 256, 0, stream>>>(Xi, ni, xi, xi + ni);
    LGC_LAUNCH_CHECK("split_tf32_kernel");
  }
  // 2. tensor maps: (dim, rows, 2 planes), box 32 fp32 x 128 rows, SWIZZLE_128B
  CUtensorMap tmU, tmI;
  for (int which = 0; which < 2; ++which) {
    const int64_t nrows = which == 0 ? n_users_total : n_items;
    cuuint64_t dims[3] = {(cuuint64_t)dim, (cuuint64_t)nrows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)dim * 4, (cuuint64_t)nrows * dim * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(which == 0 ? &tmU : &tmI, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, which == 0 ? (void*)xu : (void*)xi,
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LGC_FAIL(LGC_ERR_CUDA, "score_topk_tc: cuTensorMapEncodeTiled failed with %d", (int)r);
  }
  ScoreTcParams p{};
  p.u0 = u0; p.u1 = u1; p.n_items = (int)n_items;
  p.tiles_m = (int)ceil_div(rows, 256);
  p.tiles_n = (int)ceil_div(n_items, kTcTileN);
  p.segs = tc_segments(rows, n_items);
  p.k = k;
  p.seen_ptr = seen_ptr; p.seen_idx = seen_idx; p.fill = fill; p.exclude_seen = exclude_seen;
  p.mul = mul; p.ldmul = ldmul;
  p.cand = reinterpret_cast<unsigned long long*>(base + ws.cand);
  p.cand_cnt = reinterpret_cast<int*>(base + ws.cnt);
  const long long work = (long long)p.tiles_m * p.segs;
  int clusters = num_sms() / 2;
  if (work < clusters) clusters = (int)work;
  const int grid = 2 * clusters;
#define LGC_TC_LAUNCH(KBV, MULV)                                                                             \
  do {                                                                                                       \
    constexpr size_t smem = (size_t)(2 * KBV + kTcStages * 2 * KBV) * kTcKbBytes + 1024 + 256 + 512;         \
    static DeviceOnce attr;                                                                                  \
    if (attr.need()) {                                                                                       \
      LGC_CUDA(cudaFuncSetAttribute(score_topk_tc_kernel<KBV, MULV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                             \
      attr.mark();                                                                                           \
    }                                                                                                        \
    score_topk_tc_kernel<KBV, MULV><<<grid, kTcThreads, smem, stream>>>(tmU, tmI, p);                        \
  } while (0)
  if (dim == 64) { if (mul) LGC_TC_LAUNCH(2, true); else LGC_TC_LAUNCH(2, false); }
  else { if (mul) LGC_TC_LAUNCH(1, true); else LGC_TC_LAUNCH(1, false); }
#undef LGC_TC_LAUNCH
  LGC_LAUNCH_CHECK("score_topk_tc_kernel");
  topk_merge_kernel<kCandCap><<<(unsigned)ceil_div(rows, kMergeWarps), kMergeWarps * 32, 0, stream>>>(
      p.cand, p.cand_cnt, rows, p.segs * kTcColParts, k, out_idx, out_val);
  LGC_LAUNCH_CHECK("topk_merge_kernel");
  return LGC_OK;
}
