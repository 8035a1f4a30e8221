// spmm.cu — (P2/P3) one LightGCN propagation layer as a CSR SpMM fused with the running
// layer sum:   Y = alpha * (A_hat X + beta X0).
// Stands in for MessagePassing.propagate + message at
// /root/reference/model/LightGCN/model.py:61-63,76-84 (index_select -> norm*x_j ->
// scatter_add, three library kernels and a materialised (2E, D) message tensor) and for
// the stack/mean at :66-69 (Horner form, see lgc_propagate_mean).
//
// Mapping (HBM/L2-bound integer+fp32 gather work, no tensor cores):
//   * an embedding row is DIM fp32 = DIM/4 float4; DIM/4 lanes ("sub-group") own one
//     gathered row, so a warp keeps 32/(DIM/4) rows in flight per load instruction and
//     every LDG.128 of a sub-group is one fully used, 128-bit-per-lane coalesced segment;
//   * the (colidx, val) stream of a row is read coalesced, 32 entries per warp load, with
//     streaming (evict-first) hints, and broadcast with shuffles — the per-non-zero
//     metadata is staged in registers, never re-read;
//   * short rows (<= LGC_LONG_ROW non-zeros): one warp per row; long rows are cut into
//     LGC_CHUNK-sized chunks, one CTA per chunk; the last CTA of a row (ticket counter)
//     adds the chunk partials in chunk order, so the result is deterministic — there is no
//     floating-point atomic anywhere;
//   * the BCAST variant stores each finished row into every peer's replica (NVLink P2P
//     stores): the per-layer all-gather of the row-partitioned multi-GPU path is fused into
//     the SpMM epilogue.
#include "common.cuh"

namespace lgc {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxPeers = 8;

static int g_spmm_long_row = 0;  // override of the per-call warp-per-row threshold (lgc_spmm_long_row); 0 = none
static int g_coop_ctas_per_sm = 2;  // resident CTAs per SM of the cooperative K-layer kernel (lgc_coop_config)
static int g_spmm_unroll = 0;  // gathers in flight per lane for DIM=64; 0 = choose by grid size (lgc_spmm_config)

struct PeerPtrs {
  float* y[kMaxPeers];
};

__device__ __forceinline__ float4 ld_row4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// Sum of val[e] * X[colidx[e], :] over e in [start, end) for one warp.  On return every
// lane li of every sub-group holds the full sum for columns [4*li, 4*li+4).
// UN independent 128-bit gathers are in flight per lane, and the (colidx, val) metadata of the
// next 32 non-zeros is requested before the gathers of the current 32 are issued, so the
// dependent metadata -> gather chain is overlapped.
// MASK: `src_mask` has one bit per source row, clear = that row of X is all-zero.  Such rows are not gathered (their
// terms are exact zeros, so the sum is unchanged) and groups of SUB*UN non-zeros without a live source are skipped
// altogether: the first layer of the gradient propagation, whose input has <= 3 * batch non-zero rows, then costs the
// (colidx, val) stream instead of one 4*DIM-byte gather per non-zero.
template <int DIM, int UN, bool MASK = false>
__device__ __forceinline__ float4 warp_gather_sum(const int32_t* __restrict__ colidx,
                                                  const float* __restrict__ val,
                                                  const float* __restrict__ X, int start, int end,
                                                  int lane, const uint32_t* __restrict__ src_mask = nullptr) {
  constexpr int LPR = DIM / 4;   // lanes per gathered row
  constexpr int SUB = 32 / LPR;  // rows in flight per warp-wide load
  const int sub = lane / LPR, li = lane % LPR;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int c = 0;
  float v = 0.f;
  if (start + lane < end) {
    c = __ldcs(colidx + start + lane);
    v = __ldcs(val + start + lane);
  }
  for (int base = start; base < end; base += 32) {
    int c_next = 0;
    float v_next = 0.f;
    if (base + 32 + lane < end) {
      c_next = __ldcs(colidx + base + 32 + lane);
      v_next = __ldcs(val + base + 32 + lane);
    }
    const int n = min(32, end - base);
    unsigned live = 0xffffffffu;
    if (MASK) {
      const bool on = base + lane < end && ((__ldg(src_mask + (c >> 5)) >> (c & 31)) & 1u);
      live = __ballot_sync(0xffffffffu, on);
    }
    for (int j = 0; j < n; j += SUB * UN) {
      if (MASK && ((live >> j) & ((SUB * UN >= 32) ? 0xffffffffu : ((1u << (SUB * UN)) - 1u))) == 0u) continue;
      float4 x[UN];
      float w[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int jj = j + u * SUB + sub;
        const int cc = __shfl_sync(0xffffffffu, c, jj & 31);
        const float vv = __shfl_sync(0xffffffffu, v, jj & 31);
        const bool ok = jj < n && (!MASK || ((live >> (jj & 31)) & 1u));
        w[u] = ok ? vv : 0.f;
        x[u] = ok ? ld_row4(X + (size_t)cc * DIM + li * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        acc.x = fmaf(w[u], x[u].x, acc.x);
        acc.y = fmaf(w[u], x[u].y, acc.y);
        acc.z = fmaf(w[u], x[u].z, acc.z);
        acc.w = fmaf(w[u], x[u].w, acc.w);
      }
    }
    c = c_next;
    v = v_next;
  }
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  return acc;
}

template <int NPEER>
__device__ __forceinline__ void store_row4(float* Y, const PeerPtrs& peers, size_t off, float4 v) {
  if (NPEER == 0) {
    *reinterpret_cast<float4*>(Y + off) = v;
  } else {
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < NPEER) *reinterpret_cast<float4*>(peers.y[p] + off) = v;
  }
}

template <int DIM, int NPEER, int UN, bool MASK = false>
__global__ void __launch_bounds__(kThreads)
spmm_layer_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                  const float* __restrict__ val, const int32_t* __restrict__ chunk_row,
                  const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ row_chunk_base,
                  int chunk_begin, int n_chunk_blocks, int chunk_begin2, int n_chunk_blocks1, int row_begin, int row_end,
                  const int32_t* __restrict__ row_order,
                  const float* __restrict__ X, const float* __restrict__ X0, float alpha, float beta,
                  float* __restrict__ Y, PeerPtrs peers, float* __restrict__ partial,
                  int32_t* __restrict__ counters, int long_row, const uint32_t* __restrict__ src_mask) {
  constexpr int LPR = DIM / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if ((int)blockIdx.x >= n_chunk_blocks) {
    // ---------------- short rows: one warp per row ----------------
    // row_order (optional): the rows of [row_begin, row_end) longest first, so that the eight rows of a CTA have
    // similar lengths (its slot is not held by one straggler) and the last wave consists of the shortest rows
    const int slot = row_begin + ((int)blockIdx.x - n_chunk_blocks) * kWarpsPerBlock + warp;
    if (slot >= row_end) return;
    const int row = row_order ? __ldg(row_order + (slot - row_begin)) : slot;
    const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    if (end - start > long_row) return;  // handled by the chunk CTAs
    float4 acc = warp_gather_sum<DIM, UN, MASK>(colidx, val, X, start, end, lane, src_mask);
    if (lane < LPR) {
      const size_t off = (size_t)row * DIM + lane * 4;
      if (beta != 0.f) {
        const float4 x0 = ld_row4(X0 + off);
        acc.x = fmaf(beta, x0.x, acc.x);
        acc.y = fmaf(beta, x0.y, acc.y);
        acc.z = fmaf(beta, x0.z, acc.z);
        acc.w = fmaf(beta, x0.w, acc.w);
      }
      acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
      store_row4<NPEER>(Y, peers, off, acc);
    }
    return;
  }

  // ---------------- long rows: one CTA per LGC_CHUNK non-zeros ----------------
  __shared__ float s_part[kWarpsPerBlock][DIM];
  __shared__ int s_last;
  // the chunk CTAs cover up to two ranges of the chunk list (a rank's user rows and item rows in ONE launch)
  const int chunk = (int)blockIdx.x < n_chunk_blocks1 ? chunk_begin + (int)blockIdx.x
                                                      : chunk_begin2 + ((int)blockIdx.x - n_chunk_blocks1);
  const int row = __ldg(chunk_row + chunk);
  const int cstart = __ldg(chunk_start + chunk);
  const int rstart = __ldg(rowptr + row), rend = __ldg(rowptr + row + 1);
  if (rend - rstart <= long_row) return;  // this launch handles the row in the warp-per-row path
  const int cend = min(cstart + LGC_CHUNK, rend);
  constexpr int PER_WARP = LGC_CHUNK / kWarpsPerBlock;
  const int wstart = min(cstart + warp * PER_WARP, cend), wend = min(wstart + PER_WARP, cend);
  float4 acc = warp_gather_sum<DIM, UN, MASK>(colidx, val, X, wstart, wend, lane, src_mask);
  if (lane < LPR) *reinterpret_cast<float4*>(&s_part[warp][lane * 4]) = acc;
  __syncthreads();
  const int nch = (rend - rstart + LGC_CHUNK - 1) / LGC_CHUNK;
  const int chunk0 = __ldg(row_chunk_base + row);
  float sum = 0.f;
  if (threadIdx.x < DIM) {
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) sum += s_part[w][threadIdx.x];
  }
  if (nch > 1) {
    if (threadIdx.x < DIM) {
      __stcg(partial + (size_t)chunk * DIM + threadIdx.x, sum);
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int ticket = atomicAdd(counters + row, 1);
      s_last = (ticket == nch - 1);
      if (s_last) counters[row] = 0;  // leave the counters clean for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < DIM) {
      sum = 0.f;
      for (int c = 0; c < nch; ++c) sum += __ldcg(partial + (size_t)(chunk0 + c) * DIM + threadIdx.x);
    }
  }
  if (threadIdx.x < DIM) {
    const size_t off = (size_t)row * DIM + threadIdx.x;
    if (beta != 0.f) sum = fmaf(beta, __ldg(X0 + off), sum);
    sum *= alpha;
    if (NPEER == 0) {
      Y[off] = sum;
    } else {
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p)
        if (p < NPEER) peers.y[p][off] = sum;
    }
  }
}

// Device-side barrier over peer memory, replacing one NCCL all-reduce per layer in the multi-GPU p2p mode.
// Launched on the stream right after the layer's SpMM: the kernel boundary has completed that rank's row stores
// (local and peer); lane p then publishes `epoch` into peer p's flag array (release, system scope) and waits until
// every peer has published it into ours (acquire).  Epochs only grow, so the flags are never reset.  One rank per
// GPU: the kernels that wait on each other always run concurrently.
struct PeerFlags {
  int32_t* f[kMaxPeers];
};

__global__ void peer_barrier_kernel(const int32_t* __restrict__ local_flags, PeerFlags peers, int my_rank, int n_peers,
                                    int32_t epoch) {
  const int p = threadIdx.x;
  if (p < n_peers) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peers.f[p] + my_rank), "r"(epoch) : "memory");
    const long long t0 = clock64();
    int32_t v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(local_flags + p) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a missing peer must trap instead of hanging the GPU
        printf("lgcnhs peer barrier: rank %d never saw peer %d reach epoch %d\n", my_rank, p, epoch);
        __trap();
      }
      __nanosleep(32);
    } while (true);
  }
}

// Same barrier with the epoch kept in device memory (incremented by the kernel itself), so that a sequence of
// propagation layers + barriers can be captured in a CUDA graph and replayed: every rank replays the same sequence, so
// the epochs stay aligned without any host-side counter.
__global__ void peer_barrier_dev_kernel(const int32_t* __restrict__ local_flags, PeerFlags peers, int my_rank, int n_peers,
                                        int32_t* __restrict__ epoch_counter) {
  __shared__ int32_t s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = *epoch_counter + 1;
    *epoch_counter = s_epoch;
  }
  __syncthreads();
  const int32_t epoch = s_epoch;
  const int p = threadIdx.x;
  if (p < n_peers) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peers.f[p] + my_rank), "r"(epoch) : "memory");
    const long long t0 = clock64();
    int32_t v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(local_flags + p) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 20000000000ll) {
        printf("lgcnhs peer barrier: rank %d never saw peer %d reach epoch %d\n", my_rank, p, epoch);
        __trap();
      }
      __nanosleep(32);
    } while (true);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Small graphs: ALL K layers (+ the layer mean) in ONE cooperative launch.
// At the ML-100K / Douban shapes a layer moves 0.4 us worth of compulsory bytes; as separate launches each layer costs
// ~15 us (grid launch + drain + the serial (colidx, val) -> gather latency chain of the longest row).  Here the grid is
// resident for the whole call (cudaLaunchCooperativeKernel), work is cut by the host into warp-sized UNITS of at most 128
// non-zeros — a row, or a piece of a long row whose partial sums are combined in unit order after a grid barrier
// (deterministic, no float atomics) — and layers are separated by grid barriers instead of launches.
// Measured (profiles/r2_coop_probe.txt): with cooperative_groups' grid.sync the call was 2-3x slower than three launches;
// with the hand-written barrier below it is ~10 % faster at ML-100K and equal at the Douban shape, so it stays opt-in.
// ------------------------------------------------------------------------------------------------------------------

struct CoopParams {
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* val;
  const int32_t* unit_row;    // [n_units] row of the unit
  const int32_t* unit_start;  // [n_units] first non-zero
  const int32_t* unit_end;    // [n_units] one past the last non-zero
  const int32_t* unit_slot;   // [n_units] partial-sum slot, or -1 when the unit is a whole row
  const int32_t* split_row;   // [n_split] rows cut into several units ...
  const int32_t* split_first; // [n_split] ... their first partial slot ...
  const int32_t* split_count; // [n_split] ... and number of partials
  int n_units, n_split, n_layers;
  const float* X0;
  float* buf0;
  float* buf1;
  float* E;
  float* partial;             // [n_partials][DIM]
  unsigned int* barrier;      // [2] arrival counter of the grid barrier + exit counter, zero between launches
};

// Grid barrier of the cooperative kernel: one arrival counter that only grows during a launch (barrier b completes when
// it reaches (b + 1) * gridDim.x); the last CTA to leave the kernel resets it, so every launch starts from zero and the
// launch can be replayed from a CUDA graph.  Thread 0 of every CTA arrives with a release and spins with acquire loads;
// co-residency of the whole grid is guaranteed by cudaLaunchCooperativeKernel.
__device__ __forceinline__ void coop_grid_barrier(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * gridDim.x;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

template <int DIM>
__global__ void __launch_bounds__(kThreads)
propagate_coop_kernel(const CoopParams p) {
  unsigned int epoch = 0;
  constexpr int LPR = DIM / 4;
  const int lane = threadIdx.x & 31;
  const int gwarp = (int)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int n_warps = (int)gridDim.x * kWarpsPerBlock;
  const float* cur = p.X0;
  for (int l = 0; l < p.n_layers; ++l) {
    const bool last = l == p.n_layers - 1;
    float* out = last ? p.E : ((l & 1) ? p.buf1 : p.buf0);
    const float alpha = last ? 1.0f / (float)(p.n_layers + 1) : 1.0f;
    for (int u = gwarp; u < p.n_units; u += n_warps) {
      const int row = __ldg(p.unit_row + u), slot = __ldg(p.unit_slot + u);
      float4 acc = warp_gather_sum<DIM, 4>(p.colidx, p.val, cur, __ldg(p.unit_start + u), __ldg(p.unit_end + u), lane);
      if (lane < LPR) {
        if (slot < 0) {
          const size_t off = (size_t)row * DIM + lane * 4;
          const float4 x0 = ld_row4(p.X0 + off);
          acc.x = alpha * (acc.x + x0.x); acc.y = alpha * (acc.y + x0.y);
          acc.z = alpha * (acc.z + x0.z); acc.w = alpha * (acc.w + x0.w);
          *reinterpret_cast<float4*>(out + off) = acc;
        } else {
          __stcg(reinterpret_cast<float4*>(p.partial + (size_t)slot * DIM + lane * 4), acc);
        }
      }
    }
    if (p.n_split > 0) {
      coop_grid_barrier(p.barrier, epoch);
      for (int s = gwarp; s < p.n_split; s += n_warps) {
        if (lane < LPR) {
          const int row = __ldg(p.split_row + s), first = __ldg(p.split_first + s), cnt = __ldg(p.split_count + s);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          // partials are added in slot order (deterministic); eight loads are in flight at a time so that a row cut into
          // dozens of pieces does not pay one L2 round trip per piece
          for (int c0 = 0; c0 < cnt; c0 += 8) {
            float4 q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              q[j] = c0 + j < cnt ? __ldcg(reinterpret_cast<const float4*>(p.partial + (size_t)(first + c0 + j) * DIM + lane * 4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc.x += q[j].x; acc.y += q[j].y; acc.z += q[j].z; acc.w += q[j].w; }
          }
          const size_t off = (size_t)row * DIM + lane * 4;
          const float4 x0 = ld_row4(p.X0 + off);
          acc.x = alpha * (acc.x + x0.x); acc.y = alpha * (acc.y + x0.y);
          acc.z = alpha * (acc.z + x0.z); acc.w = alpha * (acc.w + x0.w);
          *reinterpret_cast<float4*>(out + off) = acc;
        }
      }
    }
    if (!last) coop_grid_barrier(p.barrier, epoch);
    cur = out;
  }
  // every CTA that gets here has passed the last barrier: the last one to leave puts the counters back to zero
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(p.barrier + 1, 1u) == gridDim.x - 1) {
    p.barrier[0] = 0u;
    p.barrier[1] = 0u;
    __threadfence();
  }
}

template <int NPEER>
static int launch_spmm(const int32_t* rowptr, const int32_t* colidx, const float* val,
                       const int32_t* chunk_row, const int32_t* chunk_start,
                       const int32_t* row_chunk_base, int chunk_begin, int chunk_end,
                       int64_t row_begin, int64_t row_end, const int32_t* row_order, int long_row_arg, int dim,
                       const float* X, const float* X0, float alpha, float beta, float* Y, const PeerPtrs& peers,
                       float* partial, int32_t* counters, cudaStream_t stream, int chunk_begin2 = 0, int chunk_end2 = 0,
                       const uint32_t* src_mask = nullptr) {
  const int n_chunk_blocks1 = chunk_end - chunk_begin;
  const int n_chunk_blocks = n_chunk_blocks1 + (chunk_end2 - chunk_begin2);
  const int64_t n_rows = row_end - row_begin;
  const int64_t grid = n_chunk_blocks + ceil_div(n_rows, kWarpsPerBlock);
  if (grid == 0) return LGC_OK;
#define LGC_SPMM_LAUNCH_M(D, UNR, MASKED)                                                         \
  spmm_layer_kernel<D, NPEER, UNR, MASKED><<<(unsigned)grid, kThreads, 0, stream>>>(              \
      rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, n_chunk_blocks,   \
      chunk_begin2, n_chunk_blocks1, (int)row_begin, (int)row_end, row_order, X, X0, alpha, beta, Y, peers, partial, \
      counters, long_row, src_mask)
#define LGC_SPMM_LAUNCH(D, UNR) LGC_SPMM_LAUNCH_M(D, UNR, false)
  // Rows of LGC_LONG_ROW < nnz <= long_row take the warp-per-row path in this launch: it is cheaper than chunk
  // partials + ticket + re-read (ML-20M layer: 680 -> 525 us at long_row 1024), but a lone warp streams only ~40
  // non-zeros per microsecond, so only launches with enough work hide such rows, and they must be issued first
  // (row_order).  The caller sizes long_row to the launch (about 1e-4 x its non-zeros, lgcnhs_b200/ops.py).
  // A row's fp32 summation order depends on the path it takes, so results are reproducible for a given launch
  // configuration and may differ in the last bits between configurations (1 GPU vs a row partition).
  int long_row = g_spmm_long_row ? g_spmm_long_row : long_row_arg;
  if (long_row < LGC_LONG_ROW || !row_order) long_row = LGC_LONG_ROW;
  if (long_row > 2048) long_row = 2048;
  if (src_mask) {
    // sparse-source layer: few gathers survive the mask, so the deeper unroll costs nothing and one variant is enough
    switch (dim) {
      case 32: LGC_SPMM_LAUNCH_M(32, 4, true); break;
      case 64: LGC_SPMM_LAUNCH_M(64, 4, true); break;
      default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "spmm (masked source): embedding dim %d not in {32,64}", dim);
    }
    LGC_LAUNCH_CHECK("spmm_layer_kernel (masked source)");
    return LGC_OK;
  }
  switch (dim) {
    case 32: LGC_SPMM_LAUNCH(32, 4); break;
    case 64:
      {
        // measured on B200 (tools/spmm_tune.py, tools/partition_probe.py): with more than four full waves of CTAs,
        // occupancy (32 regs, 64 warps/SM) beats per-warp ILP; below that the deeper unroll wins (per-rank launches of
        // the 4- and 8-GPU partitions: 77 vs 88 us at 4 M non-zeros)
        int un = g_spmm_unroll;
        if (un == 0) un = grid > (int64_t)num_sms() * 8 * 4 ? 2 : 4;
        if (un == 8) LGC_SPMM_LAUNCH(64, 8);
        else if (un == 2) LGC_SPMM_LAUNCH(64, 2);
        else LGC_SPMM_LAUNCH(64, 4);
      }
      break;
    case 128: LGC_SPMM_LAUNCH(128, 4); break;
    default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "spmm: embedding dim %d not in {32,64,128}", dim);
  }
#undef LGC_SPMM_LAUNCH
#undef LGC_SPMM_LAUNCH_M
  LGC_LAUNCH_CHECK("spmm_layer_kernel");
  return LGC_OK;
}

static int check_spmm_args(const int32_t* rowptr, const int32_t* colidx, const float* val,
                           const int32_t* chunk_row, const int32_t* chunk_start,
                           const int32_t* row_chunk_base, int chunk_begin, int chunk_end,
                           int64_t n_nodes, int64_t row_begin, int64_t row_end, const float* X,
                           const float* X0, float beta, float* partial, int32_t* counters) {
  LGC_REQUIRE(rowptr && X, "spmm: null rowptr / X");
  LGC_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= n_nodes, "spmm: bad row range");
  LGC_REQUIRE(0 <= chunk_begin && chunk_begin <= chunk_end, "spmm: bad chunk range");
  LGC_REQUIRE(chunk_end == chunk_begin || (chunk_row && chunk_start && row_chunk_base && partial && counters),
              "spmm: chunk list given without chunk arrays / scratch");
  LGC_REQUIRE(beta == 0.f || X0, "spmm: beta != 0 needs X0");
  LGC_REQUIRE(((uintptr_t)X & 15) == 0 && ((uintptr_t)X0 & 15) == 0, "spmm: X / X0 must be 16-byte aligned");
  (void)colidx; (void)val;
  return LGC_OK;
}

}  // namespace lgc

using namespace lgc;

extern "C" int lgc_spmm_config(int32_t unroll) {
  LGC_REQUIRE(unroll == 0 || unroll == 2 || unroll == 4 || unroll == 8, "spmm config: unroll must be 0 (auto), 2, 4 or 8");
  g_spmm_unroll = unroll;
  return LGC_OK;
}

extern "C" int lgc_coop_config(int32_t ctas_per_sm) {
  LGC_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 8, "coop config: 1..8 CTAs per SM");
  g_coop_ctas_per_sm = ctas_per_sm;
  return LGC_OK;
}

extern "C" int lgc_spmm_long_row(int32_t long_row) {
  LGC_REQUIRE(long_row == 0 || (long_row >= LGC_LONG_ROW && long_row <= 2048),
              "spmm long_row: 0 (per-call value) or in [LGC_LONG_ROW, 2048]");
  g_spmm_long_row = long_row;
  return LGC_OK;
}

// src_mask (optional): one bit per row of X, clear = the row is all-zero and is not gathered (lgc_spmm_layer_masked)
static int spmm_layer_impl(const int32_t* rowptr, const int32_t* colidx, const float* val,
                           const int32_t* chunk_row, const int32_t* chunk_start,
                           const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                           int64_t n_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                           const int32_t* row_order, int32_t long_row, const float* X, const float* X0,
                           float alpha, float beta, float* Y, float* partial, int32_t* counters,
                           const uint32_t* src_mask, lgc_stream_t stream) {
  int rc = check_spmm_args(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin,
                           chunk_end, n_nodes, row_begin, row_end, X, X0, beta, partial, counters);
  if (rc) return rc;
  LGC_REQUIRE(Y && ((uintptr_t)Y & 15) == 0, "spmm: Y null or misaligned");
  PeerPtrs none{};
  return launch_spmm<0>(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin,
                        chunk_end, row_begin, row_end, row_order, long_row, dim, X, X0, alpha, beta, Y, none,
                        partial, counters, (cudaStream_t)stream, 0, 0, src_mask);
}

extern "C" int lgc_spmm_layer(const int32_t* rowptr, const int32_t* colidx, const float* val,
                              const int32_t* chunk_row, const int32_t* chunk_start,
                              const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                              int64_t n_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                              const int32_t* row_order, int32_t long_row, const float* X, const float* X0,
                              float alpha, float beta, float* Y, float* partial, int32_t* counters,
                              lgc_stream_t stream) {
  return spmm_layer_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, n_nodes, dim,
                         row_begin, row_end, row_order, long_row, X, X0, alpha, beta, Y, partial, counters, nullptr, stream);
}

extern "C" int lgc_spmm_layer_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                     const int32_t* chunk_row, const int32_t* chunk_start,
                                     const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                                     int64_t n_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                                     const int32_t* row_order, int32_t long_row, const float* X, const float* X0,
                                     float alpha, float beta, float* Y, float* partial, int32_t* counters,
                                     const uint32_t* src_mask, lgc_stream_t stream) {
  LGC_REQUIRE(src_mask, "spmm masked: null source-row mask");
  return spmm_layer_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, n_nodes, dim,
                         row_begin, row_end, row_order, long_row, X, X0, alpha, beta, Y, partial, counters, src_mask, stream);
}

extern "C" int lgc_spmm_layer_bcast(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                    const int32_t* chunk_row, const int32_t* chunk_start,
                                    const int32_t* row_chunk_base, int32_t chunk_begin,
                                    int32_t chunk_end, int64_t n_nodes, int32_t dim,
                                    int64_t row_begin, int64_t row_end, const int32_t* row_order,
                                    int32_t long_row, const float* X, const float* X0, float alpha, float beta,
                                    float* const* peer_Y_host, int32_t n_peers, float* partial,
                                    int32_t* counters, lgc_stream_t stream) {
  int rc = check_spmm_args(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin,
                           chunk_end, n_nodes, row_begin, row_end, X, X0, beta, partial, counters);
  if (rc) return rc;
  LGC_REQUIRE(peer_Y_host && n_peers >= 1 && n_peers <= kMaxPeers, "spmm bcast: 1..8 peers");
  PeerPtrs peers{};
  for (int p = 0; p < n_peers; ++p) {
    LGC_REQUIRE(peer_Y_host[p] && ((uintptr_t)peer_Y_host[p] & 15) == 0, "spmm bcast: bad peer pointer");
    peers.y[p] = peer_Y_host[p];
  }
  cudaStream_t s = (cudaStream_t)stream;
#define LGC_BCAST(NP)                                                                              \
  case NP:                                                                                         \
    return launch_spmm<NP>(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base,            \
                           chunk_begin, chunk_end, row_begin, row_end, row_order, long_row, dim, X, X0, \
                           alpha, beta, nullptr, peers, partial, counters, s)
  switch (n_peers) {
    LGC_BCAST(1); LGC_BCAST(2); LGC_BCAST(3); LGC_BCAST(4);
    LGC_BCAST(5); LGC_BCAST(6); LGC_BCAST(7); LGC_BCAST(8);
  }
#undef LGC_BCAST
  return LGC_ERR_INVALID;
}

// One launch over an explicit LIST of rows (device int32, any subset of the nodes, longest first for best balance) and up
// to two ranges of the chunk list; every finished row is stored into n_peers replicas (n_peers = 1 with the local
// buffer = plain local SpMM).  Used by the multi-GPU partition, where a rank owns a slice of the user rows AND a slice of
// the item rows: one mixed launch keeps the short user rows and the chunked hub rows of the items in flight together,
// which two back-to-back launches do not (measured: 61 / 44 Gnnz/s apart, 63 mixed).
static int spmm_rows_bcast_impl(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                const int32_t* chunk_row, const int32_t* chunk_start, const int32_t* row_chunk_base,
                                int32_t chunk_begin, int32_t chunk_end, int32_t chunk_begin2, int32_t chunk_end2,
                                int64_t n_nodes, int32_t dim, const int32_t* row_list, int64_t n_rows, int32_t long_row,
                                const float* X, const float* X0, float alpha, float beta, float* const* peer_Y_host,
                                int32_t n_peers, float* partial, int32_t* counters, const uint32_t* src_mask,
                                lgc_stream_t stream) {
  int rc = check_spmm_args(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, n_nodes, 0,
                           n_rows, X, X0, beta, partial, counters);
  if (rc) return rc;
  LGC_REQUIRE(row_list && n_rows >= 0 && n_rows <= n_nodes, "spmm rows: row list missing / too long");
  LGC_REQUIRE(0 <= chunk_begin2 && chunk_begin2 <= chunk_end2, "spmm rows: bad second chunk range");
  LGC_REQUIRE(chunk_end2 == chunk_begin2 || (chunk_row && chunk_start && row_chunk_base && partial && counters),
              "spmm rows: chunk list given without chunk arrays / scratch");
  LGC_REQUIRE(peer_Y_host && n_peers >= 1 && n_peers <= kMaxPeers, "spmm rows: 1..8 replicas");
  PeerPtrs peers{};
  for (int p = 0; p < n_peers; ++p) {
    LGC_REQUIRE(peer_Y_host[p] && ((uintptr_t)peer_Y_host[p] & 15) == 0, "spmm rows: bad replica pointer");
    peers.y[p] = peer_Y_host[p];
  }
  cudaStream_t s = (cudaStream_t)stream;
#define LGC_ROWS(NP)                                                                                              \
  case NP:                                                                                                        \
    return launch_spmm<NP>(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, 0, \
                           n_rows, row_list, long_row, dim, X, X0, alpha, beta, nullptr, peers, partial, counters, s, \
                           chunk_begin2, chunk_end2, src_mask)
  switch (n_peers) {
    LGC_ROWS(1); LGC_ROWS(2); LGC_ROWS(3); LGC_ROWS(4);
    LGC_ROWS(5); LGC_ROWS(6); LGC_ROWS(7); LGC_ROWS(8);
  }
#undef LGC_ROWS
  return LGC_ERR_INVALID;
}

extern "C" int lgc_spmm_rows_bcast(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                   const int32_t* chunk_row, const int32_t* chunk_start, const int32_t* row_chunk_base,
                                   int32_t chunk_begin, int32_t chunk_end, int32_t chunk_begin2, int32_t chunk_end2,
                                   int64_t n_nodes, int32_t dim, const int32_t* row_list, int64_t n_rows, int32_t long_row,
                                   const float* X, const float* X0, float alpha, float beta, float* const* peer_Y_host,
                                   int32_t n_peers, float* partial, int32_t* counters, lgc_stream_t stream) {
  return spmm_rows_bcast_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, chunk_begin2,
                              chunk_end2, n_nodes, dim, row_list, n_rows, long_row, X, X0, alpha, beta, peer_Y_host, n_peers,
                              partial, counters, nullptr, stream);
}

extern "C" int lgc_spmm_rows_bcast_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                          const int32_t* chunk_row, const int32_t* chunk_start,
                                          const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                                          int32_t chunk_begin2, int32_t chunk_end2, int64_t n_nodes, int32_t dim,
                                          const int32_t* row_list, int64_t n_rows, int32_t long_row, const float* X,
                                          const float* X0, float alpha, float beta, float* const* peer_Y_host,
                                          int32_t n_peers, float* partial, int32_t* counters, const uint32_t* src_mask,
                                          lgc_stream_t stream) {
  LGC_REQUIRE(src_mask, "spmm rows masked: null source-row mask");
  return spmm_rows_bcast_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, chunk_begin, chunk_end, chunk_begin2,
                              chunk_end2, n_nodes, dim, row_list, n_rows, long_row, X, X0, alpha, beta, peer_Y_host, n_peers,
                              partial, counters, src_mask, stream);
}

extern "C" int lgc_peer_barrier(const int32_t* local_flags, int32_t* const* peer_flags_host, int32_t my_rank,
                                int32_t n_peers, int32_t epoch, lgc_stream_t stream) {
  LGC_REQUIRE(local_flags && peer_flags_host, "peer barrier: null pointer");
  LGC_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && my_rank >= 0 && my_rank < n_peers && epoch > 0,
              "peer barrier: bad rank / epoch");
  PeerFlags pf{};
  for (int p = 0; p < n_peers; ++p) {
    LGC_REQUIRE(peer_flags_host[p], "peer barrier: null peer flag array");
    pf.f[p] = peer_flags_host[p];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(local_flags, pf, my_rank, n_peers, epoch);
  LGC_LAUNCH_CHECK("peer_barrier_kernel");
  return LGC_OK;
}

extern "C" int lgc_peer_barrier_dev(const int32_t* local_flags, int32_t* const* peer_flags_host, int32_t my_rank,
                                    int32_t n_peers, int32_t* epoch_counter_dev, lgc_stream_t stream) {
  LGC_REQUIRE(local_flags && peer_flags_host && epoch_counter_dev, "peer barrier: null pointer");
  LGC_REQUIRE(n_peers >= 1 && n_peers <= kMaxPeers && my_rank >= 0 && my_rank < n_peers, "peer barrier: bad rank");
  PeerFlags pf{};
  for (int p = 0; p < n_peers; ++p) {
    LGC_REQUIRE(peer_flags_host[p], "peer barrier: null peer flag array");
    pf.f[p] = peer_flags_host[p];
  }
  peer_barrier_dev_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(local_flags, pf, my_rank, n_peers, epoch_counter_dev);
  LGC_LAUNCH_CHECK("peer_barrier_dev_kernel");
  return LGC_OK;
}

static int propagate_mean_impl(const int32_t* rowptr, const int32_t* colidx, const float* val,
                               const int32_t* chunk_row, const int32_t* chunk_start,
                               const int32_t* row_chunk_base, int32_t n_chunks, int64_t n_nodes,
                               int32_t dim, int32_t n_layers, const int32_t* row_order, int32_t long_row,
                               const float* X0, float* E, float* tmp0, float* tmp1, float* partial,
                               int32_t* counters, const uint32_t* x0_row_mask, lgc_stream_t stream) {
  LGC_REQUIRE(X0 && E && n_layers >= 0, "propagate_mean: bad arguments");
  if (n_layers == 0) {
    LGC_CUDA(cudaMemcpyAsync(E, X0, sizeof(float) * (size_t)n_nodes * dim, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return LGC_OK;
  }
  LGC_REQUIRE(n_layers == 1 || (tmp0 && (n_layers == 2 || tmp1)), "propagate_mean: scratch buffers missing");
  // Horner: S_{l+1} = A S_l + X0, E = S_K / (K+1).  The last layer writes E directly.
  const float* cur = X0;
  float* bufs[2] = {tmp0, tmp1};
  for (int l = 0; l < n_layers; ++l) {
    const bool last = (l == n_layers - 1);
    float* out = last ? E : bufs[l & 1];
    const float alpha = last ? 1.0f / (float)(n_layers + 1) : 1.0f;
    // only the first layer reads X0 as its source: that is where a sparse X0 (the BPR gradient) pays
    int rc = spmm_layer_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, 0, n_chunks,
                             n_nodes, dim, 0, n_nodes, row_order, long_row, cur, X0, alpha, 1.0f, out, partial,
                             counters, l == 0 ? x0_row_mask : nullptr, stream);
    if (rc) return rc;
    cur = out;
  }
  return LGC_OK;
}

extern "C" int lgc_propagate_mean(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                  const int32_t* chunk_row, const int32_t* chunk_start,
                                  const int32_t* row_chunk_base, int32_t n_chunks, int64_t n_nodes,
                                  int32_t dim, int32_t n_layers, const int32_t* row_order, int32_t long_row,
                                  const float* X0, float* E, float* tmp0, float* tmp1, float* partial,
                                  int32_t* counters, lgc_stream_t stream) {
  return propagate_mean_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, n_chunks, n_nodes, dim, n_layers,
                             row_order, long_row, X0, E, tmp0, tmp1, partial, counters, nullptr, stream);
}

extern "C" int lgc_propagate_mean_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                         const int32_t* chunk_row, const int32_t* chunk_start,
                                         const int32_t* row_chunk_base, int32_t n_chunks, int64_t n_nodes,
                                         int32_t dim, int32_t n_layers, const int32_t* row_order, int32_t long_row,
                                         const float* X0, float* E, float* tmp0, float* tmp1, float* partial,
                                         int32_t* counters, const uint32_t* x0_row_mask, lgc_stream_t stream) {
  LGC_REQUIRE(x0_row_mask, "propagate_mean masked: null row mask");
  return propagate_mean_impl(rowptr, colidx, val, chunk_row, chunk_start, row_chunk_base, n_chunks, n_nodes, dim, n_layers,
                             row_order, long_row, X0, E, tmp0, tmp1, partial, counters, x0_row_mask, stream);
}

// All K layers + layer mean in one cooperative launch (small graphs).  The unit lists are built once per graph by the
// host side (lgcnhs_b200/ops.py: NormGraph.coop_units): units of <= 128 non-zeros, long rows cut into several units.
extern "C" int lgc_propagate_mean_coop(const int32_t* rowptr, const int32_t* colidx, const float* val,
                                       const int32_t* unit_row, const int32_t* unit_start, const int32_t* unit_end,
                                       const int32_t* unit_slot, int32_t n_units, const int32_t* split_row,
                                       const int32_t* split_first, const int32_t* split_count, int32_t n_split,
                                       int64_t n_nodes, int32_t dim, int32_t n_layers, const float* X0, float* E,
                                       float* tmp0, float* tmp1, float* partial, uint32_t* barrier_state,
                                       lgc_stream_t stream) {
  LGC_REQUIRE(rowptr && colidx && val && unit_row && unit_start && unit_end && unit_slot && X0 && E, "propagate coop: null pointer");
  LGC_REQUIRE(n_units > 0 && n_layers >= 1 && n_nodes > 0, "propagate coop: empty problem");
  LGC_REQUIRE(n_split == 0 || (split_row && split_first && split_count && partial), "propagate coop: split lists missing");
  LGC_REQUIRE(n_layers == 1 || (tmp0 && (n_layers == 2 || tmp1)), "propagate coop: scratch buffers missing");
  LGC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "propagate coop: embedding dim not in {32,64,128}");
  LGC_REQUIRE(barrier_state, "propagate coop: barrier state missing");
  CoopParams p{rowptr, colidx, val, unit_row, unit_start, unit_end, unit_slot, split_row, split_first, split_count,
               n_units, n_split, n_layers, X0, tmp0, tmp1, E, partial, barrier_state};
  void* fn = dim == 32 ? (void*)propagate_coop_kernel<32> : dim == 64 ? (void*)propagate_coop_kernel<64>
                                                                      : (void*)propagate_coop_kernel<128>;
  int per_sm = 0;
  LGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0));
  if (per_sm < 1) LGC_FAIL(LGC_ERR_CUDA, "propagate coop: kernel does not fit an SM");
  int64_t grid = (int64_t)num_sms() * (per_sm > g_coop_ctas_per_sm ? g_coop_ctas_per_sm : per_sm);
  const int64_t want = ceil_div(n_units, kWarpsPerBlock);
  if (grid > want) grid = want;
  void* args[] = {(void*)&p};
  LGC_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kThreads), args, 0, (cudaStream_t)stream));
  note_launch();
  return LGC_OK;
}
