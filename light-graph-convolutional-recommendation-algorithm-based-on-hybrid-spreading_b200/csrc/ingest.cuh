// ingest.cuh — (N2) device primitives of the graph / format ingestion that sits in front of both hot paths:
// an LSD radix sort of 64-bit keys and an exclusive scan, hand-written for sm_100a (no CUB, no torch.unique).
// They replace the Python loops + dense (U+M)^2 detours of /root/reference/utils/graph.py:12-50 and
// /root/reference/utils/trans.py:13-80 (and PyG's per-forward gcn_norm scatter, model/LightGCN/model.py:53) behind
// lgc_csr_build, lgc_seen_csr and lgc_sort_u64.
//
// Radix sort: 8-bit digits, least significant first, only over the bytes that can be non-zero.  One pass =
//   radix_hist_kernel    : a CTA owns a tile of kSortTile keys, counts its digits, writes hist[digit][tile]
//   exclusive scan       : over the digit-major table -> global base of every (digit, tile)
//   radix_scatter_kernel : the CTA re-reads its tile; each of its 8 warps owns a contiguous sub-tile and walks it 32 keys
//                          at a time in order, so equal digits keep their input order (stable): rank inside the warp step
//                          by __match_any_sync, warp bases from a scan of the per-warp digit counts.
// HBM-bound integer work: 3 x 8 bytes per key and pass.
#pragma once
#include "common.cuh"

namespace lgc {
namespace ingest {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortPerWarp = 512;                       // keys per warp and tile (16 steps of 32)
constexpr int kSortTile = kSortWarps * kSortPerWarp;    // 4096 keys per CTA

__device__ __forceinline__ void warp_count_digits(const uint64_t* __restrict__ keys, int64_t n, int64_t wbase, int shift,
                                                  int lane, int* __restrict__ wc) {
  for (int s = 0; s < kSortPerWarp; s += 32) {
    const int64_t i = wbase + s + lane;
    const bool ok = i < n;
    const int d = ok ? (int)((keys[i] >> shift) & 255ull) : 256;   // 256 = padding class of this step
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if (ok && (peers & ((1u << lane) - 1u)) == 0u) wc[d] += __popc(peers);   // leader of the class; warp-private counters
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int n_tiles, uint32_t* __restrict__ hist) {
  __shared__ int wc[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wc[warp][d] = 0;
  __syncwarp();
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * kSortPerWarp;
  warp_count_digits(keys, n, wbase, shift, lane, wc[warp]);
  __syncthreads();
  const int d = threadIdx.x;   // 256 threads = 256 digits
  int tot = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) tot += wc[w][d];
  hist[(size_t)d * n_tiles + blockIdx.x] = (uint32_t)tot;
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t* __restrict__ out, int64_t n, int shift, int n_tiles,
                     const uint32_t* __restrict__ base) {
  __shared__ int wc[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wc[warp][d] = 0;
  __syncwarp();
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * kSortPerWarp;
  warp_count_digits(keys, n, wbase, shift, lane, wc[warp]);
  __syncthreads();
  {
    // warp bases: global base of (digit, tile) + the counts of the tile's earlier warps
    const int d = threadIdx.x;
    int run = (int)base[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const int c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  int* my = wc[warp];
  for (int s = 0; s < kSortPerWarp; s += 32) {
    const int64_t i = wbase + s + lane;
    const bool ok = i < n;
    const uint64_t k = ok ? keys[i] : 0ull;
    const int d = ok ? (int)((k >> shift) & 255ull) : 256;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int pos = 0;
    if (ok && lane == leader) {
      pos = my[d];
      my[d] = pos + __popc(peers);
    }
    pos = __shfl_sync(0xffffffffu, pos, leader);
    if (ok) out[(size_t)pos + rank] = k;
    __syncwarp();
  }
}

// ---- exclusive scan of uint32 (two levels: 2048 elements per CTA, then the CTA sums, recursively) ----
constexpr int kScanThreads = 256;
constexpr int kScanPer = 8;
constexpr int kScanTile = kScanThreads * kScanPer;

__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPer;
  uint32_t v[kScanPer];
  uint32_t tot = 0;
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) {
    v[j] = base + j < n ? in[base + j] : 0u;
    tot += v[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = tot;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  uint32_t run = inc - tot + (warp ? s_warp[warp - 1] : 0u);   // exclusive prefix of this thread inside the tile
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (tile_sums && threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = run;
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, int64_t n, const uint32_t* __restrict__ tile_base) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] += tile_base[i / kScanTile];
}

// scratch needed by exclusive_scan_u32 for n elements (uint32 words)
inline size_t scan_scratch_words(int64_t n) {
  size_t words = 0;
  while (n > kScanTile) {
    n = ceil_div(n, kScanTile);
    words += (size_t)align_up((size_t)n, 64);
  }
  return words + 64;
}

// out[i] = sum_{j<i} in[j]; in == out allowed.  Returns the number of kernels launched, or -1 on a launch error.
inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t stream) {
  if (n <= 0) return 0;
  const int64_t tiles = ceil_div(n, kScanTile);
  int launches = 0;
  if (tiles == 1) {
    scan_tile_kernel<<<1, kScanThreads, 0, stream>>>(in, out, n, nullptr);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
  }
  uint32_t* sums = scratch;
  scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, n, sums);
  if (cudaGetLastError() != cudaSuccess) return -1;
  ++launches;
  const int sub = exclusive_scan_u32(sums, sums, tiles, scratch + align_up((size_t)tiles, 64), stream);
  if (sub < 0) return -1;
  launches += sub;
  scan_add_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(out, n, sums);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return launches + 1;
}

// workspace of radix_sort_u64 in bytes: ping-pong buffer is supplied by the caller; this is the digit table + scan scratch
inline size_t sort_table_words(int64_t n) {
  const int64_t tiles = ceil_div(n > 0 ? n : 1, kSortTile);
  return (size_t)align_up((size_t)256 * tiles, 64) + scan_scratch_words(256 * tiles);
}

// Stable LSD radix sort of the low `bits` bits of 64-bit keys.  a holds the input; a and b are used alternately; returns
// the buffer that holds the sorted keys (a or b), nullptr on a launch error.  table: sort_table_words(n) uint32 words.
inline uint64_t* radix_sort_u64(uint64_t* a, uint64_t* b, int64_t n, int bits, uint32_t* table, cudaStream_t stream,
                                int* launches) {
  if (n <= 0) return a;
  const int n_tiles = (int)ceil_div(n, kSortTile);
  uint32_t* scan_scratch = table + align_up((size_t)256 * n_tiles, 64);
  uint64_t* src = a;
  uint64_t* dst = b;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<n_tiles, kSortThreads, 0, stream>>>(src, n, shift, n_tiles, table);
    if (cudaGetLastError() != cudaSuccess) return nullptr;
    const int sl = exclusive_scan_u32(table, table, (int64_t)256 * n_tiles, scan_scratch, stream);
    if (sl < 0) return nullptr;
    radix_scatter_kernel<<<n_tiles, kSortThreads, 0, stream>>>(src, dst, n, shift, n_tiles, table);
    if (cudaGetLastError() != cudaSuccess) return nullptr;
    if (launches) *launches += 2 + sl;
    uint64_t* t = src; src = dst; dst = t;
  }
  return src;
}

}  // namespace ingest
}  // namespace lgc
