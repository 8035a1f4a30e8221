// select.cuh — building blocks of the exact top-k selection shared by lgc_topk_rows and the
// fused lgc_score_topk (score_topk.cu).
//
// A candidate is the 64-bit key  (monotone fp32 value key << 32) | column : larger key = better,
// all keys of one row are distinct, so "k largest keys, descending" is the deterministic total
// order (value descending, equal values -> larger column first) that np.argsort(row)[::-1]
// produces on the reference's CPU path (/root/reference/model/SpreadMethod/recommend.py:38;
// SURVEY.md §4).  Key 0 never occurs for a real value (-inf maps to 0x007FFFFF) and marks an
// empty slot.
//
// Selection scheme ("threshold + candidate buffer"): a row keeps a buffer of CAP >= 2k keys and a
// threshold = its current k-th largest key.  Streaming values are compared against the threshold
// (one 32-bit compare in the hot loop); only the survivors are appended.  When the buffer is
// full, one warp sorts it in registers (bitonic network, CAP/32 keys per lane), keeps the k best
// and raises the threshold.  With n columns a row appends ~k ln(n/k) keys in total, so the
// selection costs almost nothing next to reading (or computing) the values.
#pragma once
#include "common.cuh"

namespace lgc {

__device__ __forceinline__ uint32_t float_key(float x) {
  const uint32_t b = __float_as_uint(x + 0.0f);       // -0.0 -> +0.0: the two zeros are ONE value (np.argsort ties)
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone: larger float -> larger key
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long make_key(uint32_t vkey, uint32_t col) {
  return ((unsigned long long)vkey << 32) | (unsigned long long)col;
}

// Sort 32*E keys held as r[e] <-> logical index e*32 + lane, DESCENDING (index 0 = largest).
template <int E>
__device__ __forceinline__ void warp_sort_desc(unsigned long long (&r)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int je = j >> 5;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & je) == 0) {
            const int e2 = e | je;
            const bool dir = ((e << 5) & k) == 0;  // block sorted "descending" when dir
            const unsigned long long a = r[e], b = r[e2];
            const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
            r[e] = dir ? hi : lo;
            r[e2] = dir ? lo : hi;
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int idx_hi = e << 5;  // k may exceed 32: the direction bit can sit in e or in lane
          const bool dir = k >= 32 ? ((idx_hi & k) == 0) : ((lane & k) == 0);
          const unsigned long long a = r[e];
          const unsigned long long b = __shfl_xor_sync(0xffffffffu, a, j);
          const bool keep_max = (dir == lower);
          r[e] = keep_max ? (a > b ? a : b) : (a > b ? b : a);
        }
      }
    }
  }
}

// One warp: keep the k largest of the first `cnt` keys of buf (cnt <= 32*E), sorted descending in
// buf[0..k); returns the new count (the non-zero keys among them).  *thr becomes the k-th largest key (0 while
// fewer than k keys exist).  Every lane must call it; buf may be written by other warps before the
// caller's preceding barrier only.
template <int E>
__device__ __forceinline__ int warp_compact(unsigned long long* buf, int cnt, int k, int lane,
                                            unsigned long long* thr) {
  __syncwarp();
  unsigned long long r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = e * 32 + lane;
    r[e] = idx < cnt ? buf[idx] : 0ull;
  }
  warp_sort_desc<E>(r, lane);
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = e * 32 + lane;
    if (idx < k) buf[idx] = r[e];
  }
  __syncwarp();
  int nz = 0;  // zero keys (empty / dropped entries) sort last and are not kept
#pragma unroll
  for (int e = 0; e < E; ++e) nz += (e * 32 + lane < k && r[e] != 0ull) ? 1 : 0;
  const int kept = __reduce_add_sync(0xffffffffu, nz);
  *thr = kept == k ? buf[k - 1] : 0ull;
  return kept;
}

// Same contract as warp_compact but WITHOUT ordering the survivors: the k-th largest key is found by an MSB-first
// radix search over the keys held in registers (one warp-wide count per bit, stopping as soon as the keys that
// share the prefix are exactly the ones still wanted), then the keys >= that threshold are packed to the front
// with a ballot / prefix.  About half the instructions of the bitonic sort for 256 keys; used for the
// intermediate compactions (the final one sorts).
template <int E>
__device__ __forceinline__ int warp_select(unsigned long long* buf, int cnt, int k, int lane,
                                           unsigned long long* thr) {
  __syncwarp();
  unsigned long long r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int idx = e * 32 + lane;
    r[e] = idx < cnt ? buf[idx] : 0ull;
  }
  unsigned long long prefix = 0ull;  // bits [63..b] of the k-th largest key found so far
  int want = k;                      // rank of the wanted key among the keys that share the prefix
#pragma unroll 1
  for (int b = 63; b >= 0; --b) {
    const unsigned long long cand = (prefix >> b) | 1ull;  // prefix extended by a 1 bit, right-aligned
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += ((r[e] >> b) == cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= want) {
      prefix |= 1ull << b;
      if (c == want) {  // exactly the wanted keys share this prefix: the threshold is their minimum
        unsigned long long mn = ~0ull;
#pragma unroll
        for (int e = 0; e < E; ++e)
          if ((r[e] >> b) == cand && r[e] < mn) mn = r[e];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, mn, off);
          mn = o < mn ? o : mn;
        }
        prefix = mn;
        break;
      }
    } else {
      want -= c;
    }
  }
  // prefix == k-th largest key (0 when fewer than k non-zero keys exist)
  __syncwarp();
  int base = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool keep = r[e] != 0ull && r[e] >= prefix;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) buf[base + __popc(m & ((1u << lane) - 1u))] = r[e];
    base += __popc(m);
  }
  __syncwarp();
  *thr = (base == k && prefix != 0ull) ? prefix : 0ull;
  return base;
}

// ---- fused top-k epilogues of the tensor-core kernels (umma_gemm.cu, score_topk_tc.cu) ----
// Every (row, segment) owns a candidate buffer of kCandCap keys in global scratch, written by ONE epilogue thread;
// a final warp-per-row kernel merges the segments' buffers and writes the k best, sorted (value descending, ties ->
// larger column first).
constexpr int kCandCap = 128;   // k <= 32, at most 64-80 values offered between two overflow checks
constexpr int kMergeWarps = 8;

template <int CAPC>
__global__ void __launch_bounds__(kMergeWarps * 32)
topk_merge_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int64_t n_rows, int segs,
                  int k, int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  constexpr int CAPM = 2 * CAPC;
  __shared__ unsigned long long s_buf[kMergeWarps][CAPM];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kMergeWarps + w;
  if (row >= n_rows) return;
  unsigned long long* buf = s_buf[w];
  unsigned long long thr;
  int cur = 0;
  for (int sg = 0; sg < segs; ++sg) {
    int c = cand_cnt[(size_t)row * segs + sg];
    c = c < CAPC ? c : CAPC;
    if (cur + c > CAPM) cur = warp_select<CAPM / 32>(buf, cur, k, lane, &thr);
    const unsigned long long* src = cand + ((size_t)row * segs + sg) * CAPC;
    for (int i = lane; i < c; i += 32) buf[cur + i] = src[i];
    cur += c;
    __syncwarp();
  }
  const int kept = warp_compact<CAPM / 32>(buf, cur, k, lane, &thr);
  for (int i = lane; i < k; i += 32) {
    const unsigned long long key = buf[i];
    out_idx[row * k + i] = i < kept ? (int64_t)(uint32_t)(key & 0xffffffffull) : -1;
    if (out_val) out_val[row * k + i] = i < kept ? key_float((uint32_t)(key >> 32)) : -INFINITY;
  }
}

}  // namespace lgc
