// score_topk.cu — (P8 / S4) dense score block + row-wise masked top-k.
//   lgc_score_block : score = Xu . Xi^T with the seen-pair fill, standing in for
//       torch.matmul + score[users, items] = -1024 at
//       /root/reference/model/LightGCN/recommend.py:86,101,111 (== evaluation.py:34,49 and
//       SpreadLightGCN/model.py:77,92,102).  fp32 FMA on CUDA cores: the contraction is only
//       dim (64) long and the result must match an fp32 SGEMM to ~1e-7, so the tensor cores
//       are not used here.
//   lgc_topk_rows   : torch.topk(score, k) (recommend.py:114) and the reference's real
//       bottleneck, np.argsort(row)[::-1] + Python membership filter
//       (/root/reference/model/SpreadMethod/recommend.py:35-47, 32.8 ms/user measured), as one
//       radix-select pass structure per row: exact, deterministic, ties -> larger index first.
#include "common.cuh"

namespace lgc {

// ------------------------------------------------------------------------------------------
// score block: 64 users x 64 items per CTA, whole K (= dim) staged in shared memory
// ------------------------------------------------------------------------------------------
constexpr int kTileU = 64, kTileI = 64, kScoreThreads = 256;

template <int DIM>
__global__ void __launch_bounds__(kScoreThreads)
score_block_kernel(const float* __restrict__ Xu, const float* __restrict__ Xi, int64_t u0, int64_t u1,
                   int64_t n_items, float* __restrict__ out, int64_t ldo) {
  // transposed staging: s[d][row], +4 padding keeps the float4 reads conflict-free
  __shared__ __align__(16) float sU[DIM][kTileU + 4];
  __shared__ __align__(16) float sI[DIM][kTileI + 4];
  const int64_t ub = u0 + (int64_t)blockIdx.y * kTileU;
  const int64_t ib = (int64_t)blockIdx.x * kTileI;
  const int tid = threadIdx.x;
  // each thread loads float4 chunks: row = tid % 64, d-chunk strided by 4 threads-groups
  for (int c = tid / 64; c < DIM / 4; c += kScoreThreads / 64) {
    const int r = tid % 64;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (ub + r < u1) a = __ldg(reinterpret_cast<const float4*>(Xu + (ub + r) * DIM + c * 4));
    if (ib + r < n_items) b = __ldg(reinterpret_cast<const float4*>(Xi + (ib + r) * DIM + c * 4));
    sU[c * 4 + 0][r] = a.x; sU[c * 4 + 1][r] = a.y; sU[c * 4 + 2][r] = a.z; sU[c * 4 + 3][r] = a.w;
    sI[c * 4 + 0][r] = b.x; sI[c * 4 + 1][r] = b.y; sI[c * 4 + 2][r] = b.z; sI[c * 4 + 3][r] = b.w;
  }
  __syncthreads();
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < DIM; ++d) {
    const float4 a = *reinterpret_cast<const float4*>(&sU[d][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&sI[d][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t u = ub + ty * 4 + i;
    if (u >= u1) continue;
    float* o = out + (u - u0) * ldo + ib + tx * 4;
    if (ib + tx * 4 + 3 < n_items && (ldo & 3) == 0 && ((uintptr_t)out & 15) == 0) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ib + tx * 4 + j < n_items) o[j] = acc[i][j];
    }
  }
}

// one warp per user row: out[u, seen items] = fill
__global__ void seen_fill_kernel(const int32_t* __restrict__ seen_ptr, const int32_t* __restrict__ seen_idx,
                                 int64_t u0, int64_t u1, float fill, float* __restrict__ out, int64_t ldo) {
  const int64_t u = u0 + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (u >= u1) return;
  const int s = seen_ptr[u], e = seen_ptr[u + 1];
  float* o = out + (u - u0) * ldo;
  for (int i = s + lane; i < e; i += 32) o[seen_idx[i]] = fill;
}

// ------------------------------------------------------------------------------------------
// row-wise masked top-k on the 64-bit composite key (ordered fp32 value, column index)
// ------------------------------------------------------------------------------------------
// ONE WARP per row, k <= 128 << n_cols, no shared memory and no barriers.  Each lane streams its
// strided share of the row once (coalesced 128 B per warp load, evict-first, 8 loads in flight)
// and keeps a sorted list of its LIST largest (value key, column) pairs.  The hot loop is one
// 32-bit compare per element: only values that beat the lane's current LIST-th best go on to the
// exclusion test (one bit of the bit-packed rows x n_cols mask) and the insertion.  Then k rounds
// of a warp-wide arg-max over the list heads — two REDUX instructions (value, then column among
// the equal values) — pop the winners in final order.  A lane whose list runs dry re-streams its
// share below its last popped pair, so the result is exact for any distribution.  (value, column)
// pairs are unique, so the result is deterministic and equal values rank the LARGER column first
// (np.argsort(row)[::-1] order).  If fewer than k columns are selectable the tail is -1.
constexpr int kTopkWarps = 8;
constexpr int kTopkMaxK = 128;
constexpr int kTopkUnroll = 8;

__device__ __forceinline__ uint32_t float_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone: larger float -> larger key
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int LIST>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_rows_kernel(const float* __restrict__ S, int n_rows, int n, int64_t lds, const uint32_t* __restrict__ mask,
                 int64_t mask_stride_bits, int64_t row_offset, int k, int64_t* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kTopkWarps + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const float* row = S + r * lds;
  const int64_t bit0 = (row_offset + r) * mask_stride_bits;
  const uint32_t* mrow = mask ? mask + (bit0 >> 5) : nullptr;
  const uint32_t boff = (uint32_t)(bit0 & 31);

  uint32_t lv[LIST], lc[LIST];  // value keys / column+1, sorted descending; lc == 0 marks an empty slot
  // keep the LIST largest pairs strictly below (bv, bc) among this lane's columns (ascending column order,
  // so among equal values the later column is the larger pair)
  auto stream = [&](uint32_t bv, uint32_t bc) {
#pragma unroll
    for (int i = 0; i < LIST; ++i) { lv[i] = 0u; lc[i] = 0u; }
    for (int c0 = lane; c0 < n; c0 += 32 * kTopkUnroll) {
      float v[kTopkUnroll];
#pragma unroll
      for (int u = 0; u < kTopkUnroll; ++u) {
        const int c = c0 + 32 * u;
        v[u] = c < n ? __ldcs(row + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kTopkUnroll; ++u) {
        const int c = c0 + 32 * u;
        const uint32_t vk = float_key(v[u]);
        // fast reject: not better than this lane's LIST-th best (an empty slot has lv == 0)
        if (c < n && (vk >= lv[LIST - 1] || lc[LIST - 1] == 0u)) {
          const bool below = vk < bv || (vk == bv && (uint32_t)(c + 1) < bc);
          bool ex = false;
          if (mrow) {
            const uint32_t b = boff + (uint32_t)c;
            ex = (__ldg(mrow + (b >> 5)) >> (b & 31)) & 1u;
          }
          if (below && !ex) {
            lv[LIST - 1] = vk;
            lc[LIST - 1] = (uint32_t)(c + 1);
            bool moving = true;  // bubble the new pair up; it stops at the first strictly larger value
#pragma unroll
            for (int i = LIST - 1; i > 0; --i) {
              moving = moving && (lc[i - 1] == 0u || lv[i] >= lv[i - 1]);
              if (moving) {
                const uint32_t tv = lv[i], tc = lc[i];
                lv[i] = lv[i - 1]; lc[i] = lc[i - 1];
                lv[i - 1] = tv; lc[i - 1] = tc;
              }
            }
          }
        }
      }
    }
  };
  stream(0xffffffffu, 0xffffffffu);

  for (int round = 0; round < k; ++round) {
    const bool has = lc[0] != 0u;
    const uint32_t mv = __reduce_max_sync(0xffffffffu, has ? lv[0] : 0u);
    const uint32_t mc = __reduce_max_sync(0xffffffffu, (has && lv[0] == mv) ? lc[0] : 0u);
    if (mc == 0u) {  // fewer than k selectable columns
      if (lane == 0)
        for (int t = round; t < k; ++t) {
          out_idx[r * k + t] = -1;
          if (out_val) out_val[r * k + t] = -INFINITY;
        }
      break;
    }
    if (has && lv[0] == mv && lc[0] == mc) {  // exactly one lane
      out_idx[r * k + round] = (int64_t)(mc - 1u);
      if (out_val) out_val[r * k + round] = key_float(mv);
#pragma unroll
      for (int i = 0; i < LIST - 1; ++i) { lv[i] = lv[i + 1]; lc[i] = lc[i + 1]; }
      lv[LIST - 1] = 0u; lc[LIST - 1] = 0u;
      if (lc[0] == 0u) stream(mv, mc);  // list ran dry: refill with this lane's pairs below the popped one
    }
  }
}

// bit-packed exclusion mask from a CSR (row r, column c -> bit r*stride + c), mask zero-filled by the caller
__global__ void mask_from_csr_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t n_rows,
                                     int64_t n_cols, int64_t stride_bits, uint32_t* __restrict__ mask) {
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  for (int i = ptr[r] + lane; i < ptr[r + 1]; i += 32) {
    const int c = idx[i];
    if (c >= 0 && c < n_cols) {
      const int64_t b = r * stride_bits + c;
      atomicOr(mask + (b >> 5), 1u << (b & 31));
    }
  }
}

}  // namespace lgc

using namespace lgc;

extern "C" int lgc_score_block(const float* Xu, const float* Xi, int64_t u0, int64_t u1, int64_t n_items,
                               int32_t dim, const int32_t* seen_ptr, const int32_t* seen_idx, float fill,
                               float* out, int64_t ldo, lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(Xu && Xi && out, "score: null pointer");
  LGC_REQUIRE(u0 >= 0 && u1 > u0 && n_items > 0 && ldo >= n_items, "score: bad extents");
  LGC_REQUIRE(((uintptr_t)Xu & 15) == 0 && ((uintptr_t)Xi & 15) == 0, "score: embeddings must be 16-byte aligned");
  LGC_REQUIRE((seen_ptr == nullptr) == (seen_idx == nullptr), "score: seen_ptr / seen_idx mismatch");
  dim3 grid((unsigned)ceil_div(n_items, kTileI), (unsigned)ceil_div(u1 - u0, kTileU));
  switch (dim) {
    case 32: score_block_kernel<32><<<grid, kScoreThreads, 0, stream>>>(Xu, Xi, u0, u1, n_items, out, ldo); break;
    case 64: score_block_kernel<64><<<grid, kScoreThreads, 0, stream>>>(Xu, Xi, u0, u1, n_items, out, ldo); break;
    default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "score: embedding dim %d not in {32,64}", dim);
  }
  LGC_LAUNCH_CHECK("score_block_kernel");
  if (seen_ptr) {
    const int64_t threads = (u1 - u0) * 32;
    seen_fill_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, stream>>>(seen_ptr, seen_idx, u0, u1, fill, out, ldo);
    LGC_LAUNCH_CHECK("seen_fill_kernel");
  }
  return LGC_OK;
}

extern "C" int lgc_topk_rows(const float* S, int64_t n_rows, int64_t n_cols, int64_t lds, const uint32_t* excl_mask,
                             int64_t mask_stride_bits, int64_t row_offset, int32_t k, int64_t* out_idx,
                             float* out_val, lgc_stream_t stream) {
  LGC_REQUIRE(S && out_idx, "topk: null pointer");
  LGC_REQUIRE(n_rows > 0 && n_cols > 0 && lds >= n_cols, "topk: bad extents");
  LGC_REQUIRE(k >= 1 && k <= kTopkMaxK && k <= n_cols, "topk: k must be in [1, min(128, n_cols)]");
  LGC_REQUIRE(n_cols < (1ll << 31) - 1 && n_rows < (1ll << 31), "topk: extents exceed int32");
  LGC_REQUIRE(!excl_mask || mask_stride_bits >= n_cols, "topk: mask stride smaller than the row");
  LGC_REQUIRE(((uintptr_t)excl_mask & 3) == 0, "topk: mask must be 4-byte aligned");
  const unsigned grid = (unsigned)ceil_div(n_rows, kTopkWarps);
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 32)
    topk_rows_kernel<4><<<grid, kTopkWarps * 32, 0, st>>>(S, (int)n_rows, (int)n_cols, lds, excl_mask, mask_stride_bits,
                                                          row_offset, k, out_idx, out_val);
  else
    topk_rows_kernel<8><<<grid, kTopkWarps * 32, 0, st>>>(S, (int)n_rows, (int)n_cols, lds, excl_mask, mask_stride_bits,
                                                          row_offset, k, out_idx, out_val);
  LGC_LAUNCH_CHECK("topk_rows_kernel");
  return LGC_OK;
}

extern "C" int lgc_mask_from_csr(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int64_t n_cols,
                                 int64_t stride_bits, uint32_t* mask, lgc_stream_t stream) {
  LGC_REQUIRE(ptr && idx && mask && n_rows > 0 && n_cols > 0 && stride_bits >= n_cols, "mask_from_csr: bad arguments");
  mask_from_csr_kernel<<<(unsigned)ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(ptr, idx, n_rows, n_cols,
                                                                                              stride_bits, mask);
  LGC_LAUNCH_CHECK("mask_from_csr_kernel");
  return LGC_OK;
}
