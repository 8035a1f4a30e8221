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
// row-wise masked top-k: radix select on the 64-bit composite key (ordered fp32 value, index)
// ------------------------------------------------------------------------------------------
// One CTA per row.  The row is read from global memory ONCE (when it has <= 256*RCACHE columns)
// into per-thread registers as order-preserving 32-bit keys with the exclusion applied.  MSB-first
// 11-bit digits of the composite key are histogrammed until the undecided bucket plus everything
// above it fits a 256-entry candidate buffer, which is then sorted (bitonic) — typically two
// digit passes for continuous scores, six only when the whole row ties.  Exact and deterministic;
// equal values rank the LARGER index first (np.argsort(row)[::-1] / CPU torch.topk order).
constexpr int kTopkThreads = 256;
constexpr int kTopkMaxK = 128;
constexpr int kCand = 256;
constexpr int kBins = 2048;

__device__ __forceinline__ uint32_t float_key(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone: larger float -> larger key
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct TopkSmem {
  unsigned int hist[kBins];
  unsigned long long cand[kCand];
  unsigned int warp_tot[kTopkThreads / 32];
  unsigned long long prefix;  // decided high bits of the k-th largest key (low bits zero)
  int k_rem;                  // how many are still to be taken from the undecided bucket
  int n_above;                // elements strictly above the undecided bucket (all selected)
  int n_eq;                   // elements in the undecided bucket
  int n_cand;
};

template <int RCACHE>
__global__ void __launch_bounds__(kTopkThreads)
topk_rows_kernel(const float* __restrict__ S, int64_t n_cols, int64_t lds, const int32_t* __restrict__ excl_ptr,
                 const int32_t* __restrict__ excl_idx, int64_t row_offset, int k, int64_t* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  extern __shared__ unsigned int s_dyn[];
  __shared__ TopkSmem sm;
  unsigned int* bitmap = s_dyn;  // n_cols bits: 1 = excluded
  const int64_t r = blockIdx.x;
  const float* row = S + r * lds;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = (int)n_cols;
  const int words = (n + 31) / 32;

  for (int w = tid; w < words; w += kTopkThreads) bitmap[w] = 0u;
  if (tid == 0) { sm.prefix = 0ull; sm.k_rem = k; sm.n_above = 0; sm.n_eq = n; sm.n_cand = 0; }
  __syncthreads();
  if (excl_ptr) {
    const int s = excl_ptr[row_offset + r], e = excl_ptr[row_offset + r + 1];
    for (int i = s + tid; i < e; i += kTopkThreads) {
      const int c = excl_idx[i];
      if (c >= 0 && c < n) atomicOr(&bitmap[c >> 5], 1u << (c & 31));
    }
  }
  __syncthreads();

  auto load_key = [&](int c) -> uint32_t {  // excluded columns get key 0 (below every real value)
    const bool ex = (bitmap[c >> 5] >> (c & 31)) & 1u;
    return ex ? 0u : float_key(__ldg(row + c));
  };
  uint32_t kreg[RCACHE > 0 ? RCACHE : 1];
  if (RCACHE > 0) {
#pragma unroll
    for (int i = 0; i < RCACHE; ++i) {
      const int c = tid + i * kTopkThreads;
      kreg[i] = c < n ? load_key(c) : 0u;
    }
  }
  // visit every column of the row: f(column, composite key)
  auto for_each = [&](auto&& f) {
    if (RCACHE > 0) {
#pragma unroll
      for (int i = 0; i < RCACHE; ++i) {
        const int c = tid + i * kTopkThreads;
        f(c, c < n, ((unsigned long long)kreg[i] << 32) | (unsigned int)c);
      }
    } else {
      for (int c0 = 0; c0 < n; c0 += kTopkThreads) {
        const int c = c0 + tid;
        f(c, c < n, c < n ? (((unsigned long long)load_key(c) << 32) | (unsigned int)c) : 0ull);
      }
    }
  };

  // digit schedule over the 64-bit composite key: 11 + 11 + 10 value bits, then the index bits
  int idx_bits = 1;
  while ((1 << idx_bits) < n) ++idx_bits;
  int shift = 64;
  for (int pass = 0; pass < 8; ++pass) {
    if (sm.n_above + sm.n_eq <= kCand) break;  // block-uniform (read after a barrier)
    int width;
    if (shift > 32) width = shift == 64 ? 11 : shift == 53 ? 11 : 10;
    else {
      const int top = shift == 32 ? idx_bits : shift;  // index bits above idx_bits are always zero
      if (shift == 32) shift = idx_bits;
      width = top < 11 ? top : 11;
    }
    if (width <= 0) break;
    shift -= width;
    for (int b = tid; b < kBins; b += kTopkThreads) sm.hist[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = sm.prefix;
    const unsigned long long hi_mask = (shift + width) >= 64 ? 0ull : (~0ull << (shift + width));
    const unsigned int dmask = (1u << width) - 1u;
    for_each([&](int, bool valid, unsigned long long key) {
      const bool act = valid && (key & hi_mask) == prefix;
      const unsigned int dg = (unsigned int)(key >> shift) & dmask;
      const unsigned int amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const unsigned int peers = __match_any_sync(amask, dg);
        if (lane == __ffs(peers) - 1) atomicAdd(&sm.hist[dg], (unsigned int)__popc(peers));
      }
    });
    __syncthreads();
    // find the bucket holding the k_rem-th largest: descending scan over the bins
    constexpr int PER = kBins / kTopkThreads;  // 8 bins per thread, thread 0 owns the TOP bins
    unsigned int mine = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) mine += sm.hist[kBins - 1 - (tid * PER + i)];
    unsigned int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) sm.warp_tot[warp] = incl;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += sm.warp_tot[w];
    const unsigned int excl_sum = before + incl - mine;  // elements in bins above this thread's bins
    const unsigned int k_rem = (unsigned int)sm.k_rem;
    __syncthreads();
    if (excl_sum < k_rem && excl_sum + mine >= k_rem) {  // exactly one thread
      unsigned int acc = excl_sum;
      for (int i = 0; i < PER; ++i) {
        const int bin = kBins - 1 - (tid * PER + i);
        const unsigned int cnt = sm.hist[bin];
        if (acc + cnt >= k_rem) {
          sm.prefix = prefix | ((unsigned long long)bin << shift);
          sm.k_rem = (int)(k_rem - acc);
          sm.n_above += (int)acc;
          sm.n_eq = (int)cnt;
          break;
        }
        acc += cnt;
      }
    }
    __syncthreads();
  }
  // everything >= prefix (at the decided precision) is a candidate: n_above + n_eq <= kCand of them,
  // or exactly k when all digits were consumed
  {
    const unsigned long long thr = sm.prefix;
    for_each([&](int, bool valid, unsigned long long key) {
      if (valid && key >= thr) {
        const int pos = atomicAdd(&sm.n_cand, 1);
        if (pos < kCand) sm.cand[pos] = key;
      }
    });
  }
  __syncthreads();
  const int n_cand = min(sm.n_cand, kCand);
  int sort_n = 32;
  while (sort_n < n_cand) sort_n <<= 1;
  for (int i = n_cand + tid; i < sort_n; i += kTopkThreads) sm.cand[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= sort_n; size <<= 1) {  // bitonic sort, descending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < sort_n / 2) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sm.cand[lo], b = sm.cand[hi];
        if ((a < b) == desc) { sm.cand[lo] = b; sm.cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (tid < k) {
    const unsigned long long key = sm.cand[tid];
    out_idx[r * k + tid] = (int64_t)(unsigned int)(key & 0xffffffffull);
    if (out_val) out_val[r * k + tid] = key_float((uint32_t)(key >> 32));
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

extern "C" int lgc_topk_rows(const float* S, int64_t n_rows, int64_t n_cols, int64_t lds,
                             const int32_t* excl_ptr, const int32_t* excl_idx, int64_t row_offset, int32_t k,
                             int64_t* out_idx, float* out_val, lgc_stream_t stream) {
  LGC_REQUIRE(S && out_idx, "topk: null pointer");
  LGC_REQUIRE(n_rows > 0 && n_cols > 0 && lds >= n_cols, "topk: bad extents");
  LGC_REQUIRE(k >= 1 && k <= kTopkMaxK && k <= n_cols, "topk: k must be in [1, min(128, n_cols)]");
  LGC_REQUIRE(n_cols < (1ll << 31) && n_rows < (1ll << 31), "topk: extents exceed int32");
  LGC_REQUIRE((excl_ptr == nullptr) == (excl_idx == nullptr), "topk: excl_ptr / excl_idx mismatch");
  const size_t dyn = (size_t)((n_cols + 31) / 32) * sizeof(unsigned int);
  LGC_REQUIRE(dyn <= 160 * 1024, "topk: more than 1.3M columns per row is not supported");
  cudaStream_t st = (cudaStream_t)stream;
#define LGC_TOPK_LAUNCH(RC)                                                                                  \
  do {                                                                                                       \
    static size_t dyn_set = 32 * 1024;                                                                       \
    if (dyn > dyn_set) {                                                                                     \
      LGC_CUDA(cudaFuncSetAttribute(topk_rows_kernel<RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
      dyn_set = dyn;                                                                                         \
    }                                                                                                        \
    topk_rows_kernel<RC><<<(unsigned)n_rows, kTopkThreads, dyn, st>>>(S, n_cols, lds, excl_ptr, excl_idx,    \
                                                                      row_offset, k, out_idx, out_val);      \
  } while (0)
  if (n_cols <= 8 * kTopkThreads) LGC_TOPK_LAUNCH(8);
  else if (n_cols <= 16 * kTopkThreads) LGC_TOPK_LAUNCH(16);
  else if (n_cols <= 32 * kTopkThreads) LGC_TOPK_LAUNCH(32);
  else LGC_TOPK_LAUNCH(0);
#undef LGC_TOPK_LAUNCH
  LGC_LAUNCH_CHECK("topk_rows_kernel");
  return LGC_OK;
}
