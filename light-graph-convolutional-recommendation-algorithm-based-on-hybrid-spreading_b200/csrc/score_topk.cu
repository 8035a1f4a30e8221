// score_topk.cu — (P8 / S4) dense score block + row-wise masked top-k.
//   lgc_score_block : score = Xu . Xi^T with the seen-pair fill, standing in for
//       torch.matmul + score[users, items] = -1024 at
//       /root/reference/model/LightGCN/recommend.py:86,101,111 (== evaluation.py:34,49 and
//       SpreadLightGCN/model.py:77,92,102).  fp32 FMA on CUDA cores: this file is the bit-exact
//       path (matches an fp32 SGEMM to ~1e-7; precise=True / LGCNHS_SCORE_FP32=1, k > 32, the
//       fusion multiplier at other dims); the default full-rank eval runs on the tensor cores
//       (score_topk_tc.cu, 3xTF32 split, 1e-5 tolerance).
//   lgc_topk_rows   : torch.topk(score, k) (recommend.py:114) and the reference's real
//       bottleneck, np.argsort(row)[::-1] + Python membership filter
//       (/root/reference/model/SpreadMethod/recommend.py:35-47, 32.8 ms/user measured), as one
//       radix-select pass structure per row: exact, deterministic, ties -> larger index first.
#include "common.cuh"
#include "select.cuh"

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
// row-wise masked top-k: ONE WARP per row, threshold + candidate buffer (select.cuh)
// ------------------------------------------------------------------------------------------
// Each lane streams its share of the row once (128-bit loads, 4 in flight per lane, evict-first)
// and the hot loop is one 32-bit compare per value against the row's current k-th best.  The few
// survivors (~k ln(n/k) per row) go through the exclusion test (one bit of the bit-packed
// rows x n_cols mask), are appended to the warp's shared-memory buffer with a ballot/prefix, and
// the buffer is compacted by an in-register bitonic sort when it fills.  Exact for any
// distribution; deterministic; equal values rank the LARGER column first.  If fewer than k
// columns are selectable the tail is (-1, -inf).
constexpr int kTopkWarps = 8;
constexpr int kTopkMaxK = 128;

template <int CAP, bool VEC>
__global__ void __launch_bounds__(kTopkWarps * 32)
topk_rows_kernel(const float* __restrict__ S, int n_rows, int n, int64_t lds, const uint32_t* __restrict__ mask,
                 int64_t mask_stride_bits, int64_t row_offset, int k, int64_t* __restrict__ out_idx,
                 float* __restrict__ out_val) {
  constexpr int E = CAP / 32;
  __shared__ unsigned long long s_buf[kTopkWarps][CAP];
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kTopkWarps + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  unsigned long long* buf = s_buf[threadIdx.x >> 5];
  const float* row = S + r * lds;
  const int64_t bit0 = (row_offset + r) * mask_stride_bits;
  const uint32_t* mrow = mask ? mask + (bit0 >> 5) : nullptr;
  const uint32_t boff = (uint32_t)(bit0 & 31);
  const uint32_t lt_mask = (1u << lane) - 1u;

  int cnt = 0;
  unsigned long long thr = 0ull;
  float thr_f = -INFINITY;  // value of the current k-th best: the hot loop is ONE float compare per element

  // warp-collective: every lane offers one (value key, column) pair; `live` = column in range
  auto offer = [&](uint32_t vk, int c, bool live) {
    const unsigned long long key = make_key(vk, (uint32_t)c);
    bool pass = live && key > thr;
    if (pass && mrow) {
      const uint32_t b = boff + (uint32_t)c;
      pass = !((__ldg(mrow + (b >> 5)) >> (b & 31)) & 1u);
    }
    uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (m == 0u) return;
    if (cnt + __popc(m) > CAP) {
      cnt = E >= 2 ? warp_select<E>(buf, cnt, k, lane, &thr) : warp_compact<E>(buf, cnt, k, lane, &thr);
      thr_f = thr ? key_float((uint32_t)(thr >> 32)) : -INFINITY;
      pass = pass && key > thr;
      m = __ballot_sync(0xffffffffu, pass);
    }
    if (pass) buf[cnt + __popc(m & lt_mask)] = key;
    cnt += __popc(m);
  };

  // Threshold seed (k <= 32): every lane takes the best selectable pair among its first values; the k-th largest
  // of the 32 lane maxima is a valid lower bound of the row's k-th best (they are 32 distinct selectable
  // elements), so the first iterations append a few dozen candidates instead of all of them.
  if (CAP == 64) {
    unsigned long long best[1] = {0ull};
    constexpr int SEED = VEC ? 16 : 8;
#pragma unroll 1
    for (int q = 0; q < SEED; ++q) {
      const int c = VEC ? (q >> 2) * 128 + lane * 4 + (q & 3) : q * 32 + lane;
      if (c < n) {
        const unsigned long long key = make_key(float_key(__ldg(row + c)), (uint32_t)c);
        if (key > best[0]) {
          bool ex = false;
          if (mrow) {
            const uint32_t b = boff + (uint32_t)c;
            ex = (__ldg(mrow + (b >> 5)) >> (b & 31)) & 1u;
          }
          if (!ex) best[0] = key;
        }
      }
    }
    warp_sort_desc<1>(best, lane);
    const unsigned long long seed = __shfl_sync(0xffffffffu, best[0], k - 1);
    if (seed != 0ull) {
      thr = seed - 1ull;  // the seed element itself must still pass
      thr_f = key_float((uint32_t)(seed >> 32));
    }
  }

  // One iteration = NV values per lane.  Fast path: one 32-bit compare per value builds the lane's pass bitmask;
  // slow path (a single, non-unrolled call site, so the sort network is instantiated once): rounds in which every
  // lane offers its next surviving value.
  constexpr int NV = VEC ? 16 : 8;
  constexpr int COLS = 32 * NV;
  for (int c0 = 0; c0 < n; c0 += COLS) {
    uint32_t vk[NV];  // raw value bits; indexed dynamically in the slow path -> lives in local memory (L1), the fast path stays in registers
    uint32_t bits = 0u;
    if (VEC) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * 128 + lane * 4;
        v[u] = c < n ? __ldcs(reinterpret_cast<const float4*>(row + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          vk[u * 4 + q] = __float_as_uint(f[q]);
          if (f[q] >= thr_f && c0 + u * 128 + lane * 4 + q < n) bits |= 1u << (u * 4 + q);
        }
      }
    } else {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 32 + lane;
        v[u] = c < n ? __ldcs(row + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        vk[u] = __float_as_uint(v[u]);
        if (v[u] >= thr_f && c0 + u * 32 + lane < n) bits |= 1u << u;
      }
    }
#pragma unroll 1
    while (__any_sync(0xffffffffu, bits != 0u)) {
      const bool has = bits != 0u;
      const int q = has ? __ffs(bits) - 1 : 0;
      bits &= bits - 1u;
      const int c = VEC ? c0 + (q >> 2) * 128 + lane * 4 + (q & 3) : c0 + q * 32 + lane;
      offer(float_key(__uint_as_float(vk[q])), c, has);
    }
  }

  cnt = warp_compact<E>(buf, cnt, k, lane, &thr);
  for (int i = lane; i < k; i += 32) {
    const unsigned long long key = buf[i];
    out_idx[r * k + i] = i < cnt ? (int64_t)(uint32_t)(key & 0xffffffffull) : -1;
    if (out_val) out_val[r * k + i] = i < cnt ? key_float((uint32_t)(key >> 32)) : -INFINITY;
  }
}

// ------------------------------------------------------------------------------------------
// fused score + seen-pair rule + top-k: the (U, M) score matrix never exists
// ------------------------------------------------------------------------------------------
// One CTA owns 64 users and walks over ALL items in tiles of 128: the tile's item embeddings are
// staged (transposed) in shared memory while the next tile's rows are already in flight in
// registers; 256 threads hold a 4 x 8 register tile of fp32-FMA dot products each.  A score that
// beats its row's current k-th best is appended to that row's candidate buffer in shared memory
// (select.cuh); full buffers are compacted by one warp each.  Seen pairs (CSR, binary search, only
// evaluated for the few survivors) either take the fill value (-1024: the reference's
// score[users, items] = -1024 followed by torch.topk) or are dropped (the filtered lists of the
// spreading family); an optional multiplier matrix fuses the Hadamard product F_new = G * F of
// /root/reference/model/SpreadLightGCN/model.py:151 into the same pass.
constexpr int kFTI = 128;
static int g_score_topk_threads = 512;  // CTA size of the 128-user tile variant: 512 threads (4 x 8 register tile, 16
                                        // warps hide the LDS latency) measured 19.1 ms vs 21.4 ms for 256 threads (8 x 8)
                                        // on the amazon-book shape (lgc_score_topk_config)

// d = a * b + c on two packed fp32 lanes (Blackwell FFMA2): each half is an IEEE round-to-nearest fmaf, so the
// result is bit-identical to the scalar loop of lgc_score_block at half the FMA-pipe issue slots (a scalar
// FFMA issues every second cycle per scheduler on sm_100: the kernel measured 26 TFLOP/s with them).
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long dup2(float a) {  // (a, a)
  unsigned long long d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(a));
  return d;
}

template <int DIM, int TU, int CAP>
constexpr size_t score_topk_smem() {
  return (size_t)DIM * (TU + 4) * 4 + 2 * (size_t)DIM * (kFTI + 4) * 4 + (size_t)TU * CAP * 8 + TU * 8 + TU * 4 + TU * 4;
}

__device__ __forceinline__ bool csr_contains(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                             int64_t row, int32_t col) {
  int lo = __ldg(ptr + row), hi = __ldg(ptr + row + 1);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t v = __ldg(idx + mid);
    if (v == col) return true;
    if (v < col) lo = mid + 1; else hi = mid;
  }
  return false;
}

// TU users per CTA (64 or 128): a thread owns TU/16 users x 8 items.  The loop is bound by the bytes shared
// memory returns to registers per FMA, so the larger tile (8 x 8: 64 B per 64 FMAs) is used whenever the
// candidate buffers of 128 rows fit next to the double-buffered item tile (k <= 32).
template <int DIM, int TU, int CAP, bool MUL, int NT>
__global__ void __launch_bounds__(NT, ((TU == 64 && CAP <= 64 && NT == 256) ? 2 : 1))
score_topk_kernel(const float* __restrict__ Xu, const float* __restrict__ Xi, int64_t u0, int64_t u1, int n_items,
                  const int32_t* __restrict__ seen_ptr, const int32_t* __restrict__ seen_idx, float fill,
                  int exclude_seen, const float* __restrict__ mul, int64_t ldmul, int k,
                  int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  constexpr int E = CAP / 32;
  constexpr int UT = TU / (NT / 16);                // users per thread
  constexpr int CH = DIM / 4;                       // float4 chunks per embedding row
  constexpr int LDI = (kFTI * CH) / NT;             // item-tile float4 loads per thread
  constexpr int NW = NT / 32;                       // warps
  constexpr int NS = UT * 8;                        // scores per thread and tile
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*sU)[TU + 4] = reinterpret_cast<float (*)[TU + 4]>(smem_raw);
  // the item tile is double-buffered: tile t+1 is staged while tile t is still being read, one barrier per tile
  float (*sI0)[kFTI + 4] = reinterpret_cast<float (*)[kFTI + 4]>(smem_raw + (size_t)DIM * (TU + 4) * 4);
  unsigned long long* cand =
      reinterpret_cast<unsigned long long*>(smem_raw + (size_t)DIM * (TU + 4) * 4 + 2 * (size_t)DIM * (kFTI + 4) * 4);
  unsigned long long* thr = cand + (size_t)TU * CAP;
  int* cnt = reinterpret_cast<int*>(thr + TU);
  int* res = cnt + TU;  // entries [0, res) of a row's buffer already went through the seen-pair rule

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t ub = u0 + (int64_t)blockIdx.x * TU;

  if (tid < TU) { thr[tid] = 0ull; cnt[tid] = 0; res[tid] = 0; }
  for (int f = tid; f < TU * CH; f += NT) {
    const int r = f % TU, c = f / TU;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ub + r < u1) a = __ldg(reinterpret_cast<const float4*>(Xu + (ub + r) * DIM + c * 4));
    sU[c * 4 + 0][r] = a.x; sU[c * 4 + 1][r] = a.y; sU[c * 4 + 2][r] = a.z; sU[c * 4 + 3][r] = a.w;
  }

  float4 pre[LDI];
  auto fetch = [&](int ib) {
#pragma unroll
    for (int m = 0; m < LDI; ++m) {
      const int f = tid + NT * m, r = f % kFTI, c = f / kFTI;
      pre[m] = ib + r < n_items ? __ldg(reinterpret_cast<const float4*>(Xi + (int64_t)(ib + r) * DIM + c * 4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&](int buf) {
    float (*sI)[kFTI + 4] = sI0 + buf * DIM;
#pragma unroll
    for (int m = 0; m < LDI; ++m) {
      const int f = tid + NT * m, r = f % kFTI, c = f / kFTI;
      sI[c * 4 + 0][r] = pre[m].x; sI[c * 4 + 1][r] = pre[m].y; sI[c * 4 + 2][r] = pre[m].z; sI[c * 4 + 3][r] = pre[m].w;
    }
  };
  fetch(0);
  stage(0);
  if (kFTI < n_items) fetch(kFTI);
  __syncthreads();

  // One warp: apply the seen-pair rule to the not yet resolved entries of a row's buffer (the binary searches of
  // the lanes overlap; candidates are appended unchecked so that no global-memory latency sits between the tiles'
  // barriers), then keep the k best and raise the row's threshold.
  auto compact_row = [&](int row, int n, bool final_sort) {
    unsigned long long* base = cand + (size_t)row * CAP;
    if (seen_ptr) {
      const int r0 = res[row];
#pragma unroll 1
      for (int idx = r0 + lane; idx < n; idx += 32) {
        const unsigned long long key = base[idx];
        const int32_t item = (int32_t)(uint32_t)(key & 0xffffffffull);
        if (csr_contains(seen_ptr, seen_idx, ub + row, item)) {
          float v = fill;
          if (MUL) v *= __ldg(mul + (ub + row - u0) * ldmul + item);
          base[idx] = exclude_seen ? 0ull : make_key(float_key(v), (uint32_t)item);
        }
      }
    }
    unsigned long long t;
    const int kept = (E >= 4 && !final_sort) ? warp_select<E>(base, n, k, lane, &t) : warp_compact<E>(base, n, k, lane, &t);
    if (lane == 0) { cnt[row] = kept; thr[row] = t; res[row] = kept; }
    __syncwarp();
    return kept;
  };

  int buf = 0;
  for (int ib = 0; ib < n_items; ib += kFTI, buf ^= 1) {
    float (*sI)[kFTI + 4] = sI0 + buf * DIM;
    unsigned long long acc2[UT][4];  // [user][item pair], packed (even item, odd item)
#pragma unroll
    for (int i = 0; i < UT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
#pragma unroll 4
    for (int d = 0; d < DIM; ++d) {
      float av[UT];
#pragma unroll
      for (int h = 0; h < UT / 4; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(&sU[d][ty * UT + h * 4]);
        av[h * 4 + 0] = a.x; av[h * 4 + 1] = a.y; av[h * 4 + 2] = a.z; av[h * 4 + 3] = a.w;
      }
      const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(&sI[d][tx * 4]);
      const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(&sI[d][64 + tx * 4]);
      const unsigned long long bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < UT; ++i) {
        const unsigned long long a2 = dup2(av[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = ffma2(a2, bv[j], acc2[i][j]);
      }
    }
    // tile t+1 (in registers since before this tile's loop) goes to the other buffer, tile t+2 takes the registers
    if (ib + kFTI < n_items) stage(buf ^ 1);
    if (ib + 2 * kFTI < n_items) fetch(ib + 2 * kFTI);

    // ---- candidates: one float compare per score against the row's threshold value; survivors (a handful per
    //      row and tile) get their exact 64-bit key and go to a small local list ----
    unsigned long long pk[NS];
    unsigned char pr[NS];
    int np = 0;
#pragma unroll
    for (int i = 0; i < UT; ++i) {
      const int row = ty * UT + i;
      const bool row_ok = ub + row < u1;
      // value of the row's k-th best so far (-inf while the buffer holds fewer than k); rows past the end never pass
      const uint32_t th = (uint32_t)(thr[row] >> 32);
      const float thf = row_ok ? (th ? key_float(th) : -INFINITY) : INFINITY;
      float mv[8];
      if (MUL) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int it0 = ib + h * 64 + tx * 4;
          const float* mp = mul + (ub + row - u0) * ldmul + it0;
#pragma unroll
          for (int q = 0; q < 4; ++q) mv[h * 4 + q] = (row_ok && it0 + q < n_items) ? __ldg(mp + q) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned long long p2 = acc2[i][j >> 1];
        const float sc = __uint_as_float((j & 1) ? (uint32_t)(p2 >> 32) : (uint32_t)(p2 & 0xffffffffull));
        const float v = MUL ? sc * mv[j] : sc;
        if (v >= thf) {
          const int item = ib + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
          if (item < n_items) {
            pk[np] = make_key(float_key(v), (uint32_t)item);
            pr[np] = (unsigned char)row;
            ++np;
          }
        }
      }
    }
    while (true) {
      int keep = 0;
      for (int q = 0; q < np; ++q) {
        const unsigned long long key = pk[q];
        const int row = pr[q];
        if (key <= thr[row]) continue;
        const int slot = atomicAdd(&cnt[row], 1);
        if (slot < CAP) { cand[(size_t)row * CAP + slot] = key; continue; }
        pk[keep] = key; pr[keep] = (unsigned char)row; ++keep;
      }
      np = keep;
      // the one barrier of the tile: every thread is past the FMA loop and has staged its part of the next tile
      if (!__syncthreads_or(np > 0)) break;
      for (int rr = 0; rr < TU / NW; ++rr) {  // each warp compacts the full buffers among its TU/8 rows
        const int row = warp * (TU / NW) + rr;
        if (cnt[row] >= CAP) compact_row(row, CAP, false);
      }
      __syncthreads();
    }
  }

  for (int rr = 0; rr < TU / NW; ++rr) {
    const int row = warp * (TU / NW) + rr;
    const int64_t u = ub + row;
    if (u >= u1) continue;
    int c = cnt[row];
    c = c < CAP ? c : CAP;
    const int kept = compact_row(row, c, true);
    for (int i = lane; i < k; i += 32) {
      const unsigned long long key = cand[(size_t)row * CAP + i];
      out_idx[(u - u0) * k + i] = i < kept ? (int64_t)(uint32_t)(key & 0xffffffffull) : -1;
      if (out_val) out_val[(u - u0) * k + i] = i < kept ? key_float((uint32_t)(key >> 32)) : -INFINITY;
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
  const bool vec = (lds & 3) == 0 && ((uintptr_t)S & 15) == 0;  // every row starts 16-byte aligned
#define LGC_TOPK_LAUNCH(CAPV, VECV)                                                                          \
  topk_rows_kernel<CAPV, VECV><<<grid, kTopkWarps * 32, 0, st>>>(S, (int)n_rows, (int)n_cols, lds, excl_mask, \
                                                                 mask_stride_bits, row_offset, k, out_idx, out_val)
  if (k <= 32) {
    if (vec) LGC_TOPK_LAUNCH(64, true); else LGC_TOPK_LAUNCH(64, false);
  } else {
    if (vec) LGC_TOPK_LAUNCH(256, true); else LGC_TOPK_LAUNCH(256, false);
  }
#undef LGC_TOPK_LAUNCH
  LGC_LAUNCH_CHECK("topk_rows_kernel");
  return LGC_OK;
}

extern "C" int lgc_score_topk_config(int32_t threads) {
  LGC_REQUIRE(threads == 256 || threads == 512, "score_topk config: threads must be 256 or 512");
  g_score_topk_threads = threads;
  return LGC_OK;
}

extern "C" int lgc_score_topk(const float* Xu, const float* Xi, int64_t u0, int64_t u1, int64_t n_items, int32_t dim,
                              const int32_t* seen_ptr, const int32_t* seen_idx, float fill, int32_t exclude_seen,
                              const float* mul, int64_t ldmul, int32_t k, int64_t* out_idx, float* out_val,
                              lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(Xu && Xi && out_idx, "score_topk: null pointer");
  LGC_REQUIRE(u0 >= 0 && u1 > u0 && n_items > 0 && n_items < (1ll << 31) - kFTI, "score_topk: bad extents");
  LGC_REQUIRE(k >= 1 && k <= kTopkMaxK && k <= n_items, "score_topk: k must be in [1, min(128, n_items)]");
  LGC_REQUIRE(((uintptr_t)Xu & 15) == 0 && ((uintptr_t)Xi & 15) == 0, "score_topk: embeddings must be 16-byte aligned");
  LGC_REQUIRE((seen_ptr == nullptr) == (seen_idx == nullptr), "score_topk: seen_ptr / seen_idx mismatch");
  LGC_REQUIRE(!mul || ldmul >= n_items, "score_topk: multiplier leading dimension smaller than the row");
#define LGC_ST_LAUNCH_NT(D, TUV, CAPV, MULV, NTV)                                                               \
  do {                                                                                                          \
    static DeviceOnce attr;                                                                                        \
    constexpr size_t smem = score_topk_smem<D, TUV, CAPV>();                                                    \
    if (attr.need()) {                                                                                                \
      LGC_CUDA(cudaFuncSetAttribute(score_topk_kernel<D, TUV, CAPV, MULV, NTV>,                                 \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
      attr.mark();                                                                                               \
    }                                                                                                           \
    score_topk_kernel<D, TUV, CAPV, MULV, NTV><<<(unsigned)ceil_div(u1 - u0, TUV), NTV, smem, stream>>>(        \
        Xu, Xi, u0, u1, (int)n_items, seen_ptr, seen_idx, fill, exclude_seen, mul, ldmul, k, out_idx, out_val); \
  } while (0)
#define LGC_ST_LAUNCH(D, TUV, CAPV, MULV)                                                                       \
  do {                                                                                                          \
    if (TUV == 128 && g_score_topk_threads == 512) LGC_ST_LAUNCH_NT(D, 128, CAPV, MULV, 512);                   \
    else LGC_ST_LAUNCH_NT(D, TUV, CAPV, MULV, 256);                                                             \
  } while (0)
  // 128-user tiles (8 x 8 register tile) while the candidate buffers fit; small problems keep 64-user tiles so
  // that the grid still covers the SMs
  const bool big = (u1 - u0) >= (int64_t)128 * num_sms();
#define LGC_ST_DIM(D)                                                                                           \
  do {                                                                                                          \
    if (k <= 32) {                                                                                              \
      if (big) { if (mul) LGC_ST_LAUNCH(D, 128, 64, true); else LGC_ST_LAUNCH(D, 128, 64, false); }             \
      else { if (mul) LGC_ST_LAUNCH(D, 64, 64, true); else LGC_ST_LAUNCH(D, 64, 64, false); }                   \
    } else {                                                                                                    \
      if (mul) LGC_ST_LAUNCH(D, 64, 256, true); else LGC_ST_LAUNCH(D, 64, 256, false);                          \
    }                                                                                                           \
  } while (0)
  switch (dim) {
    case 32: LGC_ST_DIM(32); break;
    case 64: LGC_ST_DIM(64); break;
    default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "score_topk: embedding dim %d not in {32,64}", dim);
  }
#undef LGC_ST_DIM
#undef LGC_ST_LAUNCH
#undef LGC_ST_LAUNCH_NT
  LGC_LAUNCH_CHECK("score_topk_kernel");
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
