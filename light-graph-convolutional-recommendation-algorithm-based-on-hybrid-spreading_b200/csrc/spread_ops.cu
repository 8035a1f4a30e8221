// spread_ops.cu — (S0/S2/F1) the memory-bound passes around the two spreading GEMMs:
//   hs_degrees / hs_pack_a / hs_pack_at : interaction list -> deduplicated degrees and the
//       dense K-major tensor-core operands, standing in for the dense float64 A built by a
//       Python iterrows loop at /root/reference/utils/trans.py:13-29 and for the
//       `A.T / user_degrees` operand of /root/reference/model/SpreadMethod/model.py:21-25;
//   hs_scale_w  : HybridS degree scaling (/root/reference/model/SpreadMethod/model.py:72-83,
//       three M^2 passes + two np.power in the reference) as ONE tiled pass that also emits
//       the transposed bf16 hi/mid/lo planes the F = A.W GEMM consumes;
//   hs_hadamard : F_new = G * F (/root/reference/model/SpreadLightGCN/model.py:151).
#include <cuda_bf16.h>

#include "common.cuh"

namespace lgc {

__device__ int g_degrees_bad;  // set when an interaction id is out of range

__global__ void degrees_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t nnz,
                               int64_t n_users, int64_t n_items, int32_t* __restrict__ ku, int32_t* __restrict__ ki,
                               unsigned int* __restrict__ bitmap) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int64_t u = users[e], i = items[e];
  if (u < 0 || i < 0 || u >= n_users || i >= n_items) { atomicExch(&g_degrees_bad, 1); return; }
  const int64_t bit = u * n_items + i;
  const unsigned int m = 1u << (bit & 31);
  const unsigned int old = atomicOr(bitmap + (bit >> 5), m);
  if (!(old & m)) {  // first occurrence of this (u, i) pair
    atomicAdd(ku + u, 1);
    atomicAdd(ki + i, 1);
  }
}

__global__ void pack_a_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t nnz,
                              uint16_t* __restrict__ A, int64_t ldk) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  A[(int64_t)users[e] * ldk + items[e]] = 0x3F80;  // bf16 1.0
}

__global__ void pack_a_u8_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t nnz,
                                 uint8_t* __restrict__ A, int64_t ldk) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  A[(int64_t)users[e] * ldk + items[e]] = 1;
}

__global__ void pack_at_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ items, int64_t nnz,
                               const int32_t* __restrict__ ku, int shift, int digits, uint8_t* __restrict__ At,
                               uint8_t* __restrict__ Q, int64_t ldk, int64_t plane_stride) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int64_t u = users[e], i = items[e];
  const int64_t off = i * ldk + u;
  if (At) At[off] = 1;
  if (Q) {
    const unsigned long long k = (unsigned long long)max(ku[u], 1);
    const unsigned long long q = ((1ull << shift) + k / 2) / k;  // round(2^shift / k_u)
    for (int d = 0; d < digits; ++d) Q[d * plane_stride + off] = (uint8_t)((q >> (8 * d)) & 255ull);
  }
}

// 1/x with the reference's "zero denominator -> 1" rule.  den = a_i * b_j is zero iff a factor is zero,
// i.e. iff item i or j has no interactions — and then G[i,j] is 0 as well, so G * inv(a) * inv(b) equals the
// reference's G / den (up to 2 ulp of float64, far below the single fp32 rounding that follows) while the
// per-element float64 division (the former bottleneck of this pass) becomes two multiplications.
__device__ __forceinline__ double inv_or_one(double x) { return x == 0.0 ? 1.0 : 1.0 / x; }

// W = G / (k_i^(1-l) k_j^l), 32x32 tiles, float64 arithmetic like the reference.
__global__ void __launch_bounds__(256)
scale_w_kernel(const float* __restrict__ G, int64_t ldg, int64_t n, const int32_t* __restrict__ ki, double lambda,
               float* __restrict__ W32, int64_t ldw, __nv_bfloat16* __restrict__ Wt, int64_t ldk,
               int64_t plane_stride, int planes) {
  __shared__ float tile[32][33];
  __shared__ double sa[32], sb[32];
  const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (threadIdx.x < 32) {
    const int64_t i = i0 + threadIdx.x;
    sa[threadIdx.x] = inv_or_one(i < n ? pow((double)ki[i], 1.0 - lambda) : 1.0);
  } else if (threadIdx.x < 64) {
    const int64_t j = j0 + threadIdx.x - 32;
    sb[threadIdx.x - 32] = inv_or_one(j < n ? pow((double)ki[j], lambda) : 1.0);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = i0 + r, j = j0 + tx;
    float w = 0.f;
    if (i < n && j < n) {
      w = (float)((double)G[i * ldg + j] * sa[r] * sb[tx]);
      if (W32) W32[i * ldw + j] = w;
    }
    tile[r][tx] = w;
  }
  if (!Wt) return;
  __syncthreads();
  // transposed, split write: plane[p][j, i]
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = j0 + r, i = i0 + tx;
    if (i < n && j < n) {
      float w = tile[tx][r];
      for (int p = 0; p < planes; ++p) {
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        Wt[p * plane_stride + j * ldk + i] = h;
        w -= __bfloat162float(h);  // exact: the residual fits in fp32
      }
    }
  }
}

// inv_a[i] = 1 / k_i^(1-l), inv_b[j] = 1 / k_j^l (float64, "zero -> 1"), once per lambda: keeps the slow
// float64 pow out of the M^2 passes (it used to serialise two warps per tile while six waited).
__global__ void scale_prep_kernel(const int32_t* __restrict__ ki, int64_t n, double lambda, double* __restrict__ inv_a,
                                  double* __restrict__ inv_b) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double k = (double)ki[i];
  inv_a[i] = inv_or_one(pow(k, 1.0 - lambda));
  inv_b[i] = inv_or_one(pow(k, lambda));
}

// column maxima of W = G / (k_i^(1-l) k_j^l): cmax[j] = max_i G[i,j] inv_a[i] (W >= 0; the column factor is
// applied by the consumer).  fp32 is enough — the value only selects a power-of-two scale.
// grid (column strips of 128, row chunks of 256); a warp reads 512 contiguous bytes per row (float4 per lane).
constexpr int kCmCols = 128, kCmRows = 256;

__global__ void __launch_bounds__(256)
colmax_w_kernel(const float* __restrict__ G, int64_t ldg, int64_t n, const double* __restrict__ inv_a,
                unsigned int* __restrict__ cmax_bits) {
  __shared__ float sa[kCmRows];
  __shared__ float4 smax[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * kCmCols + lane * 4;
  const int64_t i0 = (int64_t)blockIdx.y * kCmRows;
  {
    const int64_t i = i0 + threadIdx.x;
    sa[threadIdx.x] = i < n ? (float)inv_a[i] : 0.f;
  }
  __syncthreads();
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (ldg & 3) == 0 && ((uintptr_t)G & 15) == 0 && j + 3 < n;
  if (j < n) {
    // 32 rows per thread, 8 independent 128-bit loads in flight at a time (no early exit inside a group)
#pragma unroll 1
    for (int r0 = w; r0 < kCmRows; r0 += 64) {
      float4 g[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + r0 + 8 * u;
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
          if (vec) {
            g[u] = __ldg(reinterpret_cast<const float4*>(G + i * ldg + j));
          } else {
            g[u].x = G[i * ldg + j];
            g[u].y = j + 1 < n ? G[i * ldg + j + 1] : 0.f;
            g[u].z = j + 2 < n ? G[i * ldg + j + 2] : 0.f;
            g[u].w = j + 3 < n ? G[i * ldg + j + 3] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float a = sa[r0 + 8 * u];
        m.x = fmaxf(m.x, g[u].x * a); m.y = fmaxf(m.y, g[u].y * a);
        m.z = fmaxf(m.z, g[u].z * a); m.w = fmaxf(m.w, g[u].w * a);
      }
    }
  }
  smax[w][lane] = m;
  __syncthreads();
  if (w == 0 && j < n) {
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float4 o = smax[q][lane];
      m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
    }
    // non-negative floats order like their bit patterns
    atomicMax(cmax_bits + j, __float_as_uint(m.x));
    if (j + 1 < n) atomicMax(cmax_bits + j + 1, __float_as_uint(m.y));
    if (j + 2 < n) atomicMax(cmax_bits + j + 2, __float_as_uint(m.z));
    if (j + 3 < n) atomicMax(cmax_bits + j + 3, __float_as_uint(m.w));
  }
}

// W = G / (k_i^(1-l) k_j^l) quantised per column j to `digits` base-256 digits of
// q = round(W / s_j * 256^digits), s_j = the power of two strictly above the column maximum; the digit planes
// are written transposed (plane[d][j, i], K = source item i contiguous) for the u8 F = A.W GEMM, and
// cs[j] = s_j / 256^digits is the epilogue column scale.  One CTA quantises a 128 (i) x 64 (j) tile (256 B
// row segments of G), transposes it through shared memory and stores four source items per 32-bit word, so
// every plane row receives 128 contiguous bytes.
constexpr int kSwCols = 64, kSwRows = 128, kSwLd = kSwRows + 4;

__global__ void __launch_bounds__(256)
scale_w_u8_kernel(const float* __restrict__ G, int64_t ldg, int64_t n, const double* __restrict__ inv_a,
                  const double* __restrict__ inv_b, const unsigned int* __restrict__ cmax_bits,
                  float* __restrict__ W32, int64_t ldw, uint8_t* __restrict__ Wt, int64_t ldk,
                  int64_t plane_stride, int digits, float* __restrict__ cs) {
  __shared__ __align__(16) unsigned int tile[kSwCols][kSwLd];
  __shared__ float sa[kSwRows], sb[kSwCols], sinv[kSwCols];
  const int64_t j0 = (int64_t)blockIdx.x * kSwCols, i0 = (int64_t)blockIdx.y * kSwRows;
  if (threadIdx.x < kSwRows) {
    const int64_t i = i0 + threadIdx.x;
    sa[threadIdx.x] = i < n ? (float)inv_a[i] : 0.f;
  } else if (threadIdx.x < kSwRows + kSwCols) {
    const int t = threadIdx.x - kSwRows;
    const int64_t j = j0 + t;
    const float b = j < n ? (float)inv_b[j] : 1.f;
    // s_j = 2^e, the power of two strictly above the column maximum (1.0 for an all-zero column); 1 ulp of
    // head-room covers the rounding of the maximum.  sinv = 256^digits / s_j is a power of two: the scaling below
    // is exact.
    int e = 0;
    if (j < n) {
      const float cm = __uint_as_float(cmax_bits[j]) * b * 1.000001f;
      if (cm > 0.f) frexpf(cm, &e);
      if (blockIdx.y == 0 && cs) cs[j] = ldexpf(1.f, e - 8 * digits);
    }
    sb[t] = b;
    sinv[t] = ldexpf(1.f, 8 * digits - e);
  }
  __syncthreads();
  // fp32 arithmetic: W = G * (1/k_i^(1-l)) * (1/k_j^l) carries <= 4 roundings (2.4e-7 relative, the two factors
  // come from float64 pow) — the same W the fp32 output holds; the digits then represent that fp32 value to
  // 2^-(8 digits) of the column maximum (cvt.rni saturates at 2^32 - 1).
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 columns x 4 row lanes
  const int64_t j = j0 + tx;
  const float bj = sb[tx], sj = sinv[tx];
  const float qmaxf = digits >= 4 ? 4294967040.f : ldexpf(1.f, 8 * digits) - 1.f;
#pragma unroll 1
  for (int r0 = ty; r0 < kSwRows; r0 += 32) {
    float g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t i = i0 + r0 + 4 * u;
      g[u] = (i < n && j < n) ? __ldg(G + i * ldg + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + 4 * u;
      const int64_t i = i0 + r;
      const float w = g[u] * sa[r] * bj;
      if (W32 && i < n && j < n) W32[i * ldw + j] = w;
      tile[tx][r] = __float2uint_rn(fminf(w * sj, qmaxf));
    }
  }
  __syncthreads();
  // thread -> (column c, group g of four source items): lanes run over g, so a warp stores 128 contiguous bytes
#pragma unroll
  for (int t = 0; t < (kSwCols * kSwRows / 4) / 256; ++t) {
    const int p = threadIdx.x + 256 * t;
    const int c = p >> 5, g = p & 31;
    if (j0 + c < n && i0 + 4 * g < ldk) {
      const uint4 q4 = *reinterpret_cast<const uint4*>(&tile[c][4 * g]);
      uint8_t* dst = Wt + (j0 + c) * ldk + i0 + 4 * g;
      for (int d = 0; d < digits; ++d) {
        const int sh = 8 * d;
        const unsigned int word = ((q4.x >> sh) & 255u) | (((q4.y >> sh) & 255u) << 8) |
                                  (((q4.z >> sh) & 255u) << 16) | (((q4.w >> sh) & 255u) << 24);
        *reinterpret_cast<unsigned int*>(dst + d * plane_stride) = word;
      }
    }
  }
}

__global__ void hadamard_kernel(float* __restrict__ F, const float* __restrict__ Gs, int64_t rows, int64_t cols,
                                int64_t ldf, int64_t ldg) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) F[r * ldf + c] *= Gs[r * ldg + c];
}

}  // namespace lgc

using namespace lgc;

extern "C" int hs_degrees(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
                          int64_t n_items, int32_t* ku, int32_t* ki, uint8_t* dedup_bitmap,
                          lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(users && items && ku && ki && dedup_bitmap && nnz > 0, "degrees: null pointer / empty");
  LGC_REQUIRE(((uintptr_t)dedup_bitmap & 3) == 0, "degrees: bitmap must be 4-byte aligned");
  int h_bad = 0;
  LGC_CUDA(cudaMemcpyToSymbolAsync(g_degrees_bad, &h_bad, sizeof(int), 0, cudaMemcpyHostToDevice, stream));
  degrees_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, stream>>>(users, items, nnz, n_users, n_items, ku, ki,
                                                                   (unsigned int*)dedup_bitmap);
  LGC_LAUNCH_CHECK("degrees_kernel");
  LGC_CUDA(cudaMemcpyFromSymbolAsync(&h_bad, g_degrees_bad, sizeof(int), 0, cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) LGC_FAIL(LGC_ERR_INVALID, "degrees: (user, item) id out of range");
  return LGC_OK;
}

extern "C" int hs_pack_a(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
                         int64_t n_items, uint16_t* A_bf16, int64_t ldk, lgc_stream_t stream) {
  LGC_REQUIRE(users && items && A_bf16 && nnz > 0, "pack_a: null pointer / empty");
  LGC_REQUIRE(ldk >= n_items && n_users > 0, "pack_a: ldk < n_items");
  pack_a_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(users, items, nnz, A_bf16, ldk);
  LGC_LAUNCH_CHECK("pack_a_kernel");
  return LGC_OK;
}

extern "C" int hs_pack_a_u8(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
                            int64_t n_items, uint8_t* A_u8, int64_t ldk, lgc_stream_t stream) {
  LGC_REQUIRE(users && items && A_u8 && nnz > 0, "pack_a_u8: null pointer / empty");
  LGC_REQUIRE(ldk >= n_items && n_users > 0, "pack_a_u8: ldk < n_items");
  pack_a_u8_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(users, items, nnz, A_u8, ldk);
  LGC_LAUNCH_CHECK("pack_a_u8_kernel");
  return LGC_OK;
}

extern "C" int hs_pack_at(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
                          int64_t n_items, const int32_t* ku, int32_t shift, int32_t digits, uint8_t* At_u8,
                          uint8_t* Q_u8, int64_t ldk, int64_t plane_stride, lgc_stream_t stream) {
  LGC_REQUIRE(users && items && nnz > 0 && (At_u8 || Q_u8), "pack_at: null pointer / empty");
  LGC_REQUIRE(ldk >= n_users && n_items > 0, "pack_at: ldk < n_users");
  if (Q_u8) {
    LGC_REQUIRE(ku, "pack_at: digits need user degrees");
    LGC_REQUIRE(digits >= 1 && digits <= 4 && shift >= 0 && shift < 8 * digits + 31 && shift <= 62,
                "pack_at: bad digits / shift");
    LGC_REQUIRE(plane_stride >= n_items * ldk, "pack_at: plane stride overlaps planes");
  }
  pack_at_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(users, items, nnz, ku, shift, digits,
                                                                                 At_u8, Q_u8, ldk, plane_stride);
  LGC_LAUNCH_CHECK("pack_at_kernel");
  return LGC_OK;
}

extern "C" int hs_scale_w(const float* G, int64_t ldg, int64_t n, const int32_t* ki, double lambda, float* W32,
                          int64_t ldw, uint16_t* Wt_planes, int64_t ldk, int64_t plane_stride, int32_t planes,
                          lgc_stream_t stream) {
  LGC_REQUIRE(G && ki && n > 0 && ldg >= n, "scale_w: bad arguments");
  LGC_REQUIRE(W32 || Wt_planes, "scale_w: no output requested");
  LGC_REQUIRE(!W32 || ldw >= n, "scale_w: ldw < n");
  if (Wt_planes) {
    LGC_REQUIRE(planes >= 1 && planes <= 3 && ldk >= n, "scale_w: planes in 1..3, ldk >= n");
    LGC_REQUIRE(planes == 1 || plane_stride >= n * ldk, "scale_w: plane stride overlaps planes");
  }
  dim3 grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
  scale_w_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(G, ldg, n, ki, lambda, W32, ldw,
                                                         (__nv_bfloat16*)Wt_planes, ldk, plane_stride, planes);
  LGC_LAUNCH_CHECK("scale_w_kernel");
  return LGC_OK;
}

extern "C" int hs_scale_w_u8(const float* G, int64_t ldg, int64_t n, const int32_t* ki, double lambda, float* W32,
                             int64_t ldw, uint8_t* Wt_digits, int64_t ldk, int64_t plane_stride, int32_t digits,
                             float* col_scale, void* scratch, lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(G && ki && Wt_digits && col_scale && scratch && n > 0 && ldg >= n, "scale_w_u8: bad arguments");
  LGC_REQUIRE(((uintptr_t)scratch & 7) == 0, "scale_w_u8: scratch must be 8-byte aligned");
  double* inv_a = (double*)scratch;          // scratch layout: n doubles, n doubles, n uint32
  double* inv_b = inv_a + n;
  LGC_REQUIRE(digits >= 1 && digits <= 4 && ldk >= n, "scale_w_u8: digits in 1..4, ldk >= n");
  LGC_REQUIRE(digits == 1 || plane_stride >= n * ldk, "scale_w_u8: plane stride overlaps planes");
  LGC_REQUIRE(!W32 || ldw >= n, "scale_w_u8: ldw < n");
  LGC_REQUIRE((ldk & 3) == 0 && (plane_stride & 3) == 0 && ((uintptr_t)Wt_digits & 3) == 0,
              "scale_w_u8: digit planes must be 4-byte aligned with ldk and plane stride multiples of 4");
  uint32_t* colmax_scratch = (uint32_t*)(inv_b + n);
  LGC_CUDA(cudaMemsetAsync(colmax_scratch, 0, sizeof(uint32_t) * (size_t)n, stream));
  scale_prep_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ki, n, lambda, inv_a, inv_b);
  LGC_LAUNCH_CHECK("scale_prep_kernel");
  dim3 g1((unsigned)ceil_div(n, kCmCols), (unsigned)ceil_div(n, kCmRows));
  colmax_w_kernel<<<g1, 256, 0, stream>>>(G, ldg, n, inv_a, colmax_scratch);
  LGC_LAUNCH_CHECK("colmax_w_kernel");
  dim3 g2((unsigned)ceil_div(n, kSwCols), (unsigned)ceil_div(n, kSwRows));
  scale_w_u8_kernel<<<g2, 256, 0, stream>>>(G, ldg, n, inv_a, inv_b, colmax_scratch, W32, ldw, Wt_digits, ldk,
                                           plane_stride, digits, col_scale);
  LGC_LAUNCH_CHECK("scale_w_u8_kernel");
  return LGC_OK;
}

extern "C" int hs_hadamard(float* F, const float* Gscore, int64_t rows, int64_t cols, int64_t ldf, int64_t ldg,
                           lgc_stream_t stream) {
  LGC_REQUIRE(F && Gscore && rows > 0 && cols > 0 && ldf >= cols && ldg >= cols, "hadamard: bad arguments");
  dim3 grid((unsigned)ceil_div(cols, 256), (unsigned)(rows < 65535 ? rows : 65535));
  hadamard_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, Gscore, rows, cols, ldf, ldg);
  LGC_LAUNCH_CHECK("hadamard_kernel");
  return LGC_OK;
}
