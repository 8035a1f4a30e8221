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
    sa[threadIdx.x] = i < n ? pow((double)ki[i], 1.0 - lambda) : 1.0;
  } else if (threadIdx.x < 64) {
    const int64_t j = j0 + threadIdx.x - 32;
    sb[threadIdx.x - 32] = j < n ? pow((double)ki[j], lambda) : 1.0;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = i0 + r, j = j0 + tx;
    float w = 0.f;
    if (i < n && j < n) {
      double den = sa[r] * sb[tx];
      if (den == 0.0) den = 1.0;
      w = (float)((double)G[i * ldg + j] / den);
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

extern "C" int hs_hadamard(float* F, const float* Gscore, int64_t rows, int64_t cols, int64_t ldf, int64_t ldg,
                           lgc_stream_t stream) {
  LGC_REQUIRE(F && Gscore && rows > 0 && cols > 0 && ldf >= cols && ldg >= cols, "hadamard: bad arguments");
  dim3 grid((unsigned)ceil_div(cols, 256), (unsigned)(rows < 65535 ? rows : 65535));
  hadamard_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, Gscore, rows, cols, ldf, ldg);
  LGC_LAUNCH_CHECK("hadamard_kernel");
  return LGC_OK;
}
