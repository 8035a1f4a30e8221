// metrics.cu — (N1) the six evaluation metrics of the top-k lists on the device, so that a
// lambda sweep (findLambda.py:83-116) or a training-time evaluation (train.py:147-177) never
// leaves the GPU between the top-k kernel and the six numbers.
//   lgc_metrics_accuracy : Precision / Recall / NDCG sums — /root/reference/metrics/accurate.py:11-102
//       (per user: `item in items` over the k recommendations; NDCG's ideal list is k hits, :76-86);
//   lgc_metrics_hamming  : sum_i c_i (c_i - 1), c_i = number of lists containing item i — the
//       closed form of the O(U^2) pair loop of /root/reference/metrics/diversity.py:15-63
//       (H = 1 - sum / (U (U-1) k));
//   lgc_metrics_intra    : sum_u sum_{a != b in L_u} C[a,b] / sqrt(k_a k_b), C = A^T A — the closed
//       form of diversity.py:66-115 (C comes from the exact int8 tensor-core GEMM).
// All three are integer / gather work: one warp per user, the list staged in shared memory.
// Sums are accumulated in 64-bit integers where the quantity is an integer and in float64
// otherwise (the reference rounds every metric to 5 decimals).
#include "common.cuh"

namespace lgc {

constexpr int kMetWarps = 8;
constexpr int kMetMaxK = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// out[0] += hits (exact, as double of an integer), out[1] += hits_u / |pos_u|, out[2] += dcg_u / idcg,
// out[3] += 1 for every user that has at least one positive item (the keys of user_pos_items_dict)
__global__ void __launch_bounds__(kMetWarps * 32)
metrics_accuracy_kernel(const int64_t* __restrict__ rec, int64_t n_users, int k, const int32_t* __restrict__ pos_ptr,
                        const int32_t* __restrict__ pos_idx, unsigned long long* __restrict__ hits_total,
                        unsigned long long* __restrict__ users_total, double* __restrict__ out) {
  __shared__ double s_rec[kMetWarps], s_ndcg[kMetWarps];
  __shared__ int s_hits[kMetWarps], s_users[kMetWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * kMetWarps + w;
  double recall = 0.0, ndcg = 0.0;
  int hits = 0, counted = 0;
  if (u < n_users) {
    const int lo0 = pos_ptr[u], hi0 = pos_ptr[u + 1];
    if (hi0 > lo0) {
      counted = 1;
      double dcg = 0.0, idcg = 0.0;
      for (int r = lane; r < k; r += 32) {
        const int64_t item = rec[u * k + r];
        int lo = lo0, hi = hi0;
        bool hit = false;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const int32_t v = __ldg(pos_idx + mid);
          if (v == item) { hit = true; break; }
          if (v < item) lo = mid + 1; else hi = mid;
        }
        const double disc = 1.0 / log2((double)(r + 2));
        idcg += disc;
        if (hit) { dcg += disc; ++hits; }
      }
      hits = __reduce_add_sync(0xffffffffu, hits);
      dcg = warp_sum(dcg);
      idcg = warp_sum(idcg);
      recall = (double)hits / (double)(hi0 - lo0);
      ndcg = idcg > 0.0 ? dcg / idcg : 0.0;
    }
  }
  if (lane == 0) { s_rec[w] = recall; s_ndcg[w] = ndcg; s_hits[w] = hits; s_users[w] = counted; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    int h = 0, c = 0;
#pragma unroll
    for (int i = 0; i < kMetWarps; ++i) { a += s_rec[i]; b += s_ndcg[i]; h += s_hits[i]; c += s_users[i]; }
    if (c) {
      atomicAdd(hits_total, (unsigned long long)h);
      atomicAdd(users_total, (unsigned long long)c);
      atomicAdd(out + 1, a);
      atomicAdd(out + 2, b);
    }
  }
}

// c[item] += 1 for every DISTINCT item of every list (the reference intersects sets)
__global__ void __launch_bounds__(kMetWarps * 32)
metrics_hist_kernel(const int64_t* __restrict__ rec, int64_t n_users, int k, int64_t n_items, int32_t* __restrict__ counts) {
  __shared__ int64_t s_list[kMetWarps][kMetMaxK];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * kMetWarps + w;
  if (u >= n_users) return;
  for (int r = lane; r < k; r += 32) s_list[w][r] = rec[u * k + r];
  __syncwarp();
  for (int r = lane; r < k; r += 32) {
    const int64_t item = s_list[w][r];
    if (item < 0 || item >= n_items) continue;
    bool dup = false;
    for (int q = 0; q < r; ++q) dup |= s_list[w][q] == item;
    if (!dup) atomicAdd(counts + item, 1);
  }
}

__global__ void __launch_bounds__(256)
metrics_hamming_reduce_kernel(const int32_t* __restrict__ counts, int64_t n_items, unsigned long long* __restrict__ shared_pairs) {
  __shared__ unsigned long long s[8];
  unsigned long long acc = 0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long c = (unsigned long long)counts[i];
    acc += c * (c - (c ? 1ull : 0ull));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i];
    if (t) atomicAdd(shared_pairs, t);
  }
}

// sum over ordered pairs (a, b) of list positions with different item ids and non-zero degrees
__global__ void __launch_bounds__(kMetWarps * 32)
metrics_intra_kernel(const int64_t* __restrict__ rec, int64_t n_users, int k, int64_t n_items,
                     const float* __restrict__ Cmat, int64_t ldc, const int32_t* __restrict__ item_deg,
                     double* __restrict__ out) {
  __shared__ int32_t s_item[kMetWarps][kMetMaxK];
  __shared__ double s_rs[kMetWarps][kMetMaxK];  // 1 / sqrt(k_i), 0 for unusable entries
  __shared__ double s_sum[kMetWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * kMetWarps + w;
  double acc = 0.0;
  if (u < n_users) {
    for (int r = lane; r < k; r += 32) {
      const int64_t item = rec[u * k + r];
      const bool ok = item >= 0 && item < n_items;
      const int deg = ok ? item_deg[item] : 0;
      s_item[w][r] = ok ? (int32_t)item : -1;
      s_rs[w][r] = deg > 0 ? (double)deg : 0.0;
    }
    __syncwarp();
    const int pairs = k * k;
    for (int p = lane; p < pairs; p += 32) {
      const int a = p / k, b = p - a * k;
      const int ia = s_item[w][a], ib = s_item[w][b];
      const double da = s_rs[w][a], db = s_rs[w][b];
      if (ia == ib || da == 0.0 || db == 0.0) continue;  // same item id or a zero degree: skipped by the reference
      const double c = (double)__ldg(Cmat + (int64_t)ia * ldc + ib);
      acc += c / sqrt(da * db);
    }
    acc = warp_sum(acc);
  }
  if (lane == 0) s_sum[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kMetWarps; ++i) t += s_sum[i];
    if (t != 0.0) atomicAdd(out, t);
  }
}

__global__ void metrics_finish_kernel(const unsigned long long* __restrict__ ints, double* __restrict__ out) {
  // ints: [0] hits, [1] users counted, [2] shared pairs  ->  out[0], out[3], out[4] as float64
  out[0] = (double)ints[0];
  out[3] = (double)ints[1];
  out[4] = (double)ints[2];
}

}  // namespace lgc

using namespace lgc;

extern "C" int64_t lgc_metrics_scratch_bytes(int64_t n_items) {
  return (int64_t)align_up((size_t)n_items * sizeof(int32_t), 16) + 64;
}

extern "C" int lgc_metrics_topk(const int64_t* rec, int64_t n_users, int32_t k, int64_t n_items,
                                const int32_t* pos_ptr, const int32_t* pos_idx, const float* Cmat, int64_t ldc,
                                const int32_t* item_deg, double* out6, void* scratch, lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(rec && out6 && scratch, "metrics: null pointer");
  LGC_REQUIRE(n_users > 0 && n_items > 0 && k >= 1 && k <= kMetMaxK, "metrics: k must be in [1, 128]");
  LGC_REQUIRE((pos_ptr == nullptr) == (pos_idx == nullptr), "metrics: pos_ptr / pos_idx mismatch");
  LGC_REQUIRE((Cmat == nullptr) == (item_deg == nullptr), "metrics: co-occurrence matrix and item degrees go together");
  LGC_REQUIRE(!Cmat || ldc >= n_items, "metrics: ldc < n_items");
  LGC_REQUIRE(((uintptr_t)scratch & 15) == 0, "metrics: scratch must be 16-byte aligned");
  int32_t* counts = (int32_t*)scratch;
  const size_t counts_bytes = align_up((size_t)n_items * sizeof(int32_t), 16);
  unsigned long long* ints = (unsigned long long*)((char*)scratch + counts_bytes);
  LGC_CUDA(cudaMemsetAsync(scratch, 0, counts_bytes + 64, stream));
  LGC_CUDA(cudaMemsetAsync(out6, 0, 6 * sizeof(double), stream));
  const unsigned grid = (unsigned)ceil_div(n_users, kMetWarps);
  if (pos_ptr) {
    metrics_accuracy_kernel<<<grid, kMetWarps * 32, 0, stream>>>(rec, n_users, k, pos_ptr, pos_idx, ints, ints + 1, out6);
    LGC_LAUNCH_CHECK("metrics_accuracy_kernel");
  }
  metrics_hist_kernel<<<grid, kMetWarps * 32, 0, stream>>>(rec, n_users, k, n_items, counts);
  LGC_LAUNCH_CHECK("metrics_hist_kernel");
  int rb = (int)ceil_div(n_items, 256);
  if (rb > 4 * num_sms()) rb = 4 * num_sms();
  metrics_hamming_reduce_kernel<<<rb, 256, 0, stream>>>(counts, n_items, ints + 2);
  LGC_LAUNCH_CHECK("metrics_hamming_reduce_kernel");
  if (Cmat) {
    metrics_intra_kernel<<<grid, kMetWarps * 32, 0, stream>>>(rec, n_users, k, n_items, Cmat, ldc, item_deg, out6 + 5);
    LGC_LAUNCH_CHECK("metrics_intra_kernel");
  }
  metrics_finish_kernel<<<1, 1, 0, stream>>>(ints, out6);
  LGC_LAUNCH_CHECK("metrics_finish_kernel");
  return LGC_OK;
}
