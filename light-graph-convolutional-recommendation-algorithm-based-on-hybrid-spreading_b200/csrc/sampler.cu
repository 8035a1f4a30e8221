// sampler.cu — (N3) structured negative sampling on the device: the step between the forward
// pass and the loss of a LightGCN training iteration.
// Stands in for torch_geometric.utils.structured_negative_sampling as called at
// /root/reference/model/LightGCN/loss.py:58 and evaluation.py:72 (PyG 2.6.1, restated in
// oracle/lightgcn_oracle.py): for an edge (u, pos) draw neg ~ U[0, num_nodes) and re-draw while
// (u, neg) is a positive pair (and, with contains_neg_self_loops = False, while neg == u).  The
// reference moves all E edges to the host and runs np.isin over E log E keys on every training
// step; here one thread per requested triplet draws from a counter-based generator and tests
// membership by binary search in the user's sorted positive row (CSR), so only the batch is
// touched and nothing leaves the device.  The RNG stream necessarily differs from torch's (the
// reference's own batch choice is unseeded, loss.py:64); the distribution is the same.
#include "common.cuh"

namespace lgc {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void negative_sample_kernel(const int64_t* __restrict__ edge_u, const int64_t* __restrict__ edge_p,
                                       const int64_t* __restrict__ rows, int64_t n_out, int64_t n_edges,
                                       const int32_t* __restrict__ pos_ptr, const int32_t* __restrict__ pos_idx,
                                       int64_t n_users, int64_t num_nodes, int forbid_self, unsigned long long seed,
                                       int64_t* __restrict__ out_u, int64_t* __restrict__ out_p,
                                       int64_t* __restrict__ out_n, int* __restrict__ bad) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const int64_t e = rows ? rows[t] : t;
  if (e < 0 || e >= n_edges) { atomicExch(bad, 1); return; }
  const int64_t u = edge_u[e];
  if (u < 0 || u >= n_users) { atomicExch(bad, 2); return; }
  const int lo0 = pos_ptr[u], hi0 = pos_ptr[u + 1];
  unsigned long long state = splitmix64(seed ^ splitmix64((unsigned long long)t));
  int64_t neg = 0;
  for (int attempt = 0; attempt < 100000; ++attempt) {
    state = splitmix64(state);
    neg = (int64_t)__umul64hi(state, (unsigned long long)num_nodes);  // uniform on [0, num_nodes)
    if (forbid_self && neg == u) continue;
    int lo = lo0, hi = hi0;
    bool hit = false;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int32_t v = __ldg(pos_idx + mid);
      if (v == neg) { hit = true; break; }
      if (v < neg) lo = mid + 1; else hi = mid;
    }
    if (!hit) break;
  }
  out_u[t] = u;
  out_p[t] = edge_p[e];
  out_n[t] = neg;
}

}  // namespace lgc

using namespace lgc;

extern "C" int lgc_negative_sample(const int64_t* edge_u, const int64_t* edge_p, int64_t n_edges, const int64_t* rows,
                                   int64_t n_out, const int32_t* pos_ptr, const int32_t* pos_idx, int64_t n_users,
                                   int64_t num_nodes, int32_t forbid_self, uint64_t seed, int64_t* out_u,
                                   int64_t* out_p, int64_t* out_n, int32_t* status, lgc_stream_t stream) {
  LGC_REQUIRE(edge_u && edge_p && pos_ptr && pos_idx && out_u && out_p && out_n && status, "negative_sample: null pointer");
  LGC_REQUIRE(n_edges > 0 && n_out > 0 && n_users > 0 && num_nodes > 0, "negative_sample: empty problem");
  negative_sample_kernel<<<(unsigned)ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream>>>(
      edge_u, edge_p, rows, n_out, n_edges, pos_ptr, pos_idx, n_users, num_nodes, forbid_self,
      (unsigned long long)seed, out_u, out_p, out_n, status);
  LGC_LAUNCH_CHECK("negative_sample_kernel");
  return LGC_OK;
}
