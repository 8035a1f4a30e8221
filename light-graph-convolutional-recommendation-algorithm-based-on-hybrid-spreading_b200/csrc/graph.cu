// graph.cu — (P1) once-per-graph normalisation: COO -> target-keyed CSR + D^-1/2 + the
// per-edge fp32 weight, plus the long-row chunk list used by the SpMM.
// Stands in for gcn_norm(edge_index, add_self_loops=False) at
// /root/reference/model/LightGCN/model.py:53 (PyG 2.6.1), which the reference re-runs on
// every forward although the graph never changes.
//
// The device-wide key sort and scans are this library's own kernels (ingest.cuh: stable LSD radix sort of 64-bit keys,
// two-level exclusive scan) — round 1 used CUB here and torch.unique / bincount / argsort on the Python side.
#include "common.cuh"
#include "ingest.cuh"

namespace lgc {

__global__ void make_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 int64_t nnz, int64_t n_nodes, uint64_t* __restrict__ keys,
                                 int* __restrict__ bad) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int64_t s = src[e], d = dst[e];
  if (s < 0 || d < 0 || s >= n_nodes || d >= n_nodes) {
    atomicExch(bad, 1);
    s = 0;
    d = 0;
  }
  keys[e] = ((uint64_t)d << 32) | (uint32_t)s;
}

// sorted keys -> rowptr (every gap between consecutive targets is filled), colidx
__global__ void rowptr_kernel(const uint64_t* __restrict__ keys, int64_t nnz, int64_t n_nodes,
                              int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > nnz) return;
  int64_t t_prev = (e == 0) ? -1 : (int64_t)(keys[e - 1] >> 32);
  int64_t t_cur = (e == nnz) ? n_nodes : (int64_t)(keys[e] >> 32);
  for (int64_t t = t_prev + 1; t <= t_cur; ++t) rowptr[t] = (int32_t)e;
  if (e < nnz) colidx[e] = (int32_t)(keys[e] & 0xffffffffu);
}

// deg^-1/2 exactly as torch's CPU `deg.pow(-0.5)` evaluates it: 1 / sqrt(deg) with IEEE
// sqrt and divide, inf -> 0 (gcn_norm's masked_fill).
__global__ void dinv_kernel(const int32_t* __restrict__ rowptr, int64_t n_nodes,
                            float* __restrict__ dinv, int32_t* __restrict__ n_row_chunks) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  int32_t deg = rowptr[r + 1] - rowptr[r];
  dinv[r] = deg > 0 ? __fdiv_rn(1.0f, __fsqrt_rn((float)deg)) : 0.0f;
  n_row_chunks[r] = deg > LGC_LONG_ROW ? (deg + LGC_CHUNK - 1) / LGC_CHUNK : 0;
}

__global__ void edge_val_kernel(const uint64_t* __restrict__ keys, int64_t nnz,
                                const float* __restrict__ dinv, float* __restrict__ val) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  uint64_t k = keys[e];
  // norm = dis[row] * dis[col]: one fp32 multiply of the two fp32 factors (gcn_norm)
  val[e] = __fmul_rn(dinv[(uint32_t)(k & 0xffffffffu)], dinv[(uint32_t)(k >> 32)]);
}

__global__ void chunk_fill_kernel(const int32_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ row_chunk_base, int64_t n_nodes,
                                  int32_t* __restrict__ chunk_row, int32_t* __restrict__ chunk_start) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  int32_t s = rowptr[r], deg = rowptr[r + 1] - s;
  if (deg <= LGC_LONG_ROW) return;
  int32_t nch = (deg + LGC_CHUNK - 1) / LGC_CHUNK, base = row_chunk_base[r];
  for (int32_t c = 0; c < nch; ++c) {
    chunk_row[base + c] = (int32_t)r;
    chunk_start[base + c] = s + c * LGC_CHUNK;
  }
}

struct CsrWs {
  size_t keys_in, keys_out, n_row_chunks, bad, table, total;
};

static int csr_ws_layout(int64_t nnz, int64_t n_nodes, CsrWs* w) {
  size_t off = 0;
  w->keys_in = off;
  off += align_up(sizeof(uint64_t) * (size_t)nnz, 256);
  w->keys_out = off;
  off += align_up(sizeof(uint64_t) * (size_t)nnz, 256);
  w->n_row_chunks = off;
  off += align_up(sizeof(int32_t) * (size_t)(n_nodes + 1), 256);
  w->bad = off;
  off += 256;
  w->table = off;
  size_t words = ingest::sort_table_words(nnz);
  const size_t scan_words = ingest::scan_scratch_words(n_nodes + 1);
  if (scan_words > words) words = scan_words;
  off += align_up(words * sizeof(uint32_t), 256);
  w->total = off;
  return 0;
}

// (user, item) pairs -> key = user * n_items + item
__global__ void pair_keys_kernel(const int64_t* __restrict__ users, const int64_t* __restrict__ items, int64_t n,
                                 int64_t n_users, int64_t n_items, uint64_t* __restrict__ keys, int* __restrict__ bad) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  int64_t u = users[e], i = items[e];
  if (u < 0 || i < 0 || u >= n_users || i >= n_items) {
    atomicExch(bad, 1);
    u = 0;
    i = 0;
  }
  keys[e] = (uint64_t)u * (uint64_t)n_items + (uint64_t)i;
}

__global__ void head_flags_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  flags[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1u : 0u;
}

// sorted keys + exclusive scan of the head flags -> the distinct keys, ascending
__global__ void compact_unique_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos, int64_t n,
                                      uint64_t* __restrict__ out, int64_t* __restrict__ n_unique) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool head = e == 0 || keys[e] != keys[e - 1];
  if (head) out[pos[e]] = keys[e];
  if (e == n - 1) *n_unique = (int64_t)pos[e] + (head ? 1 : 0);
}

// sorted keys + exclusive scan of the head flags -> deduplicated CSR (rowptr over users, ascending item ids)
__global__ void unique_csr_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos, int64_t n,
                                  int64_t n_users, int64_t n_items, int32_t* __restrict__ rowptr, int32_t* __restrict__ idx,
                                  int64_t* __restrict__ n_unique) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e > n) return;
  if (e == n) {   // tail: rows after the last user that appears
    // pos = number of heads BEFORE an entry: the last entry adds one if it is a head itself
    const int64_t total = n ? (int64_t)pos[n - 1] + ((n == 1 || keys[n - 1] != keys[n - 2]) ? 1 : 0) : 0;
    const int64_t u_last = n ? (int64_t)(keys[n - 1] / (uint64_t)n_items) : -1;
    for (int64_t t = u_last + 1; t <= n_users; ++t) rowptr[t] = (int32_t)total;
    *n_unique = total;
    return;
  }
  const bool head = e == 0 || keys[e] != keys[e - 1];
  if (!head) return;
  const uint32_t p = pos[e];
  const uint64_t k = keys[e];
  const int64_t u = (int64_t)(k / (uint64_t)n_items);
  idx[p] = (int32_t)(k - (uint64_t)u * (uint64_t)n_items);
  const int64_t u_prev = e == 0 ? -1 : (int64_t)(keys[e - 1] / (uint64_t)n_items);
  for (int64_t t = u_prev + 1; t <= u; ++t) rowptr[t] = (int32_t)p;
}

}  // namespace lgc

using namespace lgc;

extern "C" int64_t lgc_csr_max_chunks(int64_t nnz) {
  // a long row has > LGC_LONG_ROW non-zeros, so there are < nnz/LGC_LONG_ROW of them and
  // each contributes at most one partially filled chunk
  return nnz / LGC_CHUNK + nnz / LGC_LONG_ROW + 1;
}

extern "C" int lgc_csr_build_workspace_bytes(int64_t nnz, int64_t n_nodes, size_t* bytes_host) {
  LGC_REQUIRE(bytes_host && nnz >= 0 && n_nodes > 0, "csr workspace: bad arguments");
  LGC_REQUIRE(nnz < (1ll << 31) - 2 && n_nodes < (1ll << 31) - 2, "csr: nnz / n_nodes exceed int32");
  CsrWs w;
  csr_ws_layout(nnz > 0 ? nnz : 1, n_nodes, &w);
  *bytes_host = w.total;
  return LGC_OK;
}

extern "C" int lgc_csr_build(const int64_t* src, const int64_t* dst, int64_t nnz, int64_t n_nodes,
                             int32_t* rowptr, int32_t* colidx, float* val, float* dinv,
                             int32_t* chunk_row, int32_t* chunk_start, int32_t* row_chunk_base,
                             int32_t* n_chunks_host, void* workspace, size_t workspace_bytes,
                             lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(rowptr && dinv && row_chunk_base && n_chunks_host && workspace, "csr build: null pointer");
  LGC_REQUIRE(nnz >= 0 && n_nodes > 0, "csr build: bad sizes");
  LGC_REQUIRE(nnz < (1ll << 31) - 2 && n_nodes < (1ll << 31) - 2, "csr: nnz / n_nodes exceed int32");
  CsrWs w;
  csr_ws_layout(nnz > 0 ? nnz : 1, n_nodes, &w);
  if (workspace_bytes < w.total)
    LGC_FAIL(LGC_ERR_WORKSPACE, "csr build: workspace %zu < %zu", workspace_bytes, w.total);
  char* ws = (char*)workspace;
  uint64_t* keys_in = (uint64_t*)(ws + w.keys_in);
  uint64_t* keys_out = (uint64_t*)(ws + w.keys_out);
  int32_t* n_row_chunks = (int32_t*)(ws + w.n_row_chunks);
  int* bad = (int*)(ws + w.bad);
  uint32_t* table = (uint32_t*)(ws + w.table);
  const int T = 256;

  LGC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
  if (nnz > 0) {
    LGC_REQUIRE(src && dst && colidx && val, "csr build: null edge arrays");
    make_keys_kernel<<<(unsigned)ceil_div(nnz, T), T, 0, stream>>>(src, dst, nnz, n_nodes, keys_in, bad);
    LGC_LAUNCH_CHECK("make_keys");
    // number of significant key bits: targets < n_nodes
    int hi_bits = 1;
    while ((1ll << hi_bits) < n_nodes) ++hi_bits;
    int launches = 0;
    uint64_t* sorted = ingest::radix_sort_u64(keys_in, keys_out, nnz, 32 + hi_bits, table, stream, &launches);
    if (!sorted) LGC_FAIL(LGC_ERR_CUDA, "csr build: radix sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    note_launch(launches);
    if (sorted != keys_out) {   // odd number of passes: the result sits in keys_in
      uint64_t* t = keys_in; keys_in = keys_out; keys_out = t;
    }
  }
  rowptr_kernel<<<(unsigned)ceil_div(nnz + 1, T), T, 0, stream>>>(keys_out, nnz, n_nodes, rowptr, colidx);
  LGC_LAUNCH_CHECK("rowptr");
  dinv_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(rowptr, n_nodes, dinv, n_row_chunks);
  LGC_LAUNCH_CHECK("dinv");
  if (nnz > 0) {
    edge_val_kernel<<<(unsigned)ceil_div(nnz, T), T, 0, stream>>>(keys_out, nnz, dinv, val);
    LGC_LAUNCH_CHECK("edge_val");
  }
  // chunk list: exclusive scan of per-row chunk counts (n_nodes+1 entries, last = total)
  LGC_CUDA(cudaMemsetAsync(n_row_chunks + n_nodes, 0, sizeof(int32_t), stream));
  {
    const int sl = ingest::exclusive_scan_u32((const uint32_t*)n_row_chunks, (uint32_t*)row_chunk_base, n_nodes + 1, table, stream);
    if (sl < 0) LGC_FAIL(LGC_ERR_CUDA, "csr build: scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    note_launch(sl);
  }
  if (chunk_row && chunk_start) {
    chunk_fill_kernel<<<(unsigned)ceil_div(n_nodes, T), T, 0, stream>>>(rowptr, row_chunk_base, n_nodes,
                                                                         chunk_row, chunk_start);
    LGC_LAUNCH_CHECK("chunk_fill");
  }
  int h_bad = 0;
  LGC_CUDA(cudaMemcpyAsync(n_chunks_host, row_chunk_base + n_nodes, sizeof(int32_t),
                           cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) LGC_FAIL(LGC_ERR_INVALID, "csr build: edge index out of range [0, %lld)", (long long)n_nodes);
  return LGC_OK;
}

// ---- (N2) deduplicated per-user CSR of (user, item) pairs: the mask / positive lists of both hot paths ----
extern "C" int lgc_seen_csr_workspace_bytes(int64_t n_pairs, size_t* bytes_host) {
  LGC_REQUIRE(bytes_host && n_pairs >= 0, "seen csr workspace: bad arguments");
  const int64_t n = n_pairs > 0 ? n_pairs : 1;
  size_t words = ingest::sort_table_words(n);
  const size_t sw = ingest::scan_scratch_words(n);
  if (sw > words) words = sw;
  *bytes_host = 2 * align_up(sizeof(uint64_t) * (size_t)n, 256) + align_up(sizeof(uint32_t) * (size_t)n, 256) +
                align_up(words * sizeof(uint32_t), 256) + 512;
  return LGC_OK;
}

extern "C" int lgc_seen_csr(const int64_t* users, const int64_t* items, int64_t n_pairs, int64_t n_users, int64_t n_items,
                            int32_t* rowptr, int32_t* idx, int64_t* n_unique_host, void* workspace, size_t workspace_bytes,
                            lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(rowptr && n_unique_host && workspace && n_users > 0 && n_items > 0 && n_pairs >= 0, "seen csr: bad arguments");
  LGC_REQUIRE(n_pairs < (1ll << 31) - 2 && n_users < (1ll << 31) - 2 && n_items < (1ll << 31) - 2, "seen csr: extents exceed int32");
  size_t need = 0;
  lgc_seen_csr_workspace_bytes(n_pairs, &need);
  if (workspace_bytes < need) LGC_FAIL(LGC_ERR_WORKSPACE, "seen csr: workspace %zu < %zu", workspace_bytes, need);
  const int64_t n = n_pairs;
  char* ws = (char*)workspace;
  const size_t kb = align_up(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1), 256);
  uint64_t* ka = (uint64_t*)ws;
  uint64_t* kbuf = (uint64_t*)(ws + kb);
  uint32_t* flags = (uint32_t*)(ws + 2 * kb);
  const size_t fb = align_up(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1), 256);
  uint32_t* table = (uint32_t*)(ws + 2 * kb + fb);
  size_t words = ingest::sort_table_words(n > 0 ? n : 1);
  const size_t sw = ingest::scan_scratch_words(n > 0 ? n : 1);
  if (sw > words) words = sw;
  int* bad = (int*)(ws + 2 * kb + fb + align_up(words * sizeof(uint32_t), 256));
  int64_t* n_unique_dev = (int64_t*)(bad + 2);
  const int T = 256;
  LGC_CUDA(cudaMemsetAsync(bad, 0, 64, stream));
  const uint64_t* sorted = ka;
  if (n > 0) {
    LGC_REQUIRE(users && items && idx, "seen csr: null pair arrays");
    pair_keys_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(users, items, n, n_users, n_items, ka, bad);
    LGC_LAUNCH_CHECK("pair_keys");
    int bits = 1;
    while (bits < 63 && (1ull << bits) < (uint64_t)n_users * (uint64_t)n_items) ++bits;
    int launches = 0;
    uint64_t* res = ingest::radix_sort_u64(ka, kbuf, n, bits, table, stream, &launches);
    if (!res) LGC_FAIL(LGC_ERR_CUDA, "seen csr: radix sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    note_launch(launches);
    sorted = res;
    head_flags_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(sorted, n, flags);
    LGC_LAUNCH_CHECK("head_flags");
    const int sl = ingest::exclusive_scan_u32(flags, flags, n, table, stream);
    if (sl < 0) LGC_FAIL(LGC_ERR_CUDA, "seen csr: scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    note_launch(sl);
  }
  unique_csr_kernel<<<(unsigned)ceil_div(n + 1, T), T, 0, stream>>>(sorted, flags, n, n_users, n_items, rowptr, idx, n_unique_dev);
  LGC_LAUNCH_CHECK("unique_csr");
  int h_bad = 0;
  LGC_CUDA(cudaMemcpyAsync(n_unique_host, n_unique_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaStreamSynchronize(stream));
  if (h_bad) LGC_FAIL(LGC_ERR_INVALID, "seen csr: pair out of range [0, %lld) x [0, %lld)", (long long)n_users, (long long)n_items);
  return LGC_OK;
}

// stable ascending sort of the low `bits` bits of 64-bit keys, in place (tmp: n keys of scratch)
extern "C" int lgc_sort_u64_workspace_bytes(int64_t n, size_t* bytes_host) {
  LGC_REQUIRE(bytes_host && n >= 0, "sort workspace: bad arguments");
  *bytes_host = align_up(ingest::sort_table_words(n > 0 ? n : 1) * sizeof(uint32_t), 256);
  return LGC_OK;
}

extern "C" int lgc_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int32_t bits, void* workspace, size_t workspace_bytes,
                            lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(n >= 0 && bits >= 1 && bits <= 64, "sort: bad arguments");
  if (n == 0) return LGC_OK;
  LGC_REQUIRE(keys && tmp && workspace, "sort: null pointer");
  size_t need = 0;
  lgc_sort_u64_workspace_bytes(n, &need);
  if (workspace_bytes < need) LGC_FAIL(LGC_ERR_WORKSPACE, "sort: workspace %zu < %zu", workspace_bytes, need);
  int launches = 0;
  uint64_t* res = ingest::radix_sort_u64(keys, tmp, n, bits, (uint32_t*)workspace, stream, &launches);
  if (!res) LGC_FAIL(LGC_ERR_CUDA, "sort: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  note_launch(launches);
  if (res != keys) LGC_CUDA(cudaMemcpyAsync(keys, res, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToDevice, stream));
  return LGC_OK;
}

// sorted, deduplicated copy of 64-bit keys (torch.unique of the format converters, utils/graph.py:12-50): keys is
// clobbered, out receives the distinct keys ascending, *n_unique_host their number (one D2H sync)
extern "C" int lgc_unique_u64_workspace_bytes(int64_t n, size_t* bytes_host) {
  LGC_REQUIRE(bytes_host && n >= 0, "unique workspace: bad arguments");
  const int64_t m = n > 0 ? n : 1;
  size_t words = ingest::sort_table_words(m);
  const size_t sw = ingest::scan_scratch_words(m);
  if (sw > words) words = sw;
  *bytes_host = align_up(sizeof(uint64_t) * (size_t)m, 256) + align_up(sizeof(uint32_t) * (size_t)m, 256) +
                align_up(words * sizeof(uint32_t), 256) + 256;
  return LGC_OK;
}

extern "C" int lgc_unique_u64(uint64_t* keys, uint64_t* out, int64_t n, int32_t bits, int64_t* n_unique_host, void* workspace,
                              size_t workspace_bytes, lgc_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LGC_REQUIRE(n_unique_host && n >= 0 && bits >= 1 && bits <= 64, "unique: bad arguments");
  if (n == 0) { *n_unique_host = 0; return LGC_OK; }
  LGC_REQUIRE(keys && out && workspace, "unique: null pointer");
  size_t need = 0;
  lgc_unique_u64_workspace_bytes(n, &need);
  if (workspace_bytes < need) LGC_FAIL(LGC_ERR_WORKSPACE, "unique: workspace %zu < %zu", workspace_bytes, need);
  char* ws = (char*)workspace;
  const size_t kb = align_up(sizeof(uint64_t) * (size_t)n, 256);
  const size_t fb = align_up(sizeof(uint32_t) * (size_t)n, 256);
  uint64_t* tmp = (uint64_t*)ws;
  uint32_t* flags = (uint32_t*)(ws + kb);
  uint32_t* table = (uint32_t*)(ws + kb + fb);
  size_t words = ingest::sort_table_words(n);
  const size_t sw = ingest::scan_scratch_words(n);
  if (sw > words) words = sw;
  int64_t* n_unique_dev = (int64_t*)(ws + kb + fb + align_up(words * sizeof(uint32_t), 256));
  int launches = 0;
  uint64_t* sorted = ingest::radix_sort_u64(keys, tmp, n, bits, table, stream, &launches);
  if (!sorted) LGC_FAIL(LGC_ERR_CUDA, "unique: radix sort launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  note_launch(launches);
  const int T = 256;
  head_flags_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(sorted, n, flags);
  LGC_LAUNCH_CHECK("head_flags");
  const int sl = ingest::exclusive_scan_u32(flags, flags, n, table, stream);
  if (sl < 0) LGC_FAIL(LGC_ERR_CUDA, "unique: scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  note_launch(sl);
  compact_unique_kernel<<<(unsigned)ceil_div(n, T), T, 0, stream>>>(sorted, flags, n, out, n_unique_dev);
  LGC_LAUNCH_CHECK("compact_unique");
  LGC_CUDA(cudaMemcpyAsync(n_unique_host, n_unique_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  LGC_CUDA(cudaStreamSynchronize(stream));
  return LGC_OK;
}
