/*
 * lgcnhs.h — C ABI of the B200-native LGCNHS hot path (liblgcnhs.so).
 *
 * The reference (Alex-McAvoy/LGCNHS) is pure Python and has no FFI of its own; the
 * boundary it exposes is the Python module surface model/LightGCN, model/SpreadMethod,
 * model/SpreadLightGCN(Opti).  Every entry point below replaces one or
 * more *library calls* the reference makes from those modules; the reference file:line
 * each one stands in for is cited next to it.  The Python host code under
 * light-graph-..._b200/{lgcnhs_b200,model,utils,metrics} binds these with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless
 *     the parameter name ends in _host;
 *   - the library never allocates or frees caller-visible memory: callers pass
 *     workspaces sized by the *_workspace_bytes queries;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*),
 *     returns 0 on success or a negative lgc_status, never throws;
 *   - lgc_last_error_string() gives the thread-local message of the last failure.
 */
#ifndef LGCNHS_H_
#define LGCNHS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lgc_stream_t; /* cudaStream_t */

enum lgc_status {
  LGC_OK = 0,
  LGC_ERR_INVALID = -1,   /* bad argument (shape, alignment, null) */
  LGC_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed    */
  LGC_ERR_WORKSPACE = -3, /* workspace too small                     */
  LGC_ERR_UNSUPPORTED = -4
};

int lgc_abi_version(void);
const char* lgc_last_error_string(void);
/* Number of kernels this library has launched since load / since the last reset
 * (bench.py's "gpu_launches"). */
int64_t lgc_launch_count(void);
void lgc_reset_launch_count(void);

/* ------------------------------------------------------------------------------------
 * (P1) Graph normalisation, once per graph.
 * Replaces torch_geometric gcn_norm(edge_index, add_self_loops=False) called at
 * model/LightGCN/model.py:53 (and LightGCNOpti/model.py:65) on EVERY forward, plus the
 * COO layout produced by utils/graph.py:12-35.
 *   src = edge_index[0] (message source,  x_j),  dst = edge_index[1] (aggregation target)
 *   deg[c]  = #{e : dst[e]==c}          dinv = deg^-1/2 (0 where deg==0)
 *   val[e]  = dinv[src[e]] * dinv[dst[e]]       (two fp32 factors multiplied, as PyG does)
 * Output is a CSR keyed by target row: rowptr[n_nodes+1], colidx[nnz] = sources in
 * ascending order (the order the reference's CPU scatter_add visits them), val[nnz].
 * Rows longer than LGC_LONG_ROW are additionally listed as fixed-size chunks
 * (chunk_row/chunk_start, at most lgc_csr_max_chunks(nnz) of them) so that the SpMM can
 * split them over several CTAs; *n_chunks_host receives the count (one D2H sync).
 * ---------------------------------------------------------------------------------- */
#define LGC_LONG_ROW 256  /* rows with more non-zeros than this go to the chunk path */
#define LGC_CHUNK 1024     /* non-zeros per long-row chunk (one CTA)                  */

int64_t lgc_csr_max_chunks(int64_t nnz);
int lgc_csr_build_workspace_bytes(int64_t nnz, int64_t n_nodes, size_t* bytes_host);
int lgc_csr_build(const int64_t* src, const int64_t* dst, int64_t nnz, int64_t n_nodes,
                  int32_t* rowptr, int32_t* colidx, float* val, float* dinv,
                  int32_t* chunk_row, int32_t* chunk_start, int32_t* row_chunk_base,
                  int32_t* n_chunks_host, void* workspace, size_t workspace_bytes,
                  lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (P2/P3) One propagation layer, fused with the running layer sum.
 * Replaces MessagePassing.propagate + message (model/LightGCN/model.py:61-63, 76-84:
 * index_select -> norm*x_j -> scatter_add) and one term of the stack/mean at :66-69.
 *   Y[r,:] = alpha * ( sum_{e in row r} val[e] * X[colidx[e],:]  +  beta * X0[r,:] )
 * X, X0, Y are (n_nodes, dim) fp32 row-major, dim in {32, 64, 128}.  X0 may be null when
 * beta == 0.  Deterministic (fixed summation order, no float atomics).
 * `partial` is scratch of lgc_csr_max_chunks(nnz)*dim floats, `counters` is n_nodes
 * int32 that must be zero on entry (the kernel leaves them zero).
 * row_begin/row_end restrict the launch to a row range (multi-GPU row partition);
 * row_order (may be null) lists the rows of that range in processing order, row_end - row_begin
 * entries — longest first keeps the eight rows of a CTA alike and the last wave short;
 * long_row (needs row_order; clamped to [LGC_LONG_ROW, 2048]): rows with up to this many
 * non-zeros take the warp-per-row path — about 1e-4 x the launch's non-zeros is right on B200.
 * A row's fp32 summation order depends on the path, so results are bit-reproducible for a given
 * (row range, long_row) and may differ in the last bits between configurations;
 * chunk_begin/chunk_end = row_chunk_base[row_begin], row_chunk_base[row_end] are the
 * long-row chunks of that range ([0, n_chunks) for the whole graph).
 * ---------------------------------------------------------------------------------- */
int lgc_spmm_layer(const int32_t* rowptr, const int32_t* colidx, const float* val,
                   const int32_t* chunk_row, const int32_t* chunk_start,
                   const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                   int64_t n_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                   const int32_t* row_order, int32_t long_row, const float* X, const float* X0,
                   float alpha, float beta, float* Y, float* partial, int32_t* counters,
                   lgc_stream_t stream);

/* Tuning knob: independent 128-bit gathers in flight per lane for dim 64 (2, 4 or 8). */
int lgc_spmm_config(int32_t unroll);
/* ------------------------------------------------------------------------------------
 * (N2) Format ingestion in front of both hot paths, as own kernels (stable LSD radix sort of 64-bit
 * keys + exclusive scan, csrc/ingest.cuh).
 * lgc_seen_csr: (user, item) pairs -> DEDUPLICATED int32 CSR over all users (rowptr[n_users+1],
 *   item ids ascending per row; idx needs room for n_pairs entries; *n_unique_host = entries
 *   written).  It is the per-user item list the reference builds with Python loops:
 *   getUserItemsDictByDataframe / getUserItemsDictByEdgeIndex (utils/trans.py:51-80), the exclusion
 *   pairs of model/LightGCN/recommend.py:92-111 and the positive lists np.isin searches in
 *   structured_negative_sampling (model/LightGCN/loss.py:58).  One D2H sync (the count).
 * lgc_sort_u64: stable ascending sort of the low `bits` bits of 64-bit keys, in place (tmp: n keys).
 * ---------------------------------------------------------------------------------- */
int lgc_seen_csr_workspace_bytes(int64_t n_pairs, size_t* bytes_host);
int lgc_seen_csr(const int64_t* users, const int64_t* items, int64_t n_pairs, int64_t n_users,
                 int64_t n_items, int32_t* rowptr, int32_t* idx, int64_t* n_unique_host,
                 void* workspace, size_t workspace_bytes, lgc_stream_t stream);
/* lgc_unique_u64: the distinct keys, ascending (torch.unique of the format converters utils/graph.py:12-50, which the
 * reference obtains through a dense (U+M)^2 matrix); keys is clobbered, out needs room for n keys. */
int lgc_unique_u64_workspace_bytes(int64_t n, size_t* bytes_host);
int lgc_unique_u64(uint64_t* keys, uint64_t* out, int64_t n, int32_t bits, int64_t* n_unique_host,
                   void* workspace, size_t workspace_bytes, lgc_stream_t stream);
int lgc_sort_u64_workspace_bytes(int64_t n, size_t* bytes_host);
int lgc_sort_u64(uint64_t* keys, uint64_t* tmp, int64_t n, int32_t bits, void* workspace,
                 size_t workspace_bytes, lgc_stream_t stream);

/* Tuning knob of lgc_propagate_mean_coop: resident CTAs per SM (1..8, default 2); fewer CTAs = cheaper grid barriers. */
int lgc_coop_config(int32_t ctas_per_sm);
/* Tuning knob: override of the per-call long_row for the launches that follow (in [LGC_LONG_ROW, 2048];
 * 0 = use the per-call value).  The chunk lists always cover rows > LGC_LONG_ROW. */
int lgc_spmm_long_row(int32_t long_row);

/* K-layer forward with the uniform layer mean in Horner form
 *   S_0 = X0,  S_{l+1} = A_hat S_l + X0,  E = S_K / (K+1)
 * == mean(stack([X0, A X0, ..., A^K X0])) of model/LightGCN/model.py:56-69.
 * tmp0/tmp1 are (n_nodes, dim) scratch; E may alias neither. */
int lgc_propagate_mean(const int32_t* rowptr, const int32_t* colidx, const float* val,
                       const int32_t* chunk_row, const int32_t* chunk_start,
                       const int32_t* row_chunk_base, int32_t n_chunks,
                       int64_t n_nodes, int32_t dim, int32_t n_layers,
                       const int32_t* row_order /* n_nodes entries or null */, int32_t long_row,
                       const float* X0, float* E, float* tmp0, float* tmp1,
                       float* partial, int32_t* counters, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (P6/P4) BPR loss, forward (+ optional backward with embedding-gradient scatter).
 * Replaces BPRLoss (model/LightGCN/loss.py:12-44) and the six row gathers of
 * model/LightGCN/train.py:55-57.  Reference quirks kept: loss = -mean(softplus(s+ - s-))
 * + eps * sum_b(|u0|^2 + |p0|^2 + |n0|^2)  (softplus threshold 20 as torch).
 *   E, X0        (n_users+n_items, dim) fp32, users first
 *   users/pos/neg  int64[batch]; pos/neg are ITEM ids (0-based, offset by n_users inside)
 *   loss_out     float[2]: {total loss, bpr term}
 *   gE, gX0      optional (n_nodes, dim) fp32, ACCUMULATED into (caller zero-fills)
 *   scratch      float[lgc_bpr_scratch_floats(batch)], scratch[0] zero on first use
 *                (the kernel leaves it zero)
 * ---------------------------------------------------------------------------------- */
int64_t lgc_bpr_scratch_floats(int64_t batch);
int lgc_bpr_fwd_bwd(const float* E, const float* X0, int64_t n_users, int64_t n_items,
                    int32_t dim, const int64_t* users, const int64_t* pos,
                    const int64_t* neg, int64_t batch, float eps, float grad_scale,
                    float* loss_out, float* gE, float* gX0, float* scratch,
                    lgc_stream_t stream);

/* Same loss on six PRE-GATHERED (batch, dim) row blocks — the literal signature of
 * BPRLoss(users_emb_final, users_emb_0, pos_final, pos_0, neg_final, neg_0, lambda)
 * (model/LightGCN/loss.py:12).  Gradients (optional, all six or none) are stored, not
 * accumulated. */
int lgc_bpr_rows(const float* uf, const float* u0, const float* pf, const float* p0,
                 const float* nf, const float* n0, int64_t batch, int32_t dim, float eps,
                 float grad_scale, float* loss_out, float* guf, float* gu0, float* gpf,
                 float* gp0, float* gnf, float* gn0, float* scratch, lgc_stream_t stream);

/* (P7) Adam step, torch.optim.Adam(lr, betas=(b1,b2), eps) semantics
 * (model/LightGCN/train.py:104,144): bias corrections are passed pre-computed
 * (bc1 = 1-b1^t, bc2_sqrt = sqrt(1-b2^t)). */
int lgc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  int64_t n, float lr, float beta1, float beta2, float eps, float bc1,
                  float bc2_sqrt, lgc_stream_t stream);

/* Device-resident variant for CUDA-graph capture of a whole training step: lgc_adam_hyper_step
 * increments *step_dev and writes hyper_dev = {lr / (1 - beta1^step), sqrt(1 - beta2^step)};
 * lgc_adam_step_dev reads those two scalars instead of host arguments. */
int lgc_adam_hyper_step(int64_t* step_dev, const float* lr_dev, float beta1, float beta2,
                        float* hyper_dev, lgc_stream_t stream);
int lgc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float beta1, float beta2, float eps, const float* hyper_dev,
                      lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (P8) score = Xu . Xi^T  (+ optional seen-pair fill), materialised for a block of
 * users.  Replaces torch.matmul(user_embedding, item_embedding.T) and
 * score[users, items] = -1024 (model/LightGCN/recommend.py:86,101,111;
 * evaluation.py:34,49; SpreadLightGCN/model.py:77,92,102).
 *   out[(u-u0), i] for u in [u0,u1), ld = ldo floats.
 *   seen_ptr/seen_idx: CSR over ALL users of seen item ids (may be null -> no fill).
 * ---------------------------------------------------------------------------------- */
int lgc_score_block(const float* Xu, const float* Xi, int64_t u0, int64_t u1,
                    int64_t n_items, int32_t dim, const int32_t* seen_ptr,
                    const int32_t* seen_idx, float fill, float* out, int64_t ldo,
                    lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (P8/S4) Row-wise masked top-k.  Replaces torch.topk(score, k)
 * (model/LightGCN/recommend.py:114) and the argsort + Python filter loop of
 * model/SpreadMethod/recommend.py:35-47 (== SpreadLightGCN/recommend.py:34-46).
 *   S: (n_rows, n_cols) fp32, ld = lds.  excl_mask (may be null) is a bit-packed matrix, bit
 *   (row_offset + r) * mask_stride_bits + c set = column c can never be selected for row r —
 *   the deduplicated interaction bitmap hs_degrees() produces, or lgc_mask_from_csr().
 *   Result sorted by value descending, ties -> larger index first (what the CPU reference
 *   produces for np.argsort(row)[::-1]).
 *   out_idx int64 (n_rows, k), out_val fp32 (n_rows, k) or null.  k <= 128, k <= n_cols.
 * ---------------------------------------------------------------------------------- */
int lgc_topk_rows(const float* S, int64_t n_rows, int64_t n_cols, int64_t lds,
                  const uint32_t* excl_mask, int64_t mask_stride_bits, int64_t row_offset,
                  int32_t k, int64_t* out_idx, float* out_val, lgc_stream_t stream);
/* ------------------------------------------------------------------------------------
 * (P8 / F1) FUSED score + seen-pair rule + top-k: the (U, M) score matrix is never written.
 * One call replaces torch.matmul(user_embedding, item_embedding.T), score[users, items] = -1024
 * and torch.topk(score, k) of model/LightGCN/recommend.py:86-114 (== evaluation.py:34-52), and,
 * with `mul`, the G_score * F Hadamard product + argsort/filter loop of
 * model/SpreadLightGCN/model.py:151 + recommend.py:34-46 (== SpreadLightGCNOpti).
 *   value(u, i) = (seen(u,i) ? fill : <Xu[u], Xi[i]>) * (mul ? mul[(u-u0)*ldmul + i] : 1)
 *   seen_ptr/seen_idx: CSR over ALL users of seen item ids, ascending per row (may be null).
 *   exclude_seen != 0: seen items are dropped from the ranking instead of taking `fill`.
 *   out_idx int64 (u1-u0, k), out_val fp32 (u1-u0, k) or null; sorted by value descending,
 *   ties -> larger index first.  k <= 128, k <= n_items, dim in {32, 64}.
 * ---------------------------------------------------------------------------------- */
int lgc_score_topk(const float* Xu, const float* Xi, int64_t u0, int64_t u1, int64_t n_items,
                   int32_t dim, const int32_t* seen_ptr, const int32_t* seen_idx, float fill,
                   int32_t exclude_seen, const float* mul, int64_t ldmul, int32_t k,
                   int64_t* out_idx, float* out_val, lgc_stream_t stream);
/* Same contract on the TENSOR CORES (tcgen05 kind::tf32, 3xTF32 split of both operands: hi*lo + lo*hi + hi*hi, fp32
 * accumulate in TMEM): ~10x the throughput of the fp32-FMA kernel above.  Scores agree with an fp32 SGEMM to ~1e-6
 * relative (inside the 1e-5 tolerance); ids are identical except at float near-ties.  Xu is the FULL (n_users_total,
 * dim) table, users [u0, u1) are ranked; mul (may be null) is (u1-u0, ldmul).  dim in {32, 64}, k <= 32.
 * workspace: lgc_score_topk_tc_workspace_bytes(), 256-byte aligned (hi/lo planes of both tables + candidate
 * buffers).  Replaces the same reference lines as lgc_score_topk. */
int64_t lgc_score_topk_tc_workspace_bytes(int64_t n_users_total, int64_t n_items, int64_t rows, int32_t dim);
int lgc_score_topk_tc(const float* Xu, const float* Xi, int64_t n_users_total, int64_t u0, int64_t u1,
                      int64_t n_items, int32_t dim, const int32_t* seen_ptr, const int32_t* seen_idx,
                      float fill, int32_t exclude_seen, const float* mul, int64_t ldmul, int32_t k,
                      int64_t* out_idx, float* out_val, void* workspace, int64_t workspace_bytes,
                      lgc_stream_t stream);
/* Device-side barrier over peer memory (replaces one NCCL all-reduce per propagation layer in the multi-GPU
 * p2p mode).  Every rank owns local_flags[n_peers] in IPC-mapped memory, zero-initialised; the call publishes
 * `epoch` into peer_flags_host[p][my_rank] for every p and returns (stream-ordered) once local_flags[q] >= epoch
 * for all q.  Epochs must grow monotonically.  Launch it right after the layer's lgc_spmm_layer_bcast. */
int lgc_peer_barrier(const int32_t* local_flags, int32_t* const* peer_flags_host, int32_t my_rank,
                     int32_t n_peers, int32_t epoch, lgc_stream_t stream);

/* Small graphs: all K layers + the layer mean of lgc_propagate_mean in ONE cooperative launch (layers separated by grid
 * barriers instead of kernel launches; model/LightGCN/model.py:56-69 at the ML-100K / Douban shapes, where a layer is
 * launch-latency bound).  Work units of <= 128 non-zeros: unit u covers non-zeros [unit_start[u], unit_end[u]) of row
 * unit_row[u]; unit_slot[u] = -1 for a whole row, else the partial-sum slot of a piece of a long row; split_* list the
 * long rows with their first slot and number of pieces (combined in order: deterministic).  partial: n_partials * dim
 * floats; barrier_state: 2 uint32 zeroed once by the caller (grid-barrier counters, left at zero by every launch).
 * Same result as lgc_propagate_mean up to fp32 summation order. */
int lgc_propagate_mean_coop(const int32_t* rowptr, const int32_t* colidx, const float* val,
                            const int32_t* unit_row, const int32_t* unit_start, const int32_t* unit_end,
                            const int32_t* unit_slot, int32_t n_units, const int32_t* split_row,
                            const int32_t* split_first, const int32_t* split_count, int32_t n_split,
                            int64_t n_nodes, int32_t dim, int32_t n_layers, const float* X0, float* E,
                            float* tmp0, float* tmp1, float* partial, uint32_t* barrier_state,
                            lgc_stream_t stream);

/* Same barrier with the epoch counter in device memory (the kernel increments *epoch_counter_dev and uses the new
 * value): capturable in a CUDA graph, replayable — all ranks must issue the same sequence of barriers. */
int lgc_peer_barrier_dev(const int32_t* local_flags, int32_t* const* peer_flags_host, int32_t my_rank,
                         int32_t n_peers, int32_t* epoch_counter_dev, lgc_stream_t stream);

/* Tuning knob of lgc_score_topk: CTA size (256 or 512 threads) of the 128-user tile variant. */
int lgc_score_topk_config(int32_t threads);

/* ------------------------------------------------------------------------------------
 * (N3) Structured negative sampling for BPR.  Replaces
 * torch_geometric.utils.structured_negative_sampling as used by sampleMiniBatch
 * (model/LightGCN/loss.py:46-70) and calValLoss (evaluation.py:72): for every requested edge
 * (u, pos) draw neg uniformly from [0, num_nodes) until (u, neg) is not a positive pair (and
 * neg != u when forbid_self).  rows (may be null = all edges) selects the edges; pos_ptr/pos_idx
 * is the CSR of positive items per user (ascending).  *status (device int, zeroed by the caller)
 * becomes non-zero on an out-of-range row / user id.  Counter-based generator: same seed, same
 * triplets.
 * ---------------------------------------------------------------------------------- */
int lgc_negative_sample(const int64_t* edge_u, const int64_t* edge_p, int64_t n_edges,
                        const int64_t* rows, int64_t n_out, const int32_t* pos_ptr,
                        const int32_t* pos_idx, int64_t n_users, int64_t num_nodes,
                        int32_t forbid_self, uint64_t seed, int64_t* out_u, int64_t* out_p,
                        int64_t* out_n, int32_t* status, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (N1) The six evaluation metrics of a set of top-k lists, on the device.  Replaces
 * metrics/accurate.py:11-102 (Precision / Recall / NDCG: per-user `item in items` loops) and the
 * O(U^2) / O(U k^2 U) Python loops of metrics/diversity.py:15-115 (Hamming distance, intra-list
 * similarity) by their closed forms:
 *   out6[0] = sum_u hits_u            out6[1] = sum_u hits_u / |pos_u|      out6[2] = sum_u dcg_u / idcg
 *   out6[3] = users with >= 1 positive item (the keys of user_pos_items_dict)
 *   out6[4] = sum_i c_i (c_i - 1),  c_i = lists containing item i   (H = 1 - out6[4] / (U (U-1) k))
 *   out6[5] = sum_u sum_{a != b in L_u} C[a,b] / sqrt(k_a k_b)       (I = out6[5] / (U k (k-1)))
 *   rec int64 (n_users, k) row-major; pos_ptr/pos_idx: CSR of relevant items, ascending per row
 *   (may be null: [0..3] stay 0); Cmat fp32 (n_items, ldc) = A^T A and item_deg int32[n_items]
 *   (may both be null: [5] stays 0).  scratch: lgc_metrics_scratch_bytes(n_items), 16-byte aligned.
 * ---------------------------------------------------------------------------------- */
int64_t lgc_metrics_scratch_bytes(int64_t n_items);
int lgc_metrics_topk(const int64_t* rec, int64_t n_users, int32_t k, int64_t n_items,
                     const int32_t* pos_ptr, const int32_t* pos_idx, const float* Cmat, int64_t ldc,
                     const int32_t* item_deg, double* out6, void* scratch, lgc_stream_t stream);

/* CSR (rowptr, column ids) -> bit-packed mask with the given row stride; mask zero-filled by
 * the caller, (n_rows * stride_bits + 31) / 32 words. */
int lgc_mask_from_csr(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int64_t n_cols,
                      int64_t stride_bits, uint32_t* mask, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (S0/S1) Hybrid spreading, operand packing.
 * Replaces the dense float64 A of utils/trans.py:13-29 as GEMM operand.
 * Input is the interaction list (user, item) int32 of length nnz, duplicates allowed.
 *   ku[u], ki[i] (float, optional outputs) = degrees of the *deduplicated* matrix.
 * hs_pack_a:  A  (n_users x ldk) bf16 0/1, K-major for F = A.W   (ldk >= n_items, %64==0)
 * hs_pack_at: At (n_items x ldk) uint8 0/1 and `digits` planes Q_d (n_items x ldk) uint8
 *             with Q_d[i,u] = digit_d(q_u) * A[u,i],  q_u = round(2^shift / k_u),
 *             q_u = sum_d digit_d * 256^d   (exact fixed-point 1/k_u; ldk >= n_users, %128==0)
 * All outputs must be zero-filled by the caller before the call.
 * ---------------------------------------------------------------------------------- */
int hs_degrees(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
               int64_t n_items, int32_t* ku, int32_t* ki, uint8_t* dedup_bitmap,
               lgc_stream_t stream);
int hs_pack_a(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
              int64_t n_items, uint16_t* A_bf16, int64_t ldk, lgc_stream_t stream);
int hs_pack_at(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
               int64_t n_items, const int32_t* ku, int32_t shift, int32_t digits,
               uint8_t* At_u8, uint8_t* Q_u8, int64_t ldk, int64_t plane_stride,
               lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (S1/S3) tcgen05 / TMEM tensor-core GEMM on a binary left operand:
 *     C[m, n] = rs[m] * cs[n] * scale * sum_p w_p * sum_k A[m,k] * B_p[n,k]
 * A: (M x K) K-major, B_p: `planes` matrices (N x K) K-major, plane stride in elements.
 * kind 0: bf16 operands, fp32 accumulate, w_p = 1                (planes in {1,2,3})
 * kind 1: uint8 operands, exact int32 accumulate, w_p = 256^p    (planes in {1,2,3,4})
 * Tile: 128 rows x NB columns x all planes in one MMA (N = planes*NB: 128, 256, 240, 256).
 * Replaces np.dot(A.T / user_degrees, A) (model/SpreadMethod/model.py:25) and
 * np.dot(F0, W) (model/SpreadMethod/model.py:98).
 * lda/ldb in elements, multiples of 16 bytes; pointers 16-byte aligned (TMA).
 * rs / cs may be null (== 1).  C is fp32 row-major (ldc floats).
 * ---------------------------------------------------------------------------------- */
int hs_gemm_planes(int32_t kind, const void* A, int64_t lda, const void* B, int64_t ldb,
                   int64_t plane_stride, int32_t planes, int64_t M, int64_t N, int64_t K,
                   float* C, int64_t ldc, const float* rs, const float* cs, double scale,
                   lgc_stream_t stream);
/* Same contraction for a SYMMETRIC result (kind 1 only: exact integer accumulation makes
 * C == C^T bit for bit), as G = A^T K_u^-1 A is (model/SpreadMethod/model.py:25 with A := A^T,
 * B_p := digit planes of round(2^s/k_u) A^T).  Only the tiles that touch the upper triangle
 * are computed (about half the MMAs); tiles above the diagonal blocks also store their
 * transpose.  C: (N x N) fp32.  Bit-identical to hs_gemm_planes. */
int hs_gemm_planes_sym(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                       int32_t planes, int64_t N, int64_t K, float* C, int64_t ldc, double scale,
                       lgc_stream_t stream);
/* Multi-GPU form of hs_gemm_planes_sym: a fused GEMM + all-gather.  The cluster slots of the symmetric tile schedule
 * are dealt round-robin to the n_ranks ranks; every computed tile (and its mirror) is stored into the n_store targets
 * store_C_host[0..n_store): either every rank's (N x ldc) replica mapped through CUDA IPC (n_store = n_ranks), or ONE
 * NVSwitch multicast address bound to all replicas (n_store = 1) — each rank ends with the full matrix after a barrier.
 * SURVEY 8e "W build: shard ... then all-gather W for scoring". */
int hs_gemm_planes_sym_bcast(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                             int32_t planes, int64_t N, int64_t K, float* const* store_C_host, int32_t n_store,
                             int32_t my_rank, int32_t n_ranks, int64_t ldc, double scale, lgc_stream_t stream);
/* ------------------------------------------------------------------------------------
 * (S3+S4 fused) F = A . W and the per-user filtered top-k in ONE pass: the F tiles never
 * leave TMEM/registers.  Replaces np.dot(A, W) (model/SpreadMethod/model.py:98) followed by
 * the per-user np.argsort + Python `not in` filter loop (model/SpreadMethod/recommend.py:35-47).
 *   value(u, j) = cs[j] * scale * sum_p 256^p sum_i A[u,i] * B_p[j,i]      (kind 1, exact int32)
 *   excl_mask  : bit-packed (rows x N) matrix of pairs that must not be returned, bit
 *                (row_offset + u) * mask_stride_bits + j; may be null (unfiltered ranking)
 *   out_idx int64 (M, k), out_val fp32 (M, k) or null: sorted by value descending, ties ->
 *                larger column first (what np.argsort(row)[::-1] gives); (-1, -inf) padding.
 * Work item = (256-row block, segment of the column tiles); each epilogue thread keeps the
 * threshold / fill count of its row in registers across the item's tiles and appends the
 * few survivors to a private candidate buffer in `scratch`; a final warp-per-row kernel
 * merges the segments.  planes in {3,4}, k <= 32, M > 128.  Lists are identical to
 * hs_gemm_planes + lgc_topk_rows.  scratch: hs_resource_topk_scratch_bytes(), 256-B aligned.
 * ---------------------------------------------------------------------------------- */
int64_t hs_resource_topk_scratch_bytes(int64_t M, int64_t N, int32_t planes);
int hs_resource_topk(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t plane_stride,
                     int32_t planes, int64_t M, int64_t N, int64_t K, const float* cs, double scale,
                     const uint32_t* excl_mask, int64_t mask_stride_bits, int64_t row_offset,
                     int32_t k, int64_t* out_idx, float* out_val, void* scratch,
                     int64_t scratch_bytes, lgc_stream_t stream);
/* tcgen05 kind::f16 accumulates with truncation (measured: tools/probe_umma_numerics.py), so
 * the bf16 kind drains its TMEM accumulator into round-to-nearest fp32 registers every
 * chunk_kb K-blocks of 64 elements (default 8): error <= 4*chunk_kb*2^-23 per output,
 * independent of K.  Tuning/diagnostic knob. */
int hs_gemm_config(int32_t chunk_kb);
/* 1 (default): CTA-pair kernel (tcgen05 cta_group::2, UMMA M=256, each CTA stages half of B);
 * 0: single-CTA kernel (M=128).  Results are bit-identical. */
int hs_gemm_use_cta_pair(int32_t on);
/* Plain CUDA-core fp32-accumulate version of the same contract, used by the GPU tests as
 * an on-device cross-check of the tcgen05 path (never by the product path). */
int hs_gemm_planes_simt(int32_t kind, const void* A, int64_t lda, const void* B,
                        int64_t ldb, int64_t plane_stride, int32_t planes, int64_t M,
                        int64_t N, int64_t K, float* C, int64_t ldc, const float* rs,
                        const float* cs, double scale, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (S2) HybridS degree scaling:  W[i,j] = G[i,j] / den,  den = k_i^(1-lambda) * k_j^lambda,
 * den == 0 -> 1  (model/SpreadMethod/model.py:72-83), evaluated in float64 like the
 * reference and rounded once to fp32.
 *   W32   optional (n x ldw) fp32 output (row i = source item)
 *   Wt_planes optional bf16 planes p=0..planes-1 of W^T: plane[p][j, i] = split_p(W[i,j])
 *             (hi / mid / lo), (n x ldk) each, K-major operand for F = A.W
 * ---------------------------------------------------------------------------------- */
int hs_scale_w(const float* G, int64_t ldg, int64_t n, const int32_t* ki, double lambda,
               float* W32, int64_t ldw, uint16_t* Wt_planes, int64_t ldk,
               int64_t plane_stride, int32_t planes, lgc_stream_t stream);

/* HybridS scaling with a FIXED-POINT result for the exact int8 F = A.W GEMM: per column j,
 * q[i,j] = round(W[i,j] / s_j * 256^digits) with s_j the power of two strictly above max_i W[i,j];
 * digit planes plane[d][j, i] (uint8, K = source item i contiguous) and col_scale[j] = s_j / 256^digits
 * (the GEMM's epilogue column scale).  Products and int32 sums are exact, so
 * |F[u,j] - ref| <= k_u * 2^-(8*digits+1) * s_j plus one fp32 rounding.  W must be >= 0 (it is: G >= 0).
 * scratch: 20*n bytes, 8-byte aligned.  A must be packed as uint8 (hs_pack_a_u8). */
int hs_pack_a_u8(const int32_t* users, const int32_t* items, int64_t nnz, int64_t n_users,
                 int64_t n_items, uint8_t* A_u8, int64_t ldk, lgc_stream_t stream);
int hs_scale_w_u8(const float* G, int64_t ldg, int64_t n, const int32_t* ki, double lambda,
                  float* W32, int64_t ldw, uint8_t* Wt_digits, int64_t ldk, int64_t plane_stride,
                  int32_t digits, float* col_scale, void* scratch, lgc_stream_t stream);

/* (F1) Fusion  F_new = G_score * F  (model/SpreadLightGCN/model.py:151), elementwise,
 * in place on F. */
int hs_hadamard(float* F, const float* Gscore, int64_t rows, int64_t cols, int64_t ldf,
                int64_t ldg, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Multi-GPU support for the fused SpMM + all-gather (rows written straight into every
 * peer's replica over NVLink): CUDA IPC handle export / import for a device buffer.
 * ---------------------------------------------------------------------------------- */
/* blob = 64-byte cudaIpcMemHandle_t of the containing allocation + 8-byte offset of dptr in it */
int lgc_ipc_get_handle(void* dptr, uint8_t* handle72_host);
int lgc_ipc_open_handle(const uint8_t* handle72_host, void** dptr_host);
int lgc_ipc_close_handle(void* base_dptr);
/* Same as lgc_spmm_layer for rows [row_begin,row_end) but every finished row is stored
 * into n_peers output replicas (peer_Y_host[p] = device pointer valid in this process). */
int lgc_spmm_layer_bcast(const int32_t* rowptr, const int32_t* colidx, const float* val,
                         const int32_t* chunk_row, const int32_t* chunk_start,
                         const int32_t* row_chunk_base, int32_t chunk_begin,
                         int32_t chunk_end, int64_t n_nodes,
                         int32_t dim, int64_t row_begin, int64_t row_end,
                         const int32_t* row_order, int32_t long_row, const float* X,
                         const float* X0, float alpha, float beta,
                         float* const* peer_Y_host, int32_t n_peers, float* partial,
                         int32_t* counters, lgc_stream_t stream);

/* lgc_bpr_fwd_bwd with a DETERMINISTIC gradient scatter (bit-reproducible training): every (triplet, role) writes its
 * gradient row to a compact per-entry array, then the first entry of every distinct table row adds its duplicates in
 * ascending entry order and stores the sums with plain stores — no floating-point atomics.  gE / gX0 must be zero on
 * entry in the rows the batch does not touch (they are not written).  workspace: lgc_bpr_det_workspace_bytes(). */
int64_t lgc_bpr_det_workspace_bytes(int64_t batch, int32_t dim);
int lgc_bpr_fwd_bwd_det(const float* E, const float* X0, int64_t n_users, int64_t n_items, int32_t dim,
                        const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch, float eps,
                        float grad_scale, float* loss_out, float* gE, float* gX0, float* scratch,
                        void* workspace, int64_t workspace_bytes, lgc_stream_t stream);
/* Adam (torch.optim.Adam.step, model/LightGCN/train.py:104,144) on elements [offset, offset + n) of the parameter
 * table with the step-dependent scalars in device memory (lgc_adam_hyper_step), the gradient given as grad + grad2
 * (grad2 may be null: the propagated gradient and the sparse direct rows need no separate add pass), and the updated
 * parameters stored into n_peers replicas of the table (peer_tables_host[r]: rank r's copy mapped through CUDA IPC;
 * n_peers = 0: param_table only).  Multi-GPU (SURVEY 8e "gradient rows live where the shard lives"): the owner of a row
 * range is the only rank that updates it and pushes the new rows to every replica over NVLink. */
int lgc_adam_step_fused(float* param_table, float* const* peer_tables_host, int32_t n_peers, const float* grad,
                        const float* grad2, float* exp_avg, float* exp_avg_sq, int64_t offset, int64_t n,
                        float beta1, float beta2, float eps, const float* hyper_dev, lgc_stream_t stream);
/* Zero the rows users[b], n_users + pos[b], n_users + neg[b] of up to two (N, dim) gradient tables (a, b; either may
 * be null): the clean-up after a step, instead of a memset of the whole tables. */
int lgc_zero_rows(float* a, float* b, int32_t dim, const int64_t* users, const int64_t* pos, const int64_t* neg,
                  int64_t batch, int64_t n_users, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Measurement probe (bench.py): uniformly random whole-row reads (dim fp32 per row, 128-bit
 * per lane) from an (n_rows, dim) table at full occupancy — the gather pattern of the
 * propagation SpMM with everything else removed.  Its bytes / time is the L2 gather peak the
 * SpMM's gather fraction is stated against.  out: lgc_probe_gather_threads() floats of
 * scratch; *gathers_done_host receives the number of rows actually read (>= n_gathers).
 * No reference counterpart (the reference has no benchmark harness).
 * ---------------------------------------------------------------------------------- */
int64_t lgc_probe_gather_threads(void);
int lgc_probe_gather(const float* table, int64_t n_rows, int32_t dim, int64_t n_gathers,
                     uint32_t seed, float* out, int64_t* gathers_done_host, lgc_stream_t stream);

/* Measurement probe (tools/mcast_store_probe.py): the EXCHANGE half of the row-partitioned propagation with the gather
 * removed — n_rows rows of 64 fp32 (row_list = device int32 row ids, or null for 0..n_rows-1) are copied from `src` to the
 * same offsets of `dst`, which may be local, a CUDA-IPC peer mapping or an NVSwitch multicast address.  mode: 0 = 16 lanes x
 * 16-byte stores per row (what lgc_spmm_rows_bcast issues), 1 = 32 lanes (two rows per instruction), 2 = one 256-byte
 * cp.async.bulk (TMA) store per row from shared memory, 3 = one 8 KB cp.async.bulk store per 32 consecutive rows. */
int lgc_probe_row_store(const float* src, float* dst, const int32_t* row_list, int64_t n_rows, int32_t mode,
                        lgc_stream_t stream);
/* The same copy by a PERSISTENT grid of n_ctas CTAs x 1024 threads walking the row list `passes` times; exclusive != 0
 * launches it with enough dynamic shared memory that no other CTA shares its SMs (a dedicated-sender emulation). */
int lgc_probe_row_store_persistent(const float* src, float* dst, const int32_t* row_list, int64_t n_rows,
                                   int32_t n_ctas, int32_t passes, int32_t exclusive, lgc_stream_t stream);

/* lgc_spmm_layer over an explicit LIST of rows (device int32[n_rows], any subset of the nodes; longest rows first gives
 * the best balance) plus up to two ranges of the long-row chunk list, in ONE launch; every finished row is stored into
 * n_peers replicas (n_peers = 1 with the local buffer: a plain local SpMM).  The multi-GPU partition gives a rank a slice
 * of the user rows and a slice of the item rows: one mixed launch keeps both row classes in flight together. */
int lgc_spmm_rows_bcast(const int32_t* rowptr, const int32_t* colidx, const float* val,
                        const int32_t* chunk_row, const int32_t* chunk_start,
                        const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                        int32_t chunk_begin2, int32_t chunk_end2, int64_t n_nodes, int32_t dim,
                        const int32_t* row_list, int64_t n_rows, int32_t long_row, const float* X,
                        const float* X0, float alpha, float beta, float* const* peer_Y_host,
                        int32_t n_peers, float* partial, int32_t* counters, lgc_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sparse-source layers.  autograd's backward of the propagation
 * (/root/reference/model/LightGCN/train.py:140, loss.backward()) starts from dL/dE, which has at most 3 * batch non-zero rows
 * (the users, positives and negatives of the mini-batch) out of U + M: the first gradient layer would gather nnz rows of
 * which ~98 % are exact zeros.  src_mask / x0_row_mask = one bit per row of the source matrix, bit (r & 31) of word r >> 5,
 * CLEAR = the row is all-zero; such rows are not gathered and groups of non-zeros without a live source are skipped.  The
 * result is the one the unmasked call gives (the skipped terms are exact zeros).  dim 32 or 64.
 *   lgc_row_mask_batch       : set (set != 0) the bits of rows users[b], n_users + pos[b], n_users + neg[b], or clear them
 *                              again (set == 0; whole words — every set bit of the mask must come from the same batch).
 *   lgc_spmm_layer_masked    : lgc_spmm_layer with the mask on X.
 *   lgc_spmm_rows_bcast_masked: lgc_spmm_rows_bcast with the mask on X.
 *   lgc_propagate_mean_masked: lgc_propagate_mean with the mask on X0, applied in the first layer (the only one whose
 *                              source is X0 itself).
 * ---------------------------------------------------------------------------------- */
int lgc_row_mask_batch(uint32_t* mask, const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch,
                       int64_t n_users, int32_t set, lgc_stream_t stream);
int lgc_spmm_layer_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                          const int32_t* chunk_row, const int32_t* chunk_start,
                          const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                          int64_t n_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                          const int32_t* row_order, int32_t long_row, const float* X, const float* X0,
                          float alpha, float beta, float* Y, float* partial, int32_t* counters,
                          const uint32_t* src_mask, lgc_stream_t stream);
int lgc_spmm_rows_bcast_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                               const int32_t* chunk_row, const int32_t* chunk_start,
                               const int32_t* row_chunk_base, int32_t chunk_begin, int32_t chunk_end,
                               int32_t chunk_begin2, int32_t chunk_end2, int64_t n_nodes, int32_t dim,
                               const int32_t* row_list, int64_t n_rows, int32_t long_row, const float* X,
                               const float* X0, float alpha, float beta, float* const* peer_Y_host,
                               int32_t n_peers, float* partial, int32_t* counters, const uint32_t* src_mask,
                               lgc_stream_t stream);
int lgc_propagate_mean_masked(const int32_t* rowptr, const int32_t* colidx, const float* val,
                              const int32_t* chunk_row, const int32_t* chunk_start,
                              const int32_t* row_chunk_base, int32_t n_chunks, int64_t n_nodes,
                              int32_t dim, int32_t n_layers, const int32_t* row_order, int32_t long_row,
                              const float* X0, float* E, float* tmp0, float* tmp1, float* partial,
                              int32_t* counters, const uint32_t* x0_row_mask, lgc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCNHS_H_ */
