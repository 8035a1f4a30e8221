// train_ops.cu — (P4/P6/P7) the per-step training math around the propagation:
//   * fused BPR forward + backward with the six embedding-row gathers and the
//     embedding-gradient scatter folded in, standing in for
//     /root/reference/model/LightGCN/loss.py:12-44 (BPRLoss, ~12 library kernels) and
//     /root/reference/model/LightGCN/train.py:55-57 (six index_select) plus their autograd
//     backward (index_add / elementwise);
//   * Adam, standing in for torch.optim.Adam.step at
//     /root/reference/model/LightGCN/train.py:104,144.
// Both are tiny HBM-bound elementwise passes (3.1 MB at B=1024, 7*N*4D bytes for Adam).
#include "common.cuh"

namespace lgc {

constexpr int kBprThreads = 256;
constexpr int kBprMaxGrid = 2048;

struct BprPtrs {
  const float *uf, *u0, *pf, *p0, *nf, *n0;  // row bases (forward operands)
  float *guf, *gu0, *gpf, *gp0, *gnf, *gn0;  // gradient bases (null -> forward only)
  const int64_t *iu, *ip, *in;               // row indices, null -> row b (pre-gathered rows)
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
template <bool ATOMIC>
__device__ __forceinline__ void add4(float* p, float4 v) {
  if (ATOMIC) {
    atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
  } else {
    *reinterpret_cast<float4*>(p) = v;
  }
}

// torch.nn.functional.softplus(x) with beta=1, threshold=20 and its derivative
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_t(float x) {
  if (x > 20.f) return 1.f;
  const float z = expf(x);
  return z / (z + 1.f);
}

// One sub-group of DIM/4 lanes per (user, pos, neg) triplet.
//   ATOMIC: gradients are scatter-added (duplicate rows in a batch); otherwise stored.
template <int DIM, bool ATOMIC>
__global__ void __launch_bounds__(kBprThreads)
bpr_kernel(BprPtrs P, int64_t batch, float eps, float grad_scale, float* __restrict__ loss_out,
           float* __restrict__ scratch) {
  constexpr int LPR = DIM / 4;
  constexpr int TPB = kBprThreads / LPR;  // triplets per block per pass
  const int sub = threadIdx.x / LPR, li = threadIdx.x % LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float sp_sum = 0.f, reg_sum = 0.f;
  const float inv_b = 1.0f / (float)batch;

  for (int64_t b0 = (int64_t)blockIdx.x * TPB; b0 < batch; b0 += (int64_t)gridDim.x * TPB) {
    const int64_t b = b0 + sub;
    const bool ok = b < batch;
    float4 uf = make_float4(0, 0, 0, 0), pf = uf, nf = uf, u0 = uf, p0 = uf, n0 = uf;
    size_t ou = 0, op = 0, on = 0;
    if (ok) {
      ou = (size_t)(P.iu ? P.iu[b] : b) * DIM + li * 4;
      op = (size_t)(P.ip ? P.ip[b] : b) * DIM + li * 4;
      on = (size_t)(P.in ? P.in[b] : b) * DIM + li * 4;
      uf = ld4(P.uf + ou); pf = ld4(P.pf + op); nf = ld4(P.nf + on);
      u0 = ld4(P.u0 + ou); p0 = ld4(P.p0 + op); n0 = ld4(P.n0 + on);
    }
    float sp = dot4(uf, pf), sn = dot4(uf, nf);
    float rg = dot4(u0, u0) + dot4(p0, p0) + dot4(n0, n0);
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) {
      sp += __shfl_xor_sync(0xffffffffu, sp, off);
      sn += __shfl_xor_sync(0xffffffffu, sn, off);
      rg += __shfl_xor_sync(0xffffffffu, rg, off);
    }
    const float x = sp - sn;
    if (ok && li == 0) {
      sp_sum += softplus_t(x);
      reg_sum += rg;
    }
    if (P.guf != nullptr && ok) {
      // d loss / d x_b = -(1/B) * softplus'(x_b)
      const float g = -grad_scale * inv_b * softplus_grad_t(x);
      const float r = 2.f * eps * grad_scale;
      add4<ATOMIC>(P.guf + ou, make_float4(g * (pf.x - nf.x), g * (pf.y - nf.y), g * (pf.z - nf.z), g * (pf.w - nf.w)));
      add4<ATOMIC>(P.gpf + op, make_float4(g * uf.x, g * uf.y, g * uf.z, g * uf.w));
      add4<ATOMIC>(P.gnf + on, make_float4(-g * uf.x, -g * uf.y, -g * uf.z, -g * uf.w));
      add4<ATOMIC>(P.gu0 + ou, make_float4(r * u0.x, r * u0.y, r * u0.z, r * u0.w));
      add4<ATOMIC>(P.gp0 + op, make_float4(r * p0.x, r * p0.y, r * p0.z, r * p0.w));
      add4<ATOMIC>(P.gn0 + on, make_float4(r * n0.x, r * n0.y, r * n0.z, r * n0.w));
    }
  }

  // block reduce (fixed order), then the last block adds the block partials in block order
  __shared__ float s_sp[kBprThreads / 32], s_rg[kBprThreads / 32];
  __shared__ int s_last;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    sp_sum += __shfl_xor_sync(0xffffffffu, sp_sum, off);
    reg_sum += __shfl_xor_sync(0xffffffffu, reg_sum, off);
  }
  if (lane == 0) { s_sp[warp] = sp_sum; s_rg[warp] = reg_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < kBprThreads / 32; ++w) { a += s_sp[w]; c += s_rg[w]; }
    float* part = scratch + 2;
    __stcg(part + 2 * blockIdx.x, a);
    __stcg(part + 2 * blockIdx.x + 1, c);
    __threadfence();
    const int ticket = atomicAdd(reinterpret_cast<int*>(scratch), 1);
    s_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float a = 0.f, c = 0.f;
    const float* part = scratch + 2;
    for (unsigned i = 0; i < gridDim.x; ++i) { a += __ldcg(part + 2 * i); c += __ldcg(part + 2 * i + 1); }
    const float bpr = -a * inv_b;
    loss_out[0] = bpr + eps * c;
    loss_out[1] = bpr;
    *reinterpret_cast<int*>(scratch) = 0;
  }
}

// ---- deterministic gradient scatter (bit-reproducible training) ------------------------------------------------------
// The atomicAdd scatter above sums the contributions of duplicate rows of a batch in an order that changes from run to
// run.  Deterministic variant, two kernels:
//   1. bpr_compact_kernel: the same forward/backward math, but every (triplet, role) writes its gradient row into
//      compact per-entry arrays (entry e = 3 b + role; role 0 = user, 1 = positive item, 2 = negative item);
//   2. bpr_reduce_kernel: one sub-group of DIM/4 lanes per entry; the FIRST entry of every distinct table row (no
//      earlier entry with the same row) adds the rows of all its duplicates in ascending entry order and stores the
//      sums into gE / gX0 with plain stores.  O(entries^2 / 16) integer compares in L1 — 0.6 M for B = 1024.
template <int DIM>
__global__ void __launch_bounds__(kBprThreads)
bpr_compact_kernel(const float* __restrict__ E, const float* __restrict__ X0, int64_t n_users,
                   const int64_t* __restrict__ users, const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                   int64_t batch, float eps, float grad_scale, float* __restrict__ loss_out, float* __restrict__ scratch,
                   int32_t* __restrict__ ent_row, float* __restrict__ ent_gE, float* __restrict__ ent_gX) {
  constexpr int LPR = DIM / 4;
  constexpr int TPB = kBprThreads / LPR;
  const int sub = threadIdx.x / LPR, li = threadIdx.x % LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float sp_sum = 0.f, reg_sum = 0.f;
  const float inv_b = 1.0f / (float)batch;
  for (int64_t b0 = (int64_t)blockIdx.x * TPB; b0 < batch; b0 += (int64_t)gridDim.x * TPB) {
    const int64_t b = b0 + sub;
    const bool ok = b < batch;
    float4 uf = make_float4(0, 0, 0, 0), pf = uf, nf = uf, u0 = uf, p0 = uf, n0 = uf;
    int64_t ru = 0, rp = 0, rn = 0;
    if (ok) {
      ru = users[b]; rp = n_users + pos[b]; rn = n_users + neg[b];
      uf = ld4(E + ru * DIM + li * 4); pf = ld4(E + rp * DIM + li * 4); nf = ld4(E + rn * DIM + li * 4);
      u0 = ld4(X0 + ru * DIM + li * 4); p0 = ld4(X0 + rp * DIM + li * 4); n0 = ld4(X0 + rn * DIM + li * 4);
    }
    float sp = dot4(uf, pf), sn = dot4(uf, nf);
    float rg = dot4(u0, u0) + dot4(p0, p0) + dot4(n0, n0);
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) {
      sp += __shfl_xor_sync(0xffffffffu, sp, off);
      sn += __shfl_xor_sync(0xffffffffu, sn, off);
      rg += __shfl_xor_sync(0xffffffffu, rg, off);
    }
    const float x = sp - sn;
    if (ok && li == 0) {
      sp_sum += softplus_t(x);
      reg_sum += rg;
    }
    if (ok) {
      const float g = -grad_scale * inv_b * softplus_grad_t(x);
      const float r = 2.f * eps * grad_scale;
      const size_t e0 = (size_t)(3 * b) * DIM + li * 4;
      if (li == 0) { ent_row[3 * b] = (int32_t)ru; ent_row[3 * b + 1] = (int32_t)rp; ent_row[3 * b + 2] = (int32_t)rn; }
      *reinterpret_cast<float4*>(ent_gE + e0) = make_float4(g * (pf.x - nf.x), g * (pf.y - nf.y), g * (pf.z - nf.z), g * (pf.w - nf.w));
      *reinterpret_cast<float4*>(ent_gE + e0 + DIM) = make_float4(g * uf.x, g * uf.y, g * uf.z, g * uf.w);
      *reinterpret_cast<float4*>(ent_gE + e0 + 2 * DIM) = make_float4(-g * uf.x, -g * uf.y, -g * uf.z, -g * uf.w);
      *reinterpret_cast<float4*>(ent_gX + e0) = make_float4(r * u0.x, r * u0.y, r * u0.z, r * u0.w);
      *reinterpret_cast<float4*>(ent_gX + e0 + DIM) = make_float4(r * p0.x, r * p0.y, r * p0.z, r * p0.w);
      *reinterpret_cast<float4*>(ent_gX + e0 + 2 * DIM) = make_float4(r * n0.x, r * n0.y, r * n0.z, r * n0.w);
    }
  }
  __shared__ float s_sp[kBprThreads / 32], s_rg[kBprThreads / 32];
  __shared__ int s_last;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    sp_sum += __shfl_xor_sync(0xffffffffu, sp_sum, off);
    reg_sum += __shfl_xor_sync(0xffffffffu, reg_sum, off);
  }
  if (lane == 0) { s_sp[warp] = sp_sum; s_rg[warp] = reg_sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < kBprThreads / 32; ++w) { a += s_sp[w]; c += s_rg[w]; }
    float* part = scratch + 2;
    __stcg(part + 2 * blockIdx.x, a);
    __stcg(part + 2 * blockIdx.x + 1, c);
    __threadfence();
    const int ticket = atomicAdd(reinterpret_cast<int*>(scratch), 1);
    s_last = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float a = 0.f, c = 0.f;
    const float* part = scratch + 2;
    for (unsigned i = 0; i < gridDim.x; ++i) { a += __ldcg(part + 2 * i); c += __ldcg(part + 2 * i + 1); }
    const float bpr = -a * inv_b;
    loss_out[0] = bpr + eps * c;
    loss_out[1] = bpr;
    *reinterpret_cast<int*>(scratch) = 0;
  }
}

template <int DIM>
__global__ void __launch_bounds__(kBprThreads)
bpr_reduce_kernel(const int32_t* __restrict__ ent_row, const float* __restrict__ ent_gE, const float* __restrict__ ent_gX,
                  int n_ent, float* __restrict__ gE, float* __restrict__ gX) {
  constexpr int LPR = DIM / 4;
  const int e = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR);
  const int li = threadIdx.x % LPR;
  if (e >= n_ent) return;                         // whole sub-groups leave together (LPR divides the block size)
  const int row = __ldg(ent_row + e);
  // a sub-group is LPR consecutive lanes of one warp: its lanes vote with a sub-group mask
  const unsigned lane = threadIdx.x & 31u;
  const unsigned gmask = (LPR == 32 ? 0xffffffffu : ((1u << LPR) - 1u) << (lane & ~(unsigned)(LPR - 1)));
  bool dup = false;
  for (int j0 = 0; j0 < e; j0 += LPR) {
    const int j = j0 + li;
    dup |= (j < e) && (__ldg(ent_row + j) == row);
  }
  if (__any_sync(gmask, dup)) return;             // not the first entry of its row
  float4 a = *reinterpret_cast<const float4*>(ent_gE + (size_t)e * DIM + li * 4);
  float4 x = *reinterpret_cast<const float4*>(ent_gX + (size_t)e * DIM + li * 4);
  for (int j0 = e + 1; j0 < n_ent; j0 += LPR) {
    const int j = j0 + li;
    const bool hit = (j < n_ent) && (__ldg(ent_row + j) == row);
    unsigned m = (__ballot_sync(gmask, hit) & gmask) >> (lane & ~(unsigned)(LPR - 1));
    while (m) {                                   // duplicates in ascending entry order
      const int jj = j0 + (__ffs(m) - 1);
      m &= m - 1u;
      const float4 b = *reinterpret_cast<const float4*>(ent_gE + (size_t)jj * DIM + li * 4);
      const float4 y = *reinterpret_cast<const float4*>(ent_gX + (size_t)jj * DIM + li * 4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    }
  }
  *reinterpret_cast<float4*>(gE + (size_t)row * DIM + li * 4) = a;
  *reinterpret_cast<float4*>(gX + (size_t)row * DIM + li * 4) = x;
}

static int bpr_grid(int64_t batch, int dim) {
  const int tpb = kBprThreads / (dim / 4);
  int64_t g = ceil_div(batch, tpb);
  int64_t cap = (int64_t)num_sms() * 8;
  if (cap > kBprMaxGrid) cap = kBprMaxGrid;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <bool ATOMIC>
static int launch_bpr(const BprPtrs& P, int dim, int64_t batch, float eps, float grad_scale,
                      float* loss_out, float* scratch, cudaStream_t stream) {
  const int grid = bpr_grid(batch, dim);
  switch (dim) {
    case 32: bpr_kernel<32, ATOMIC><<<grid, kBprThreads, 0, stream>>>(P, batch, eps, grad_scale, loss_out, scratch); break;
    case 64: bpr_kernel<64, ATOMIC><<<grid, kBprThreads, 0, stream>>>(P, batch, eps, grad_scale, loss_out, scratch); break;
    case 128: bpr_kernel<128, ATOMIC><<<grid, kBprThreads, 0, stream>>>(P, batch, eps, grad_scale, loss_out, scratch); break;
    default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "bpr: embedding dim %d not in {32,64,128}", dim);
  }
  LGC_LAUNCH_CHECK("bpr_kernel");
  return LGC_OK;
}

__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                            float4* __restrict__ v, int64_t n4, float* __restrict__ pt,
                            const float* __restrict__ gt, float* __restrict__ mt, float* __restrict__ vt,
                            int tail, float beta1, float beta2, float eps, float step_size, float bc2_sqrt,
                            const float* __restrict__ hyper) {
  if (hyper) {  // step-dependent scalars kept on the device, so the whole step can live in a CUDA graph
    step_size = hyper[0];
    bc2_sqrt = hyper[1];
  }
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    mm = mm + (gg - mm) * (1.f - beta1);          // exp_avg.lerp_(grad, 1-beta1)
    vv = vv * beta2 + (1.f - beta2) * gg * gg;    // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);           // param.addcdiv_(exp_avg, denom, -step_size)
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
    const int i = threadIdx.x;
    float pp = pt[i], mm = mt[i], vv = vt[i];
    upd(pp, gt[i], mm, vv);
    pt[i] = pp; mt[i] = mm; vt[i] = vv;
  }
}

// Adam on a slice of the parameter table with (1) the gradient given as the sum of two buffers — the propagated
// gradient and the sparse direct (regulariser) rows — so that no separate add pass is needed, and (2) the updated
// parameters stored into every replica of the table (multi-GPU: peers[r] = rank r's copy of the table, mapped through
// CUDA IPC; the owner of a row range is the only rank that updates it, and pushes the result over NVLink).
struct ParamPeers {
  float* p[8];
};

__global__ void adam_fused_kernel(const float4* __restrict__ p_local, ParamPeers peers, int n_peers,
                                  const float4* __restrict__ g1, const float4* __restrict__ g2, float4* __restrict__ m,
                                  float4* __restrict__ v, int64_t off4, int64_t n4, float beta1, float beta2, float eps,
                                  const float* __restrict__ hyper) {
  const float step_size = hyper[0], bc2_sqrt = hyper[1];
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    mm = mm + (gg - mm) * (1.f - beta1);
    vv = vv * beta2 + (1.f - beta2) * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = off4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < off4 + n4; i += stride) {
    float4 pp = p_local[i], gg = g1[i], mm = m[i], vv = v[i];
    if (g2) {
      const float4 h = g2[i];
      gg.x += h.x; gg.y += h.y; gg.z += h.z; gg.w += h.w;
    }
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    m[i] = mm; v[i] = vv;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < n_peers) reinterpret_cast<float4*>(peers.p[r])[i] = pp;
  }
}

// zero the rows a mini-batch touched in up to two (N, dim) gradient buffers (instead of memset of the whole tables)
__global__ void zero_rows_kernel(float* __restrict__ a, float* __restrict__ b, int dim, const int64_t* __restrict__ users,
                                 const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int64_t batch,
                                 int64_t n_users) {
  const int lpr = dim / 4;
  const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / lpr;
  const int li = threadIdx.x % lpr;
  if (t >= 3 * batch) return;
  const int64_t bi = t / 3;
  const int which = (int)(t % 3);
  const int64_t row = which == 0 ? users[bi] : n_users + (which == 1 ? pos[bi] : neg[bi]);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a) *reinterpret_cast<float4*>(a + row * dim + li * 4) = z;
  if (b) *reinterpret_cast<float4*>(b + row * dim + li * 4) = z;
}

// one bit per node: set (or clear again) the bits of the rows a mini-batch touches — the non-zero rows of dL/dE, which the
// first gradient-propagation layer uses to skip the gathers of all-zero rows (lgc_propagate_mean_masked).  Clearing stores
// whole words: every set bit of the mask comes from the same batch.
__global__ void row_mask_batch_kernel(uint32_t* __restrict__ mask, const int64_t* __restrict__ users,
                                      const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int64_t batch,
                                      int64_t n_users, int set) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= 3 * batch) return;
  const int64_t bi = t / 3;
  const int which = (int)(t % 3);
  const int64_t row = which == 0 ? users[bi] : n_users + (which == 1 ? pos[bi] : neg[bi]);
  if (set) atomicOr(mask + (row >> 5), 1u << (row & 31));
  else mask[row >> 5] = 0u;
}

// ++step; hyper = {lr / (1 - beta1^step), sqrt(1 - beta2^step)}  — torch.optim.Adam's bias corrections
// (float64 like the Python side of torch's single-tensor path, rounded once to fp32)
__global__ void adam_hyper_kernel(long long* __restrict__ step, const float* __restrict__ lr, double beta1, double beta2,
                                  float* __restrict__ hyper) {
  const long long t = *step + 1;
  *step = t;
  const float bc1 = (float)(1.0 - pow(beta1, (double)t));
  hyper[0] = lr[0] / bc1;
  hyper[1] = (float)sqrt(1.0 - pow(beta2, (double)t));
}

}  // namespace lgc

using namespace lgc;

extern "C" int64_t lgc_bpr_scratch_floats(int64_t batch) {
  (void)batch;
  return 2 + 2 * (int64_t)kBprMaxGrid;  // ticket counter + one (softplus, reg) pair per block
}

extern "C" int lgc_bpr_fwd_bwd(const float* E, const float* X0, int64_t n_users, int64_t n_items,
                               int32_t dim, const int64_t* users, const int64_t* pos,
                               const int64_t* neg, int64_t batch, float eps, float grad_scale,
                               float* loss_out, float* gE, float* gX0, float* scratch,
                               lgc_stream_t stream) {
  LGC_REQUIRE(E && X0 && users && pos && neg && loss_out && scratch, "bpr: null pointer");
  LGC_REQUIRE(batch > 0 && n_users > 0 && n_items > 0, "bpr: empty batch or tables");
  LGC_REQUIRE((gE == nullptr) == (gX0 == nullptr), "bpr: gE and gX0 must both be given or both be null");
  const size_t ioff = (size_t)n_users * dim;
  BprPtrs P{};
  P.uf = E; P.pf = E + ioff; P.nf = E + ioff;
  P.u0 = X0; P.p0 = X0 + ioff; P.n0 = X0 + ioff;
  if (gE) {
    P.guf = gE; P.gpf = gE + ioff; P.gnf = gE + ioff;
    P.gu0 = gX0; P.gp0 = gX0 + ioff; P.gn0 = gX0 + ioff;
  }
  P.iu = users; P.ip = pos; P.in = neg;
  return launch_bpr<true>(P, dim, batch, eps, grad_scale, loss_out, scratch, (cudaStream_t)stream);
}

extern "C" int lgc_bpr_rows(const float* uf, const float* u0, const float* pf, const float* p0,
                            const float* nf, const float* n0, int64_t batch, int32_t dim, float eps,
                            float grad_scale, float* loss_out, float* guf, float* gu0, float* gpf,
                            float* gp0, float* gnf, float* gn0, float* scratch, lgc_stream_t stream) {
  LGC_REQUIRE(uf && u0 && pf && p0 && nf && n0 && loss_out && scratch, "bpr rows: null pointer");
  LGC_REQUIRE(batch > 0, "bpr rows: empty batch");
  const bool bwd = guf != nullptr;
  LGC_REQUIRE(!bwd || (gu0 && gpf && gp0 && gnf && gn0), "bpr rows: partial gradient outputs");
  BprPtrs P{};
  P.uf = uf; P.u0 = u0; P.pf = pf; P.p0 = p0; P.nf = nf; P.n0 = n0;
  P.guf = guf; P.gu0 = gu0; P.gpf = gpf; P.gp0 = gp0; P.gnf = gnf; P.gn0 = gn0;
  return launch_bpr<false>(P, dim, batch, eps, grad_scale, loss_out, scratch, (cudaStream_t)stream);
}

extern "C" int lgc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t n, float lr, float beta1, float beta2, float eps, float bc1,
                             float bc2_sqrt, lgc_stream_t stream) {
  LGC_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "adam: null pointer / empty");
  LGC_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "adam: buffers must be 16-byte aligned");
  LGC_REQUIRE(bc1 > 0.f && bc2_sqrt > 0.f, "adam: bias corrections must be positive");
  const int64_t n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  int64_t grid = ceil_div(n4 > 0 ? n4 : 1, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (grid > cap) grid = cap;
  adam_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4, param + n4 * 4,
      grad + n4 * 4, exp_avg + n4 * 4, exp_avg_sq + n4 * 4, tail, beta1, beta2, eps, lr / bc1, bc2_sqrt, nullptr);
  LGC_LAUNCH_CHECK("adam_kernel");
  return LGC_OK;
}

extern "C" int lgc_adam_hyper_step(int64_t* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_dev,
                                   lgc_stream_t stream) {
  LGC_REQUIRE(step_dev && lr_dev && hyper_dev, "adam hyper: null pointer");
  adam_hyper_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)step_dev, lr_dev, (double)beta1, (double)beta2, hyper_dev);
  LGC_LAUNCH_CHECK("adam_hyper_kernel");
  return LGC_OK;
}

extern "C" int lgc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 float beta1, float beta2, float eps, const float* hyper_dev, lgc_stream_t stream) {
  LGC_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper_dev && n > 0, "adam: null pointer / empty");
  LGC_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "adam: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  int64_t grid = ceil_div(n4 > 0 ? n4 : 1, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (grid > cap) grid = cap;
  adam_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      (float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4, param + n4 * 4,
      grad + n4 * 4, exp_avg + n4 * 4, exp_avg_sq + n4 * 4, tail, beta1, beta2, eps, 0.f, 1.f, hyper_dev);
  LGC_LAUNCH_CHECK("adam_kernel");
  return LGC_OK;
}

extern "C" int lgc_adam_step_fused(float* param_table, float* const* peer_tables_host, int32_t n_peers, const float* grad,
                                   const float* grad2, float* exp_avg, float* exp_avg_sq, int64_t offset, int64_t n,
                                   float beta1, float beta2, float eps, const float* hyper_dev, lgc_stream_t stream) {
  LGC_REQUIRE(param_table && grad && exp_avg && exp_avg_sq && hyper_dev && n > 0 && offset >= 0, "adam fused: null pointer / empty");
  LGC_REQUIRE((((uintptr_t)param_table | (uintptr_t)grad | (uintptr_t)grad2 | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "adam fused: buffers must be 16-byte aligned");
  LGC_REQUIRE((offset & 3) == 0 && (n & 3) == 0, "adam fused: offset and count must be multiples of 4 elements");
  LGC_REQUIRE(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_tables_host), "adam fused: 0..8 replicas");
  ParamPeers peers{};
  int np = n_peers;
  if (np == 0) {
    peers.p[0] = param_table;
    np = 1;
  } else {
    for (int r = 0; r < n_peers; ++r) {
      LGC_REQUIRE(peer_tables_host[r] && ((uintptr_t)peer_tables_host[r] & 15) == 0, "adam fused: bad replica pointer");
      peers.p[r] = peer_tables_host[r];
    }
  }
  const int64_t n4 = n / 4;
  int64_t grid = ceil_div(n4, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (grid > cap) grid = cap;
  adam_fused_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      (const float4*)param_table, peers, np, (const float4*)grad, (const float4*)grad2, (float4*)exp_avg, (float4*)exp_avg_sq,
      offset / 4, n4, beta1, beta2, eps, hyper_dev);
  LGC_LAUNCH_CHECK("adam_fused_kernel");
  return LGC_OK;
}

extern "C" int lgc_zero_rows(float* a, float* b, int32_t dim, const int64_t* users, const int64_t* pos, const int64_t* neg,
                             int64_t batch, int64_t n_users, lgc_stream_t stream) {
  LGC_REQUIRE((a || b) && users && pos && neg && batch > 0, "zero_rows: null pointer / empty batch");
  LGC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "zero_rows: dim must be 32, 64 or 128");
  LGC_REQUIRE((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "zero_rows: buffers must be 16-byte aligned");
  const int64_t threads = 3 * batch * (dim / 4);
  zero_rows_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(a, b, dim, users, pos, neg, batch, n_users);
  LGC_LAUNCH_CHECK("zero_rows_kernel");
  return LGC_OK;
}

extern "C" int lgc_row_mask_batch(uint32_t* mask, const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch,
                                  int64_t n_users, int32_t set, lgc_stream_t stream) {
  LGC_REQUIRE(mask && users && pos && neg && batch > 0 && n_users > 0, "row_mask_batch: null pointer / empty batch");
  row_mask_batch_kernel<<<(unsigned)ceil_div(3 * batch, 256), 256, 0, (cudaStream_t)stream>>>(mask, users, pos, neg, batch,
                                                                                             n_users, set);
  LGC_LAUNCH_CHECK("row_mask_batch_kernel");
  return LGC_OK;
}

extern "C" int64_t lgc_bpr_det_workspace_bytes(int64_t batch, int32_t dim) {
  if (batch <= 0 || dim <= 0) return 0;
  return (int64_t)(align_up((size_t)3 * batch * sizeof(int32_t), 256) + 2 * align_up((size_t)3 * batch * dim * sizeof(float), 256));
}

extern "C" int lgc_bpr_fwd_bwd_det(const float* E, const float* X0, int64_t n_users, int64_t n_items, int32_t dim,
                                   const int64_t* users, const int64_t* pos, const int64_t* neg, int64_t batch, float eps,
                                   float grad_scale, float* loss_out, float* gE, float* gX0, float* scratch,
                                   void* workspace, int64_t workspace_bytes, lgc_stream_t stream_) {
  LGC_REQUIRE(E && X0 && users && pos && neg && loss_out && scratch && gE && gX0 && workspace, "bpr det: null pointer");
  LGC_REQUIRE(batch > 0 && batch <= 65536 && n_users > 0 && n_items > 0 && n_users + n_items < (1ll << 31),
              "bpr det: batch must be in [1, 65536] and the tables addressable with int32");
  LGC_REQUIRE(workspace_bytes >= lgc_bpr_det_workspace_bytes(batch, dim) && ((uintptr_t)workspace & 255) == 0,
              "bpr det: workspace too small or misaligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  char* base = reinterpret_cast<char*>(workspace);
  int32_t* ent_row = reinterpret_cast<int32_t*>(base);
  const size_t o1 = align_up((size_t)3 * batch * sizeof(int32_t), 256);
  const size_t o2 = o1 + align_up((size_t)3 * batch * dim * sizeof(float), 256);
  float* ent_gE = reinterpret_cast<float*>(base + o1);
  float* ent_gX = reinterpret_cast<float*>(base + o2);
  const int grid = bpr_grid(batch, dim);
  const int n_ent = (int)(3 * batch);
  const unsigned rgrid = (unsigned)ceil_div((int64_t)n_ent * (dim / 4), kBprThreads);
  switch (dim) {
#define LGC_BPR_DET(D)                                                                                                \
  case D:                                                                                                             \
    bpr_compact_kernel<D><<<grid, kBprThreads, 0, stream>>>(E, X0, n_users, users, pos, neg, batch, eps, grad_scale,   \
                                                            loss_out, scratch, ent_row, ent_gE, ent_gX);              \
    LGC_LAUNCH_CHECK("bpr_compact_kernel");                                                                           \
    bpr_reduce_kernel<D><<<rgrid, kBprThreads, 0, stream>>>(ent_row, ent_gE, ent_gX, n_ent, gE, gX0);                  \
    LGC_LAUNCH_CHECK("bpr_reduce_kernel");                                                                            \
    break;
    LGC_BPR_DET(32) LGC_BPR_DET(64) LGC_BPR_DET(128)
#undef LGC_BPR_DET
    default: LGC_FAIL(LGC_ERR_UNSUPPORTED, "bpr det: embedding dim %d not in {32,64,128}", dim);
  }
  return LGC_OK;
}
