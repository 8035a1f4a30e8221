// probe.cu — measurement probes used by bench.py to obtain roofline DENOMINATORS on the box it runs on.
//
// lgc_probe_gather: the access pattern that bounds the propagation SpMM (spmm.cu) with everything else removed —
// uniformly random reads of whole embedding rows (dim fp32 = 256 B at dim 64, DIM/4 lanes x LDG.128 per row) from a
// table that is L2-resident on B200 (42 MB at the ML-20M shape, L2 = 126 MB), at full occupancy with 8 independent
// gathers in flight per lane, no metadata stream, no output but one float4 per warp.  Its throughput is the
// "L2 gather peak" against which bench.py states the SpMM's gather fraction: a physical, <= 1 bound for a kernel
// whose 8 GB of per-layer row gathers are served by L2, not HBM.
#include "common.cuh"

namespace lgc {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int DIM>
__global__ void __launch_bounds__(256)
probe_gather_kernel(const float* __restrict__ table, uint32_t n_rows, long long gathers_per_group, uint32_t seed,
                    float* __restrict__ out) {
  constexpr int LPR = DIM / 4;
  constexpr int UN = 8;
  const int lane = threadIdx.x & 31;
  const int li = lane % LPR;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;   // one group gathers one row per step
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t state = hash32(seed ^ (uint32_t)group * 0x9E3779B9u);
  for (long long g = 0; g < gathers_per_group; g += UN) {
    float4 x[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      state = state * 1664525u + 1013904223u;
      const uint32_t row = __umulhi(hash32(state), n_rows);
      x[u] = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * DIM + li * 4));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
  }
  // one store per thread keeps the loads alive; negligible next to the gathers
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  out[t] = acc.x + acc.y + acc.z + acc.w;
}


// ---- lgc_probe_row_store: the exchange half of the row-partitioned propagation with the gather removed — every
// finished 256-byte row (dim 64 fp32) is copied from a local table to `dst`, which may be a peer or an NVSwitch
// multicast address.  The modes differ only in how the row leaves the SM:
//   0  16 lanes x st.global.v4.f32 per row (what spmm_layer_kernel does), one warp per row
//   1  32 lanes x st.global.v4.f32 = two rows per warp instruction
//   2  row staged in shared memory, one cp.async.bulk (TMA) store of 256 B per row
//   3  32 consecutive rows staged in shared memory, one cp.async.bulk store of 8 KB (row_list ignored)
__device__ __forceinline__ void bulk_store(void* dst, const void* smem_src, uint32_t bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256)
probe_row_store_kernel(const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ row_list,
                       long long n_rows) {
  constexpr int DIM = 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (MODE == 0) {
    const long long slot = (long long)blockIdx.x * 8 + warp;
    if (slot >= n_rows) return;
    const long long row = row_list ? row_list[slot] : slot;
    if (lane < 16) {
      const size_t off = (size_t)row * DIM + lane * 4;
      *reinterpret_cast<float4*>(dst + off) = __ldg(reinterpret_cast<const float4*>(src + off));
    }
  } else if (MODE == 1) {
    const long long slot = ((long long)blockIdx.x * 8 + warp) * 2 + (lane >> 4);
    if (slot >= n_rows) return;
    const long long row = row_list ? row_list[slot] : slot;
    const size_t off = (size_t)row * DIM + (lane & 15) * 4;
    *reinterpret_cast<float4*>(dst + off) = __ldg(reinterpret_cast<const float4*>(src + off));
  } else if (MODE == 2) {
    __shared__ __align__(128) float stage[8][DIM];
    const long long slot = (long long)blockIdx.x * 8 + warp;
    if (slot >= n_rows) return;
    const long long row = row_list ? row_list[slot] : slot;
    const size_t off = (size_t)row * DIM;
    if (lane < 16) *reinterpret_cast<float4*>(&stage[warp][lane * 4]) = __ldg(reinterpret_cast<const float4*>(src + off + lane * 4));
    __syncwarp();
    if (lane == 0) {
      bulk_store(dst + off, stage[warp], DIM * 4);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    __shared__ __align__(128) float stage[32 * DIM];
    const long long row0 = (long long)blockIdx.x * 32;
    if (row0 >= n_rows) return;
    const int rows = (int)min(32ll, n_rows - row0);
    const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)row0 * DIM);
    for (int i = threadIdx.x; i < rows * 16; i += 256) reinterpret_cast<float4*>(stage)[i] = __ldg(s4 + i);
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(dst + (size_t)row0 * DIM, stage, (uint32_t)rows * DIM * 4);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

// Persistent variant: n_ctas CTAs of 1024 threads walk the row list `passes` times, two rows per warp instruction and
// four instructions in flight.  exclusive = launched with enough dynamic shared memory that nothing else fits on the SM:
// the probe for "does a saturated NVLink store stream hold up the OTHER memory traffic of the SMs it is issued from?"
__global__ void __launch_bounds__(1024)
probe_row_store_persistent_kernel(const float* __restrict__ src, float* __restrict__ dst, const int32_t* __restrict__ row_list,
                                  long long n_rows, int passes) {
  constexpr int DIM = 64;
  const int lane = threadIdx.x & 31;
  const long long n_warps = (long long)gridDim.x * 32, gw = (long long)blockIdx.x * 32 + (threadIdx.x >> 5);
  const long long pairs = (n_rows + 1) / 2;
  for (int pass = 0; pass < passes; ++pass) {
    for (long long p0 = gw; p0 < pairs; p0 += 4 * n_warps) {
      float4 v[4];
      size_t off[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long slot = (p0 + u * n_warps) * 2 + (lane >> 4);
        ok[u] = p0 + u * n_warps < pairs && slot < n_rows;
        const long long row = ok[u] ? (row_list ? row_list[slot] : slot) : 0;
        off[u] = (size_t)row * DIM + (lane & 15) * 4;
        if (ok[u]) v[u] = __ldg(reinterpret_cast<const float4*>(src + off[u]));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) *reinterpret_cast<float4*>(dst + off[u]) = v[u];
    }
  }
}

}  // namespace lgc

using namespace lgc;

extern "C" int64_t lgc_probe_gather_threads(void) { return (int64_t)num_sms() * 8 * 256; }

extern "C" int lgc_probe_gather(const float* table, int64_t n_rows, int32_t dim, int64_t n_gathers, uint32_t seed,
                                float* out, int64_t* gathers_done_host, lgc_stream_t stream) {
  LGC_REQUIRE(table && out && gathers_done_host, "probe_gather: null pointer");
  LGC_REQUIRE(n_rows > 0 && n_rows < (1ll << 32) && n_gathers > 0, "probe_gather: bad extents");
  LGC_REQUIRE(((uintptr_t)table & 15) == 0, "probe_gather: table must be 16-byte aligned");
  const int grid = num_sms() * 8;     // 8 CTAs x 256 threads = 64 warps per SM: full occupancy at 32 registers
  const int lpr = dim / 4;
  LGC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "probe_gather: dim must be 32, 64 or 128");
  const long long groups = (long long)grid * 256 / lpr;
  long long per_group = (n_gathers + groups - 1) / groups;
  per_group = (per_group + 7) / 8 * 8;
  *gathers_done_host = per_group * groups;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dim) {
    case 32: probe_gather_kernel<32><<<grid, 256, 0, st>>>(table, (uint32_t)n_rows, per_group, seed, out); break;
    case 64: probe_gather_kernel<64><<<grid, 256, 0, st>>>(table, (uint32_t)n_rows, per_group, seed, out); break;
    default: probe_gather_kernel<128><<<grid, 256, 0, st>>>(table, (uint32_t)n_rows, per_group, seed, out); break;
  }
  LGC_LAUNCH_CHECK("probe_gather_kernel");
  return LGC_OK;
}

extern "C" int lgc_probe_row_store(const float* src, float* dst, const int32_t* row_list, int64_t n_rows, int32_t mode,
                                   lgc_stream_t stream) {
  LGC_REQUIRE(src && dst && n_rows > 0, "probe_row_store: null pointer / no rows");
  LGC_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 127) == 0, "probe_row_store: tables must be 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g8 = (unsigned)((n_rows + 7) / 8);
  switch (mode) {
    case 0: probe_row_store_kernel<0><<<g8, 256, 0, st>>>(src, dst, row_list, n_rows); break;
    case 1: probe_row_store_kernel<1><<<(g8 + 1) / 2, 256, 0, st>>>(src, dst, row_list, n_rows); break;
    case 2: probe_row_store_kernel<2><<<g8, 256, 0, st>>>(src, dst, row_list, n_rows); break;
    case 3: probe_row_store_kernel<3><<<(unsigned)((n_rows + 31) / 32), 256, 0, st>>>(src, dst, nullptr, n_rows); break;
    default: LGC_REQUIRE(false, "probe_row_store: mode must be 0..3");
  }
  LGC_LAUNCH_CHECK("probe_row_store_kernel");
  return LGC_OK;
}

extern "C" int lgc_probe_row_store_persistent(const float* src, float* dst, const int32_t* row_list, int64_t n_rows,
                                              int32_t n_ctas, int32_t passes, int32_t exclusive, lgc_stream_t stream) {
  LGC_REQUIRE(src && dst && n_rows > 0 && n_ctas > 0 && passes > 0, "probe_row_store_persistent: bad arguments");
  LGC_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "probe_row_store_persistent: tables must be 16-byte aligned");
  constexpr int kExclusiveSmem = 160 * 1024;   // more than half an SM's shared memory: one CTA per SM, nothing beside it
  static DeviceOnce once;
  if (once.need()) {
    LGC_CUDA(cudaFuncSetAttribute(probe_row_store_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExclusiveSmem));
    once.mark();
  }
  probe_row_store_persistent_kernel<<<n_ctas, 1024, exclusive ? kExclusiveSmem : 0, (cudaStream_t)stream>>>(src, dst, row_list,
                                                                                                           n_rows, passes);
  LGC_LAUNCH_CHECK("probe_row_store_persistent_kernel");
  return LGC_OK;
}
