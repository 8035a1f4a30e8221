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
