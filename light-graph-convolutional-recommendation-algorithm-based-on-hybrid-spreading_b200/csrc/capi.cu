// capi.cu — library-wide C-ABI plumbing: version, thread-local error string, launch
// accounting, and CUDA IPC handle exchange for the multi-GPU peer-store path.
#include <cuda.h>

#include "common.cuh"

#include <atomic>
#include <string.h>

namespace lgc {
static thread_local char g_err[kErrBufLen] = "";
static std::atomic<long long> g_launches{0};
char* err_buf() { return g_err; }
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace lgc

using namespace lgc;

extern "C" int lgc_abi_version(void) { return 1; }
extern "C" const char* lgc_last_error_string(void) { return err_buf(); }
extern "C" int64_t lgc_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void lgc_reset_launch_count(void) { g_launches.store(0); }

// An IPC blob is the 64-byte cudaIpcMemHandle_t of the containing allocation followed by the
// 8-byte offset of the pointer inside it (torch sub-allocates from cudaMalloc segments).
typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

extern "C" int lgc_ipc_get_handle(void* dptr, uint8_t* handle72_host) {
  LGC_REQUIRE(dptr && handle72_host, "ipc: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  static GetRangeFn get_range = nullptr;
  if (!get_range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      LGC_FAIL(LGC_ERR_CUDA, "ipc: cuMemGetAddressRange entry point not available");
    get_range = (GetRangeFn)fp;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (CUdeviceptr)dptr) != CUDA_SUCCESS)
    LGC_FAIL(LGC_ERR_CUDA, "ipc: cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  LGC_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  const uint64_t off = (uint64_t)((CUdeviceptr)dptr - base);
  memcpy(handle72_host, &h, 64);
  memcpy(handle72_host + 64, &off, 8);
  return LGC_OK;
}

extern "C" int lgc_ipc_open_handle(const uint8_t* handle72_host, void** dptr_host) {
  LGC_REQUIRE(handle72_host && dptr_host, "ipc: null pointer");
  cudaIpcMemHandle_t h;
  uint64_t off = 0;
  memcpy(&h, handle72_host, 64);
  memcpy(&off, handle72_host + 64, 8);
  void* base = nullptr;
  LGC_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr_host = (char*)base + off;
  return LGC_OK;
}

extern "C" int lgc_ipc_close_handle(void* base_dptr) {
  LGC_REQUIRE(base_dptr, "ipc: null pointer");
  LGC_CUDA(cudaIpcCloseMemHandle(base_dptr));
  return LGC_OK;
}
