// capi.cu — library-wide C-ABI plumbing: version, thread-local error string, launch
// accounting, and CUDA IPC handle exchange for the multi-GPU peer-store path.
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

extern "C" int lgc_ipc_get_handle(void* dptr, uint8_t* handle64_host) {
  LGC_REQUIRE(dptr && handle64_host, "ipc: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  LGC_CUDA(cudaIpcGetMemHandle(&h, dptr));
  memcpy(handle64_host, &h, 64);
  return LGC_OK;
}

extern "C" int lgc_ipc_open_handle(const uint8_t* handle64_host, void** dptr_host) {
  LGC_REQUIRE(handle64_host && dptr_host, "ipc: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  LGC_CUDA(cudaIpcOpenMemHandle(dptr_host, h, cudaIpcMemLazyEnablePeerAccess));
  return LGC_OK;
}

extern "C" int lgc_ipc_close_handle(void* dptr) {
  LGC_REQUIRE(dptr, "ipc: null pointer");
  LGC_CUDA(cudaIpcCloseMemHandle(dptr));
  return LGC_OK;
}
