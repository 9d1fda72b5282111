// Error reporting, launch accounting.
#include <string>
#include <atomic>
#include "common.cuh"
#include "../../include/fs2_b200.h"

#include <stdlib.h>

static int pdl_from_env() {
  const char* e = getenv("FS2_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}
int g_fs2_pdl = pdl_from_env();

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

void fs2_set_error(const char* msg) { g_err = msg ? msg : ""; }

int fs2_check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    fs2_set_error(cudaGetErrorString(e));
    return FS2_ERR_CUDA;
  }
  return FS2_OK;
}

extern "C" const char* fs2_last_error(void) { return g_err.c_str(); }
extern "C" int fs2_abi_version(void) { return FS2_ABI_VERSION; }
extern "C" long long fs2_launch_count(void) { return g_launches.load(); }
