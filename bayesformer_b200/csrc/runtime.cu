// Error state, pointer validation and launch accounting shared by every entry point.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace bfb {

static thread_local char g_error[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch() { ++g_launches; }

int check_device_ptr(const void* p, const char* name) {
  if (p == nullptr) {
    set_error("%s is NULL", name);
    return BFB200_ERR_INVALID_ARG;
  }
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cudaPointerGetAttributes failed (%s) — no CUDA device? this library has no CPU path",
              name, cudaGetErrorString(e));
    return BFB200_ERR_NOT_DEVICE_PTR;
  }
  if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) {
    set_error("%s is not a CUDA device pointer (this library has no CPU path)", name);
    return BFB200_ERR_NOT_DEVICE_PTR;
  }
  return BFB200_OK;
}

}  // namespace bfb

extern "C" {

int bfb200_abi_version(void) { return BFB200_ABI_VERSION; }
const char* bfb200_last_error(void) { return bfb::g_error; }
int64_t bfb200_launch_count(void) { return bfb::g_launches; }
void bfb200_reset_launch_count(void) { bfb::g_launches = 0; }

}  // extern "C"
