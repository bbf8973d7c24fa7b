// Error state, pointer validation and launch accounting shared by every entry point.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

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

// Per-kernel device timing (bench.py's roofline leg): while enabled, an event is recorded on the
// launching stream after every kernel; launch i's duration is event[i] - event[i-1] (stream order), so a
// launch-bound gap in front of a kernel is charged to it.
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> events;      // events[0] = start marker
  std::vector<const char*> names;       // names[i] belongs to events[i + 1]
  size_t used = 0;
};
static thread_local Profiler g_prof;
static constexpr size_t kMaxProfiledLaunches = 1 << 16;

void count_launch(const char* name, cudaStream_t st) {
  ++g_launches;
  if (!g_prof.on || g_prof.used + 1 >= kMaxProfiledLaunches) return;
  if (g_prof.used + 1 >= g_prof.events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { g_prof.on = false; return; }
    g_prof.events.push_back(e);
  }
  cudaEventRecord(g_prof.events[g_prof.used + 1], st);
  g_prof.names.push_back(name);
  ++g_prof.used;
}

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

int bfb200_profile_begin(void* stream) {
  using namespace bfb;
  if (g_prof.events.empty()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) {
      set_error("cudaEventCreate failed");
      return BFB200_ERR_CUDA;
    }
    g_prof.events.push_back(e);
  }
  g_prof.names.clear();
  g_prof.used = 0;
  cudaEventRecord(g_prof.events[0], (cudaStream_t)stream);
  g_prof.on = true;
  return BFB200_OK;
}

int bfb200_profile_end(char* json_out, size_t json_bytes) {
  using namespace bfb;
  g_prof.on = false;
  if (json_out == nullptr || json_bytes == 0) return BFB200_OK;
  std::map<std::string, std::pair<int64_t, double>> agg;  // name -> (launches, total ms)
  if (g_prof.used > 0) {
    if (cudaEventSynchronize(g_prof.events[g_prof.used]) != cudaSuccess) {
      set_error("cudaEventSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
      return BFB200_ERR_CUDA;
    }
    for (size_t i = 0; i < g_prof.used; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_prof.events[i], g_prof.events[i + 1]);
      auto& a = agg[g_prof.names[i]];
      a.first += 1;
      a.second += ms;
    }
  }
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             (long long)kv.second.first, kv.second.second);
    js += buf;
    first = false;
  }
  js += "}";
  if (js.size() + 1 > json_bytes) {
    set_error("profile buffer too small: need %zu bytes", js.size() + 1);
    return BFB200_ERR_WORKSPACE;
  }
  memcpy(json_out, js.c_str(), js.size() + 1);
  return BFB200_OK;
}

}  // extern "C"
