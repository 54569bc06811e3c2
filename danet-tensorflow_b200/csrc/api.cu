// Library identity, error string, device check.
#include <stdarg.h>
#include "common.cuh"

namespace danet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace danet

extern "C" int danet_version(void) { return 100; }

extern "C" const char* danet_last_error_string(void) { return danet::g_err; }

extern "C" int danet_check_device(void) {
  int dev = 0, major = 0;
  DANET_CUDA(cudaGetDevice(&dev));
  DANET_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DANET_REQUIRE(major == 10, DANET_E_ARCH, "device compute capability %d.x, need 10.x (sm_100a)", major);
  return DANET_OK;
}

// profiling aid: one thread stores %globaltimer (ns) -- lets a host script reconstruct the timeline of a
// multi-stream CUDA graph, where CUDA events cannot be read back
__global__ void timestamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

extern "C" int danet_timestamp(unsigned long long* slot, void* stream) {
  DANET_REQUIRE(slot, DANET_E_ARG, "timestamp: null pointer");
  timestamp_kernel<<<1, 1, 0, danet::as_stream(stream)>>>(slot);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
