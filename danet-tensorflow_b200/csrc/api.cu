// Library identity, error string, device check.
#include <stdarg.h>
#include <atomic>
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
  static std::atomic<int> cached{0};
  int n = cached.load(std::memory_order_relaxed);
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached.store(n, std::memory_order_relaxed);
    else
      return 148;
  }
  return n;
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

// host helper of the checkpoint writer / reader (tf_bundle.py): CRC32C (Castagnoli, reflected 0x82F63B78) as
// TensorFlow's tensor bundles store it per tensor and per table block; slicing-by-8, ~1 GB/s on one core
namespace {
struct Crc32cTables {
  unsigned int t[8][256];
  Crc32cTables() {
    for (unsigned int i = 0; i < 256; ++i) {
      unsigned int c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (unsigned int i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
}  // namespace

extern "C" unsigned int danet_crc32c(const void* data, size_t n, unsigned int crc) {
  static const Crc32cTables tables;                    // function-local static: initialised once, thread-safe
  const unsigned int (*tab)[256] = tables.t;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  crc = ~crc;
  while (n >= 8) {
    const unsigned int lo = crc ^ (p[0] | (p[1] << 8) | (p[2] << 16) | ((unsigned int)p[3] << 24));
    const unsigned int hi = p[4] | (p[5] << 8) | (p[6] << 16) | ((unsigned int)p[7] << 24);
    crc = tab[7][lo & 0xFF] ^ tab[6][(lo >> 8) & 0xFF] ^ tab[5][(lo >> 16) & 0xFF] ^ tab[4][lo >> 24] ^
          tab[3][hi & 0xFF] ^ tab[2][(hi >> 8) & 0xFF] ^ tab[1][(hi >> 16) & 0xFF] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = tab[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

// Clears a small buffer (the completion flags of danet_gemm_split_pipelined) with a one-block KERNEL: inside a captured
// graph a cudaMemsetAsync becomes a memset node, and sixteen of those per step measured 30 us slower than sixteen tiny
// kernels (2.424 vs 2.394 ms per step, profiles/r02_ab_switches.txt).
namespace danet {
__global__ void zero_words_kernel(unsigned int* p, size_t n) {
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
}
}  // namespace danet

extern "C" int danet_zero_async(void* ptr, size_t bytes, void* stream) {
  DANET_REQUIRE(ptr || bytes == 0, DANET_E_ARG, "zero_async: null pointer");
  if (bytes == 0) return DANET_OK;
  DANET_REQUIRE(bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 3) == 0, DANET_E_ALIGN, "zero_async: 4-byte granularity");
  if (bytes <= 64 * 1024) {
    danet::zero_words_kernel<<<1, 256, 0, danet::as_stream(stream)>>>(reinterpret_cast<unsigned int*>(ptr), bytes / 4);
    DANET_LAUNCH_CHECK();
  } else {
    DANET_CUDA(cudaMemsetAsync(ptr, 0, bytes, danet::as_stream(stream)));
  }
  return DANET_OK;
}
