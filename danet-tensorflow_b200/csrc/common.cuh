// Shared helpers for libdanet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/danet.h"

namespace danet {

void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define DANET_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      danet::set_error(__VA_ARGS__);              \
      return (code);                              \
    }                                             \
  } while (0)

#define DANET_CUDA(call)                                                        \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      danet::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                       cudaGetErrorString(e__));                                \
      return DANET_E_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define DANET_LAUNCH_CHECK() DANET_CUDA(cudaGetLastError())

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr float kEps = 1e-7f;   // default.json:17

}  // namespace danet
