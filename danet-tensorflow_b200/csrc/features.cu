// Mixture features (main.py:233-240) and per-utterance mean centring
// (app/modules.py:209-210, 244-245).  Pure streaming kernels: float2/float4
// coalesced accesses, deterministic two-stage reduction for the mean.
#include "common.cuh"

namespace danet {

template <int C_MAX>
__global__ void __launch_bounds__(256)
mix_features_kernel(const float2* __restrict__ src, int C, long long TF,
                    float2* __restrict__ mix, float* __restrict__ src_pwr,
                    float* __restrict__ mix_pwr, float* __restrict__ logmag) {
  const int b = blockIdx.y;
  const float2* s = src + (size_t)b * C * TF;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < TF;
       i += (long long)gridDim.x * blockDim.x) {
    float2 acc = make_float2(0.f, 0.f);
    for (int c = 0; c < C; ++c) {
      float2 v = __ldg(s + (size_t)c * TF + i);
      acc.x += v.x;
      acc.y += v.y;
      if (src_pwr) src_pwr[((size_t)b * C + c) * TF + i] = sqrtf(v.x * v.x + v.y * v.y);
    }
    float p = sqrtf(acc.x * acc.x + acc.y * acc.y);
    size_t o = (size_t)b * TF + i;
    if (mix) mix[o] = acc;
    if (mix_pwr) mix_pwr[o] = p;
    if (logmag) logmag[o] = log1pf(p);
  }
}

constexpr int kCenterParts = 64;

__global__ void __launch_bounds__(256)
center_partial_kernel(const float* __restrict__ x, long long n_per, float* __restrict__ part) {
  __shared__ float s_red[8];
  const int b = blockIdx.y, p = blockIdx.x;
  const float* xb = x + (size_t)b * n_per;
  const long long chunk = (n_per + kCenterParts - 1) / kCenterParts;
  const long long lo = p * chunk, hi = min(n_per, lo + chunk);
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) acc += __ldg(xb + i);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    part[b * kCenterParts + p] = t;
  }
}

__global__ void __launch_bounds__(256)
center_apply_kernel(const float* __restrict__ x, long long n_per, const float* __restrict__ part,
                    float* __restrict__ y) {
  __shared__ float s_mean;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int p = 0; p < kCenterParts; ++p) t += part[b * kCenterParts + p];
    s_mean = t / (float)n_per;
  }
  __syncthreads();
  const float m = s_mean;
  const float* xb = x + (size_t)b * n_per;
  float* yb = y + (size_t)b * n_per;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_per; i += (long long)gridDim.x * 256)
    yb[i] = xb[i] - m;
}

__global__ void mean_final_kernel(const float* __restrict__ part, int B, long long n_per, float* __restrict__ mean) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float t = 0.f;
  for (int p = 0; p < kCenterParts; ++p) t += part[b * kCenterParts + p];     // same order as center_apply_kernel
  mean[b] = t / (float)n_per;
}

__global__ void __launch_bounds__(256)
leaky_relu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float leak) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = x[i];
    y[i] = leak == 0.f ? fmaxf(v, 0.f) : fmaxf(v * leak, v);     // app/ops.py:103-107
  }
}

__global__ void __launch_bounds__(256)
leaky_relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long n,
                      float leak) {
  // the activation is monotone, so the sign of the OUTPUT tells the branch
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    dx[i] = y[i] > 0.f ? dy[i] : dy[i] * leak;
}

}  // namespace danet

using namespace danet;

extern "C" int danet_leaky_relu_bwd(const float* y, const float* dy, float* dx, long long n, float leak, void* stream) {
  DANET_REQUIRE(y && dy && dx, DANET_E_ARG, "leaky_relu_bwd: null pointer");
  DANET_REQUIRE(n >= 0, DANET_E_SHAPE, "leaky_relu_bwd: n %lld", n);
  if (n == 0) return DANET_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  leaky_relu_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(y, dy, dx, n, leak);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_leaky_relu_fwd(const float* x, float* y, long long n, float leak, void* stream) {
  DANET_REQUIRE(x && y, DANET_E_ARG, "leaky_relu: null pointer");
  DANET_REQUIRE(n >= 0, DANET_E_SHAPE, "leaky_relu: n %lld", n);
  if (n == 0) return DANET_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  leaky_relu_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, y, n, leak);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_mix_features_fwd(const float* src_c64, int B, int C, int TF, float* mix_c64,
                                      float* src_pwr, float* mix_pwr, float* logmag, void* stream) {
  DANET_REQUIRE(src_c64, DANET_E_ARG, "mix_features: null src");
  DANET_REQUIRE(B >= 0 && C >= 1 && TF >= 0, DANET_E_SHAPE, "mix_features: B %d C %d TF %d", B, C, TF);
  DANET_REQUIRE(aligned8(src_c64) && aligned8(mix_c64), DANET_E_ALIGN, "mix_features: complex buffers must be 8-byte aligned");
  if (B == 0 || TF == 0) return DANET_OK;
  DANET_REQUIRE(B <= 65535, DANET_E_SHAPE, "mix_features: B %d > 65535", B);
  int gx = (TF + 255) / 256;
  int cap = max(1, (num_sms() * 8 + B - 1) / B);
  if (gx > cap) gx = cap;
  mix_features_kernel<8><<<dim3(gx, B), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(src_c64), C, TF, reinterpret_cast<float2*>(mix_c64), src_pwr,
      mix_pwr, logmag);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" size_t danet_center_workspace_bytes(int B) {
  return (size_t)(B > 0 ? B : 0) * kCenterParts * sizeof(float);
}

extern "C" int danet_center_fwd(const float* x, int B, long long n_per, float* y, float* workspace,
                                void* stream) {
  DANET_REQUIRE(x && y && workspace, DANET_E_ARG, "center: null pointer");
  DANET_REQUIRE(B >= 0 && n_per >= 1, DANET_E_SHAPE, "center: B %d n_per %lld", B, n_per);
  if (B == 0) return DANET_OK;
  DANET_REQUIRE(B <= 65535, DANET_E_SHAPE, "center: B %d > 65535", B);
  center_partial_kernel<<<dim3(kCenterParts, B), 256, 0, as_stream(stream)>>>(x, n_per, workspace);
  DANET_LAUNCH_CHECK();
  long long gx = (n_per + 255) / 256;
  long long cap = max(1, (num_sms() * 8 + B - 1) / B);
  if (gx > cap) gx = cap;
  center_apply_kernel<<<dim3((unsigned)gx, B), 256, 0, as_stream(stream)>>>(x, n_per, workspace, y);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_mean_fwd(const float* x, int B, long long n_per, float* mean, float* workspace, void* stream) {
  DANET_REQUIRE(x && mean && workspace, DANET_E_ARG, "mean: null pointer");
  DANET_REQUIRE(B >= 0 && B <= 65535 && n_per >= 1, DANET_E_SHAPE, "mean: B %d n_per %lld", B, n_per);
  if (B == 0) return DANET_OK;
  center_partial_kernel<<<dim3(kCenterParts, B), 256, 0, as_stream(stream)>>>(x, n_per, workspace);
  DANET_LAUNCH_CHECK();
  mean_final_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(workspace, B, n_per, mean);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
