// a16 tail: bias-gradient column sums and the fused clip-by-value + Adam update
// (main.py:359-363; app/ozers.py:15-18 -> tf.train.AdamOptimizer defaults).
#include "common.cuh"

namespace danet {

constexpr int kColParts = 128;

__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, long long ld, long long rows, int n, float* __restrict__ part) {
  // grid (ceil(n/256), kColParts): coalesced along n, rows strided over the parts
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= n) return;
  const long long chunk = (rows + kColParts - 1) / kColParts;
  const long long lo = blockIdx.y * chunk, hi = min(rows, lo + chunk);
  // eight loads in flight per thread (one dependent load per iteration left the pass at 1.5 TB/s)
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  long long r = lo;
  for (; r + 8 <= hi; r += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += __ldg(x + (r + j) * ld + col);
  }
  for (; r < hi; ++r) a[0] += __ldg(x + r * ld + col);
  part[(size_t)blockIdx.y * n + col] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int n, float* __restrict__ out, int accumulate) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= n) return;
  float acc = accumulate ? out[col] : 0.f;
  for (int p = 0; p < kColParts; ++p) acc += part[(size_t)p * n + col];
  out[col] = acc;
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, long long n, float grad_scale, float clip, float lr_t, float b1, float b2,
                 float eps) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float g = grad[i] * grad_scale;
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);      // tf.clip_by_value (main.py:359-362)
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    param[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void __launch_bounds__(256)
clip_sgd_kernel(float* __restrict__ param, const float* __restrict__ grad, long long n, float grad_scale, float clip,
                float lr) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float g = grad[i] * grad_scale;
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    param[i] -= lr * g;
  }
}

}  // namespace danet

using namespace danet;

extern "C" int danet_clip_sgd(float* param, const float* grad, long long n, float grad_scale, float clip, float lr,
                              void* stream) {
  DANET_REQUIRE(param && grad, DANET_E_ARG, "clip_sgd: null pointer");
  DANET_REQUIRE(n >= 0, DANET_E_SHAPE, "clip_sgd: n %lld", n);
  if (n == 0) return DANET_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  clip_sgd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, n, grad_scale, clip, lr);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" size_t danet_colsum_workspace_bytes(int n) { return (size_t)(n > 0 ? n : 1) * kColParts * sizeof(float); }

extern "C" int danet_colsum(const float* x, long long ld, long long rows, int n, float* out, int accumulate,
                            void* workspace, size_t workspace_bytes, void* stream) {
  DANET_REQUIRE(x && out && workspace, DANET_E_ARG, "colsum: null pointer");
  DANET_REQUIRE(rows >= 0 && n >= 1 && ld >= n, DANET_E_SHAPE, "colsum: rows %lld n %d ld %lld", rows, n, ld);
  DANET_REQUIRE(workspace_bytes >= danet_colsum_workspace_bytes(n), DANET_E_WORKSPACE, "colsum: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  dim3 g((n + 255) / 256, kColParts);
  colsum_partial_kernel<<<g, 256, 0, as_stream(stream)>>>(x, ld, rows, n, part);
  DANET_LAUNCH_CHECK();
  colsum_final_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(part, n, out, accumulate);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_clip_adam(float* param, const float* grad, float* m, float* v, long long n, float grad_scale,
                               float clip, float lr, float beta1, float beta2, float eps, int step, void* stream) {
  DANET_REQUIRE(param && grad && m && v, DANET_E_ARG, "clip_adam: null pointer");
  DANET_REQUIRE(n >= 0 && step >= 1, DANET_E_SHAPE, "clip_adam: n %lld step %d", n, step);
  if (n == 0) return DANET_OK;
  const double lr_t = (double)lr * sqrt(1. - pow((double)beta2, step)) / (1. - pow((double)beta1, step));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  clip_adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, m, v, n, grad_scale, clip, (float)lr_t,
                                                                    beta1, beta2, eps);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
