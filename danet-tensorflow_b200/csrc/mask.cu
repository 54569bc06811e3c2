// K4a: dot-product mask x mixture.
//   logits[b,tf,c] = <V[b,tf,:], A[b,c,:]>; softmax over C (app/modules.py:585-603) or
//   sigmoid (:556-574); sep_pwr = |mix| * mask; re-phased output mask * mix
//   (main.py:281-284).  One pass over the embedding, HBM-bound.
#include "common.cuh"

namespace danet {

constexpr int kMaxC = 8;
constexpr int kMaxE = 128;

template <bool VEC4>
__global__ void __launch_bounds__(256)
mask_cmul_kernel(const float* __restrict__ embed, const float* __restrict__ attractors,
                 const float2* __restrict__ mix, const float* __restrict__ mix_pwr,
                 float* __restrict__ sep_pwr,
                 float2* __restrict__ sep, float* __restrict__ masks, int C, long long TF, int E,
                 int kind) {
  __shared__ float s_att[kMaxC * kMaxE];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C * E; i += blockDim.x) s_att[i] = attractors[(size_t)b * C * E + i];
  __syncthreads();
  const float* Vb = embed + (size_t)b * TF * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < TF;
       i += (long long)gridDim.x * blockDim.x) {
    float logit[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) logit[c] = 0.f;
    const float* v = Vb + (size_t)i * E;
    if (VEC4) {
      const float4* v4 = reinterpret_cast<const float4*>(v);
      for (int e4 = 0; e4 < E / 4; ++e4) {
        float4 x = __ldg(v4 + e4);
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
          if (c < C) {
            const float* a = s_att + c * E + 4 * e4;
            logit[c] = fmaf(x.x, a[0], logit[c]);
            logit[c] = fmaf(x.y, a[1], logit[c]);
            logit[c] = fmaf(x.z, a[2], logit[c]);
            logit[c] = fmaf(x.w, a[3], logit[c]);
          }
      }
    } else {
      for (int e = 0; e < E; ++e) {
        float x = __ldg(v + e);
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
          if (c < C) logit[c] = fmaf(x, s_att[c * E + e], logit[c]);
      }
    }
    float m[kMaxC];
    if (kind == 0) {
      float mx = logit[0];
#pragma unroll
      for (int c = 1; c < kMaxC; ++c)
        if (c < C) mx = fmaxf(mx, logit[c]);
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) {
          m[c] = expf(logit[c] - mx);
          den += m[c];
        }
      float inv = 1.f / den;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) m[c] *= inv;
    } else {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < C) m[c] = sigmoidf_(logit[c]);
    }
    float2 z = make_float2(0.f, 0.f);
    float p;
    if (mix) {
      z = __ldg(mix + (size_t)b * TF + i);
      p = mix_pwr ? __ldg(mix_pwr + (size_t)b * TF + i) : sqrtf(z.x * z.x + z.y * z.y);
    } else {
      p = __ldg(mix_pwr + (size_t)b * TF + i);
    }
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        size_t o = ((size_t)b * C + c) * TF + i;
        if (sep_pwr) sep_pwr[o] = p * m[c];
        if (sep) sep[o] = make_float2(z.x * m[c], z.y * m[c]);
        if (masks) masks[((size_t)b * TF + i) * C + c] = m[c];
      }
  }
}

}  // namespace danet

using namespace danet;

extern "C" int danet_mask_cmul_fwd(const float* embed, const float* attractors, const float* mix_c64,
                                   const float* mix_pwr, float* sep_pwr, float* sep_c64, float* masks, int B, int C, int TF,
                                   int E, int kind, void* stream) {
  DANET_REQUIRE(embed && attractors && (mix_c64 || mix_pwr), DANET_E_ARG, "mask_cmul: null pointer");
  DANET_REQUIRE(mix_c64 || !sep_c64, DANET_E_ARG, "mask_cmul: complex output needs the complex mixture");
  DANET_REQUIRE(B >= 0 && TF >= 0 && C >= 1 && C <= kMaxC && E >= 1 && E <= kMaxE, DANET_E_SHAPE,
                "mask_cmul: B %d C %d (<=%d) TF %d E %d (<=%d)", B, C, kMaxC, TF, E, kMaxE);
  DANET_REQUIRE(kind == 0 || kind == 1, DANET_E_ARG, "mask_cmul: kind %d", kind);
  DANET_REQUIRE(aligned8(mix_c64) && aligned8(sep_c64), DANET_E_ALIGN, "mask_cmul: complex buffers must be 8-byte aligned");
  if (B == 0 || TF == 0) return DANET_OK;
  DANET_REQUIRE(B <= 65535, DANET_E_SHAPE, "mask_cmul: B %d > 65535", B);
  int gx = (TF + 255) / 256;
  int cap = max(1, (num_sms() * 8 + B - 1) / B);
  if (gx > cap) gx = cap;
  dim3 grid(gx, B);
  const bool vec = (E % 4 == 0) && aligned16(embed);
  auto mixp = reinterpret_cast<const float2*>(mix_c64);
  auto sepp = reinterpret_cast<float2*>(sep_c64);
  if (vec)
    mask_cmul_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(embed, attractors, mixp, mix_pwr, sep_pwr, sepp, masks, C, TF, E, kind);
  else
    mask_cmul_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(embed, attractors, mixp, mix_pwr, sep_pwr, sepp, masks, C, TF, E, kind);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
