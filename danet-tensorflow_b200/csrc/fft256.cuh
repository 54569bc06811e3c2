// 256-point complex FFT held by 16 cooperating threads (a half warp): 16 x 16
// Cooley-Tukey, each stage a 16-point FFT in registers, one padded shared-memory
// transpose in between.  Two real frames ride one complex transform (real part =
// frame A, imaginary part = frame B), which is how both the STFT (app/utils.py:117)
// and the inverse transform of utils.istft (app/utils.py:71) are evaluated here.
#pragma once
#include <cuda_runtime.h>

namespace danet {

constexpr int kFft = 256;
constexpr int kFftPad = 16 * 17;   // padded transpose tile, in float2

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 4-point DFT, forward uses W4 = -i, inverse +i
template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  float2 s0 = make_float2(a.x + c.x, a.y + c.y);
  float2 s1 = make_float2(a.x - c.x, a.y - c.y);
  float2 s2 = make_float2(b.x + d.x, b.y + d.y);
  float2 s3 = make_float2(b.x - d.x, b.y - d.y);
  // -i * s3 (forward) or +i * s3 (inverse)
  float2 r3 = INV ? make_float2(-s3.y, s3.x) : make_float2(s3.y, -s3.x);
  a = make_float2(s0.x + s2.x, s0.y + s2.y);
  b = make_float2(s1.x + r3.x, s1.y + r3.y);
  c = make_float2(s0.x - s2.x, s0.y - s2.y);
  d = make_float2(s1.x - r3.x, s1.y - r3.y);
}

// in-register 16-point DFT, natural order in and out.
// n = 4*n1 + n2, k = k1 + 4*k2:  X[k] = sum_n2 W16^(n2 k1) W4^(n2 k2) sum_n1 v[4 n1+n2] W4^(n1 k1)
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f;   // cos(pi/8)
  constexpr float s1 = 0.38268343236508977f;   // sin(pi/8)
  constexpr float c2 = 0.70710678118654752f;   // cos(pi/4)
  const float sg = INV ? 1.f : -1.f;
  // W16^m = cos(2 pi m/16) + sg * i sin(2 pi m/16)
  const float2 w1 = make_float2(c1, sg * s1), w2 = make_float2(c2, sg * c2),
               w3 = make_float2(s1, sg * c1), w4 = make_float2(0.f, sg),
               w6 = make_float2(-c2, sg * c2), w9 = make_float2(-c1, -sg * s1);
  // step 1: for each n2, DFT4 over n1 (elements n2, 4+n2, 8+n2, 12+n2) -> slot 4*k1+n2
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // step 2: twiddle W16^(n2*k1) on slot 4*k1 + n2
  v[5] = cmul(v[5], w1);  v[6] = cmul(v[6], w2);   v[7] = cmul(v[7], w3);
  v[9] = cmul(v[9], w2);  v[10] = cmul(v[10], w4); v[11] = cmul(v[11], w6);
  v[13] = cmul(v[13], w3); v[14] = cmul(v[14], w6); v[15] = cmul(v[15], w9);
  // step 3: for each k1, DFT4 over n2 (slots 4*k1 .. 4*k1+3) -> X[k1 + 4*k2] in slot 4*k1+k2
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // slot 4*k1 + k2 holds X[k1 + 4*k2]: transpose the 4x4 to natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      float2 t = v[4 * a + b];
      v[4 * a + b] = v[4 * b + a];
      v[4 * b + a] = t;
    }
}

// 256-point transform across a 16-thread group.  On entry thread j holds
// v[m] = z[16*m + j]; on exit thread j holds v[k2] = Z[j + 16*k2].
// tw[q] = exp(-/+ 2 pi i q / 256) (sign by INV), buf = this group's padded tile.
template <bool INV>
__device__ __forceinline__ void fft256_group(float2 (&v)[16], int j, const float2* __restrict__ tw,
                                             float2* buf, unsigned group_mask) {
  fft16<INV>(v);                                        // over m -> k1
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    float2 t = tw[(j * k1) & 255];
    if (INV) t.y = -t.y;
    buf[k1 * 17 + j] = cmul(v[k1], t);
  }
  __syncwarp(group_mask);
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) v[jj] = buf[j * 17 + jj];   // thread j now plays k1 = j
  __syncwarp(group_mask);
  fft16<INV>(v);                                        // over jj -> k2
}

}  // namespace danet
