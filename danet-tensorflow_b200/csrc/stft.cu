// K1: batched STFT (+ log-magnitude) and the reference-semantics iSTFT.
//   STFT : scipy.signal.stft as called at app/utils.py:117-122 (sqrt-hann 256, hop 64,
//          zero boundary, padded tail, 1/sum(w) scaling), transposed to [T,129].
//   iSTFT: utils.istft (app/utils.py:53-75): frames 0..T-5, irfft * w overlap-add,
//          divide by sum(w^2) where non-zero, length 64*T.
// Both are HBM-bound; each block stages its waveform chunk / spectra once, runs
// sixteen 256-point complex FFTs (two real frames each) in registers + one
// shared-memory transpose, and writes coalesced rows.
#include <math.h>
#include "common.cuh"
#include "fft256.cuh"

namespace danet {

constexpr int kFramesPerBlock = 32;                // 16 groups x 2 frames
constexpr int kHop = DANET_FFT_STRIDE;
constexpr int kBins = DANET_FEATURE;

static double window_sum() {
  // sum of float32(sqrt(hann_sym(256))) == sum of float32(sin(pi k / 255))
  double s = 0.;
  for (int k = 0; k < kFft; ++k) s += (double)(float)sin(M_PI * k / 255.);
  return s;
}

__device__ __forceinline__ float window_at(int k) { return sinpif((float)k * (1.f / 255.f)); }

__global__ void __launch_bounds__(256)
stft_kernel(const float* __restrict__ wav, int n_samples, int T, float inv_wsum,
            float2* __restrict__ spec, float* __restrict__ logmag) {
  __shared__ float s_wav[kHop * (kFramesPerBlock - 1) + kFft];   // 2240 samples
  __shared__ float s_win[kFft];
  __shared__ float2 s_tw[kFft];
  __shared__ float2 s_buf[16][kFftPad];

  const int tid = threadIdx.x;
  const int sig = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerBlock;
  const float* w_in = wav + (size_t)sig * n_samples;

  {
    float s, c;
    sincospif(-(float)tid * (2.f / 256.f), &s, &c);
    s_tw[tid] = make_float2(c, s);
    s_win[tid] = window_at(tid);
  }
  const long long first = (long long)kHop * t0 - kFft / 2;      // sample index of s_wav[0]
  for (int i = tid; i < kHop * (kFramesPerBlock - 1) + kFft; i += 256) {
    long long g = first + i;
    s_wav[i] = (g >= 0 && g < n_samples) ? __ldg(w_in + g) : 0.f;
  }
  __syncthreads();

  const int g = tid >> 4, j = tid & 15;
  const unsigned gmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
  const int tA = t0 + 2 * g, tB = tA + 1;
  if (tA < T) {   // uniform across the 16-thread group
    float2 v[16];
    const float* fa = s_wav + kHop * (2 * g);
    const float* fb = fa + kHop;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      int n = 16 * m + j;
      float w = s_win[n];
      v[m] = make_float2(fa[n] * w, fb[n] * w);
    }
    float2* buf = s_buf[g];
    fft256_group<false>(v, j, s_tw, buf, gmask);
    // thread j holds Z[j + 16*k2]; share through buf in natural order
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) buf[j + 16 * k2] = v[k2];
    __syncwarp(gmask);
    float2* outA = spec + ((size_t)sig * T + tA) * kBins;
    float2* outB = outA + kBins;
    float* lmA = logmag ? logmag + ((size_t)sig * T + tA) * kBins : nullptr;
    const bool hasB = tB < T;
    for (int k = j; k <= kFft / 2; k += 16) {
      float2 zk = buf[k];
      float2 zn = buf[(kFft - k) & (kFft - 1)];
      // A = (Zk + conj(Zn))/2 ; B = (Zk - conj(Zn))/(2i)
      float2 a = make_float2(0.5f * (zk.x + zn.x) * inv_wsum, 0.5f * (zk.y - zn.y) * inv_wsum);
      float2 b = make_float2(0.5f * (zk.y + zn.y) * inv_wsum, -0.5f * (zk.x - zn.x) * inv_wsum);
      outA[k] = a;
      if (lmA) lmA[k] = log1pf(sqrtf(a.x * a.x + a.y * a.y));
      if (hasB) {
        outB[k] = b;
        if (lmA) lmA[kBins + k] = log1pf(sqrtf(b.x * b.x + b.y * b.y));
      }
    }
  }
}

// Each block owns 29 output hops [h0, h0+29) of one signal and recomputes the
// 3 halo frames before them: frames f0 = h0-3 .. f0+31.
constexpr int kHopsPerBlock = kFramesPerBlock - 3;

__global__ void __launch_bounds__(256)
istft_kernel(const float2* __restrict__ spec, int T, float* __restrict__ wav) {
  __shared__ float s_win[kFft];
  __shared__ float2 s_tw[kFft];
  __shared__ float2 s_buf[16][kFftPad];   // reused as 2 x 256 real frames per group

  const int tid = threadIdx.x;
  const int sig = blockIdx.y;
  const int h0 = blockIdx.x * kHopsPerBlock;
  const int f0 = h0 - 3;
  const int last_frame = T - 5;            // range(0, 64*T - 256, 64) -> n = 0 .. T-5
  {
    float s, c;
    sincospif(-(float)tid * (2.f / 256.f), &s, &c);
    s_tw[tid] = make_float2(c, s);
    s_win[tid] = window_at(tid);
  }
  __syncthreads();

  const int g = tid >> 4, j = tid & 15;
  const unsigned gmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
  const int fA = f0 + 2 * g, fB = fA + 1;
  const bool okA = fA >= 0 && fA <= last_frame;
  const bool okB = fB >= 0 && fB <= last_frame;
  float* fr = reinterpret_cast<float*>(s_buf[g]);   // [2][256] floats after the transform
  if (okA || okB) {
    const float2* XA = spec + ((size_t)sig * T + (okA ? fA : fB)) * kBins;
    const float2* XB = spec + ((size_t)sig * T + (okB ? fB : fA)) * kBins;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      int k = 16 * m + j;
      int kk = k <= 128 ? k : 256 - k;
      float2 a = okA ? __ldg(XA + kk) : make_float2(0.f, 0.f);
      float2 b = okB ? __ldg(XB + kk) : make_float2(0.f, 0.f);
      if (kk == 0 || kk == 128) { a.y = 0.f; b.y = 0.f; }   // irfft ignores these
      // k <= 128: A + iB ; else conj(A) + i conj(B)
      v[m] = k <= 128 ? make_float2(a.x - b.y, a.y + b.x) : make_float2(a.x + b.y, -a.y + b.x);
    }
    fft256_group<true>(v, j, s_tw, s_buf[g], gmask);
    __syncwarp(gmask);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      int n = j + 16 * k2;
      float w = s_win[n] * (1.f / 256.f);
      fr[n] = v[k2].x * w;
      fr[256 + n] = v[k2].y * w;
    }
  } else {
    for (int n = j; n < 512; n += 16) fr[n] = 0.f;
  }
  __syncthreads();

  // gather: sample i = 64*h + r sums frames h-3..h at offsets r + 64*(h - n)
  float* out = wav + (size_t)sig * kHop * T;
  const float* frames = reinterpret_cast<const float*>(s_buf);
  for (int idx = tid; idx < kHopsPerBlock * kHop; idx += 256) {
    int h = h0 + idx / kHop, r = idx % kHop;
    if (h >= T) break;
    float acc = 0.f, ws = 0.f;
#pragma unroll
    for (int d = 3; d >= 0; --d) {         // ascending frame order n = h-3 .. h, as the reference loop
      int n = h - d;
      if (n >= 0 && n <= last_frame) {
        int slot = n - f0;                 // 0..31
        int grp = slot >> 1, half = slot & 1;
        float wv = s_win[r + kHop * d];
        acc += frames[(size_t)grp * (2 * kFftPad) + half * 256 + r + kHop * d];
        ws += wv * wv;
      }
    }
    out[(size_t)kHop * h + r] = ws != 0.f ? acc / ws : acc;
  }
}

}  // namespace danet

using namespace danet;

extern "C" int danet_stft_num_frames(int n_samples) {
  if (n_samples < kFft) return DANET_E_SHAPE;
  return (n_samples + kHop - 1) / kHop + 1;
}

extern "C" int danet_stft_fwd(const float* wav, int n_sig, int n_samples, float* spec_c64,
                              float* logmag, void* stream) {
  DANET_REQUIRE(n_sig >= 0, DANET_E_SHAPE, "stft: n_sig %d", n_sig);
  if (n_sig == 0 && n_samples >= kFft) return DANET_OK;
  DANET_REQUIRE(wav && spec_c64, DANET_E_ARG, "stft: null pointer");
  DANET_REQUIRE(n_samples >= kFft, DANET_E_SHAPE,
                "stft: window is longer than input signal (%d < 256)", n_samples);
  DANET_REQUIRE(aligned8(spec_c64), DANET_E_ALIGN, "stft: spec must be 8-byte aligned");
  if (n_sig == 0) return DANET_OK;
  DANET_REQUIRE(n_sig <= 65535, DANET_E_SHAPE, "stft: n_sig %d > 65535", n_sig);
  static const float inv_wsum = (float)(1.0 / window_sum());
  const int T = danet_stft_num_frames(n_samples);
  dim3 grid((T + kFramesPerBlock - 1) / kFramesPerBlock, n_sig);
  stft_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      wav, n_samples, T, inv_wsum, reinterpret_cast<float2*>(spec_c64), logmag);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_istft_fwd(const float* spec_c64, int n_sig, int T, float* wav, void* stream) {
  DANET_REQUIRE(n_sig >= 0 && T >= 1, DANET_E_SHAPE, "istft: n_sig %d T %d", n_sig, T);
  if (n_sig == 0) return DANET_OK;
  DANET_REQUIRE(spec_c64 && wav, DANET_E_ARG, "istft: null pointer");
  DANET_REQUIRE(aligned8(spec_c64), DANET_E_ALIGN, "istft: spec must be 8-byte aligned");
  if (n_sig == 0) return DANET_OK;
  DANET_REQUIRE(n_sig <= 65535, DANET_E_SHAPE, "istft: n_sig %d > 65535", n_sig);
  dim3 grid((T + kHopsPerBlock - 1) / kHopsPerBlock, n_sig);
  istft_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec_c64), T, wav);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
