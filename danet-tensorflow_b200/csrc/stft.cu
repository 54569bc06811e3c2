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

// log(1 + |z|) (main.py:239-240) on the special-function unit: sqrt.approx + lg2.approx (7 instructions against ~30 for
// sqrtf + log1pf; the kernel is bound by instruction issue, not by HBM).  Absolute error <= 2e-7 (the sum 1 + |z| rounds
// to 6e-8; lg2.approx 2^-21.4 near 1, 3 ulp elsewhere) against values of up to ~10 and the 5e-6 gate of the STFT tests.
__device__ __forceinline__ float log1p_abs(float re, float im) {
  float m;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(re * re + im * im));
  return __logf(1.f + m);
}

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
  {
    // all nine loads of a thread in flight before the first store: rolled, every iteration waited out its own global
    // load (ncu source page: half of the kernel's stall samples sat on this store -- nine DRAM latencies in series)
    constexpr int kN = kHop * (kFramesPerBlock - 1) + kFft, kIt = (kN + 255) / 256;
    float v[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = tid + 256 * it;
      const long long g = first + i;
      v[it] = (i < kN && g >= 0 && g < n_samples) ? __ldg(w_in + g) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int i = tid + 256 * it;
      if (i < kN) s_wav[i] = v[it];
    }
  }
  __syncthreads();

  const int g = tid >> 4, j = tid & 15;
  // The transform needs its 16-thread group only, but the two groups of a warp run it in lockstep under FULL-warp
  // barriers: with a half-warp mask per group (first build) the halves are separate convergence groups and every
  // instruction is issued twice.  Frames beyond T are transformed too (their samples are zeros inside s_wav) and only
  // their stores are skipped, so the whole warp always takes part.
  const unsigned gmask = 0xFFFFFFFFu;
  const int tA = t0 + 2 * g, tB = tA + 1;
  {
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
    for (int k = j; k <= kFft / 2 && tA < T; k += 16) {
      float2 zk = buf[k];
      float2 zn = buf[(kFft - k) & (kFft - 1)];
      // A = (Zk + conj(Zn))/2 ; B = (Zk - conj(Zn))/(2i)
      float2 a = make_float2(0.5f * (zk.x + zn.x) * inv_wsum, 0.5f * (zk.y - zn.y) * inv_wsum);
      float2 b = make_float2(0.5f * (zk.y + zn.y) * inv_wsum, -0.5f * (zk.x - zn.x) * inv_wsum);
      outA[k] = a;
      if (lmA) lmA[k] = log1p_abs(a.x, a.y);
      if (hasB) {
        outB[k] = b;
        if (lmA) lmA[kBins + k] = log1p_abs(b.x, b.y);
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


// K4 in one kernel: dot-product mask (app/modules.py:548-603) x complex mixture (main.py:281-284) -> utils.istft
// (app/utils.py:53-75) for every source.  A block owns the same 29 hops + 3 halo frames of ONE mixture as
// istft_kernel; it first evaluates mask_c(t,f) * mix(t,f) for its 32 frames and all C sources into shared memory (one
// read of the embedding tile, the same arithmetic as mask_cmul_kernel), then runs the inverse transform and the
// overlap-add once per source from there.  The separated spectra (8*C*T*F bytes per mixture) never exist in HBM.
constexpr int kFusedMaxC = 4;
constexpr int kFusedMaxE = 64;
constexpr int kFusedThreads = 512;     // phase 1: 512 threads of loads in flight; phase 2: two sources transformed at once

// EQ = E / 4 as a compile-time constant (5: E = 20, 10: E = 40; 0 = any E): the embedding of a bin is then fetched by EQ
// back-to-back 16-byte loads, and with the bin loop unrolled 4x a thread has 20-40 loads in flight instead of one dependent
// load per FMA group -- phase 1 is latency-bound (one 140 KB block per SM).
template <int EQ>
__global__ void __launch_bounds__(kFusedThreads)
mask_istft_kernel(const float* __restrict__ embed, const float* __restrict__ attractors, const float2* __restrict__ mix,
                  int C, int T, int E, int kind, float* __restrict__ wav) {
  extern __shared__ __align__(16) uint8_t fused_smem[];
  float2* s_buf = reinterpret_cast<float2*>(fused_smem);                    // [2 halves][16][kFftPad]
  float2* s_tw = s_buf + 2 * 16 * kFftPad;                                  // [256]
  float* s_win = reinterpret_cast<float*>(s_tw + kFft);                     // [256]
  float* s_att = s_win + kFft;                                              // [C][E]
  float2* s_spec = reinterpret_cast<float2*>(s_att + kFusedMaxC * kFusedMaxE);   // [C][32][kBins]

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int h0 = blockIdx.x * kHopsPerBlock;
  const int f0 = h0 - 3;
  const int last_frame = T - 5;            // range(0, 64*T - 256, 64) -> n = 0 .. T-5
  if (tid < kFft) {
    float sn, cs;
    sincospif(-(float)tid * (2.f / 256.f), &sn, &cs);
    s_tw[tid] = make_float2(cs, sn);
    s_win[tid] = window_at(tid);
  }
  for (int i = tid; i < C * E; i += kFusedThreads) s_att[i] = attractors[(size_t)b * C * E + i];
  __syncthreads();

  // ---- phase 1: masked spectra of frames f0 .. f0+31 for every source -> shared memory
  const long long TF = (long long)T * kBins;
  const float* Vb = embed + (size_t)b * TF * E;
  const float2* Mb = mix + (size_t)b * TF;
#pragma unroll 4
  for (int idx = tid; idx < kFramesPerBlock * kBins; idx += kFusedThreads) {
    const int slot = idx / kBins, k = idx - slot * kBins;
    const int fr = f0 + slot;
    float m[kFusedMaxC];
    float2 z = make_float2(0.f, 0.f);
    if (fr >= 0 && fr <= last_frame) {
      const long long i = (long long)fr * kBins + k;
      float logit[kFusedMaxC];
#pragma unroll
      for (int c = 0; c < kFusedMaxC; ++c) logit[c] = 0.f;
      const float4* v4 = reinterpret_cast<const float4*>(Vb + (size_t)i * E);
      z = __ldg(Mb + i);
      auto dot4 = [&](const float4 x, int e4) {
#pragma unroll
        for (int c = 0; c < kFusedMaxC; ++c)
          if (c < C) {
            const float* a = s_att + c * E + 4 * e4;
            logit[c] = fmaf(x.x, a[0], logit[c]);
            logit[c] = fmaf(x.y, a[1], logit[c]);
            logit[c] = fmaf(x.z, a[2], logit[c]);
            logit[c] = fmaf(x.w, a[3], logit[c]);
          }
      };
      if (EQ > 0) {
        float4 xs[EQ > 0 ? EQ : 1];
#pragma unroll
        for (int e4 = 0; e4 < EQ; ++e4) xs[e4] = __ldg(v4 + e4);
#pragma unroll
        for (int e4 = 0; e4 < EQ; ++e4) dot4(xs[e4], e4);
      } else {
        for (int e4 = 0; e4 < E / 4; ++e4) dot4(__ldg(v4 + e4), e4);
      }
      if (kind == 0) {
        float mx = logit[0];
#pragma unroll
        for (int c = 1; c < kFusedMaxC; ++c)
          if (c < C) mx = fmaxf(mx, logit[c]);
        float den = 0.f;
#pragma unroll
        for (int c = 0; c < kFusedMaxC; ++c)
          if (c < C) {
            m[c] = __expf(logit[c] - mx);          // ex2.approx: 2 ulp on arguments <= 0, masks move by < 1e-6
            den += m[c];
          }
        const float inv = __fdividef(1.f, den);        // den in [1, C]
#pragma unroll
        for (int c = 0; c < kFusedMaxC; ++c)
          if (c < C) m[c] *= inv;
      } else {
#pragma unroll
        for (int c = 0; c < kFusedMaxC; ++c)
          if (c < C) m[c] = __fdividef(1.f, 1.f + __expf(-fmaxf(logit[c], -80.f)));
      }
    } else {
#pragma unroll
      for (int c = 0; c < kFusedMaxC; ++c) m[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < kFusedMaxC; ++c)
      if (c < C) s_spec[((size_t)c * kFramesPerBlock + slot) * kBins + k] = make_float2(z.x * m[c], z.y * m[c]);
  }
  __syncthreads();

  // ---- phase 2: the inverse transform and overlap-add of istft_kernel from shared memory, two sources at a time
  // (threads 0-255 take source c0, threads 256-511 source c0 + 1)
  const int half = tid >> 8, t8 = tid & 255;
  const int g = t8 >> 4, j = t8 & 15;
  const unsigned gmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
  const int fA = f0 + 2 * g, fB = fA + 1;
  const bool okA = fA >= 0 && fA <= last_frame;
  const bool okB = fB >= 0 && fB <= last_frame;
  float2* my_buf = s_buf + (size_t)half * 16 * kFftPad;
  float* fr = reinterpret_cast<float*>(my_buf + (size_t)g * kFftPad);   // [2][256] floats after the transform
  for (int c0 = 0; c0 < C; c0 += 2) {
    const int c = c0 + half;
    const bool live = c < C;
    if (live && (okA || okB)) {
      const float2* XA = s_spec + ((size_t)c * kFramesPerBlock + 2 * g) * kBins;
      const float2* XB = XA + kBins;
      float2 v[16];
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int k = 16 * m + j;
        const int kk = k <= 128 ? k : 256 - k;
        float2 a = XA[kk];                                  // out-of-range frames were stored as zeros
        float2 bq = XB[kk];
        if (kk == 0 || kk == 128) { a.y = 0.f; bq.y = 0.f; }   // irfft ignores these
        v[m] = k <= 128 ? make_float2(a.x - bq.y, a.y + bq.x) : make_float2(a.x + bq.y, -a.y + bq.x);
      }
      fft256_group<true>(v, j, s_tw, my_buf + (size_t)g * kFftPad, gmask);
      __syncwarp(gmask);
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int n = j + 16 * k2;
        const float w = s_win[n] * (1.f / 256.f);
        fr[n] = v[k2].x * w;
        fr[256 + n] = v[k2].y * w;
      }
    } else if (live) {
      for (int n = j; n < 512; n += 16) fr[n] = 0.f;
    }
    __syncthreads();
    if (live) {
      float* out = wav + ((size_t)b * C + c) * kHop * T;
      const float* frames = reinterpret_cast<const float*>(my_buf);
      for (int idx = t8; idx < kHopsPerBlock * kHop; idx += 256) {
        const int h = h0 + idx / kHop, r = idx % kHop;
        if (h >= T) break;
        float acc = 0.f, ws = 0.f;
#pragma unroll
        for (int d = 3; d >= 0; --d) {         // ascending frame order n = h-3 .. h, as the reference loop
          const int n = h - d;
          if (n >= 0 && n <= last_frame) {
            const int slot = n - f0;
            const int grp = slot >> 1, hf = slot & 1;
            const float wv = s_win[r + kHop * d];
            acc += frames[(size_t)grp * (2 * kFftPad) + hf * 256 + r + kHop * d];
            ws += wv * wv;
          }
        }
        out[(size_t)kHop * h + r] = ws != 0.f ? acc / ws : acc;
      }
    }
    __syncthreads();                         // s_buf is reused by the next pair of sources
  }
}

static size_t mask_istft_smem_bytes(int C) {
  return (size_t)2 * 16 * kFftPad * 8 + kFft * 8 + kFft * 4 + (size_t)kFusedMaxC * kFusedMaxE * 4 +
         (size_t)C * kFramesPerBlock * kBins * 8;
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

extern "C" int danet_mask_cmul_istft_fwd(const float* embed, const float* attractors, const float* mix_c64, float* wav,
                                         int B, int C, int T, int E, int kind, void* stream) {
  DANET_REQUIRE(B >= 0 && T >= 1 && C >= 1 && C <= kFusedMaxC && E >= 4 && E % 4 == 0 && E <= kFusedMaxE, DANET_E_SHAPE,
                "mask_cmul_istft: B %d C %d (<=%d) T %d E %d (multiple of 4, <=%d)", B, C, kFusedMaxC, T, E, kFusedMaxE);
  DANET_REQUIRE(kind == 0 || kind == 1, DANET_E_ARG, "mask_cmul_istft: kind %d", kind);
  if (B == 0) return DANET_OK;
  DANET_REQUIRE(embed && attractors && mix_c64 && wav, DANET_E_ARG, "mask_cmul_istft: null pointer");
  DANET_REQUIRE(aligned16(embed) && aligned8(mix_c64), DANET_E_ALIGN,
                "mask_cmul_istft: embed must be 16-byte, mix 8-byte aligned");
  DANET_REQUIRE(B <= 65535, DANET_E_SHAPE, "mask_cmul_istft: B %d > 65535", B);
  const size_t smem = mask_istft_smem_bytes(C);
  dim3 grid((T + kHopsPerBlock - 1) / kHopsPerBlock, B);
  const float2* mix = reinterpret_cast<const float2*>(mix_c64);
#define DANET_K4_LAUNCH(EQ)                                                                                         \
  do {                                                                                                              \
    DANET_CUDA(cudaFuncSetAttribute(mask_istft_kernel<EQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mask_istft_kernel<EQ><<<grid, kFusedThreads, smem, as_stream(stream)>>>(embed, attractors, mix, C, T, E, kind, wav); \
  } while (0)
  if (E == 20) DANET_K4_LAUNCH(5);
  else if (E == 40) DANET_K4_LAUNCH(10);
  else DANET_K4_LAUNCH(0);
#undef DANET_K4_LAUNCH
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
