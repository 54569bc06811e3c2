// K5: permutation-invariant MSE + SNR.
//   ops.pit_mse_loss (app/ops.py:374-431): X[b,i,j] = mean_{t,f} |x_i - y_j|^2,
//   L[b,p] = sum_i X[b,i,perm_p(i)] with perms in itertools.permutations order,
//   first argmin, loss = mean_b min_p L.  ops.batch_snr (app/ops.py:191-222) on the
//   aligned estimate: noise power = L[b,p*] / C, signal power = mean |x|^2.
// One streaming pass over x and y (HBM-bound) + a tiny deterministic finalize.
#include "common.cuh"

namespace danet {

constexpr int kPitMaxC = 4;
constexpr int kPitParts = 32;

template <bool CPLX>
__global__ void __launch_bounds__(256)
pit_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, int C, long long TF,
                   float* __restrict__ part) {
  __shared__ float s_red[8][kPitMaxC * kPitMaxC + kPitMaxC];
  const int b = blockIdx.y, p = blockIdx.x;
  const long long chunk = (TF + kPitParts - 1) / kPitParts;
  const long long lo = p * chunk, hi = min(TF, lo + chunk);
  float acc[kPitMaxC * kPitMaxC + kPitMaxC];
#pragma unroll
  for (int k = 0; k < kPitMaxC * kPitMaxC + kPitMaxC; ++k) acc[k] = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    float2 xv[kPitMaxC], yv[kPitMaxC];
#pragma unroll
    for (int c = 0; c < kPitMaxC; ++c)
      if (c < C) {
        size_t o = ((size_t)b * C + c) * TF + i;
        if (CPLX) {
          xv[c] = __ldg(reinterpret_cast<const float2*>(x) + o);
          yv[c] = __ldg(reinterpret_cast<const float2*>(y) + o);
        } else {
          xv[c] = make_float2(__ldg(x + o), 0.f);
          yv[c] = make_float2(__ldg(y + o), 0.f);
        }
      }
#pragma unroll
    for (int i1 = 0; i1 < kPitMaxC; ++i1)
      if (i1 < C) {
        acc[kPitMaxC * kPitMaxC + i1] += xv[i1].x * xv[i1].x + xv[i1].y * xv[i1].y;
#pragma unroll
        for (int j1 = 0; j1 < kPitMaxC; ++j1)
          if (j1 < C) {
            float dr = xv[i1].x - yv[j1].x, di = xv[i1].y - yv[j1].y;
            acc[i1 * kPitMaxC + j1] += dr * dr + di * di;
          }
      }
  }
#pragma unroll
  for (int k = 0; k < kPitMaxC * kPitMaxC + kPitMaxC; ++k) {
    float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kPitMaxC * kPitMaxC + kPitMaxC) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
    part[((size_t)b * kPitParts + p) * (kPitMaxC * kPitMaxC + kPitMaxC) + threadIdx.x] = t;
  }
}

// One block: warp w takes utterances w, w + 8, ...; lane k < 20 adds up the partial sums of entry k in part order (the
// single-thread version of round 1 spent 0.2 ms on 20 K dependent loads at B = 32), lane 0 walks the C! permutations;
// thread 0 then adds the B minima in utterance order.  Same arithmetic order as before: bit-identical and deterministic.
__global__ void __launch_bounds__(256)
pit_finalize_kernel(const float* __restrict__ part, int B, int C, long long TF,
                    float* __restrict__ cross, float* __restrict__ perm_losses,
                    int* __restrict__ perm_idx, float* __restrict__ loss,
                    float* __restrict__ snr, float* __restrict__ best_ws) {
  constexpr int kN = kPitMaxC * kPitMaxC + kPitMaxC;
  __shared__ float s_x[8][kN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int nperm = 1;
  for (int c = 2; c <= C; ++c) nperm *= c;
  const float inv_n = 1.f / (float)TF;
  for (int b = warp; b < B; b += 8) {
    if (lane < kN) {
      float t = 0.f;
      for (int p = 0; p < kPitParts; ++p) t += part[((size_t)b * kPitParts + p) * kN + lane];
      s_x[warp][lane] = t * inv_n;
    }
    __syncwarp();
    if (lane == 0) {
      const float* X = s_x[warp];
      for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) cross[((size_t)b * C + i) * C + j] = X[i * kPitMaxC + j];
      int perm[kPitMaxC];
      for (int c = 0; c < C; ++c) perm[c] = c;
      float best = 0.f;
      int best_p = 0;
      for (int p = 0; p < nperm; ++p) {
        float L = 0.f;
        for (int i = 0; i < C; ++i) L += X[i * kPitMaxC + perm[i]];
        if (perm_losses) perm_losses[(size_t)b * nperm + p] = L;
        if (p == 0 || L < best) { best = L; best_p = p; }
        // next lexicographic permutation (itertools.permutations order)
        int i = C - 2;
        while (i >= 0 && perm[i] > perm[i + 1]) --i;
        if (i >= 0) {
          int j = C - 1;
          while (perm[j] < perm[i]) --j;
          int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
          for (int l = i + 1, r = C - 1; l < r; ++l, --r) { t = perm[l]; perm[l] = perm[r]; perm[r] = t; }
        }
      }
      if (perm_idx) perm_idx[b] = best_p;
      best_ws[b] = best;
      if (snr) {
        float sp = 0.f;
        for (int c = 0; c < C; ++c) sp += X[kPitMaxC * kPitMaxC + c];
        sp /= (float)C;
        float np = best / (float)C;
        snr[b] = 4.342944819f * (logf(sp + kEps) - logf(np + kEps));
      }
    }
    __syncwarp();
  }
  __syncthreads();                         // best_ws was written by this block's own threads
  if (threadIdx.x == 0 && loss) {
    float total = 0.f;
    for (int b2 = 0; b2 < B; ++b2) total += best_ws[b2];
    loss[0] = total / (float)B;
  }
}

}  // namespace danet

using namespace danet;

extern "C" size_t danet_pit_workspace_bytes(int B, int C) {
  (void)C;
  return (size_t)(B > 0 ? B : 0) * (kPitParts * (kPitMaxC * kPitMaxC + kPitMaxC) + 1) * sizeof(float);
}

extern "C" int danet_pit_mse_fwd(const float* x, const float* y, int B, int C, int TF, int is_complex,
                                 float* cross, float* perm_losses, int* perm_idx, float* loss,
                                 float* snr, void* workspace, size_t workspace_bytes, void* stream) {
  DANET_REQUIRE(x && y && cross && workspace, DANET_E_ARG, "pit_mse: null pointer");
  DANET_REQUIRE(B >= 1 && TF >= 1 && C >= 1 && C <= kPitMaxC, DANET_E_SHAPE,
                "pit_mse: B %d C %d (<=%d) TF %d", B, C, kPitMaxC, TF);
  DANET_REQUIRE(workspace_bytes >= danet_pit_workspace_bytes(B, C), DANET_E_WORKSPACE,
                "pit_mse: workspace %zu < %zu", workspace_bytes, danet_pit_workspace_bytes(B, C));
  DANET_REQUIRE(B <= 65535, DANET_E_SHAPE, "pit_mse: B %d > 65535", B);
  if (is_complex) DANET_REQUIRE(aligned8(x) && aligned8(y), DANET_E_ALIGN, "pit_mse: complex buffers must be 8-byte aligned");
  float* part = reinterpret_cast<float*>(workspace);
  dim3 grid(kPitParts, B);
  if (is_complex)
    pit_partial_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, y, C, TF, part);
  else
    pit_partial_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, y, C, TF, part);
  DANET_LAUNCH_CHECK();
  pit_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(part, B, C, TF, cross, perm_losses, perm_idx, loss, snr,
                                                        part + (size_t)B * kPitParts * (kPitMaxC * kPitMaxC + kPitMaxC));
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
