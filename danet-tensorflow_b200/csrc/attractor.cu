// K3: attractor estimation -- one streaming pass over the embedding per call.
//   truth family  app/modules.py:390-412 / 425-450 / 462-487
//   anchor        app/modules.py:501-545 + app/ops.py:273-292
//   k-means       new plugin (README.md:216-217 lists it as unimplemented)
// All three are "per-bin row weights, then a weighted sum of embedding vectors":
//   acc[r][e] += Wt[tf][r] * V[tf][e],  den[r] += Wt[tf][r]
// with R rows = C (truth, k-means: one-hot class x weight) or P*C (anchor: softmax over
// every anchor subset).  A block stages a tile of V in shared memory with coalesced
// float4 loads, phase 1 computes Wt for the tile (one thread per bin), phase 2 is a small
// register-tiled product.  Partials go to the workspace; a finalize kernel reduces them in
// a fixed order (deterministic, no atomics) and applies the estimator's epilogue.
#include <stdlib.h>
#include "common.cuh"
#include "mma_sync.cuh"

namespace danet {

constexpr int kTile = 128;        // bins per tile
constexpr int kAttMaxC = 4;
constexpr int kAttMaxE = 128;
constexpr int kAttMaxRows = 80;   // C(6,3)*3 = 60
constexpr int kMaxAnchor = 8;
constexpr int kParts = 32;        // blocks per utterance
constexpr int kMaxAccPerThread = 4;

enum { MODE_TRUTH = 0, MODE_ANCHOR = 1, MODE_KMEANS = 2, MODE_ANCHOR2 = 3 };

struct AttParams {
  const float* embed;     // [B][TF][E]
  const float* src_pwr;   // truth: [B][C][TF]
  const float* mix_pwr;   // truth modes 1,2: [B][TF]
  const float* aux;       // anchor: anchors [A][E]; kmeans: centroids [B][C][E]
  float* part;            // [B][kParts][R][nQ*4]
  long long TF;
  int C, E, R, nQ, ldv, n_anchor, n_sub, truth_mode;
  int subsets[20 * kAttMaxC];   // anchor index table [P][C]
};

// phase 1 for one time-frequency bin: the R row weights Wt[tf][:]
// EC: the embedding size when it is known at compile time (register-tiled fast path), 0 = read it from p
template <int MODE, int EC = 0>
__device__ __forceinline__ void row_weights(const AttParams& p, int b, long long tf, bool in_range,
                                            const float* __restrict__ v, const float* __restrict__ sAux,
                                            float* __restrict__ w) {
  const int E = EC ? EC : p.E, C = p.C, R = p.R;
  const long long TF = p.TF;
  if (!in_range) {
    for (int r = 0; r < R; ++r) w[r] = 0.f;
  } else if (MODE == MODE_TRUTH) {
    const float* sp = p.src_pwr + (size_t)b * C * TF + tf;
    int k = 0;
    float best = __ldg(sp);
    for (int c = 1; c < C; ++c) {          // first maximum on ties (modules.py:396)
      const float x = __ldg(sp + (size_t)c * TF);
      if (x > best) { best = x; k = c; }
    }
    float wt = 1.f;
    if (p.truth_mode == 1) wt = __ldg(p.mix_pwr + (size_t)b * TF + tf) > 5.f ? 1.f : 0.f;
    if (p.truth_mode == 2) wt = __ldg(p.mix_pwr + (size_t)b * TF + tf);
    for (int c = 0; c < C; ++c) w[c] = c == k ? wt : 0.f;
  } else if (MODE == MODE_ANCHOR) {
    float logit[kMaxAnchor];
#pragma unroll
    for (int a = 0; a < kMaxAnchor; ++a) logit[a] = 0.f;
    for (int e = 0; e < E; e += 4) {
      const float4 x = *reinterpret_cast<const float4*>(v + e);
#pragma unroll
      for (int a = 0; a < kMaxAnchor; ++a)
        if (a < p.n_anchor) {
          const float4 an = *reinterpret_cast<const float4*>(sAux + a * E + e);
          logit[a] = fmaf(x.x, an.x, logit[a]);
          logit[a] = fmaf(x.y, an.y, logit[a]);
          logit[a] = fmaf(x.z, an.z, logit[a]);
          logit[a] = fmaf(x.w, an.w, logit[a]);
        }
    }
    for (int s = 0; s < p.n_sub; ++s) {     // eq.6: softmax over the subset's anchors
      float l[kAttMaxC];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kAttMaxC; ++c)
        if (c < C) {
          const int a = p.subsets[s * kAttMaxC + c];
          float lv = logit[0];
#pragma unroll
          for (int q = 1; q < kMaxAnchor; ++q) lv = a == q ? logit[q] : lv;
          l[c] = lv;
          mx = fmaxf(mx, lv);
        }
      float den = 0.f;
#pragma unroll
      for (int c = 0; c < kAttMaxC; ++c)
        if (c < C) { l[c] = expf(l[c] - mx); den += l[c]; }
      const float inv = 1.f / den;
#pragma unroll
      for (int c = 0; c < kAttMaxC; ++c)
        if (c < C) w[s * C + c] = l[c] * inv;
    }
  } else if (MODE == MODE_ANCHOR2) {
    // two sources: softmax over a pair (a,b) is S_a = 1 / (1 + exp(l_b - l_a)), S_b = 1 - S_a.  Only S_a is
    // accumulated; row P carries the constant weight 1, and the finalize kernel recovers the second member of
    // every pair as (sum V - sum S_a V) / (N - sum S_a): half the eq.7 products.  Pairs in
    // itertools.combinations order, indices compile-time after unrolling (no register-select chains).
    float logit[kMaxAnchor];
#pragma unroll
    for (int a = 0; a < kMaxAnchor; ++a) logit[a] = 0.f;
    for (int e = 0; e < E; e += 4) {
      const float4 x = *reinterpret_cast<const float4*>(v + e);
#pragma unroll
      for (int a = 0; a < kMaxAnchor; ++a)
        if (a < p.n_anchor) {
          const float4 an = *reinterpret_cast<const float4*>(sAux + a * E + e);
          logit[a] = fmaf(x.x, an.x, logit[a]);
          logit[a] = fmaf(x.y, an.y, logit[a]);
          logit[a] = fmaf(x.z, an.z, logit[a]);
          logit[a] = fmaf(x.w, an.w, logit[a]);
        }
    }
    int s = 0;
#pragma unroll
    for (int a = 0; a < kMaxAnchor; ++a)
#pragma unroll
      for (int bq = a + 1; bq < kMaxAnchor; ++bq)
        if (bq < p.n_anchor) w[s++] = __fdividef(1.f, 1.f + __expf(logit[bq] - logit[a]));
    w[s] = 1.f;
  } else {   // k-means: nearest centroid, first minimum on ties
    int k = 0;
    float best = INFINITY;
    for (int c = 0; c < C; ++c) {
      float d = 0.f;
      for (int e = 0; e < E; ++e) {
        const float t = v[e] - sAux[c * E + e];
        d = fmaf(t, t, d);
      }
      if (d < best) { best = d; k = c; }
    }
    for (int c = 0; c < C; ++c) w[c] = c == k ? 1.f : 0.f;
  }
}

// Fast path, E = 4*NQ known at compile time: 256 bins per tile, phase 1 one bin per thread, phase 2
// thread (row r, group g) keeps the whole row accumulator (E sums + the weight sum) in registers and
// walks the bins g, g+G, ...: 1 weight load + NQ broadcast float4 loads per 4*NQ+1 FMAs.
constexpr int kTileRT = 256;

template <int MODE, int NQ>
__global__ void __launch_bounds__(256)
attractor_partial_rt_kernel(const AttParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int E = 4 * NQ;
  const int C = p.C, R = p.R, ldv = p.ldv;
  float* sV = smem;                                  // [kTileRT][ldv]
  const int ldw = R | 1;                             // odd row stride: the per-bin weight stores are conflict-free
  float* sW = sV + kTileRT * ldv;                    // [kTileRT][ldw]
  float* sAux = sW + kTileRT * ldw;                  // anchors / centroids (16-byte aligned: ldw*256 floats)
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x;
  const long long TF = p.TF;
  const float* Vb = p.embed + (size_t)b * TF * E;
  const int n_aux = (MODE == MODE_ANCHOR || MODE == MODE_ANCHOR2) ? p.n_anchor * E : (MODE == MODE_KMEANS ? C * E : 0);
  for (int i = tid; i < n_aux; i += 256)
    sAux[i] = MODE == MODE_KMEANS ? p.aux[(size_t)b * C * E + i] : p.aux[i];

  const int G = 256 / R;                             // bin-interleaved groups per row (R <= 256)
  const int r = tid % R, g = tid / R;
  const bool active = g < G;
  float acc[E];
  float den = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;

  const long long n_tiles = (TF + kTileRT - 1) / kTileRT;
  const long long tiles_per = (n_tiles + kParts - 1) / kParts;
  const long long t_lo = part * tiles_per;
  const long long t_hi = t_lo + tiles_per < n_tiles ? t_lo + tiles_per : n_tiles;
  for (long long tile = t_lo; tile < t_hi; ++tile) {
    const long long tf0 = tile * kTileRT;
    const int n_here = (int)(TF - tf0 < kTileRT ? TF - tf0 : kTileRT);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(Vb + (size_t)tf0 * E);
      for (int i = tid; i < n_here * NQ; i += 256) {
        const float4 x = __ldg(src + i);
        *reinterpret_cast<float4*>(sV + (i / NQ) * ldv + 4 * (i % NQ)) = x;
      }
    }
    __syncthreads();
    row_weights<MODE, E>(p, b, tf0 + tid, tid < n_here, sV + tid * ldv, sAux, sW + tid * ldw);
    __syncthreads();
    if (active) {
      for (int tfl = g; tfl < n_here; tfl += G) {
        const float wv = sW[tfl * ldw + r];
        const float* vr = sV + tfl * ldv;
        den += wv;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(vr + 4 * q);
          acc[4 * q + 0] = fmaf(wv, x.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(wv, x.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(wv, x.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(wv, x.w, acc[4 * q + 3]);
        }
      }
    }
  }
  // fixed-order reduction of the G groups through shared memory
  __syncthreads();
  constexpr int LD = E + 4;                          // partial row layout: E sums, weight sum, 3 zeros
  float* red = smem;                                 // [G][R][LD]
  if (active) {
    float* o = red + ((size_t)g * R + r) * LD;
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] = acc[e];
    o[E] = den; o[E + 1] = 0.f; o[E + 2] = 0.f; o[E + 3] = 0.f;
  }
  __syncthreads();
  float* dst = p.part + ((size_t)b * kParts + part) * R * LD;
  for (int i = tid; i < R * LD; i += 256) {
    float sum = red[i];
    for (int gg = 1; gg < G; ++gg) sum += red[(size_t)gg * R * LD + i];
    dst[i] = sum;
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
attractor_partial_kernel(const AttParams p) {
  extern __shared__ __align__(16) float smem[];
  const int E = p.E, C = p.C, R = p.R, nQ = p.nQ, ldv = p.ldv;
  float* sV = smem;                                  // [kTile][ldv]: E values, then (1,0,0,0)
  float* sW = sV + kTile * ldv;                      // [kTile][R]
  float* sAux = sW + kTile * R;                      // anchors / centroids
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x;
  const long long TF = p.TF;
  const float* Vb = p.embed + (size_t)b * TF * E;

  const int n_aux = (MODE == MODE_ANCHOR || MODE == MODE_ANCHOR2) ? p.n_anchor * E : (MODE == MODE_KMEANS ? C * E : 0);
  for (int i = tid; i < n_aux; i += 256)
    sAux[i] = MODE == MODE_KMEANS ? p.aux[(size_t)b * C * E + i] : p.aux[i];
  for (int i = tid; i < kTile; i += 256) {   // constant "ones" quad feeding the denominators
    float* q = sV + i * ldv + E;
    q[0] = 1.f; q[1] = 0.f; q[2] = 0.f; q[3] = 0.f;
  }

  // phase-2 mapping: output quads (r, q) split over G bin-interleaved groups
  const int n_out = R * nQ;
  const int G = n_out >= 256 ? 1 : 256 / n_out;
  float4 acc[kMaxAccPerThread];
#pragma unroll
  for (int i = 0; i < kMaxAccPerThread; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  const long long n_tiles = (TF + kTile - 1) / kTile;
  const long long tiles_per = (n_tiles + kParts - 1) / kParts;
  const long long t_lo = part * tiles_per;
  const long long t_hi = t_lo + tiles_per < n_tiles ? t_lo + tiles_per : n_tiles;
  const int eq = E / 4;
  for (long long tile = t_lo; tile < t_hi; ++tile) {
    const long long tf0 = tile * kTile;
    const int n_here = (int)(TF - tf0 < kTile ? TF - tf0 : kTile);
    __syncthreads();   // previous tile fully consumed (also orders the prologue stores)
    {   // stage V tile: contiguous float4 stream
      const float4* src = reinterpret_cast<const float4*>(Vb + (size_t)tf0 * E);
      for (int i = tid; i < n_here * eq; i += 256) {
        const float4 v = __ldg(src + i);
        *reinterpret_cast<float4*>(sV + (i / eq) * ldv + 4 * (i % eq)) = v;
      }
    }
    __syncthreads();
    if (tid < kTile)   // phase 1: row weights of bin tf0 + tid
      row_weights<MODE>(p, b, tf0 + tid, tid < n_here, sV + tid * ldv, sAux, sW + tid * R);
    __syncthreads();
    // phase 2: acc[r][q] += Wt[tf][r] * V[tf][4q..4q+3]
#pragma unroll
    for (int i = 0; i < kMaxAccPerThread; ++i) {
      const int o = tid + i * 256;
      if (o < G * n_out) {
        const int g = o / n_out, rq = o % n_out, r = rq / nQ, q = rq % nQ;
        float4 a = acc[i];
        for (int tfl = g; tfl < n_here; tfl += G) {
          const float wv = sW[tfl * R + r];
          const float4 x = *reinterpret_cast<const float4*>(sV + tfl * ldv + 4 * q);
          a.x = fmaf(wv, x.x, a.x); a.y = fmaf(wv, x.y, a.y);
          a.z = fmaf(wv, x.z, a.z); a.w = fmaf(wv, x.w, a.w);
        }
        acc[i] = a;
      }
    }
  }
  // reduce the G groups through shared memory (reuse sV), fixed order
  __syncthreads();
  float4* red = reinterpret_cast<float4*>(smem);
#pragma unroll
  for (int i = 0; i < kMaxAccPerThread; ++i) {
    const int o = tid + i * 256;
    if (o < G * n_out) red[o] = acc[i];
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(p.part) + ((size_t)b * kParts + part) * n_out;
  for (int rq = tid; rq < n_out; rq += 256) {
    float4 s = red[rq];
    for (int g = 1; g < G; ++g) {
      const float4 t = red[g * n_out + rq];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    dst[rq] = s;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Anchor estimator, two sources, E = 20: the benchmark configuration (app/modules.py:501-545 with C = 2).
// Same arithmetic as MODE_ANCHOR2 above (first-member sigmoids + an all-ones row), restructured around what the
// round-1 profile showed (18 % of the HBM peak, 21 % warps active, three block barriers per 256-bin tile):
//   * a WARP owns a chunk of 32 consecutive bins (2560 contiguous bytes) and never synchronises with another warp
//     until the final reduction; the chunks arrive by cp.async.bulk into a per-warp double buffer, so the next
//     chunk's HBM latency hides under the current chunk's arithmetic;
//   * the weighted sums  acc[r][e] += S[bin][r] * V[bin][e]  (eq.7; 16 rows x 21 columns x 32 bins per chunk) are
//     a small matrix product and run on the tensor cores: mma.sync m16n8k8 on TF32 hi/lo splits of BOTH operands,
//     three products per term ("3xTF32": ~fp32 accuracy, far inside the 5e-5 the unit test asks for) -- 36 MMAs in
//     place of 10 752 FMAs per chunk; column 20 of the B operand is the constant 1, so the weight sums (the
//     denominators of eq.7) fall out of the same product;
//   * the six anchor logits of a bin are computed once by the lane that owns the bin and shared through 1 KB of
//     shared memory; every lane then evaluates the sigmoids of exactly the (row, bin) pairs its A fragment holds.
// Deterministic: chunk -> warp assignment and every summation order are fixed.
constexpr int kMmaE = 20;
constexpr int kMmaChunk = 32;                       // bins per warp step = MMA K (4 x k8)
constexpr int kMmaWarps = 8;
constexpr int kMmaLd = kMmaE + 4;                   // partial row: 20 sums, weight sum, 3 zeros (= finalize's ld)

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                 "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra W_%=;\n\t}"
      ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(32 * kMmaWarps)
attractor_anchor2_mma_kernel(const float* __restrict__ embed, const float* __restrict__ anchors, float* __restrict__ part,
                             long long TF, int n_anchor, int n_sub) {
  extern __shared__ __align__(128) uint8_t mma_smem[];
  typedef float VBuf[2][kMmaChunk * kMmaE];
  typedef float LBuf[kMmaChunk][8];
  VBuf* sV = reinterpret_cast<VBuf*>(mma_smem);                               // per-warp double buffer: [bin][e]
  LBuf* sL = reinterpret_cast<LBuf*>(sV + kMmaWarps);                         // anchor logits of the chunk's bins
  float* sA = reinterpret_cast<float*>(sL + kMmaWarps);                       // anchors [n_anchor][E]
  typedef uint64_t Bars[2];
  Bars* bars = reinterpret_cast<Bars*>(sA + kMaxAnchor * kMmaE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, blk = blockIdx.x;
  const int gid = lane >> 2, tig = lane & 3;
  const float* Vb = embed + (size_t)b * TF * kMmaE;

  for (int i = tid; i < n_anchor * kMmaE; i += 32 * kMmaWarps) sA[i] = anchors[i];
  if (lane == 0) {
    uint32_t a0 = (uint32_t)__cvta_generic_to_shared(&bars[warp][0]), a1 = (uint32_t)__cvta_generic_to_shared(&bars[warp][1]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // the rows of this lane's A fragments: r0 = gid, r1 = gid + 8; row r < n_sub is pair r in combinations order (a < b),
  // row n_sub is the all-ones row, rows above are empty
  int pa[2], pb[2], kind[2];                           // kind 0: pair, 1: ones, 2: empty
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = gid + 8 * h;
    pa[h] = 0; pb[h] = 1;
    kind[h] = r < n_sub ? 0 : (r == n_sub ? 1 : 2);
    if (r < n_sub) {
      int s = 0;
      for (int a = 0; a < n_anchor; ++a)
        for (int bq = a + 1; bq < n_anchor; ++bq, ++s)
          if (s == r) { pa[h] = a; pb[h] = bq; }
    }
  }

  // chunk c of the utterance goes to warp slot (c mod n_slots); a slot walks c, c + n_slots, ...
  const long long n_chunks = (TF + kMmaChunk - 1) / kMmaChunk;
  const long long n_slots = (long long)gridDim.x * kMmaWarps;
  const long long slot = (long long)blk * kMmaWarps + warp;
  float acc[3][4];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;

  auto issue = [&](long long c, int buf) {
    const long long bin0 = c * kMmaChunk;
    const int n_here = (int)(TF - bin0 < kMmaChunk ? TF - bin0 : kMmaChunk);
    const uint32_t bytes = (uint32_t)n_here * kMmaE * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(&bars[warp][buf])), "r"(bytes) : "memory");
    bulk_g2s(&sV[warp][buf][0], Vb + (size_t)bin0 * kMmaE, bytes, &bars[warp][buf]);
  };

  if (slot < n_chunks && lane == 0) issue(slot, 0);
  int it = 0;
  for (long long c = slot; c < n_chunks; c += n_slots, ++it) {
    const int buf = it & 1;
    if (c + n_slots < n_chunks && lane == 0) issue(c + n_slots, buf ^ 1);     // that buffer was drained an iteration ago
    mbar_wait_parity(&bars[warp][buf], (uint32_t)(it >> 1) & 1);
    const long long bin0 = c * kMmaChunk;
    const int n_here = (int)(TF - bin0 < kMmaChunk ? TF - bin0 : kMmaChunk);
    float* tile = &sV[warp][buf][0];
    if (n_here < kMmaChunk) {                              // last chunk of the utterance: clear the rows that were not copied
      for (int i = n_here * kMmaE + lane; i < kMmaChunk * kMmaE; i += 32) tile[i] = 0.f;
      __syncwarp();
    }

    // ---- the six anchor logits of bin (bin0 + lane): 5 x LDS.128 of the bin's row (row stride 20 floats: conflict-free)
    {
      float lg[kMaxAnchor];
#pragma unroll
      for (int a = 0; a < kMaxAnchor; ++a) lg[a] = 0.f;
      if (lane < n_here) {
#pragma unroll
        for (int e = 0; e < kMmaE; e += 4) {
          const float4 x = *reinterpret_cast<const float4*>(tile + lane * kMmaE + e);
#pragma unroll
          for (int a = 0; a < kMaxAnchor; ++a)
            if (a < n_anchor) {
              const float4 an = *reinterpret_cast<const float4*>(sA + a * kMmaE + e);
              lg[a] = fmaf(x.x, an.x, lg[a]);
              lg[a] = fmaf(x.y, an.y, lg[a]);
              lg[a] = fmaf(x.z, an.z, lg[a]);
              lg[a] = fmaf(x.w, an.w, lg[a]);
            }
        }
      }
      *reinterpret_cast<float4*>(&sL[warp][lane][0]) = make_float4(lg[0], lg[1], lg[2], lg[3]);
      *reinterpret_cast<float4*>(&sL[warp][lane][4]) = make_float4(lg[4], lg[5], lg[6], lg[7]);
    }
    __syncwarp();

    // ---- acc[16 x 24] += S^T[16 x 32] * [V | 1 | 0][32 x 24], four k8 steps.  The 16 sigmoids of this lane's A fragments
    // (a0 (row gid, bin tig), a1 (row gid+8, bin tig), a2 (row gid, bin tig+4), a3 (row gid+8, bin tig+4) per step) are
    // evaluated first as independent chains, then the 36 MMAs follow.
    float sv[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 8 * ks + tig + 4 * j;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // softmax over the pair (a, b): S_a = 1 / (1 + exp(l_b - l_a))   (eq.6 with C = 2)
          const float d = sL[warp][k][pb[h]] - sL[warp][k][pa[h]];
          float x = __fdividef(1.f, 1.f + __expf(d));
          x = kind[h] == 0 ? x : (kind[h] == 1 ? 1.f : 0.f);
          sv[ks][2 * j + h] = k < n_here ? x : 0.f;
        }
      }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(sv[ks][i], ahi[i], alo[i]);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        // B fragment: b0 (bin tig, column 8 nt + gid), b1 (bin tig + 4, same column); column 20 is the constant 1.
        // Bins beyond the end carry a zero weight in A and zeros here (the tail chunk's buffer was cleared above).
        uint32_t bhi[2], blo[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = 8 * ks + tig + 4 * j;
          float x;
          if (nt < 2) x = tile[k * kMmaE + 8 * nt + gid];
          else x = gid < 4 ? tile[k * kMmaE + 16 + gid] : (gid == 4 ? 1.f : 0.f);
          split_tf32(x, bhi[j], blo[j]);
        }
        mma_tf32(acc[nt], alo, bhi[0], bhi[1]);        // small terms first
        mma_tf32(acc[nt], ahi, blo[0], blo[1]);
        mma_tf32(acc[nt], ahi, bhi[0], bhi[1]);
      }
    }
    __syncwarp();                                        // sL / this buffer are rewritten next iteration
  }

  // ---- fixed-order reduction of the eight warps' fragments, then one partial per block
  // accumulator fragment: c0 (row gid, col 2 tig), c1 (row gid, col 2 tig + 1), c2 / c3 the same for row gid + 8
  // (the fragments go through each warp's own, fully drained, chunk buffer: 16 x 24 floats of its 2 x 640)
  float* red = &sV[warp][0][0];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) {
    const int col = 8 * nt + 2 * tig;
    red[gid * kMmaLd + col] = acc[nt][0];
    red[gid * kMmaLd + col + 1] = acc[nt][1];
    red[(gid + 8) * kMmaLd + col] = acc[nt][2];
    red[(gid + 8) * kMmaLd + col + 1] = acc[nt][3];
  }
  __syncthreads();
  float* dst = part + ((size_t)b * kParts + blk) * (size_t)(n_sub + 1) * kMmaLd;
  for (int i = tid; i < (n_sub + 1) * kMmaLd; i += 32 * kMmaWarps) {
    float sum = sV[0][0][i];
#pragma unroll
    for (int w = 1; w < kMmaWarps; ++w) sum += sV[w][0][i];
    dst[i] = sum;
  }
}

static int launch_anchor2_mma(const float* embed, const float* anchors, float* part, int B, long long TF, int n_anchor,
                              int n_sub, cudaStream_t st) {
  constexpr size_t smem = (size_t)kMmaWarps * (2 * kMmaChunk * kMmaE + kMmaChunk * 8) * 4 + kMaxAnchor * kMmaE * 4 +
                          kMmaWarps * 2 * 8;
  DANET_CUDA(cudaFuncSetAttribute(attractor_anchor2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attractor_anchor2_mma_kernel<<<dim3(kParts, B), 32 * kMmaWarps, smem, st>>>(embed, anchors, part, TF, n_anchor, n_sub);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

// one block per utterance: sum the partials, then the estimator epilogue
template <int MODE>
__global__ void __launch_bounds__(256)
attractor_finalize_kernel(const float* __restrict__ part, int C, int E, int R, int nQ, int n_sub,
                          float denom_add, float* __restrict__ attractors,
                          float* __restrict__ attractor_sets, float* __restrict__ sims,
                          int* __restrict__ choice, float* __restrict__ den_out, int halved, int n_parts) {
  __shared__ float s_sum[kAttMaxRows * (kAttMaxE + 4)];
  __shared__ float s_sim[32];
  __shared__ int s_choice;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ld = nQ * 4, n = R * ld;
  if (halved) {
    // partial rows: n_sub first-member sums, then the all-ones row; expand to the [n_sub][2] rows of eq.7
    const int np = (n_sub + 1) * ld;
    for (int i = tid; i < n_sub * ld; i += 256) {
      const int s = i / ld, e = i % ld;
      float first = 0.f, total = 0.f;
      const float* pp = part + (size_t)b * n_parts * np;
#pragma unroll 8
      for (int pt = 0; pt < n_parts; ++pt) {               // independent loads: unrolled so they are all in flight
        first += __ldg(pp + (size_t)pt * np + s * ld + e);
        total += __ldg(pp + (size_t)pt * np + n_sub * ld + e);
      }
      s_sum[(2 * s) * ld + e] = first;
      s_sum[(2 * s + 1) * ld + e] = total - first;
    }
  } else {
    for (int i = tid; i < n; i += 256) {
      float s = 0.f;
      const float* pp = part + (size_t)b * n_parts * n + i;
#pragma unroll 8
      for (int pt = 0; pt < n_parts; ++pt) s += __ldg(pp + (size_t)pt * n);
      s_sum[i] = s;
    }
  }
  __syncthreads();
  if (den_out)   // raw weight sums per row, kept for the backward pass
    for (int r = tid; r < R; r += 256) den_out[(size_t)b * R + r] = s_sum[r * ld + E];
  if (MODE == MODE_TRUTH) {
    for (int i = tid; i < C * E; i += 256) {
      const int c = i / E, e = i % E;
      attractors[(size_t)b * C * E + i] = s_sum[c * ld + e] / (s_sum[c * ld + E] + denom_add);
    }
  } else if (MODE == MODE_KMEANS) {
    for (int i = tid; i < C * E; i += 256) {   // empty cluster keeps its previous centroid
      const int c = i / E, e = i % E;
      const float cnt = s_sum[c * ld + E];
      if (cnt > 0.f) attractors[(size_t)b * C * E + i] = s_sum[c * ld + e] / cnt;
    }
  } else {
    for (int i = tid; i < R * E; i += 256) {   // eq.7
      const int r = i / E, e = i % E;
      const float v = s_sum[r * ld + e] / s_sum[r * ld + E];
      s_sum[r * ld + e] = v;                   // each element owned by one thread
      if (attractor_sets) attractor_sets[(size_t)b * R * E + i] = v;
    }
    __syncthreads();
    if (tid < n_sub) {                         // eq.8: max over the full C x C Gram (diagonal included)
      float mx = -INFINITY;
      for (int c1 = 0; c1 < C; ++c1)
        for (int c2 = 0; c2 < C; ++c2) {
          float d = 0.f;
          for (int e = 0; e < E; ++e)
            d = fmaf(s_sum[(tid * C + c1) * ld + e], s_sum[(tid * C + c2) * ld + e], d);
          mx = fmaxf(mx, d);
        }
      s_sim[tid] = mx;
      if (sims) sims[(size_t)b * n_sub + tid] = mx;
    }
    __syncthreads();
    if (tid == 0) {                            // eq.9: first argmin
      int best = 0;
      for (int s = 1; s < n_sub; ++s)
        if (s_sim[s] < s_sim[best]) best = s;
      s_choice = best;
      if (choice) choice[b] = best;
    }
    __syncthreads();
    for (int i = tid; i < C * E; i += 256) {
      const int c = i / E, e = i % E;
      attractors[(size_t)b * C * E + i] = s_sum[(s_choice * C + c) * ld + e];
    }
  }
}

static int n_choose_k(int n, int k) {
  if (k < 0 || k > n) return 0;
  long long r = 1;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return (int)r;
}

static int quad_stride(int E) { return ((E / 4 + 1) | 1) * 4; }   // odd quad count: conflict-free float4 rows

static size_t att_smem_bytes(int E, int R, int n_aux) {
  size_t tile = (size_t)kTile * quad_stride(E) + (size_t)kTile * R + n_aux;
  size_t red = (size_t)256 * kMaxAccPerThread * 4;
  return (tile > red ? tile : red) * sizeof(float);
}

template <int MODE, int NQ>
static int launch_partial_rt(AttParams& p, int B, int n_aux, cudaStream_t st) {
  size_t tile = (size_t)kTileRT * p.ldv + (size_t)kTileRT * (p.R | 1) + n_aux;
  size_t red = (size_t)256 * (4 * NQ + 4);
  const size_t smem = (tile > red ? tile : red) * sizeof(float);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "attractor: %zu B of shared memory needed", smem);
  DANET_CUDA(cudaFuncSetAttribute(attractor_partial_rt_kernel<MODE, NQ>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attractor_partial_rt_kernel<MODE, NQ><<<dim3(kParts, B), 256, smem, st>>>(p);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

template <int MODE>
static int launch_partial(AttParams& p, int B, int n_aux, cudaStream_t st) {
  if (p.R <= 256) {   // register-tiled fast path for the embedding sizes in use
    switch (p.E) {
      case 4: return launch_partial_rt<MODE, 1>(p, B, n_aux, st);
      case 12: return launch_partial_rt<MODE, 3>(p, B, n_aux, st);
      case 20: return launch_partial_rt<MODE, 5>(p, B, n_aux, st);
      case 40: return launch_partial_rt<MODE, 10>(p, B, n_aux, st);
      default: break;
    }
  }
  const size_t smem = att_smem_bytes(p.E, p.R, n_aux);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "attractor: %zu B of shared memory needed", smem);
  DANET_CUDA(cudaFuncSetAttribute(attractor_partial_kernel<MODE>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attractor_partial_kernel<MODE><<<dim3(kParts, B), 256, smem, st>>>(p);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

// eq.7 normalisation, eq.8 similarity, eq.9 selection on partial sums produced elsewhere (the fused output projection,
// gemm_tc.cu): `n_parts` partials of [(n_sub + 1)][E + 4] per utterance in the halved two-source layout
int attractor_anchor_finalize(const float* part, int n_parts, int B, int E, int n_sub, float* attractors, float* sets,
                              float* sims, int* choice, float* den, cudaStream_t stream) {
  DANET_REQUIRE(E % 4 == 0 && E <= kAttMaxE && 2 * n_sub <= kAttMaxRows && n_sub <= 20, DANET_E_SHAPE,
                "attractor finalize: E %d n_sub %d", E, n_sub);
  attractor_finalize_kernel<MODE_ANCHOR><<<B, 256, 0, stream>>>(part, 2, E, 2 * n_sub, E / 4 + 1, n_sub, 0.f, attractors, sets,
                                                                sims, choice, den, 1, n_parts);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

static int check_common(const char* who, const float* embed, int B, int C, int TF, int E, int R,
                        void* ws, size_t ws_bytes) {
  DANET_REQUIRE(embed && ws, DANET_E_ARG, "%s: null pointer", who);
  DANET_REQUIRE(B >= 0 && B <= 65535 && C >= 1 && C <= kAttMaxC && TF >= 1 && E >= 4 && E % 4 == 0 &&
                    E <= kAttMaxE && R <= kAttMaxRows,
                DANET_E_SHAPE, "%s: B %d C %d (<=%d) TF %d E %d (multiple of 4, <=%d) rows %d", who, B, C,
                kAttMaxC, TF, E, kAttMaxE, R);
  DANET_REQUIRE((size_t)R * (E / 4 + 1) <= 256 * kMaxAccPerThread, DANET_E_SHAPE,
                "%s: %d rows x E %d exceeds the accumulator budget", who, R, E);
  DANET_REQUIRE(aligned16(embed), DANET_E_ALIGN, "%s: embed must be 16-byte aligned", who);
  DANET_REQUIRE(ws_bytes >= danet_attractor_workspace_bytes(B, R, E), DANET_E_WORKSPACE,
                "%s: workspace %zu < %zu", who, ws_bytes, danet_attractor_workspace_bytes(B, R, E));
  DANET_REQUIRE(aligned16(ws), DANET_E_ALIGN, "%s: workspace must be 16-byte aligned", who);
  return DANET_OK;
}

}  // namespace danet

using namespace danet;

extern "C" size_t danet_attractor_workspace_bytes(int B, int n_acc_rows, int E) {
  if (B < 1 || n_acc_rows < 1 || E < 4) return 256;
  return (size_t)B * kParts * n_acc_rows * (E / 4 + 1) * 4 * sizeof(float);
}

extern "C" int danet_anchor_num_subsets(int n_anchor, int C) { return n_choose_k(n_anchor, C); }

extern "C" int danet_attractor_truth_fwd(const float* embed, const float* src_pwr, const float* mix_pwr,
                                         float* attractors, float* den, int B, int C, int TF, int E, int mode,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("attractor_truth", embed, B, C, TF, E, C, workspace, workspace_bytes);
  if (rc) return rc;
  DANET_REQUIRE(src_pwr && attractors, DANET_E_ARG, "attractor_truth: null pointer");
  DANET_REQUIRE(mode >= 0 && mode <= 2, DANET_E_ARG, "attractor_truth: mode %d", mode);
  DANET_REQUIRE(mode == 0 || mix_pwr, DANET_E_ARG, "attractor_truth: mode %d needs mix_pwr", mode);
  if (B == 0) return DANET_OK;
  AttParams p = {};
  p.embed = embed; p.src_pwr = src_pwr; p.mix_pwr = mix_pwr; p.aux = nullptr;
  p.part = reinterpret_cast<float*>(workspace);
  p.TF = TF; p.C = C; p.E = E; p.R = C; p.nQ = E / 4 + 1; p.ldv = quad_stride(E); p.truth_mode = mode;
  rc = launch_partial<MODE_TRUTH>(p, B, 0, as_stream(stream));
  if (rc) return rc;
  attractor_finalize_kernel<MODE_TRUTH><<<B, 256, 0, as_stream(stream)>>>(
      p.part, C, E, p.R, p.nQ, 0, mode == 0 ? 1.f : kEps, attractors, nullptr, nullptr, nullptr, den, 0, kParts);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_attractor_anchor_fwd(const float* embed, const float* anchors, float* attractors,
                                          float* attractor_sets, float* similarities, int* choice,
                                          float* den, int B, int C, int TF, int E, int n_anchor, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  DANET_REQUIRE(n_anchor >= C && n_anchor <= kMaxAnchor, DANET_E_SHAPE,
                "attractor_anchor: n_anchor %d (C %d .. %d)", n_anchor, C, kMaxAnchor);
  const int P = n_choose_k(n_anchor, C);
  DANET_REQUIRE(P >= 1 && P <= 20, DANET_E_SHAPE, "attractor_anchor: %d subsets (max 20)", P);
  int rc = check_common("attractor_anchor", embed, B, C, TF, E, P * C, workspace, workspace_bytes);
  if (rc) return rc;
  DANET_REQUIRE(anchors && attractors, DANET_E_ARG, "attractor_anchor: null pointer");
  if (B == 0) return DANET_OK;
  AttParams p = {};
  p.embed = embed; p.aux = anchors; p.part = reinterpret_cast<float*>(workspace);
  p.TF = TF; p.C = C; p.E = E; p.R = P * C; p.nQ = E / 4 + 1; p.ldv = quad_stride(E);
  p.n_anchor = n_anchor; p.n_sub = P;
  {   // itertools.combinations order (app/ops.py:287-290)
    int idx[kAttMaxC];
    for (int c = 0; c < C; ++c) idx[c] = c;
    for (int s = 0; s < P; ++s) {
      for (int c = 0; c < C; ++c) p.subsets[s * kAttMaxC + c] = idx[c];
      int i = C - 1;
      while (i >= 0 && idx[i] == n_anchor - C + i) --i;
      if (i < 0) break;
      ++idx[i];
      for (int j = i + 1; j < C; ++j) idx[j] = idx[j - 1] + 1;
    }
  }
  const bool halved = C == 2 && (E == 4 || E == 12 || E == 20 || E == 40);   // fast path instantiations
  // tensor-core path for the benchmark configuration (two sources, E = 20, at most 15 pairs); DANET_ATTRACTOR_SIMT=1
  // forces the register-tiled SIMT kernel (A/B runs, cross-check)
  const bool mma = halved && E == kMmaE && P + 1 <= 16 && !getenv("DANET_ATTRACTOR_SIMT");
  if (mma) {
    rc = launch_anchor2_mma(embed, anchors, p.part, B, TF, n_anchor, P, as_stream(stream));
  } else if (halved) {
    p.R = P + 1;
    rc = launch_partial<MODE_ANCHOR2>(p, B, n_anchor * E, as_stream(stream));
    p.R = P * C;
  } else {
    rc = launch_partial<MODE_ANCHOR>(p, B, n_anchor * E, as_stream(stream));
  }
  if (rc) return rc;
  attractor_finalize_kernel<MODE_ANCHOR><<<B, 256, 0, as_stream(stream)>>>(
      p.part, C, E, p.R, p.nQ, P, 0.f, attractors, attractor_sets, similarities, choice, den, halved ? 1 : 0, kParts);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_attractor_kmeans_fwd(const float* embed, float* centroids, int B, int C, int TF,
                                          int E, int n_iter, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  int rc = check_common("attractor_kmeans", embed, B, C, TF, E, C, workspace, workspace_bytes);
  if (rc) return rc;
  DANET_REQUIRE(centroids && n_iter >= 0, DANET_E_ARG, "attractor_kmeans: centroids %p n_iter %d",
                (void*)centroids, n_iter);
  if (B == 0) return DANET_OK;
  AttParams p = {};
  p.embed = embed; p.aux = centroids; p.part = reinterpret_cast<float*>(workspace);
  p.TF = TF; p.C = C; p.E = E; p.R = C; p.nQ = E / 4 + 1; p.ldv = quad_stride(E);
  for (int it = 0; it < n_iter; ++it) {
    rc = launch_partial<MODE_KMEANS>(p, B, C * E, as_stream(stream));
    if (rc) return rc;
    attractor_finalize_kernel<MODE_KMEANS><<<B, 256, 0, as_stream(stream)>>>(
        p.part, C, E, p.R, p.nQ, 0, 0.f, centroids, nullptr, nullptr, nullptr, nullptr, 0, kParts);
    DANET_LAUNCH_CHECK();
  }
  return DANET_OK;
}
