// K6 (part): backward of the separation head -- what TF autodiff (main.py:357-358) derives for
//   loss = pit_mse(src, mask * mix)                       main.py:281-289, app/ops.py:406-431
//   mask = softmax_c / sigmoid (V . A_c)                  app/modules.py:556-603
//   A    = estimator(V)                                   app/modules.py:390-545
// Gradients do not flow through argmax / argmin / the permutation choice, only through the
// gathered values.  Two streaming passes over the embedding:
//   pass 1  d_attr[b,c,:]  = sum_tf dlogit_c(tf) V[tf,:]                       (reduction)
//   pass 2  d_embed[tf,:]  = sum_c dlogit_c A_c  +  estimator path(d_attr)     (one write of dV)
//           d_anchors      = sum_b sum_tf dl_c(tf) V[tf,:] scattered by the chosen subset
// Both reuse the tile skeleton of attractor.cu: V tile -> smem, one thread per bin computes the
// per-bin coefficients, a register-tiled product does the reduction, deterministic two-stage sums.
#include "common.cuh"

namespace danet {

constexpr int kGTile = 256;
constexpr int kGParts = 32;
constexpr int kGMaxC = 4;
constexpr int kGMaxAnchor = 8;

struct GradParams {
  const float* embed;      // [B][TF][E]
  const float* attr;       // [B][C][E]
  const float2* mix;       // [B][TF]
  const float2* src;       // [B][C][TF]
  const int* perm_idx;     // [B]
  const float* d_attr;     // pass 2: [B][C][E]
  const float* src_pwr;    // truth family: [B][C][TF]
  const float* mix_pwr;    // truth modes 1, 2: [B][TF]
  const float* anchors;    // anchor: [A][E]
  const int* choice;       // anchor: [B]
  const float* den;        // truth: [B][C]; anchor: [B][P][C]
  float* d_embed;          // pass 2 out [B][TF][E]
  float* part;             // partial sums [B][kGParts][C][E+4]
  long long TF;
  int B, C, E, kind, est_mode, n_anchor, n_sub;
  float scale;             // 2 / (B * TF)
  int subsets[20 * kGMaxC];
};

// p-th permutation of 0..C-1 in itertools.permutations (lexicographic) order
__device__ __forceinline__ void nth_permutation(int p, int C, int (&perm)[kGMaxC]) {
  int avail[kGMaxC] = {0, 1, 2, 3};
  int fact = 1;
  for (int i = 2; i < C; ++i) fact *= i;          // (C-1)!
  for (int i = 0; i < C; ++i) {
    const int q = p / fact;
    p -= q * fact;
    perm[i] = avail[q];
    for (int j = q; j + 1 < kGMaxC; ++j) avail[j] = avail[j + 1];
    if (C - 1 - i > 0) fact /= (C - 1 - i);
  }
}

// dlogit_c for one bin: masks from the logits, dL/dmask from the PIT-aligned complex error
template <int NQ>
__device__ __forceinline__ void mask_dlogit(const GradParams& p, int b, long long tf, const float* __restrict__ v,
                                            const float* __restrict__ sA, const int (&inv_perm)[kGMaxC],
                                            float (&dl)[kGMaxC]) {
  constexpr int E = 4 * NQ;
  const int C = p.C;
  float logit[kGMaxC];
#pragma unroll
  for (int c = 0; c < kGMaxC; ++c) logit[c] = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(v + 4 * q);
#pragma unroll
    for (int c = 0; c < kGMaxC; ++c)
      if (c < C) {
        const float4 a = *reinterpret_cast<const float4*>(sA + c * E + 4 * q);
        logit[c] = fmaf(x.x, a.x, logit[c]); logit[c] = fmaf(x.y, a.y, logit[c]);
        logit[c] = fmaf(x.z, a.z, logit[c]); logit[c] = fmaf(x.w, a.w, logit[c]);
      }
  }
  float m[kGMaxC];
  if (p.kind == 0) {
    float mx = logit[0];
#pragma unroll
    for (int c = 1; c < kGMaxC; ++c)
      if (c < C) mx = fmaxf(mx, logit[c]);
    float den = 0.f;
#pragma unroll
    for (int c = 0; c < kGMaxC; ++c)
      if (c < C) { m[c] = expf(logit[c] - mx); den += m[c]; }
    const float inv = 1.f / den;
#pragma unroll
    for (int c = 0; c < kGMaxC; ++c)
      if (c < C) m[c] *= inv;
  } else {
#pragma unroll
    for (int c = 0; c < kGMaxC; ++c)
      if (c < C) m[c] = sigmoidf_(logit[c]);
  }
  // estimate j pairs with source i = inv_perm[j]:  dL/dm_j = scale * (m_j |mix|^2 - Re(conj(mix) src_i))
  const float2 z = __ldg(p.mix + (size_t)b * p.TF + tf);
  const float zz = z.x * z.x + z.y * z.y;
  float dm[kGMaxC];
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < kGMaxC; ++c)
    if (c < C) {
      const float2 s = __ldg(p.src + ((size_t)b * C + inv_perm[c]) * p.TF + tf);
      dm[c] = p.scale * (m[c] * zz - (z.x * s.x + z.y * s.y));
      dot = fmaf(m[c], dm[c], dot);
    }
#pragma unroll
  for (int c = 0; c < kGMaxC; ++c)
    if (c < C) dl[c] = p.kind == 0 ? m[c] * (dm[c] - dot) : dm[c] * m[c] * (1.f - m[c]);
}

// PASS: 1 = d_attr reduction; 2 = d_embed write (+ d_anchors reduction in anchor mode)
template <int PASS, int NQ>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const GradParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int E = 4 * NQ;
  constexpr int LD = E + 4;
  const int C = p.C;
  const int ldv = ((NQ + 1) | 1) * 4;
  const int J = 3 * kGMaxC;                          // coefficient slots per bin (pass 2)
  float* sV = smem;                                  // [kGTile][ldv]
  float* sW = sV + kGTile * ldv;                     // [kGTile][C]   reduction weights
  float* sCoef = sW + kGTile * kGMaxC;               // [kGTile][J]   pass 2: d_embed coefficients
  float* sBasis = sCoef + kGTile * J;                // [J][E]        pass 2: d_embed basis vectors
  float* sA = sBasis + J * E;                        // [C][E] attractors
  float* sAn = sA + kGMaxC * E;                      // [C][E] chosen anchors (anchor mode)
  float* sK = sAn + kGMaxC * E;                      // [C] d_den_c = -(d_attr_c . A_c) / den_c
  const int tid = threadIdx.x, b = blockIdx.y, part = blockIdx.x;
  const long long TF = p.TF;
  const float* Vb = p.embed + (size_t)b * TF * E;

  int inv_perm[kGMaxC] = {0, 1, 2, 3};
  {
    int perm[kGMaxC];
    nth_permutation(p.perm_idx[b], C, perm);
    for (int i = 0; i < C; ++i) inv_perm[perm[i]] = i;
  }
  const bool anchor = p.est_mode == 3;
  int sub[kGMaxC] = {0, 0, 0, 0};
  if (PASS == 2 && anchor) {
    const int ch = p.choice[b];
    for (int c = 0; c < C; ++c) sub[c] = p.subsets[ch * kGMaxC + c];
  }
  for (int i = tid; i < C * E; i += 256) {
    sA[i] = p.attr[(size_t)b * C * E + i];
    if (PASS == 2) {
      const int c = i / E, e = i % E;
      // basis slot c: A_c (coefficient dlogit_c); slot C+c: d_num_c = d_attr_c / den_c (coefficient S_c or
      // weight*onehot); slot 2C+c: chosen anchor (coefficient dl_c, anchor mode)
      const float dAc = p.d_attr[(size_t)b * C * E + i];
      float den;
      if (anchor) den = p.den[((size_t)b * p.n_sub + p.choice[b]) * C + c];
      else den = p.den[(size_t)b * C + c] + (p.est_mode == 0 ? 1.f : kEps);
      sBasis[c * E + e] = sA[i];
      sBasis[(C + c) * E + e] = dAc / den;
      if (anchor) {
        const float an = p.anchors[(size_t)sub[c] * E + e];
        sAn[i] = an;
        sBasis[(2 * C + c) * E + e] = an;
      }
    }
  }
  __syncthreads();
  if (PASS == 2 && anchor && tid < C) {
    // d_den_c = -(d_attr_c . A_c) / den_c
    float d = 0.f;
    for (int e = 0; e < E; ++e) d = fmaf(sBasis[(C + tid) * E + e], sA[tid * E + e], d);
    sK[tid] = -d;
  }

  const bool reduce = PASS == 1 || anchor;
  const int G = 256 / C;
  const int r = tid % C, g = tid / C;
  const bool active = reduce && g < G;
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;

  const long long n_tiles = (TF + kGTile - 1) / kGTile;
  const long long tiles_per = (n_tiles + kGParts - 1) / kGParts;
  const long long t_lo = part * tiles_per;
  const long long t_hi = t_lo + tiles_per < n_tiles ? t_lo + tiles_per : n_tiles;
  for (long long tile = t_lo; tile < t_hi; ++tile) {
    const long long tf0 = tile * kGTile;
    const int n_here = (int)(TF - tf0 < kGTile ? TF - tf0 : kGTile);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(Vb + (size_t)tf0 * E);
      for (int i = tid; i < n_here * NQ; i += 256)
        *reinterpret_cast<float4*>(sV + (i / NQ) * ldv + 4 * (i % NQ)) = __ldg(src + i);
    }
    __syncthreads();
    if (tid < n_here) {
      const long long tf = tf0 + tid;
      const float* v = sV + tid * ldv;
      float dl[kGMaxC];
      mask_dlogit<NQ>(p, b, tf, v, sA, inv_perm, dl);
      if (PASS == 1) {
        for (int c = 0; c < C; ++c) sW[tid * C + c] = dl[c];
      } else {
        float* cf = sCoef + tid * J;
        for (int c = 0; c < C; ++c) cf[c] = dl[c];
        if (anchor) {
          // eq.6 softmax over the chosen subset, then its Jacobian against dS_c = d_num_c . V + d_den_c
          float l[kGMaxC], S[kGMaxC], dS[kGMaxC];
          float mx = -INFINITY;
          for (int c = 0; c < C; ++c) {
            float a = 0.f, d = 0.f;
            for (int e = 0; e < E; ++e) {
              a = fmaf(v[e], sAn[c * E + e], a);
              d = fmaf(v[e], sBasis[(C + c) * E + e], d);
            }
            l[c] = a;
            dS[c] = d + sK[c];
            mx = fmaxf(mx, a);
          }
          float den = 0.f;
          for (int c = 0; c < C; ++c) { S[c] = expf(l[c] - mx); den += S[c]; }
          float dot = 0.f;
          for (int c = 0; c < C; ++c) { S[c] /= den; dot = fmaf(S[c], dS[c], dot); }
          for (int c = 0; c < C; ++c) {
            const float dlc = S[c] * (dS[c] - dot);
            cf[C + c] = S[c];
            cf[2 * C + c] = dlc;
            sW[tid * C + c] = dlc;
          }
        } else {
          // truth family: A_c = sum w [k=c] V / (sum w [k=c] + add)  ->  dV += w [k=c] d_attr_c / den_c
          const float* sp = p.src_pwr + (size_t)b * C * TF + tf;
          int k = 0;
          float best = __ldg(sp);
          for (int c = 1; c < C; ++c) {
            const float x = __ldg(sp + (size_t)c * TF);
            if (x > best) { best = x; k = c; }
          }
          float wt = 1.f;
          if (p.est_mode == 1) wt = __ldg(p.mix_pwr + (size_t)b * TF + tf) > 5.f ? 1.f : 0.f;
          if (p.est_mode == 2) wt = __ldg(p.mix_pwr + (size_t)b * TF + tf);
          for (int c = 0; c < C; ++c) { cf[C + c] = c == k ? wt : 0.f; cf[2 * C + c] = 0.f; }
        }
      }
    } else if (tid < kGTile) {
      for (int c = 0; c < C; ++c) sW[tid * C + c] = 0.f;
    }
    __syncthreads();
    if (active) {
      for (int tfl = g; tfl < n_here; tfl += G) {
        const float wv = sW[tfl * C + r];
        const float* vr = sV + tfl * ldv;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(vr + 4 * q);
          acc[4 * q + 0] = fmaf(wv, x.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(wv, x.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(wv, x.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wv, x.w, acc[4 * q + 3]);
        }
      }
    }
    if (PASS == 2) {
      // d_embed tile: float4 number i of the tile <-> (bin i / NQ, quad i % NQ): coalesced stores
      float4* dst = reinterpret_cast<float4*>(p.d_embed + ((size_t)b * TF + tf0) * E);
      const int nj = anchor ? 3 * C : 2 * C;
      for (int i = tid; i < n_here * NQ; i += 256) {
        const int tfl = i / NQ, q = i % NQ;
        const float* cf = sCoef + tfl * J;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < nj; ++j) {
          const float cj = cf[j];
          const float4 bv = *reinterpret_cast<const float4*>(sBasis + j * E + 4 * q);
          o.x = fmaf(cj, bv.x, o.x); o.y = fmaf(cj, bv.y, o.y); o.z = fmaf(cj, bv.z, o.z); o.w = fmaf(cj, bv.w, o.w);
        }
        dst[i] = o;
      }
    }
  }
  if (!reduce) return;
  __syncthreads();
  float* red = smem;                                 // [G][C][LD]
  if (active) {
    float* o = red + ((size_t)g * C + r) * LD;
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] = acc[e];
    o[E] = 0.f; o[E + 1] = 0.f; o[E + 2] = 0.f; o[E + 3] = 0.f;
  }
  __syncthreads();
  float* dst = p.part + ((size_t)b * kGParts + part) * C * LD;
  for (int i = tid; i < C * LD; i += 256) {
    float sum = red[i];
    for (int gg = 1; gg < G; ++gg) sum += red[(size_t)gg * C * LD + i];
    dst[i] = sum;
  }
}

// pass 1 epilogue: d_attr[b,c,:] = sum over parts
__global__ void __launch_bounds__(256)
head_bwd_sum_kernel(const float* __restrict__ part, int C, int E, float* __restrict__ d_attr) {
  const int b = blockIdx.x, LD = E + 4;
  for (int i = threadIdx.x; i < C * E; i += 256) {
    const int c = i / E, e = i % E;
    float s = 0.f;
    for (int pt = 0; pt < kGParts; ++pt) s += part[(((size_t)b * kGParts + pt) * C + c) * LD + e];
    d_attr[(size_t)b * C * E + i] = s;
  }
}

// pass 2 epilogue (anchor mode): d_anchors[a,:] = sum_b sum_{c : subset_b[c] == a} sum over parts
__global__ void __launch_bounds__(256)
anchor_grad_sum_kernel(const float* __restrict__ part, const int* __restrict__ choice, GradParams p,
                       float* __restrict__ d_anchors) {
  const int E = p.E, C = p.C, LD = E + 4;
  for (int i = threadIdx.x; i < p.n_anchor * E; i += 256) {
    const int a = i / E, e = i % E;
    float s = 0.f;
    for (int b = 0; b < p.B; ++b) {
      const int ch = choice[b];
      for (int c = 0; c < C; ++c)
        if (p.subsets[ch * kGMaxC + c] == a)
          for (int pt = 0; pt < kGParts; ++pt) s += part[(((size_t)b * kGParts + pt) * C + c) * LD + e];
    }
    d_anchors[i] = s;
  }
}

static size_t head_smem_bytes(int E) {
  const int NQ = E / 4, ldv = ((NQ + 1) | 1) * 4, J = 3 * kGMaxC;
  size_t tile = (size_t)kGTile * ldv + (size_t)kGTile * kGMaxC + (size_t)kGTile * J + (size_t)J * E +
                2 * (size_t)kGMaxC * E + 2 * kGMaxC;
  size_t red = (size_t)256 * (E + 4);
  return (tile > red ? tile : red) * sizeof(float);
}

template <int PASS>
static int launch_head(const GradParams& p, cudaStream_t st) {
  const size_t smem = head_smem_bytes(p.E);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "head_bwd: %zu B of shared memory needed", smem);
#define DANET_HEAD_CASE(NQV)                                                                                \
  case 4 * NQV:                                                                                             \
    DANET_CUDA(cudaFuncSetAttribute(head_bwd_kernel<PASS, NQV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                            \
    head_bwd_kernel<PASS, NQV><<<dim3(kGParts, p.B), 256, smem, st>>>(p);                                   \
    break;
  switch (p.E) {
    DANET_HEAD_CASE(1)
    DANET_HEAD_CASE(3)
    DANET_HEAD_CASE(5)
    DANET_HEAD_CASE(10)
    default:
      set_error("head_bwd: embedding size %d not instantiated (4, 12, 20, 40)", p.E);
      return DANET_E_SHAPE;
  }
#undef DANET_HEAD_CASE
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

static int n_choose_k2(int n, int k) {
  if (k < 0 || k > n) return 0;
  long long r = 1;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return (int)r;
}

static int fill_common(GradParams& p, const float* embed, const float* attr, const float* mix, const float* src,
                       const int* perm_idx, int B, int C, int TF, int E, int kind, void* ws, size_t ws_bytes) {
  DANET_REQUIRE(embed && attr && mix && src && perm_idx && ws, DANET_E_ARG, "head_bwd: null pointer");
  DANET_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && C <= kGMaxC && TF >= 1 && E >= 4 && E % 4 == 0, DANET_E_SHAPE,
                "head_bwd: B %d C %d TF %d E %d", B, C, TF, E);
  DANET_REQUIRE(kind == 0 || kind == 1, DANET_E_ARG, "head_bwd: kind %d", kind);
  DANET_REQUIRE(aligned16(embed) && aligned8(mix) && aligned8(src) && aligned16(ws), DANET_E_ALIGN,
                "head_bwd: alignment");
  DANET_REQUIRE(ws_bytes >= danet_head_bwd_workspace_bytes(B, C, E), DANET_E_WORKSPACE, "head_bwd: workspace %zu < %zu",
                ws_bytes, danet_head_bwd_workspace_bytes(B, C, E));
  p = GradParams();
  p.embed = embed; p.attr = attr;
  p.mix = reinterpret_cast<const float2*>(mix);
  p.src = reinterpret_cast<const float2*>(src);
  p.perm_idx = perm_idx;
  p.part = reinterpret_cast<float*>(ws);
  p.TF = TF; p.B = B; p.C = C; p.E = E; p.kind = kind;
  p.scale = 2.f / ((float)B * (float)TF);
  return DANET_OK;
}

}  // namespace danet

using namespace danet;

extern "C" size_t danet_head_bwd_workspace_bytes(int B, int C, int E) {
  if (B < 1 || C < 1 || E < 4) return 256;
  return (size_t)B * kGParts * C * (E + 4) * sizeof(float);
}

extern "C" int danet_head_bwd_attractors(const float* embed, const float* attractors, const float* mix_c64,
                                         const float* src_c64, const int* perm_idx, float* d_attractors, int B,
                                         int C, int TF, int E, int kind, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  GradParams p;
  int rc = fill_common(p, embed, attractors, mix_c64, src_c64, perm_idx, B, C, TF, E, kind, workspace, workspace_bytes);
  if (rc) return rc;
  DANET_REQUIRE(d_attractors, DANET_E_ARG, "head_bwd_attractors: null output");
  rc = launch_head<1>(p, as_stream(stream));
  if (rc) return rc;
  head_bwd_sum_kernel<<<B, 256, 0, as_stream(stream)>>>(p.part, C, E, d_attractors);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_head_bwd_embed(const float* embed, const float* attractors, const float* mix_c64,
                                    const float* src_c64, const int* perm_idx, const float* d_attractors,
                                    int est_mode, const float* src_pwr, const float* mix_pwr, const float* anchors,
                                    const int* choice, const float* den, float* d_embed, float* d_anchors, int B,
                                    int C, int TF, int E, int kind, int n_anchor, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  GradParams p;
  int rc = fill_common(p, embed, attractors, mix_c64, src_c64, perm_idx, B, C, TF, E, kind, workspace, workspace_bytes);
  if (rc) return rc;
  DANET_REQUIRE(d_attractors && d_embed && den, DANET_E_ARG, "head_bwd_embed: null pointer");
  DANET_REQUIRE(est_mode >= 0 && est_mode <= 3, DANET_E_ARG, "head_bwd_embed: est_mode %d", est_mode);
  DANET_REQUIRE(aligned16(d_embed), DANET_E_ALIGN, "head_bwd_embed: d_embed must be 16-byte aligned");
  p.d_attr = d_attractors; p.den = den; p.d_embed = d_embed; p.est_mode = est_mode;
  if (est_mode == 3) {
    DANET_REQUIRE(anchors && choice && d_anchors, DANET_E_ARG, "head_bwd_embed: anchor mode needs anchors, choice, d_anchors");
    DANET_REQUIRE(n_anchor >= C && n_anchor <= kGMaxAnchor, DANET_E_SHAPE, "head_bwd_embed: n_anchor %d", n_anchor);
    const int P = n_choose_k2(n_anchor, C);
    DANET_REQUIRE(P >= 1 && P <= 20, DANET_E_SHAPE, "head_bwd_embed: %d subsets", P);
    p.anchors = anchors; p.choice = choice; p.n_anchor = n_anchor; p.n_sub = P;
    int idx[kGMaxC];
    for (int c = 0; c < C; ++c) idx[c] = c;
    for (int s = 0; s < P; ++s) {     // itertools.combinations order (app/ops.py:287-290)
      for (int c = 0; c < C; ++c) p.subsets[s * kGMaxC + c] = idx[c];
      int i = C - 1;
      while (i >= 0 && idx[i] == n_anchor - C + i) --i;
      if (i < 0) break;
      ++idx[i];
      for (int j = i + 1; j < C; ++j) idx[j] = idx[j - 1] + 1;
    }
  } else {
    DANET_REQUIRE(src_pwr && (est_mode == 0 || mix_pwr), DANET_E_ARG, "head_bwd_embed: truth modes need src_pwr (and mix_pwr)");
    p.src_pwr = src_pwr; p.mix_pwr = mix_pwr;
  }
  rc = launch_head<2>(p, as_stream(stream));
  if (rc) return rc;
  if (est_mode == 3) {
    anchor_grad_sum_kernel<<<1, 256, 0, as_stream(stream)>>>(p.part, choice, p, d_anchors);
    DANET_LAUNCH_CHECK();
  }
  return DANET_OK;
}
