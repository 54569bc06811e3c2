// Dense layers on the 5th-generation tensor cores (backend 1 of danet_linear_fwd):
//   C[M,N] = A[M,K] * W[K,N] (+ bias)   -- app/ops.py:72-89 (lyr_linear)
// fp32 parity (1e-3 end to end after 2004 recurrent steps) rules out plain bf16/tf32 operands
// (measured 5e-3 / 6e-4 on the embedding), so operands are split x = hi + lo into two bf16 and
// three products hi*hi + hi*lo + lo*hi accumulate in fp32 in TMEM ("bf16x3", ~1e-5).
//   1. split kernels write A -> [hi;lo] bf16 [2M, Kp] and W^T -> [hi;lo] bf16 [2N, Kp]
//      (K-major, K zero-padded to a multiple of 64) into the caller's workspace;
//   2. gemm kernel: 128x128 output tiles, TMA (SWIZZLE_128B) -> 3-stage smem ring -> tcgen05.mma kind::f16
//      M128 N128 K16 into one of TWO 128-column TMEM accumulators, epilogue warps tcgen05.ld -> + bias -> global
//      (optionally [B,T] -> [T,B] row remap) while the next tile's MMAs run; CTAs pick up further tiles through
//      cluster launch control (try_cancel of pending CTAs of the same grid).
#include <cuda.h>
#include <atomic>
#include "common.cuh"
#include "tc_common.cuh"
#include "mma_sync.cuh"

namespace danet {

using namespace tc;

constexpr int kTM = 128, kTN = 128, kTK = 64;
constexpr int kMaxOrderedTiles = 64;
constexpr int kStages = 3;
constexpr int kTileBytes = kTM * kTK * 2;                 // 16 KB: one bf16 operand tile
constexpr int kStageBytes = 4 * kTileBytes;               // A_hi, A_lo, B_hi, B_lo
constexpr int kEpiLd = 36;                                // floats per row of an epilogue warp's 32 x 32 transposition tile
constexpr int kEpiTileBytes = 32 * kEpiLd * 4;
constexpr int kGemmSmem = kStages * kStageBytes + 1024 /* alignment slack */ + 512 /* barriers, CLC responses */ +
                          4 * kEpiTileBytes /* one transposition tile per epilogue warp */;

static inline int pad_k(int K) { return (K + kTK - 1) / kTK * kTK; }

// ---- operand split --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ A, long long lda, int M, int K, int Kp, long long lo_row, int row_perm_T,
                  __nv_bfloat16* __restrict__ out) {
  // out[0..M) = hi rows, out[lo_row..lo_row+M) = lo rows, each Kp wide; row_perm_T > 0: source row b*T + t -> row t*B + b
  const long long total = (long long)M * (Kp / 2);
  const int nb = row_perm_T > 0 ? M / row_perm_T : 0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int m = (int)(i / (Kp / 2)), k = (int)(i % (Kp / 2)) * 2;
    const float x0 = k < K ? __ldg(A + (size_t)m * lda + k) : 0.f;
    const float x1 = k + 1 < K ? __ldg(A + (size_t)m * lda + k + 1) : 0.f;
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(x0, h0, l0);
    split_bf16(x1, h1, l1);
    const int mo = row_perm_T > 0 ? (m % row_perm_T) * nb + m / row_perm_T : m;
    *reinterpret_cast<__nv_bfloat162*>(out + (size_t)mo * Kp + k) = __nv_bfloat162(h0, h1);
    *reinterpret_cast<__nv_bfloat162*>(out + (size_t)(lo_row + mo) * Kp + k) = __nv_bfloat162(l0, l1);
  }
}

// src is [K rows, R cols] row-major (reduction index k along rows) -> out[r][k] hi rows [0,R), lo rows
// [R,2R); 32x32 tiles through smem.  perm_T > 0: reduction index k = t*nb + b (time-major) reads source row
// b*perm_T + t + shift (batch-major source), zero when t + shift falls outside [0, perm_T): this pairs a
// [B,T,*] tensor (optionally shifted by one step) with a [T,B,*] one.
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ W, long long ldw, int K, int N, int Kp, int perm_T, int shift,
                       long long lo_row, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nb = perm_T > 0 ? K / perm_T : 0;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    float v = 0.f;
    if (k < K && n < N) {
      if (perm_T > 0) {
        const int t = k / nb + shift, b = k % nb;
        if (t >= 0 && t < perm_T) v = __ldg(W + ((size_t)b * perm_T + t) * ldw + n);
      } else {
        v = __ldg(W + (size_t)k * ldw + n);
      }
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    if (n < N && k < Kp) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][r], h, l);
      out[(size_t)n * Kp + k] = h;
      out[(size_t)(lo_row + n) * Kp + k] = l;
    }
  }
}

// ---- tcgen05 GEMM ------------------------------------------------------------------------
struct GemmParams {
  const float* bias;
  float* C;
  long long ldc;
  int M, N, n_kblocks, T, nb;   // T > 0: logical row b*T+t is stored at row t*nb+b
  int accumulate;               // C += instead of C =
  const float* row_mu;          // nullable [M / rows_per_mu]: C[row, n] -= row_mu[row / rows_per_mu] * col_s[n]
  const float* col_s;           //   (mean-centring of A folded into the epilogue: (x - mu) W = x W - mu colsum(W))
  int rows_per_mu;
  int kb_per_split;             // K blocks per blockIdx.z (split-K: partial sums are added atomically)
  int atomic;                   // 1 when gridDim.z > 1
  // Pipelined hand-over to a consumer kernel that runs CONCURRENTLY (the recurrence that reads these input projections,
  // danet_gemm_split_pipelined): row tile blockIdx.y is m_order[blockIdx.y] -- the tiles the consumer needs first come
  // first -- and every epilogue warp bumps tile_flags[row tile] after its stores (release), so that
  // tile_flags[m] == 4 * column tiles means "rows 128 m .. 128 m + 127 of C are complete and visible".
  int* tile_flags;              // nullable
  unsigned char m_order[kMaxOrderedTiles];
};

// Persistent over output tiles through CLUSTER LAUNCH CONTROL: the grid still has one CTA per tile, but a CTA that
// finishes its tile cancels a not-yet-launched CTA of the same grid (clusterlaunchcontrol.try_cancel) and takes
// over its tile.  Barriers, the TMEM allocation and the tensor-map prefetch are paid once per SM instead of once
// per tile, and with TWO accumulators in TMEM the epilogue of tile k (tcgen05.ld -> global) overlaps the MMAs of
// tile k+1.  Unlike a fixed `tile += gridDim.x` loop this needs no assumption about how many CTAs are resident:
// the recurrent clusters of other stream groups hold whole SMs for ~0.5 ms, and tiles simply go to whichever CTAs run.
constexpr int kClcSlots = 4;

__device__ __forceinline__ void clc_try_cancel(uint32_t resp_addr, uint32_t mbar_addr) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
               ::"r"(resp_addr), "r"(mbar_addr) : "memory");
}
// decode a response: true + the cancelled CTA's blockIdx when a pending CTA was taken over
__device__ __forceinline__ bool clc_query(uint32_t resp_addr, int& x, int& y, int& z) {
  uint32_t ok = 0;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%4];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %3, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n\t}"
      : "+r"(x), "+r"(y), "+r"(z), "=r"(ok)
      : "r"(resp_addr)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(256, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by an offset into the __shared__ array: derived pointers keep the shared address space (LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint64_t* clc_full_bar = tmem_empty_bar + 2;          // [kClcSlots]
  uint64_t* clc_empty_bar = clc_full_bar + kClcSlots;   // [kClcSlots]
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(clc_empty_bar + kClcSlots);     // [kClcSlots] x 16 B (16-byte aligned)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(clc_resp + 16 * kClcSlots);
  float* epi_tiles = reinterpret_cast<float*>(smem + kStages * kStageBytes + 512);      // [4 warps][32][kEpiLd]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar + s, 1);
      mbar_init(tmem_empty_bar + s, 4);          // one arrival per epilogue warp
    }
    for (int s = 0; s < kClcSlots; ++s) {
      mbar_init(clc_full_bar + s, 1);
      mbar_init(clc_empty_bar + s, 5);           // MMA thread + 4 epilogue warps have read the response
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * kTN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every role walks the same tile sequence: tile 0 = this CTA's own block index, tile k+1 = response k
  int tx = blockIdx.x, ty = blockIdx.y, tz = blockIdx.z;

  if (warp == 0) {
    if (elect_one_sync()) {   // ===== TMA producer + tile scheduler =====
      uint32_t kbc = 0;       // K blocks issued so far (ring position)
      for (uint32_t it = 0;; ++it) {
        const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
        // ask for the next tile now, so the answer is here when this tile's loads have been issued
        mbar_wait(clc_empty_bar + slot, cph ^ 1);
        mbar_arrive_expect_tx(clc_full_bar + slot, 16);
        clc_try_cancel(smem_u32(clc_resp + 16 * slot), smem_u32(clc_full_bar + slot));
        const int m0 = (p.tile_flags ? (int)p.m_order[ty] : ty) * kTM, n0 = tx * kTN, kb0 = tz * p.kb_per_split;
        const int nkb = min(p.kb_per_split, p.n_kblocks - kb0);
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int s = kbc % kStages;
          const uint32_t ph = (kbc / kStages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          mbar_arrive_expect_tx(full_bar + s, kStageBytes);
          const int kc = (kb0 + kb) * kTK;
          tma_load_2d(st, &map_a, full_bar + s, kc, m0);                          // A hi
          tma_load_2d(st + kTileBytes, &map_a, full_bar + s, kc, p.M + m0);       // A lo
          tma_load_2d(st + 2 * kTileBytes, &map_b, full_bar + s, kc, n0);         // W^T hi
          tma_load_2d(st + 3 * kTileBytes, &map_b, full_bar + s, kc, p.N + n0);   // W^T lo
        }
        mbar_wait(clc_full_bar + slot, cph);
        const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
        fence_proxy_async_smem();
        if (!more) break;
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {   // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(kTM, kTN);
      uint32_t kbc = 0;
      for (uint32_t it = 0;; ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        const uint32_t tmem_acc = tmem_base + acc * kTN;
        const int kb0 = tz * p.kb_per_split;
        const int nkb = min(p.kb_per_split, p.n_kblocks - kb0);
        mbar_wait(tmem_empty_bar + acc, aph ^ 1);        // the epilogue drained this accumulator (tile it-2)
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int s = kbc % kStages;
          const uint32_t ph = (kbc / kStages) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * kStageBytes);
          const uint64_t a_hi = umma_desc_k_sw128(base), a_lo = umma_desc_k_sw128(base + kTileBytes);
          const uint64_t b_hi = umma_desc_k_sw128(base + 2 * kTileBytes), b_lo = umma_desc_k_sw128(base + 3 * kTileBytes);
#pragma unroll
          for (int k = 0; k < kTK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);   // 32 bytes per K16 step inside the swizzle row
            umma_bf16(tmem_acc, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_bf16(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1);
            umma_bf16(tmem_acc, a_lo + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(empty_bar + s);                    // smem slot reusable once these MMAs retire
        }
        umma_commit(tmem_full_bar + acc);                // accumulator complete
        const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
        mbar_wait(clc_full_bar + slot, cph);
        const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
        fence_proxy_async_smem();
        mbar_arrive(clc_empty_bar + slot);
        if (!more) break;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp q owns TMEM lanes 32q..32q+31 = output rows m0+32q+lane =====
    const int q = warp - 4;
    const bool vec = (p.ldc & 3) == 0;
    for (uint32_t it = 0;; ++it) {
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      const uint32_t tmem_acc = tmem_base + acc * kTN;
      const int tym = p.tile_flags ? (int)p.m_order[ty] : ty;
      const int m0 = tym * kTM, n0 = tx * kTN;
      mbar_wait(tmem_full_bar + acc, aph);
      tc_fence_after();
      const int row = m0 + 32 * q + lane;
      const bool row_ok = row < p.M;
      const size_t orow = p.T > 0 ? (size_t)(row % p.T) * p.nb + row / p.T : (size_t)row;
      float* crow = p.C + orow * p.ldc;
      const float mu = (p.row_mu && row_ok) ? __ldg(p.row_mu + row / p.rows_per_mu) : 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kTN; c0 += 32) {
        if (n0 + c0 >= p.N) break;                       // warp-uniform
        float v[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(32 * q) << 16) + c0, v);
        if (p.atomic) {
          if (row_ok) {
            // split-K: every K slice adds its partial tile (C was zeroed, or holds the value to accumulate onto)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c0 + j;
              if (n < p.N) atomicAdd(crow + n, v[j] + ((p.bias && tz == 0) ? __ldg(p.bias + n) : 0.f));
            }
          }
        } else if (vec && n0 + c0 + 32 <= p.N) {
          // Coalesced stores: a lane holds 32 columns of ITS row, and writing them directly makes every store
          // instruction touch 32 rows (32 half-used sectors).  Transposed through a 32 x 32 shared-memory tile, one
          // instruction writes 4 rows x 128 contiguous bytes: 8x fewer memory transactions per tile.
          float* tile = epi_tiles + (size_t)q * 32 * kEpiLd;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tile + lane * kEpiLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int cj = 4 * (lane & 7);                         // my 4 columns of the chunk
          float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), bb = ss;
          if (p.row_mu) ss = __ldg(reinterpret_cast<const float4*>(p.col_s + n0 + c0 + cj));
          if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + cj));
          const unsigned long long my_ptr = reinterpret_cast<unsigned long long>(crow);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);                   // row of the warp's 32 handled by this lane now
            const unsigned long long rp = __shfl_sync(0xffffffffu, my_ptr, r);
            const float rmu = __shfl_sync(0xffffffffu, mu, r);
            const int rok = __shfl_sync(0xffffffffu, (int)row_ok, r);
            float4 o = *reinterpret_cast<const float4*>(tile + r * kEpiLd + cj);
            o.x = fmaf(-rmu, ss.x, o.x) + bb.x; o.y = fmaf(-rmu, ss.y, o.y) + bb.y;
            o.z = fmaf(-rmu, ss.z, o.z) + bb.z; o.w = fmaf(-rmu, ss.w, o.w) + bb.w;
            if (rok) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(rp) + n0 + c0 + cj);
              if (p.accumulate) {
                const float4 cc = *dst;
                o.x += cc.x; o.y += cc.y; o.z += cc.z; o.w += cc.w;
              }
              *dst = o;
            }
          }
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n < p.N)
              crow[n] = v[j] - (p.row_mu ? mu * __ldg(p.col_s + n) : 0.f) + (p.bias ? __ldg(p.bias + n) : 0.f) +
                        (p.accumulate ? crow[n] : 0.f);
          }
        }
      }
      // accumulator drained (tcgen05.wait::ld ran inside every tmem_ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar + acc);
      if (p.tile_flags) {
        // this warp's 32 rows of the tile are stored: publish (every lane fences its own stores, one lane counts)
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.tile_flags + tym, 1);
      }
      const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
      mbar_wait(clc_full_bar + slot, cph);
      const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(clc_empty_bar + slot);
      if (!more) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * kTN);
}


// ================================================================================================================
// Output projection FUSED with the anchor estimator's accumulation:
//   V[b,t,f,:] = (x[b,t,:] - mean_b) W[:, f E .. f E + E)          app/modules.py:244-259
//   S[b,p,t,f] = softmax over the anchor pair p of <V[b,t,f], anchor>     :513-519 (eq.6, two sources)
//   sum_{t,f} S V  and  sum_{t,f} S  per utterance and pair               :520-523 (eq.7's numerator / denominator)
// in ONE kernel: the embedding tile never has to be read back for the estimator -- it is still in the epilogue warp's
// registers / shared memory when the weighted sums are taken.  Same tcgen05 pipeline as gemm_bf16x3_kernel with
//   * tiles of 128 rows x 160 columns: 160 = 8 whole time-frequency bins of E = 20, so a tile holds complete embedding
//     vectors (UMMA N = 160; two 160-column accumulators in tensor memory);
//   * row tiles cut per utterance (tile (b, mt) = frames 128 mt .. of utterance b; rows beyond T masked), so a tile's
//     partial sums belong to one utterance;
//   * eight epilogue warps: warp (q, hh) owns TMEM lanes 32 q .. and bins 4 hh .. 4 hh + 3 of the tile, two bins at a time:
//     tcgen05.ld -> centring term -> the warp's 32 x 40 shared-memory tile -> (a) coalesced store of V, (b) six anchor
//     logits per (row, bin), (c) the weighted sums on the tensor cores (mma.sync 3xTF32, as attractor_anchor2_mma_kernel:
//     A = pair sigmoids evaluated in fragment layout, B = the tile with a constant-1 column for the denominators);
//   * per tile a fixed-order reduction of the eight warps' fragments and ONE partial [16][24] written to
//     part[b][mt][tx] -- indexed by tile, not by CTA, so the result does not depend on which CTA ran which tile
//     (the tile scheduler is cluster launch control, i.e. dynamic); attractor_finalize sums the partials in order.
constexpr int kPN = 160;
constexpr int kPStages = 2;
constexpr int kPTileA = kTM * kTK * 2;                        // 16 KB
constexpr int kPTileB = kPN * kTK * 2;                        // 20 KB
constexpr int kPStageBytes = 2 * kPTileA + 2 * kPTileB;       // 72 KB: A_hi, A_lo, B_hi, B_lo
constexpr int kPE = 20;
constexpr int kPPair = 2 * kPE;                               // columns handled at once: 2 bins
constexpr int kPEpiWarps = 8;
constexpr int kPThreads = 128 + 32 * kPEpiWarps;
constexpr int kPLdL = 24;                                     // floats per row of the logit tile ([2 bins][8], padded)
constexpr int kPWarpTile = 32 * kPPair * 4;                   // 5 KB
constexpr int kPWarpL = 32 * kPLdL * 4;                       // 3 KB
constexpr int kPRows = 16, kPLd = kPE + 4;                    // partial: 16 rows x 24
constexpr int kPSmem = kPStages * kPStageBytes + 1024 + 512 + kPEpiWarps * (kPWarpTile + kPWarpL) + 8 * kPE * 4;

struct ProjAnchorParams {
  float* V;                 // [B*T][N]
  const float* row_mu;      // [B] mean of x over (T, K), or null
  const float* col_s;       // [N] column sums of W (with row_mu)
  const float* anchors;     // [n_anchor][E]
  float* part;              // [B][mtiles][ntiles][(n_sub+1)][24]
  int B, T, N, F, M, n_kblocks, n_anchor, n_sub, mtiles, ntiles;
};

__global__ void __launch_bounds__(kPThreads, 1)
proj_anchor_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const ProjAnchorParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by an OFFSET into the __shared__ array (not by integer arithmetic on the pointer value), so every pointer
  // derived below keeps its shared address space and compiles to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPStages * kPStageBytes);
  uint64_t* empty_bar = full_bar + kPStages;
  uint64_t* tmem_full_bar = empty_bar + kPStages;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2]
  uint64_t* clc_full_bar = tmem_empty_bar + 2;          // [kClcSlots]
  uint64_t* clc_empty_bar = clc_full_bar + kClcSlots;   // [kClcSlots]
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(clc_empty_bar + kClcSlots);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(clc_resp + 16 * kClcSlots);
  uint8_t* epi = smem + kPStages * kPStageBytes + 512;
  float* sA = reinterpret_cast<float*>(epi + kPEpiWarps * (kPWarpTile + kPWarpL));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full_bar + s, 1);
      mbar_init(tmem_empty_bar + s, kPEpiWarps);
    }
    for (int s = 0; s < kClcSlots; ++s) {
      mbar_init(clc_full_bar + s, 1);
      mbar_init(clc_empty_bar + s, 1 + kPEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.n_anchor * kPE; i += kPThreads) sA[i] = p.anchors[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int tx = blockIdx.x, ty = blockIdx.y, tz = 0;

  if (warp == 0) {
    if (elect_one_sync()) {   // ===== TMA producer + tile scheduler =====
      uint32_t kbc = 0;
      for (uint32_t it = 0;; ++it) {
        const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
        mbar_wait(clc_empty_bar + slot, cph ^ 1);
        mbar_arrive_expect_tx(clc_full_bar + slot, 16);
        clc_try_cancel(smem_u32(clc_resp + 16 * slot), smem_u32(clc_full_bar + slot));
        const int b = ty / p.mtiles, mt = ty % p.mtiles;
        const int m0 = b * p.T + mt * kTM, n0 = tx * kPN;
        for (int kb = 0; kb < p.n_kblocks; ++kb, ++kbc) {
          const int s = kbc % kPStages;
          const uint32_t ph = (kbc / kPStages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* st = smem + s * kPStageBytes;
          mbar_arrive_expect_tx(full_bar + s, kPStageBytes);
          const int kc = kb * kTK;
          tma_load_2d(st, &map_a, full_bar + s, kc, m0);                                  // x hi
          tma_load_2d(st + kPTileA, &map_a, full_bar + s, kc, p.M + m0);                  // x lo
          tma_load_2d(st + 2 * kPTileA, &map_b, full_bar + s, kc, n0);                    // W^T hi
          tma_load_2d(st + 2 * kPTileA + kPTileB, &map_b, full_bar + s, kc, p.N + n0);    // W^T lo
        }
        mbar_wait(clc_full_bar + slot, cph);
        const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
        fence_proxy_async_smem();
        if (!more) break;
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {   // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_bf16(kTM, kPN);
      uint32_t kbc = 0;
      for (uint32_t it = 0;; ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        const uint32_t tmem_acc = tmem_base + acc * kPN;
        mbar_wait(tmem_empty_bar + acc, aph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.n_kblocks; ++kb, ++kbc) {
          const int s = kbc % kPStages;
          const uint32_t ph = (kbc / kPStages) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + s * kPStageBytes);
          const uint64_t a_hi = umma_desc_k_sw128(base), a_lo = umma_desc_k_sw128(base + kPTileA);
          const uint64_t b_hi = umma_desc_k_sw128(base + 2 * kPTileA), b_lo = umma_desc_k_sw128(base + 2 * kPTileA + kPTileB);
#pragma unroll
          for (int k = 0; k < kTK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
            umma_bf16(tmem_acc, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_bf16(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1);
            umma_bf16(tmem_acc, a_lo + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(empty_bar + s);
        }
        umma_commit(tmem_full_bar + acc);
        const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
        mbar_wait(clc_full_bar + slot, cph);
        const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
        fence_proxy_async_smem();
        mbar_arrive(clc_empty_bar + slot);
        if (!more) break;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp (q, hh): TMEM lanes 32 q .. 32 q + 31 = rows, columns 80 hh .. 80 hh + 79 = bins 4 hh .. 4 hh + 3 =====
    const int w = warp - 4, q = w & 3, hh = w >> 2;
    const int gid = lane >> 2, tig = lane & 3;
    float* tile = reinterpret_cast<float*>(epi + (size_t)w * (kPWarpTile + kPWarpL));     // [32][40]
    float* sL = tile + 32 * kPPair;                                                       // [32][24]: logits [bin][8]
    // rows of this lane's A fragments (see attractor_anchor2_mma_kernel): pair index -> (first, second) anchor
    int pa[2], pb[2], kind[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = gid + 8 * h;
      pa[h] = 0; pb[h] = 1;
      kind[h] = r < p.n_sub ? 0 : (r == p.n_sub ? 1 : 2);
      if (r < p.n_sub) {
        int s = 0;
        for (int a = 0; a < p.n_anchor; ++a)
          for (int bq = a + 1; bq < p.n_anchor; ++bq, ++s)
            if (s == r) { pa[h] = a; pb[h] = bq; }
      }
    }
    const int part_ld = (p.n_sub + 1) * kPLd;
    for (uint32_t it = 0;; ++it) {
      const uint32_t acc_i = it & 1, aph = (it >> 1) & 1;
      const uint32_t tmem_acc = tmem_base + acc_i * kPN;
      const int b = ty / p.mtiles, mt = ty % p.mtiles;
      const int n0 = tx * kPN;
      const int t0 = mt * kTM + 32 * q;                       // first frame of this warp's rows
      const int n_here = min(32, max(0, p.T - t0));           // valid rows form a prefix
      const size_t g0 = (size_t)b * p.T + t0;                 // global row of the warp's row 0
      const float mu = p.row_mu ? __ldg(p.row_mu + b) : 0.f;
      float acc[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
      mbar_wait(tmem_full_bar + acc_i, aph);
      tc_fence_after();
#pragma unroll 1
      for (int pp = 0; pp < 2; ++pp) {
        const int c0 = 80 * hh + kPPair * pp;
        const int bin0 = (n0 + c0) / kPE;
        if (bin0 >= p.F) break;                               // warp-uniform: nothing left in this tile for the warp
        const int nbin = min(2, p.F - bin0);
        // the centring term's column sums for this pair: issued before the tensor-memory load so their latency hides under it
        float4 cs[kPPair / 4];
        if (p.row_mu) {
#pragma unroll
          for (int j = 0; j < kPPair; j += 4)
            cs[j / 4] = n0 + c0 + j < p.N ? __ldg(reinterpret_cast<const float4*>(p.col_s + n0 + c0 + j))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float v[kPPair];
        {
          float v32[32], v8[8];
          tmem_ld_32x32(tmem_acc + ((uint32_t)(32 * q) << 16) + c0, v32);
          tmem_ld_32x8(tmem_acc + ((uint32_t)(32 * q) << 16) + c0 + 32, v8);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v32[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[32 + j] = v8[j];
        }
        if (p.row_mu) {
#pragma unroll
          for (int j = 0; j < kPPair; j += 4) {
            v[j] = fmaf(-mu, cs[j / 4].x, v[j]); v[j + 1] = fmaf(-mu, cs[j / 4].y, v[j + 1]);
            v[j + 2] = fmaf(-mu, cs[j / 4].z, v[j + 2]); v[j + 3] = fmaf(-mu, cs[j / 4].w, v[j + 3]);
          }
        }
        __syncwarp();                                         // the previous pair's readers are done with tile / sL
#pragma unroll
        for (int j = 0; j < kPPair; j += 4)
          *reinterpret_cast<float4*>(tile + lane * kPPair + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        // six anchor logits per (row, bin).  (Measured and dropped: the same logits as [32 x 20] x [20 x 8] on mma.sync
        // m16n8k8 3xTF32 -- 18 MMAs per bin in place of 120 FMAs per lane -- 78.9 -> 82.9 us per group of 8.)
#pragma unroll
        for (int bs = 0; bs < 2; ++bs) {
          float lg[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) lg[a] = 0.f;
#pragma unroll
          for (int e = 0; e < kPE; e += 4) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
              if (a < p.n_anchor) {
                const float4 an = *reinterpret_cast<const float4*>(sA + a * kPE + e);
                lg[a] = fmaf(v[kPE * bs + e], an.x, lg[a]);
                lg[a] = fmaf(v[kPE * bs + e + 1], an.y, lg[a]);
                lg[a] = fmaf(v[kPE * bs + e + 2], an.z, lg[a]);
                lg[a] = fmaf(v[kPE * bs + e + 3], an.w, lg[a]);
              }
          }
          *reinterpret_cast<float4*>(sL + lane * kPLdL + 8 * bs) = make_float4(lg[0], lg[1], lg[2], lg[3]);
          *reinterpret_cast<float4*>(sL + lane * kPLdL + 8 * bs + 4) = make_float4(lg[4], lg[5], lg[6], lg[7]);
        }
        __syncwarp();
        // (a) the embedding leaves as whole 160-byte row segments (10 lanes per row)
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int idx = i * 32 + lane, r = idx / 10, part4 = idx - r * 10;
          if (r < n_here && n0 + c0 + 4 * part4 < p.N)
            *reinterpret_cast<float4*>(p.V + (g0 + r) * (size_t)p.N + n0 + c0 + 4 * part4) =
                *reinterpret_cast<const float4*>(tile + r * kPPair + 4 * part4);
        }
        // (b) acc[16 x 24] += S^T[16 x 32 rows] * [V_bin | 1 | 0][32 rows x 24] for each bin of the pair.
        // The 16 pair sigmoids this lane's A fragments hold (4 k8 steps x 4) are evaluated first, as 16 independent
        // chains (the epilogue has two warps per scheduler: instruction-level parallelism is what hides the MUFU and
        // shared-memory latencies), then the 36 MMAs follow with their B fragments.
        for (int bs = 0; bs < nbin; ++bs) {
          float sv[4][4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int k = 8 * ks + tig + 4 * j;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                // softmax over the pair (a, b): S_a = 1 / (1 + exp(l_b - l_a))   (eq.6 with C = 2)
                const float d = sL[k * kPLdL + 8 * bs + pb[h]] - sL[k * kPLdL + 8 * bs + pa[h]];
                float x = __fdividef(1.f, 1.f + __expf(d));
                x = kind[h] == 0 ? x : (kind[h] == 1 ? 1.f : 0.f);
                sv[ks][2 * j + h] = k < n_here ? x : 0.f;
              }
            }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // ONE TF32 product per term (operands rounded to nearest): the sums run over thousands of bins, so the
            // rounding errors average out -- 1e-5 of the attractor scale at T = 501, 4e-5 at 12 frames (emulated on the
            // oracle), against 1e-4 already in the embedding; the 3xTF32 form is kept where the sums are the product
            // (attractor.cu).  A third of the legacy-MMA issue slots, no lo terms to form.
            uint32_t a[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = to_tf32(sv[ks][i]);
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
              // rows beyond the utterance carry a zero weight in A and finite values here: no mask needed
              uint32_t b[2];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int k = 8 * ks + tig + 4 * j;
                float x;
                if (nt < 2) x = tile[k * kPPair + kPE * bs + 8 * nt + gid];
                else x = gid < 4 ? tile[k * kPPair + kPE * bs + 16 + gid] : (gid == 4 ? 1.f : 0.f);
                b[j] = to_tf32(x);
              }
              mma_tf32(acc[nt], a, b[0], b[1]);
            }
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp before the (TMEM-free) reduction
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar + acc_i);
      // fixed-order reduction of the eight warps' fragments -> one partial per tile
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const int col = 8 * nt + 2 * tig;
        tile[gid * kPLd + col] = acc[nt][0];
        tile[gid * kPLd + col + 1] = acc[nt][1];
        tile[(gid + 8) * kPLd + col] = acc[nt][2];
        tile[(gid + 8) * kPLd + col + 1] = acc[nt][3];
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kPEpiWarps) : "memory");
      {
        float* dst = p.part + (((size_t)b * p.mtiles + mt) * p.ntiles + tx) * part_ld;
        for (int i = threadIdx.x - 128; i < part_ld; i += 32 * kPEpiWarps) {
          float sum = 0.f;
#pragma unroll
          for (int ww = 0; ww < kPEpiWarps; ++ww)
            sum += reinterpret_cast<const float*>(epi + (size_t)ww * (kPWarpTile + kPWarpL))[i];
          dst[i] = sum;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kPEpiWarps) : "memory");
      const uint32_t slot = it % kClcSlots, cph = (it / kClcSlots) & 1;
      mbar_wait(clc_full_bar + slot, cph);
      const bool more = clc_query(smem_u32(clc_resp + 16 * slot), tx, ty, tz);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(clc_empty_bar + slot);
      if (!more) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn fn = cached.load(std::memory_order_relaxed);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
      cached.store(fn, std::memory_order_relaxed);
    }
  }
  return fn;
}

// bf16 [rows, Kp] row-major, box = 64 (K) x box_rows, 128-byte swizzle
int make_tensor_map_bf16(CUtensorMap* map, const void* base, long long rows, int Kp, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  DANET_REQUIRE(enc, DANET_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DANET_REQUIRE(r == CUDA_SUCCESS, DANET_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DANET_OK;
}

size_t linear_tc_workspace_bytes(int M, int N, int K) {
  const size_t Kp = pad_k(K);
  return ((size_t)2 * M * Kp * 2 + 1023) / 1024 * 1024 + (size_t)2 * N * Kp * 2 + 1024;
}

struct GemmOperand {
  const float* ptr;
  long long ld;
  int trans;     // 0: stored [rows = own index, cols = K]; 1: stored [rows = K, cols = own index]
  int perm_T;    // trans == 1: K index is time-major over a batch-major source (see split_transpose_kernel);
                 // trans == 0: source row b*perm_T + t is written to row t*B + b (time-major rows of a batch-major source)
  int shift;     // with perm_T (trans == 1): time shift of the source row, zero filled
};

static int split_operand(const GemmOperand& op, int R, int K, int Kp, __nv_bfloat16* out, long long lo_row,
                         cudaStream_t stream) {
  if (op.trans == 0) {
    const long long total = (long long)R * (Kp / 2);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(op.ptr, op.ld, R, K, Kp, lo_row, op.perm_T, out);
  } else {
    dim3 g(Kp / 32, (R + 31) / 32);
    split_transpose_kernel<<<g, 256, 0, stream>>>(op.ptr, op.ld, K, R, Kp, op.perm_T, op.shift, lo_row, out);
  }
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

// the product on operands that are already split: A2 [2M, Kp], B2 [2N, Kp] bf16 (hi rows, then lo rows)
int gemm_tc_split(const __nv_bfloat16* A2, const __nv_bfloat16* B2, const float* bias, float* C, long long ldc,
                  int M, int N, int K, int out_perm_T, int accumulate, cudaStream_t stream, const float* row_mu = nullptr,
                  const float* col_s = nullptr, int rows_per_mu = 1, int* tile_flags = nullptr,
                  const unsigned char* m_order = nullptr) {
  DANET_REQUIRE(aligned16(C) && (!bias || aligned16(bias)), DANET_E_ALIGN, "gemm: C and bias must be 16-byte aligned");
  DANET_REQUIRE(aligned16(A2) && aligned16(B2), DANET_E_ALIGN, "gemm: split operands must be 16-byte aligned");
  const int Kp = pad_k(K);
  CUtensorMap map_a, map_b;
  int rc = make_tensor_map_bf16(&map_a, A2, 2ll * M, Kp, kTM);
  if (rc) return rc;
  rc = make_tensor_map_bf16(&map_b, B2, 2ll * N, Kp, kTN);
  if (rc) return rc;
  GemmParams p;
  p.bias = bias; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.n_kblocks = Kp / kTK;
  p.T = out_perm_T; p.nb = out_perm_T > 0 ? M / out_perm_T : 0;
  p.accumulate = accumulate;
  p.row_mu = row_mu; p.col_s = col_s; p.rows_per_mu = rows_per_mu > 0 ? rows_per_mu : 1;
  p.tile_flags = tile_flags;
  for (int i = 0; i < kMaxOrderedTiles; ++i) p.m_order[i] = m_order ? m_order[i] : (unsigned char)i;
  DANET_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
  dim3 grid((N + kTN - 1) / kTN, (M + kTM - 1) / kTM);
  DANET_REQUIRE(grid.y <= 65535, DANET_E_SHAPE, "gemm: M %d too large", M);
  // split-K when the output has too few tiles to fill the SMs and the reduction is long (dW = X^T dY:
  // 30-50 tiles, K = B*T): slices of >= 8 K blocks, partial tiles added with red.global.add.f32
  {
    const int tiles = (int)(grid.x * grid.y), sms = num_sms();
    int splits = 1;
    if (tiles * 4 <= sms * 3 && p.n_kblocks >= 16 && !row_mu && !tile_flags) {
      splits = (sms + tiles / 2) / tiles;
      const int max_splits = p.n_kblocks / 8;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
    p.kb_per_split = (p.n_kblocks + splits - 1) / splits;
    splits = (p.n_kblocks + p.kb_per_split - 1) / p.kb_per_split;
    grid.z = splits;
    p.atomic = splits > 1;
    if (p.atomic && !accumulate) {
      if (ldc == N) {
        DANET_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), stream));
      } else {
        DANET_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, stream));
      }
    }
  }
  gemm_bf16x3_kernel<<<grid, 256, kGemmSmem, stream>>>(map_a, map_b, p);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

// C[M,N] (ldc) (+)= A'[M,K] * B'[K,N] (+ bias), A' / B' described by GemmOperand
int gemm_tc(const GemmOperand& A, const GemmOperand& B, const float* bias, float* C, long long ldc, int M, int N,
            int K, int out_perm_T, int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DANET_REQUIRE(workspace, DANET_E_ARG, "gemm: the tcgen05 backend needs a workspace");
  DANET_REQUIRE(workspace_bytes >= linear_tc_workspace_bytes(M, N, K), DANET_E_WORKSPACE,
                "gemm: workspace %zu < %zu", workspace_bytes, linear_tc_workspace_bytes(M, N, K));
  const int Kp = pad_k(K);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* A2 = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* B2 = reinterpret_cast<__nv_bfloat16*>(ws + ((size_t)2 * M * Kp * 2 + 1023) / 1024 * 1024);
  int rc = split_operand(A, M, K, Kp, A2, M, stream);
  if (rc) return rc;
  // B' is [K,N]: "own index" N; stored [K,N] means rows = K, i.e. the transposing split
  GemmOperand Bt = B;
  Bt.trans = B.trans ? 0 : 1;
  rc = split_operand(Bt, N, K, Kp, B2, N, stream);
  if (rc) return rc;
  return gemm_tc_split(A2, B2, bias, C, ldc, M, N, K, out_perm_T, accumulate, stream);
}

int linear_tc_fwd(const float* A, long long lda, const float* W, long long ldw, const float* bias, float* C,
                  int M, int N, int K, int time_major_T, void* workspace, size_t workspace_bytes,
                  cudaStream_t stream) {
  GemmOperand a = {A, lda, 0, 0, 0}, b = {W, ldw, 0, 0, 0};
  return gemm_tc(a, b, bias, C, N, M, N, K, time_major_T, 0, workspace, workspace_bytes, stream);
}

int attractor_anchor_finalize(const float* part, int n_parts, int B, int E, int n_sub, float* attractors, float* sets,
                              float* sims, int* choice, float* den, cudaStream_t stream);

static inline int proj_anchor_mtiles(int T) { return (T + kTM - 1) / kTM; }
static inline int proj_anchor_ntiles(int N) { return (N + kPN - 1) / kPN; }

}  // namespace danet

using namespace danet;

extern "C" size_t danet_proj_anchor_workspace_bytes(int B, int T, int F, int E) {
  if (B < 1 || T < 1 || F < 1 || E < 1) return 256;
  return (size_t)B * proj_anchor_mtiles(T) * proj_anchor_ntiles(F * E) * kPRows * kPLd * sizeof(float);
}

extern "C" int danet_proj_anchor_fwd(const void* A2, const void* W2, const float* row_mu, const float* col_s,
                                     const float* anchors, float* embed, float* attractors, float* attractor_sets,
                                     float* similarities, int* choice, int B, int T, int F, int E, int K, int n_anchor,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  DANET_REQUIRE(A2 && W2 && anchors && embed && attractors && workspace, DANET_E_ARG, "proj_anchor: null pointer");
  DANET_REQUIRE(E == kPE, DANET_E_SHAPE, "proj_anchor: the fused kernel is built for E = %d (got %d); use danet_gemm_split + "
                "danet_attractor_anchor_fwd", kPE, E);
  DANET_REQUIRE(B >= 0 && T >= 1 && F >= 1 && K >= 1, DANET_E_SHAPE, "proj_anchor: B %d T %d F %d K %d", B, T, F, K);
  const int n_sub = n_anchor * (n_anchor - 1) / 2;
  DANET_REQUIRE(n_anchor >= 2 && n_anchor <= 6 && n_sub + 1 <= kPRows, DANET_E_SHAPE,
                "proj_anchor: n_anchor %d (two sources, 2 .. 6 anchors)", n_anchor);
  DANET_REQUIRE(!row_mu || (col_s && aligned16(col_s)), DANET_E_ARG, "proj_anchor: row_mu needs col_s (16-byte aligned)");
  DANET_REQUIRE(aligned16(A2) && aligned16(W2) && aligned16(embed) && aligned16(workspace), DANET_E_ALIGN,
                "proj_anchor: operands, embed and workspace must be 16-byte aligned");
  DANET_REQUIRE(workspace_bytes >= danet_proj_anchor_workspace_bytes(B, T, F, E), DANET_E_WORKSPACE,
                "proj_anchor: workspace %zu < %zu", workspace_bytes, danet_proj_anchor_workspace_bytes(B, T, F, E));
  if (B == 0) return DANET_OK;
  const int M = B * T, N = F * E, Kp = pad_k(K);
  CUtensorMap map_a, map_b;
  int rc = make_tensor_map_bf16(&map_a, A2, 2ll * M, Kp, kTM);
  if (rc) return rc;
  rc = make_tensor_map_bf16(&map_b, W2, 2ll * N, Kp, kPN);
  if (rc) return rc;
  ProjAnchorParams p;
  p.V = embed; p.row_mu = row_mu; p.col_s = col_s; p.anchors = anchors;
  p.part = reinterpret_cast<float*>(workspace);
  p.B = B; p.T = T; p.N = N; p.F = F; p.M = M; p.n_kblocks = Kp / kTK; p.n_anchor = n_anchor; p.n_sub = n_sub;
  p.mtiles = proj_anchor_mtiles(T); p.ntiles = proj_anchor_ntiles(N);
  DANET_REQUIRE((long long)B * p.mtiles <= 65535, DANET_E_SHAPE, "proj_anchor: B %d x %d row tiles exceeds the grid", B, p.mtiles);
  DANET_CUDA(cudaFuncSetAttribute(proj_anchor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
  dim3 grid(p.ntiles, B * p.mtiles);
  proj_anchor_kernel<<<grid, kPThreads, kPSmem, as_stream(stream)>>>(map_a, map_b, p);
  DANET_LAUNCH_CHECK();
  return attractor_anchor_finalize(p.part, p.mtiles * p.ntiles, B, E, n_sub, attractors, attractor_sets, similarities,
                                   choice, nullptr, as_stream(stream));
}

namespace danet {

}  // namespace danet

using namespace danet;

extern "C" size_t danet_gemm_workspace_bytes(int M, int N, int K) {
  if (M < 1 || N < 1 || K < 1) return 256;
  return linear_tc_workspace_bytes(M, N, K);
}

extern "C" int danet_gemm(const float* A, long long lda, int transA, int permA_T, int shiftA, const float* B,
                          long long ldb, int transB, const float* bias, float* C, long long ldc, int M, int N,
                          int K, int out_perm_T, int accumulate, void* workspace, size_t workspace_bytes,
                          void* stream) {
  DANET_REQUIRE(A && B && C, DANET_E_ARG, "gemm: null pointer");
  DANET_REQUIRE(M >= 0 && N >= 1 && K >= 1 && ldc >= N, DANET_E_SHAPE, "gemm: M %d N %d K %d ldc %lld", M, N, K, ldc);
  DANET_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N), DANET_E_SHAPE, "gemm: lda %lld ldb %lld", lda, ldb);
  DANET_REQUIRE(permA_T == 0 || (transA && K % permA_T == 0 && shiftA >= -1 && shiftA <= 1), DANET_E_ARG,
                "gemm: permA_T %d needs transA and K %% T == 0, shift in [-1,1]", permA_T);
  DANET_REQUIRE(permA_T > 0 || shiftA == 0, DANET_E_ARG, "gemm: shiftA needs permA_T");
  DANET_REQUIRE(out_perm_T >= 0 && (out_perm_T == 0 || M % out_perm_T == 0), DANET_E_SHAPE,
                "gemm: M %d is not a multiple of out_perm_T %d", M, out_perm_T);
  if (M == 0) return DANET_OK;
  GemmOperand a = {A, lda, transA ? 1 : 0, permA_T, shiftA}, b = {B, ldb, transB ? 1 : 0, 0, 0};
  return gemm_tc(a, b, bias, C, ldc, M, N, K, out_perm_T, accumulate ? 1 : 0, workspace, workspace_bytes,
                 as_stream(stream));
}

extern "C" size_t danet_split_operand_bytes(int rows, int K) {
  if (rows < 1 || K < 1) return 256;
  return (size_t)2 * rows * pad_k(K) * 2;
}

extern "C" int danet_split_operand(const float* X, long long ld, int stored_k_major_rows, int rows, int K,
                                   void* out_bf16, int row0, int rows_total, void* stream) {
  DANET_REQUIRE(X && out_bf16, DANET_E_ARG, "split_operand: null pointer");
  DANET_REQUIRE(rows >= 1 && K >= 1 && row0 >= 0 && rows_total >= row0 + rows, DANET_E_SHAPE,
                "split_operand: rows %d K %d row0 %d rows_total %d", rows, K, row0, rows_total);
  DANET_REQUIRE(aligned16(out_bf16), DANET_E_ALIGN, "split_operand: out must be 16-byte aligned");
  const int Kp = pad_k(K);
  GemmOperand op = {X, ld, stored_k_major_rows ? 1 : 0, 0, 0};
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16) + (size_t)row0 * Kp;
  return split_operand(op, rows, K, Kp, out, rows_total, as_stream(stream));
}

// danet_split_operand for a batch-major activation [B*T, K] whose operand rows are wanted TIME-major (row t*B + b): the A
// operand of danet_gemm_split_pipelined(rows_time_major = 1) for the first recurrent layer
extern "C" int danet_split_operand_time_major(const float* X, long long ld, int rows, int K, int T, void* out_bf16,
                                              void* stream) {
  DANET_REQUIRE(X && out_bf16, DANET_E_ARG, "split_operand_time_major: null pointer");
  DANET_REQUIRE(rows >= 1 && K >= 1 && T >= 1 && rows % T == 0 && ld >= K, DANET_E_SHAPE,
                "split_operand_time_major: rows %d K %d T %d ld %lld", rows, K, T, ld);
  DANET_REQUIRE(aligned16(out_bf16), DANET_E_ALIGN, "split_operand_time_major: out must be 16-byte aligned");
  GemmOperand op = {X, ld, 0, T, 0};
  return split_operand(op, rows, K, pad_k(K), reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, as_stream(stream));
}

extern "C" int danet_split_operand_paired(const float* X, long long ld, int rows, int K, int perm_T, int shift,
                                          void* out_bf16, int row0, int rows_total, void* stream) {
  DANET_REQUIRE(X && out_bf16, DANET_E_ARG, "split_operand_paired: null pointer");
  DANET_REQUIRE(rows >= 1 && K >= 1 && row0 >= 0 && rows_total >= row0 + rows && ld >= rows, DANET_E_SHAPE,
                "split_operand_paired: rows %d K %d row0 %d rows_total %d ld %lld", rows, K, row0, rows_total, ld);
  DANET_REQUIRE(perm_T >= 1 && K % perm_T == 0 && shift >= -1 && shift <= 1, DANET_E_ARG,
                "split_operand_paired: K %d must be a multiple of perm_T %d, shift in [-1,1]", K, perm_T);
  DANET_REQUIRE(aligned16(out_bf16), DANET_E_ALIGN, "split_operand_paired: out must be 16-byte aligned");
  const int Kp = pad_k(K);
  GemmOperand op = {X, ld, 1, perm_T, shift};
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16) + (size_t)row0 * Kp;
  return split_operand(op, rows, K, Kp, out, rows_total, as_stream(stream));
}

extern "C" int danet_gemm_split(const void* A2, const void* B2, const float* bias, const float* row_mu,
                                const float* col_s, int rows_per_mu, float* C, long long ldc, int M, int N, int K,
                                int out_perm_T, int accumulate, void* stream) {
  DANET_REQUIRE(A2 && B2 && C, DANET_E_ARG, "gemm_split: null pointer");
  DANET_REQUIRE(M >= 0 && N >= 1 && K >= 1 && ldc >= N, DANET_E_SHAPE, "gemm_split: M %d N %d K %d ldc %lld", M, N, K, ldc);
  DANET_REQUIRE(out_perm_T >= 0 && (out_perm_T == 0 || M % out_perm_T == 0), DANET_E_SHAPE,
                "gemm_split: M %d is not a multiple of out_perm_T %d", M, out_perm_T);
  DANET_REQUIRE(!row_mu || (col_s && rows_per_mu >= 1 && aligned16(col_s)), DANET_E_ARG,
                "gemm_split: row_mu needs col_s (16-byte aligned) and rows_per_mu >= 1");
  if (M == 0) return DANET_OK;
  return gemm_tc_split(reinterpret_cast<const __nv_bfloat16*>(A2), reinterpret_cast<const __nv_bfloat16*>(B2), bias, C,
                       ldc, M, N, K, out_perm_T, accumulate ? 1 : 0, as_stream(stream), row_mu, col_s, rows_per_mu);
}

// Input projections of a recurrent layer, handed to the recurrence tile by tile (see GemmParams::tile_flags): same product
// as danet_gemm_split with out_perm_T = T (rows b*T + t of A land at row t*B + b of C), the row tiles issued in the order
// a forward AND a backward scan over time consume them (tiles holding the first / last frames of an utterance first), and
// tile_flags[m] counting the finished (column tile, epilogue warp) pairs of row tile m.
// rows_time_major: A's rows are already t*B + b (danet_lstm_seq_fwd_pipelined's out_split_time_major, or
// danet_split_operand_time_major): C = A B^T row for row, tiles issued alternately from the two ends of time -- the scans
// can start once TWO row tiles are done instead of every tile that holds some utterance's first or last frame.
extern "C" int danet_gemm_split_pipelined(const void* A2, const void* B2, const float* bias, float* C, long long ldc, int M,
                                          int N, int K, int T, int rows_time_major, int* tile_flags, int* flag_need,
                                          void* stream) {
  DANET_REQUIRE(A2 && B2 && C && tile_flags && flag_need, DANET_E_ARG, "gemm_split_pipelined: null pointer");
  DANET_REQUIRE(M >= 1 && N >= 1 && K >= 1 && ldc >= N && T >= 1 && M % T == 0, DANET_E_SHAPE,
                "gemm_split_pipelined: M %d N %d K %d ldc %lld T %d", M, N, K, ldc, T);
  const int mtiles = (M + kTM - 1) / kTM;
  DANET_REQUIRE(mtiles <= kMaxOrderedTiles, DANET_E_SHAPE, "gemm_split_pipelined: %d row tiles (max %d, i.e. M <= %d)", mtiles,
                kMaxOrderedTiles, kMaxOrderedTiles * kTM);
  // key of a tile = the earliest scan step (from either end of time) that touches one of its rows
  int key[kMaxOrderedTiles];
  unsigned char order[kMaxOrderedTiles];
  for (int m = 0; m < kMaxOrderedTiles; ++m) order[m] = (unsigned char)m;
  const int nb = M / T;
  for (int m = 0; m < mtiles; ++m) {
    int best = T;
    for (int r = m * kTM; r < (m + 1) * kTM && r < M; ++r) {
      const int t = rows_time_major ? r / nb : r % T, d = t < T - 1 - t ? t : T - 1 - t;
      if (d < best) best = d;
    }
    key[m] = best;
  }
  for (int i = 1; i < mtiles; ++i) {            // stable insertion sort by key
    const unsigned char v = order[i];
    int j = i - 1;
    while (j >= 0 && key[order[j]] > key[v]) { order[j + 1] = order[j]; --j; }
    order[j + 1] = v;
  }
  *flag_need = 4 * ((N + kTN - 1) / kTN);
  return gemm_tc_split(reinterpret_cast<const __nv_bfloat16*>(A2), reinterpret_cast<const __nv_bfloat16*>(B2), bias, C, ldc,
                       M, N, K, rows_time_major ? 0 : T, 0, as_stream(stream), nullptr, nullptr, 1, tile_flags, order);
}
