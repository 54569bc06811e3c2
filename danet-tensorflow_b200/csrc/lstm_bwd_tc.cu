// K6 on tcgen05: backward-through-time of the (Bi)LSTM layer as a persistent cluster kernel (the twin of
// lstm_tc.cu; math in lstm_bwd.cu's header).  Per step the recurrent term is  dh = da_{t+1} Wh^T  with a
// 4H-long reduction.  CTA r of a cluster owns hidden units [32r, 32r+32) exactly as in the forward kernel, so
// the 128 gate gradients da[:, own 4x32 gate columns] are produced locally and never travel.  The product
// is split along that reduction instead:
//   partial_r[all units, utterances] = Wh[:, own gate cols] (A: [ceil(H/128)*128, 128], TMEM resident)
//                                      * da_own^T (B: [128, 16], shared memory, written by this CTA)
// (72 tcgen05.mma per step at H = 300: 3 M tiles x 8 K16 steps x bf16x3), and the partials are reduce-
// scattered: warp w reads TMEM lanes 32w.. of M tile t = the 32 units of peer 4t + w and ships that 1 KB
// slice with one cp.async.bulk (DSMEM); every CTA sums the ncta slices it receives for its own units.
// Same bytes on the wire per step as the forward broadcast, no global-memory round trip, no cluster barrier.
// With 8 utterances per cluster the hi and lo halves of da share one B tile along N ([lo | hi | zero] atoms, as in
// lstm_tc.cu generation 2): 48 MMAs per step instead of 72.  Global stores of da are issued AFTER the step's last
// proxy fence (a fence right behind them waits for their acknowledgement: measured 1500 cycles per step), operands of
// the next step are prefetched one step ahead.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace danet {

using namespace tc;

constexpr int kBwUnits = 32;
constexpr int kBwMaxCta = 12;          // H <= 384: 3 M tiles -> 96 accumulator + 384 operand TMEM columns
constexpr int kBwThreads = 160 + 32 * kBwMaxCta;
constexpr int kBwEpi = 128;
constexpr int kBwBBlock = 2048;        // one K-block (32 gate rows) of the B operand, layout as lstm_tc.cu

struct LstmBwdTcParams {
  const float* d_out;      // [B][T][n_dir*H]
  float* gates;            // [n_dir][T][B][4H]: in gates [g|i|f|o], out da
  const float* cell_seq;   // [n_dir][T][B][H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  long long ldw;
  int n_dir, T, B, H;
  long long* prof;         // nullable: per-step phase stamps of CTA (0,0,0) (DANET_LSTM_PROFILE=1), 16 slots per step
};

#define DANET_BPROF(slot)                                                      \
  do {                                                                         \
    if (prof_on) p.prof[(size_t)n * 16 + (slot)] = clock64();                  \
  } while (0)

template <int NB>
__global__ void __launch_bounds__(kBwThreads, 1)
lstm_bwd_tc_kernel(const LstmBwdTcParams p) {
  constexpr int UPT = NB / 4;                       // hidden units per epilogue thread
  constexpr uint32_t kSlice = 32 * NB * 4;          // fp32 partial slice for one peer: 32 units x NB utterances
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;
  const int rank = (int)cluster_ctarank();
  const int bt = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mt = (ncta * kBwUnits + 127) / 128;   // M tiles over all hidden units

  uint8_t* sB = smem;                                              // [4 K-blocks][kBwBBlock]  da_own^T, bf16 hi/lo
  float* sRecv = reinterpret_cast<float*>(sB + 4 * kBwBBlock);     // [2][ncta][32][NB]  partial slices for my units
  float* sStage = sRecv + 2 * kBwMaxCta * 32 * NB;                 // [2][ncta][32][NB]  slices to ship
  uint64_t* p_full = reinterpret_cast<uint64_t*>(sStage + 2 * kBwMaxCta * 32 * NB);   // [2]
  uint64_t* b_full = p_full + 2;
  uint64_t* acc_full = b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int unit0 = rank * kBwUnits;
  const int b0 = bt * NB;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || warp == 4);

  if (tid == 0) {
    mbar_init(p_full + 0, ncta);
    mbar_init(p_full + 1, ncta);
    mbar_init(b_full, kBwEpi / 32);            // one arrival per epilogue warp
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 4 * kBwBBlock / 16; i += kBwThreads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;                               // tile t: columns [32t, 32t+16)
  const uint32_t tmem_a_hi = tmem_base + 32 * (uint32_t)n_mt;        // tile t: 64 columns at +64t
  const uint32_t tmem_a_lo = tmem_a_hi + 64 * (uint32_t)n_mt;

  // ---- one-time: A[unit_out][k = 4u+g] = Wh[unit_out][g*H + unit0 + u] -> packed bf16 hi/lo in TMEM ----
  if (warp < 4) {
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    for (int t = 0; t < n_mt; ++t) {
      const int unit_out = 128 * t + tid;
      const bool row_ok = unit_out < H;
      const float* wrow = p.Wh[dir] + (size_t)(row_ok ? unit_out : 0) * p.ldw + unit0;
      for (int u4 = 0; u4 < kBwUnits; u4 += 4) {       // 4 local units -> 16 reduction indices -> 8 columns
        float w[4][4];                                 // [gate][unit]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && unit0 + u4 < H) v = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)g * H + u4));
          w[g][0] = v.x; w[g][1] = v.y; w[g][2] = v.z; w[g][3] = v.w;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu)
#pragma unroll
          for (int gp = 0; gp < 2; ++gp) {             // k = 4u + 2gp, 4u + 2gp + 1
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(w[2 * gp][uu], h0, l0);
            split_bf16(w[2 * gp + 1][uu], h1, l1);
            hi[2 * uu + gp] = pack_bf16(h0, h1);
            lo[2 * uu + gp] = pack_bf16(l0, l1);
          }
        const uint32_t col = (uint32_t)(64 * t + 2 * u4);            // (4*u4)/2 columns into the tile
        tmem_st_32x8(tmem_a_hi + lane_sel + col, hi);
        tmem_st_32x8(tmem_a_lo + lane_sel + col, lo);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();

  if (warp == 4) {
    // ================= MMA issuer: one partial product per processed step =================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 16);
      // NB 8: K-block = [lo atom | hi atom | zero atom], SBO 512: A_hi x [lo | hi], A_lo x [hi | 0] (columns j and 8 + j
      // of the accumulator add up to the product).  NB 16: [hi | lo] per 8-row atom, SBO 1024, three products.
      const uint64_t bd = NB == 8 ? (umma_desc_k_sw64(smem_u32(sB)) & ~(0x3FFFull << 32)) | ((uint64_t)(512 >> 4) << 32)
                                  : umma_desc_k_sw64(smem_u32(sB));
      for (int n = 0; n + 1 < T; ++n) {
        mbar_wait(b_full, n & 1);                          // da of this step is staged in sB
        DANET_BPROF(8);
        tc_fence_after();
        for (int t = 0; t < n_mt; ++t) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t b_first = bd + (uint64_t)((kb * kBwBBlock) >> 4);      // NB 8: lo|hi   NB 16: hi
            const uint64_t b_second = b_first + (uint64_t)(512 >> 4);             // NB 8: hi|0    NB 16: lo
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t ac = (uint32_t)(64 * t + kb * 16 + k * 8);
              const uint64_t adv = (uint64_t)(k * 2);
              if (NB == 8) {
                umma_bf16_ts(tmem_acc + 32 * t, tmem_a_hi + ac, b_first + adv, idesc, (kb | k) != 0);
                umma_bf16_ts(tmem_acc + 32 * t, tmem_a_lo + ac, b_second + adv, idesc, 1);
              } else {
                umma_bf16_ts(tmem_acc + 32 * t, tmem_a_hi + ac, b_first + adv, idesc, (kb | k) != 0);
                umma_bf16_ts(tmem_acc + 32 * t, tmem_a_hi + ac, b_second + adv, idesc, 1);
                umma_bf16_ts(tmem_acc + 32 * t, tmem_a_lo + ac, b_first + adv, idesc, 1);
              }
            }
          }
        }
        umma_commit(acc_full);
        DANET_BPROF(9);
      }
    }
  } else if (warp < 4) {
    // ================= epilogue warps =================
    const int bl = lane % NB, ub = 8 * warp + UPT * (lane / NB);      // utterance, first of my UPT units
    const int b = b0 + bl, unit = unit0 + ub;
    const bool valid = b < B && unit < H;
    const int outw = p.n_dir * H;
    const int G4 = 4 * H;
    float dc_next[UPT];
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu) dc_next[uu] = 0.f;

    // recurrence-independent operands of one step: gates, c_t, c_{t-1}, d_out
    struct Operands { float gt[4][UPT], cc[UPT], cp[UPT], dout[UPT]; };
    auto gates_row = [&](int n) -> float* {
      const int s = T - 1 - n;
      const int to = dir ? T - 1 - s : s;
      return p.gates + (((size_t)dir * T + to) * B + (valid ? b : 0)) * G4 + (valid ? unit : 0);
    };
    auto load_step = [&](int n, Operands& o) {
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) { o.cc[uu] = 0.f; o.cp[uu] = 0.f; o.dout[uu] = 0.f; }
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) o.gt[g][uu] = 0.f;
      if (!valid || n >= T) return;
      const int s = T - 1 - n;                               // processing index of lstm_bwd.cu
      const int to = dir ? T - 1 - s : s;
      const int tp = dir ? to + 1 : to - 1;
      const float* grow = gates_row(n);
      const float* crow = p.cell_seq + (((size_t)dir * T + to) * B + b) * H + unit;
      const float* prow = p.cell_seq + (((size_t)dir * T + (s > 0 ? tp : to)) * B + b) * H + unit;
      const float* drow = p.d_out + ((size_t)b * T + to) * outw + dir * H + unit;
      if (UPT == 4) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(grow + g * H));
          o.gt[g][0] = v.x; o.gt[g][1] = v.y; o.gt[g][UPT - 2] = v.z; o.gt[g][UPT - 1] = v.w;
        }
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(crow));
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(prow));
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(drow));
        o.cc[0] = c4.x; o.cc[1] = c4.y; o.cc[UPT - 2] = c4.z; o.cc[UPT - 1] = c4.w;
        o.cp[0] = p4.x; o.cp[1] = p4.y; o.cp[UPT - 2] = p4.z; o.cp[UPT - 1] = p4.w;
        o.dout[0] = d4.x; o.dout[1] = d4.y; o.dout[UPT - 2] = d4.z; o.dout[UPT - 1] = d4.w;
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(grow + g * H));
          o.gt[g][0] = v.x; o.gt[g][1] = v.y;
        }
        const float2 c2 = __ldg(reinterpret_cast<const float2*>(crow));
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(prow));
        const float2 d2 = __ldg(reinterpret_cast<const float2*>(drow));
        o.cc[0] = c2.x; o.cc[1] = c2.y; o.cp[0] = p2.x; o.cp[1] = p2.y; o.dout[0] = d2.x; o.dout[1] = d2.y;
      }
      // (c_{-1} = 0 at s == 0 is applied where cp is USED: a select here would wait for the prefetched load)
    };
    Operands cur, nxt;
    load_step(0, cur);

    for (int n = 0; n < T; ++n) {
      DANET_BPROF(0);
      load_step(n + 1, nxt);                                 // one step ahead: HBM latency exceeds the wait below
      DANET_BPROF(1);
      // everything that does not depend on the recurrent term is computed while the partial slices are still in flight:
      //   dc = dht * kA + dc_next,  da = (dc * kI, dc * kB, dc * kC, dht * kD),  dc_next' = dc * kF
      float kA[UPT], kB[UPT], kC[UPT], kD[UPT], kI[UPT], kF[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float gg = cur.gt[0][uu], ig = cur.gt[1][uu], fg = cur.gt[2][uu], og = cur.gt[3][uu];
        const float e2 = __expf(-2.f * fabsf(cur.cc[uu]));             // tanh from ex2/rcp: ~3e-7 absolute
        const float th = copysignf(__fdividef(1.f - e2, 1.f + e2), cur.cc[uu]);
        kA[uu] = og * (1.f - th * th);
        kI[uu] = ig;                                                    // candidate has no tanh
        kB[uu] = gg * ig * (1.f - ig);
        kC[uu] = n + 1 < T ? cur.cp[uu] * fg * (1.f - fg) : 0.f;        // c_{-1} = 0 at the first time step
        kD[uu] = th * og * (1.f - og);
        kF[uu] = fg;
      }
      // dh_rec = sum over source CTAs of their partial slice for my units
      float dh[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) dh[uu] = 0.f;
      if (n > 0) {
        mbar_wait(p_full + ((n - 1) & 1), ((n - 1) >> 1) & 1);
        DANET_BPROF(2);
        const float* rv = sRecv + (size_t)((n - 1) & 1) * kBwMaxCta * 32 * NB + (size_t)ub * NB + bl;
        float part[kBwMaxCta][UPT];                          // all loads in flight, then a fixed-order sum
#pragma unroll
        for (int src = 0; src < kBwMaxCta; ++src)
#pragma unroll
          for (int uu = 0; uu < UPT; ++uu) part[src][uu] = src < ncta ? rv[((size_t)src * 32 + uu) * NB] : 0.f;
#pragma unroll
        for (int src = 0; src < kBwMaxCta; ++src)
#pragma unroll
          for (int uu = 0; uu < UPT; ++uu) dh[uu] += part[src][uu];
      }
      DANET_BPROF(3);
      float da[4][UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float dht = dh[uu] + cur.dout[uu];
        const float dc = fmaf(dht, kA[uu], dc_next[uu]);
        dc_next[uu] = dc * kF[uu];
        da[0][uu] = valid ? dc * kI[uu] : 0.f;
        da[1][uu] = valid ? dc * kB[uu] : 0.f;
        da[2][uu] = valid ? dc * kC[uu] : 0.f;
        da[3][uu] = valid ? dht * kD[uu] : 0.f;
      }
      DANET_BPROF(4);
      // da leaves for HBM only AFTER the staging fence (a proxy fence right behind global stores waits for their
      // acknowledgement: measured ~1500 cycles per step); the MMA phase that follows hides them from the second fence
      auto store_da = [&]() {
        if (!valid) return;
        float* grow = gates_row(n);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (UPT == 4) *reinterpret_cast<float4*>(grow + g * H) = make_float4(da[g][0], da[g][1], da[g][UPT - 2], da[g][UPT - 1]);
          else *reinterpret_cast<float2*>(grow + g * H) = make_float2(da[g][0], da[g][1]);
        }
      };
      if (n + 1 < T) {
        // stage da_own^T (row = utterance, k = 4u + g: the 4 gates of a unit are 8 contiguous bytes)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) {
          uint2 vh, vl;
          split2_bf16(da[0][uu], da[1][uu], vh.x, vl.x);
          split2_bf16(da[2][uu], da[3][uu], vh.y, vl.y);
          const int ul = ub + uu;                                        // local unit 0..31
          uint8_t* dst = sB + (ul >> 3) * kBwBBlock + sw64_offset(bl, 4 * (ul & 7));
          if (NB == 8) {                                                 // [lo | hi | zero]
            *reinterpret_cast<uint2*>(dst) = vl;
            *reinterpret_cast<uint2*>(dst + 512) = vh;
          } else {                                                       // [hi | lo] per 8-row atom
            *reinterpret_cast<uint2*>(dst) = vh;
            *reinterpret_cast<uint2*>(dst + 512) = vl;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full);
        store_da();
        DANET_BPROF(5);
        // partial products of this step -> one slice per peer
        mbar_wait(acc_full, n & 1);
        DANET_BPROF(6);
        tc_fence_after();
        float* stg = sStage + (size_t)(n & 1) * kBwMaxCta * 32 * NB;
        const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
        if (NB == 8) {
          // n_mt <= 3 tiles: tiles 0 and 1 behind one wait, then tile 2 (register budget: 96 per thread)
#pragma unroll
          for (int t0 = 0; t0 < 3; t0 += 2) {
            uint32_t r[2][16];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              const int t = t0 + tt;
              if (t < 3 && t < n_mt && 4 * t + warp < ncta) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[tt][0]), "=r"(r[tt][1]), "=r"(r[tt][2]), "=r"(r[tt][3]), "=r"(r[tt][4]), "=r"(r[tt][5]),
                      "=r"(r[tt][6]), "=r"(r[tt][7]), "=r"(r[tt][8]), "=r"(r[tt][9]), "=r"(r[tt][10]), "=r"(r[tt][11]),
                      "=r"(r[tt][12]), "=r"(r[tt][13]), "=r"(r[tt][14]), "=r"(r[tt][15])
                    : "r"(tmem_acc + 32 * t + lane_sel)
                    : "memory");
              }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              const int t = t0 + tt;
              const int peer = 4 * t + warp;                             // TMEM lanes 32w.. of tile t = peer's units
              if (t < 3 && t < n_mt && peer < ncta) {
                // the slice for my own units goes straight into my receive table (no transport needed)
                float* base = peer == rank ? sRecv + (size_t)(n & 1) * kBwMaxCta * 32 * NB : stg;
                float4* o = reinterpret_cast<float4*>(base + ((size_t)peer * 32 + lane) * NB);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  o[j] = make_float4(__uint_as_float(r[tt][4 * j]) + __uint_as_float(r[tt][8 + 4 * j]),
                                     __uint_as_float(r[tt][4 * j + 1]) + __uint_as_float(r[tt][9 + 4 * j]),
                                     __uint_as_float(r[tt][4 * j + 2]) + __uint_as_float(r[tt][10 + 4 * j]),
                                     __uint_as_float(r[tt][4 * j + 3]) + __uint_as_float(r[tt][11 + 4 * j]));
              }
            }
          }
        } else {
          for (int t = 0; t < n_mt; ++t) {
            const int peer = 4 * t + warp;
            if (peer < ncta) {
              float v[16];
              tmem_ld_32x16(tmem_acc + 32 * t + lane_sel, v);
              float* base = peer == rank ? sRecv + (size_t)(n & 1) * kBwMaxCta * 32 * NB : stg;
              float4* o = reinterpret_cast<float4*>(base + ((size_t)peer * 32 + lane) * NB);
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        asm volatile("bar.arrive 1, %0;" ::"r"(kBwEpi + 32 * ncta) : "memory");
        DANET_BPROF(7);
      }
      if (n + 1 >= T) store_da();
      cur = nxt;
    }
  }
  if (warp >= 5 && warp - 5 < ncta) {
    // ================= sender warps: warp 5+i ships the slice for peer rank+i =================
    const int peer = (rank + (warp - 5)) % ncta;
    const uint32_t peer_dst = mapa(smem_u32(sRecv + (size_t)rank * 32 * NB), peer);    // my slot in the peer's table
    const uint32_t peer_bar = mapa(smem_u32(p_full), peer);
    for (int n = 0; n + 1 < T; ++n) {
      asm volatile("bar.sync 1, %0;" ::"r"(kBwEpi + 32 * ncta) : "memory");
      if (elect_one_sync()) {
        if (peer == rank) {
          mbar_arrive(p_full + (n & 1));                        // own slice: written in place by the epilogue
        } else {
          const uint32_t boff = (uint32_t)(n & 1) * kBwMaxCta * 32 * NB * 4;
          const uint32_t bar = peer_bar + (uint32_t)(n & 1) * 8;
          mbar_arrive_expect_tx_cluster(bar, kSlice);
          dsmem_bulk_copy(peer_dst + boff, smem_u32(sStage + (size_t)(n & 1) * kBwMaxCta * 32 * NB + (size_t)peer * 32 * NB),
                          kSlice, bar);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

template <int NB>
static int launch_lstm_bwd_tc(const LstmBwdTcParams& p, int ncta, cudaStream_t stream) {
  const size_t smem = 4 * kBwBBlock + (size_t)2 * 2 * kBwMaxCta * 32 * NB * 4 + 64 + 1024;
  DANET_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ncta > 8)
    DANET_CUDA(cudaFuncSetAttribute(lstm_bwd_tc_kernel<NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncta, (p.B + NB - 1) / NB, p.n_dir);
  cfg.blockDim = dim3(kBwThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DANET_CUDA(cudaLaunchKernelEx(&cfg, lstm_bwd_tc_kernel<NB>, p));
  return DANET_OK;
}

bool lstm_bwd_tc_supported(int H) { return H % 4 == 0 && (H + kBwUnits - 1) / kBwUnits <= kBwMaxCta; }

int lstm_bwd_tc(const float* d_out, float* gates, const float* cell_seq, const float* const* host_Wh, long long ldw,
                int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DANET_REQUIRE(lstm_bwd_tc_supported(H), DANET_E_SHAPE, "lstm_seq_bwd: tcgen05 backend needs H <= %d", kBwMaxCta * kBwUnits);
  DANET_REQUIRE(aligned16(d_out) && aligned16(gates) && aligned16(cell_seq) && aligned16(host_Wh[0]), DANET_E_ALIGN,
                "lstm_seq_bwd: buffers must be 16-byte aligned");
  const int ncta = (H + kBwUnits - 1) / kBwUnits;
  LstmBwdTcParams p;
  p.d_out = d_out; p.gates = gates; p.cell_seq = cell_seq;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.ldw = ldw; p.n_dir = n_dir; p.T = T; p.B = B; p.H = H;
  p.prof = nullptr;
  if (getenv("DANET_LSTM_PROFILE") && workspace && workspace_bytes >= (size_t)T * 16 * sizeof(long long)) {
    p.prof = reinterpret_cast<long long*>(workspace);
    DANET_CUDA(cudaMemsetAsync(p.prof, 0, (size_t)T * 16 * sizeof(long long), stream));
  }
  const int clusters8 = n_dir * ((B + 7) / 8);
  const char* force = getenv("DANET_LSTM_NB");
  const int nb = force ? atoi(force) : (clusters8 * (ncta + 2) <= num_sms() ? 8 : 16);
  return nb == 8 ? launch_lstm_bwd_tc<8>(p, ncta, stream) : launch_lstm_bwd_tc<16>(p, ncta, stream);
}

}  // namespace danet
