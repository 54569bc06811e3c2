// K6 for WIDE layers (384 < H <= 608: the `lstm-orig` encoder, app/modules.py:140-196) on tcgen05 -- backend 2 of
// danet_lstm_seq_bwd.  The twin of lstm_wide_tc.cu (groups of ncta = ceil(H/32) CTAs launched cooperatively, exchange
// through L2 with flag-in-data words) for the decomposition of lstm_bwd_tc.cu (math in lstm_bwd.cu's header):
//   * CTA r of a group owns hidden units [32r, 32r+32): the 128 gate gradients da[:, own 4 x 32 gate columns] are produced
//     locally and never travel.  The recurrent term  dh = da_{t+1} Wh^T  is split along its 4H-long reduction:
//       partial_r[all units, 8 utterances] = Wh[:, own gate cols] (A: [5 x 128 rows, 128], TENSOR MEMORY resident)
//                                            * da_own^T           (B: [128, 16], shared memory, written by this CTA)
//     40 tcgen05.mma (M128 N16 K16) per step at H = 600;
//   * A holds the weights as ONE fp16 value per element (5 M tiles x 64 columns + 5 accumulators = 480 of the 512 TMEM
//     columns: a lo image does not fit); da travels into the product as an fp16 hi/lo pair ([lo | hi | zero] atoms along N,
//     as in lstm_tc.cu) of da * 2^k, k chosen per CTA and step from the largest |da| (gradients span far more than fp16's
//     exponent range; the partial sums are scaled back before they leave).  Backward products therefore see the weights
//     rounded to 2^-12 relative -- this backend is selected where the forward already carries its state as fp16
//     (Model.train_recurrent_fp16); backend 0 is the exact fp32 kernel;
//   * reduce-scatter through L2: warp w reads TMEM lanes 32w.. of M tile t = the 32 units of peer 4t + w and publishes
//     that slice (32 units x 8 utterances, fp32) as 256 8-byte words {value, step number}; every CTA gathers the ncta slices
//     for its own units with 16-byte loads, polling in rounds, and sums them in source order (deterministic).  Two
//     buffers by step parity: a slot of step n+2 is only rewritten after its producer has gathered every slice of step
//     n+1, whose producers had gathered step n before they published.  Every spin is bounded (NaN instead of a hang).
//     Measured (tools/lstm_wide_bwd_profile.py, B = 32): 7208 cycles per step -- gather 4764, da / scale / staging 772,
//     accumulators 511, TMEM loads + publishing 1160; the last source's slice arrives 2940 cycles after CTA 0 has finished
//     publishing: the CTAs whose units sit in the last M tile get every slice at the END of the sources' publishing windows,
//     lag by that much, and everybody waits for their partial sums.  Tried and dropped: watching one word per source before the full
//     gather (7904), publishing the tiles in an order rotated by the source's rank (7671).
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "lstm_tc_common.cuh"

namespace danet {

using namespace tc;

constexpr int kWbUnits = 32;
constexpr int kWbMinCta = 13, kWbMaxCta = 19;
constexpr int kWbNB = 8;                 // utterances per group
constexpr int kWbEpi = 128;              // 4 epilogue warps: one TMEM lane quadrant each
constexpr int kWbThreads = kWbEpi + 32;  // + the MMA warp
constexpr int kWbBBlock = 2048;          // one K-block (32 reduction indices) of the B operand: [lo 512 | hi 512 | zero 512 | pad]
constexpr uint32_t kWbPollLimit = 1u << 19;

struct LstmWideBwdParams {
  const float* d_out;      // [B][T][n_dir*H]
  float* gates;            // [n_dir][T][B][4H]: in gates [g|i|f|o], out da
  const float* cell_seq;   // [n_dir][T][B][H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  long long ldw;
  uint2* xch;              // [n_dir][groups of this launch][2 parities][dest ncta][src ncta][8 utterances][32 units], zeroed
  int n_dir, T, B, H;
  int group0;
  long long* prof;         // nullable (DANET_LSTM_PROFILE): CTA (0,0,0), epilogue thread 0: cycles summed over the steps spent in
                           // [0] the gather, [1] da / scale / staging, [2] waiting for the accumulators, [3] TMEM loads + publishing;
                           // [4] from the end of its publishing to the arrival of the last source's slice
};

__device__ __forceinline__ void wb_ll_store(uint2* dst, float value, uint32_t flag) {
  const unsigned long long v = ((unsigned long long)flag << 32) | __float_as_uint(value);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 wb_ll_load2(const uint2* src) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}

__global__ void __launch_bounds__(kWbThreads, 1)
lstm_wide_bwd_kernel(const LstmWideBwdParams p) {
  constexpr int NB = kWbNB, UPT = 2;                 // two hidden units per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;
  const int rank = blockIdx.x;
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mt = (ncta * kWbUnits + 127) / 128;    // M tiles over all hidden units (5 at H = 600)

  uint8_t* sB = smem;                                              // [4 K-blocks][kWbBBlock]  (da_own * 2^k)^T, fp16 hi/lo
  uint64_t* b_full = reinterpret_cast<uint64_t*>(sB + 4 * kWbBBlock);
  uint64_t* acc_full = b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_max = reinterpret_cast<float*>(tmem_slot + 2);         // [2][4]: the warps' largest |da| of a step (parity)

  const int unit0 = rank * kWbUnits;
  const int b0 = (p.group0 + grp) * NB;
  uint2* xch = p.xch + ((size_t)dir * gridDim.y + grp) * 2 * (size_t)ncta * ncta * 256;

  if (tid == 0) {
    mbar_init(b_full, kWbEpi / 32);            // one arrival per epilogue warp
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 4 * kWbBBlock / 16; i += kWbThreads) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;                               // tile t: columns [32t, 32t+16)
  const uint32_t tmem_a = tmem_base + 32 * (uint32_t)n_mt;           // tile t: 64 columns at +64t

  // ---- one-time: A[unit_out][k = 4u+g] = Wh[unit_out][g*H + unit0 + u] -> packed fp16 pairs in TMEM ----
  if (warp < 4) {
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    for (int t = 0; t < n_mt; ++t) {
      const int unit_out = 128 * t + tid;
      const bool row_ok = unit_out < H;
      const float* wrow = p.Wh[dir] + (size_t)(row_ok ? unit_out : 0) * p.ldw + unit0;
      for (int u4 = 0; u4 < kWbUnits; u4 += 4) {       // 4 local units -> 16 reduction indices -> 8 columns
        float w[4][4];                                 // [gate][unit]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && unit0 + u4 < H) v = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)g * H + u4));
          w[g][0] = v.x; w[g][1] = v.y; w[g][2] = v.z; w[g][3] = v.w;
        }
        uint32_t hi[8];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu)
#pragma unroll
          for (int gp = 0; gp < 2; ++gp) {             // k = 4u + 2gp, 4u + 2gp + 1
            const __half2 h2 = __floats2half2_rn(w[2 * gp][uu], w[2 * gp + 1][uu]);
            hi[2 * uu + gp] = *reinterpret_cast<const uint32_t*>(&h2);
          }
        tmem_st_32x8(tmem_a + lane_sel + (uint32_t)(64 * t + 2 * u4), hi);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4) {
    // ================= MMA issuer: one partial product per processed step =================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 16);
      const uint64_t bd = umma_desc_k_sw64_sbo512(smem_u32(sB));       // atoms [lo | hi | zero]: columns j and 8 + j add up
      for (int n = 0; n + 1 < T; ++n) {
        mbar_wait(b_full, n & 1);                          // da of this step is staged in sB
        tc_fence_after();
        for (int t = 0; t < n_mt; ++t) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t bk = bd + (uint64_t)((kb * kWbBBlock) >> 4);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16_ts(tmem_acc + 32 * t, tmem_a + (uint32_t)(64 * t + kb * 16 + k * 8), bk + (uint64_t)(k * 2), idesc,
                           (kb | k) != 0);
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int bl = lane % NB, ub = 8 * warp + UPT * (lane / NB);      // utterance, first of my two units
    const int b = b0 + bl, unit = unit0 + ub;
    const bool valid = b < B && unit < H;
    const int outw = p.n_dir * H;
    const int G4 = 4 * H;
    float dc_next[UPT] = {0.f, 0.f};
    bool dead = false;

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
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(grow + g * H));
        o.gt[g][0] = v.x; o.gt[g][1] = v.y;
      }
      const float2 c2 = __ldg(reinterpret_cast<const float2*>(crow));
      const float2 p2 = __ldg(reinterpret_cast<const float2*>(prow));
      const float2 d2 = __ldg(reinterpret_cast<const float2*>(drow));
      o.cc[0] = c2.x; o.cc[1] = c2.y; o.cp[0] = p2.x; o.cp[1] = p2.y; o.dout[0] = d2.x; o.dout[1] = d2.y;
    };
    Operands cur, nxt;
    load_step(0, cur);
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;

    const bool prof_on = p.prof != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    long long pt[4] = {0, 0, 0, 0};
    long long arrive_sum = 0, t_pub_end = 0;
    for (int n = 0; n < T; ++n) {
      const long long c0 = prof_on ? clock64() : 0;
      load_step(n + 1, nxt);                                 // one step ahead
      // everything that does not depend on the recurrent term first:
      //   dc = dht * kA + dc_next,  da = (dc * kI, dc * kB, dc * kC, dht * kD),  dc_next' = dc * kF
      float kA[UPT], kB[UPT], kC[UPT], kD[UPT], kI[UPT], kF[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float gg = cur.gt[0][uu], ig = cur.gt[1][uu], fg = cur.gt[2][uu], og = cur.gt[3][uu];
        const float e2 = __expf(-2.f * fabsf(cur.cc[uu]));
        const float th = copysignf(__fdividef(1.f - e2, 1.f + e2), cur.cc[uu]);
        kA[uu] = og * (1.f - th * th);
        kI[uu] = ig;                                                    // candidate has no tanh
        kB[uu] = gg * ig * (1.f - ig);
        kC[uu] = n + 1 < T ? cur.cp[uu] * fg * (1.f - fg) : 0.f;        // c_{-1} = 0 at the first time step
        kD[uu] = th * og * (1.f - og);
        kF[uu] = fg;
      }
      // dh_rec = sum over source CTAs of their partial slice for my units (published at step n - 1 with flag n)
      float dh[UPT] = {0.f, 0.f};
      if (n > 0) {
        const uint2* src = xch + (size_t)((n - 1) & 1) * ncta * ncta * 256 + (size_t)rank * ncta * 256 + bl * 32 + ub;
        const uint32_t want = (uint32_t)n;
        if (p.prof != nullptr && warp == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
          // profile only: when has the LAST source's slice for this CTA arrived, counted from the end of this CTA's own
          // publishing (lane s watches the last word source s stores for us; the warp reconverges behind the slowest lane):
          // summed over the steps in prof[4]
          if (lane < ncta) {
            const uint2* canary = src + (size_t)lane * 256 + (7 - bl) * 32 + (31 - ub);
            unsigned long long c = 0;
            for (int spin = 0; spin < (1 << 16); ++spin) {
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(canary) : "memory");
              if ((uint32_t)(c >> 32) == want) break;
            }
            arrive_sum += clock64() - t_pub_end;
          }
          __syncwarp();
        }
        uint4 v[kWbMaxCta];
        uint32_t pend = 0;
#pragma unroll
        for (int s = 0; s < kWbMaxCta; ++s)
          if (s < ncta) {
            v[s] = wb_ll_load2(src + (size_t)s * 256);
            pend |= 1u << s;
          }
        uint32_t spins = 0;
        while (pend && !dead) {
#pragma unroll
          for (int s = 0; s < kWbMaxCta; ++s)
            if (((pend >> s) & 1u) && v[s].y == want && v[s].w == want) pend &= ~(1u << s);
          if (!pend) break;
          if (++spins > kWbPollLimit) { dead = true; break; }
#pragma unroll
          for (int s = 0; s < kWbMaxCta; ++s)
            if ((pend >> s) & 1u) v[s] = wb_ll_load2(src + (size_t)s * 256);
        }
#pragma unroll
        for (int s = 0; s < kWbMaxCta; ++s)
          if (s < ncta) {
            dh[0] += __uint_as_float(v[s].x);
            dh[1] += __uint_as_float(v[s].z);
          }
        if (dead) dh[0] = dh[1] = __int_as_float(0x7fc00000);          // a peer never came: visibly invalid
      }
      const long long c1 = prof_on ? clock64() : 0;
      float da[4][UPT];
      float amax = 0.f;
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float dht = dh[uu] + cur.dout[uu];
        const float dc = fmaf(dht, kA[uu], dc_next[uu]);
        dc_next[uu] = dc * kF[uu];
        da[0][uu] = valid ? dc * kI[uu] : 0.f;
        da[1][uu] = valid ? dc * kB[uu] : 0.f;
        da[2][uu] = valid ? dc * kC[uu] : 0.f;
        da[3][uu] = valid ? dht * kD[uu] : 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) amax = fmaxf(amax, fabsf(da[g][uu]));
      }
      auto store_da = [&]() {
        if (!valid) return;
        float* grow = gates_row(n);
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<float2*>(grow + g * H) = make_float2(da[g][0], da[g][1]);
      };
      if (n + 1 < T) {
        // common power-of-two scale of this CTA's da for this step: the largest |da| lands in [2^13, 2^14)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        float* sm = s_max + 4 * (n & 1);
        if (lane == 0) sm[warp] = amax;
        asm volatile("bar.sync 1, %0;" ::"n"(kWbEpi) : "memory");
        const float m = fmaxf(fmaxf(sm[0], sm[1]), fmaxf(sm[2], sm[3]));
        int k = 0;
        if (m > 0.f && m < 3.0e38f) {                        // NaN / Inf pass through unscaled
          k = 14 - (int)((__float_as_uint(m) >> 23) & 0xff) + 126;      // 13 - floor(log2 m) for normal m
          k = max(-60, min(60, k));
        }
        const float scale = __uint_as_float((uint32_t)(127 + k) << 23), inv = __uint_as_float((uint32_t)(127 - k) << 23);
        // stage (da * scale)^T (row = utterance, k = 4u + g: the 4 gates of a unit are 8 contiguous bytes)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) {
          uint2 vh, vl;
          split2_f16(da[0][uu] * scale, da[1][uu] * scale, vh.x, vl.x);
          split2_f16(da[2][uu] * scale, da[3][uu] * scale, vh.y, vl.y);
          const int ul = ub + uu;                                        // local unit 0..31
          uint8_t* dst = sB + (ul >> 3) * kWbBBlock + sw64_offset(bl, 4 * (ul & 7));
          *reinterpret_cast<uint2*>(dst) = vl;                           // [lo | hi | zero]
          *reinterpret_cast<uint2*>(dst + 512) = vh;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full);
        store_da();
        const long long c2 = prof_on ? clock64() : 0;
        // partial products of this step -> one slice per peer, published with flag n + 1
        mbar_wait(acc_full, n & 1);
        const long long c3 = prof_on ? clock64() : 0;
        tc_fence_after();
        uint2* pub = xch + (size_t)(n & 1) * ncta * ncta * 256 + (size_t)rank * 256 + lane;      // + dest * ncta * 256 + utt * 32
        for (int t = 0; t < n_mt; ++t) {
          const int peer = 4 * t + warp;                               // TMEM lanes 32w.. of tile t = peer's units
          if (peer < ncta) {                                           // warp-uniform
            float r[16];
            tmem_ld_32x16(tmem_acc + 32 * t + lane_sel, r);
            uint2* dstw = pub + (size_t)peer * ncta * 256;
#pragma unroll
            for (int j = 0; j < 8; ++j) wb_ll_store(dstw + j * 32, (r[j] + r[8 + j]) * inv, (uint32_t)(n + 1));
          }
        }
        tc_fence_before();
        if (p.prof != nullptr) t_pub_end = clock64();
        if (prof_on) {
          const long long c4 = clock64();
          pt[0] += c1 - c0; pt[1] += c2 - c1; pt[2] += c3 - c2; pt[3] += c4 - c3;
        }
      } else {
        store_da();
      }
      cur = nxt;
    }
    if (prof_on) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p.prof[i] = pt[i];
    }
    if (prof_on) p.prof[4] = arrive_sum;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

static int wb_ncta(int H) { return (H + kWbUnits - 1) / kWbUnits; }

bool lstm_wide_bwd_supported(int H) { return H % 4 == 0 && wb_ncta(H) >= kWbMinCta && wb_ncta(H) <= kWbMaxCta; }

size_t lstm_wide_bwd_workspace_bytes(int n_dir, int B, int H) {
  const size_t ncta = wb_ncta(H);
  return (size_t)n_dir * ((B + kWbNB - 1) / kWbNB) * 2 * ncta * ncta * 256 * sizeof(uint2) + 256;
}

int lstm_wide_bwd(const float* d_out, float* gates, const float* cell_seq, const float* const* host_Wh, long long ldw,
                  int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DANET_REQUIRE(lstm_wide_bwd_supported(H), DANET_E_SHAPE, "lstm_seq_bwd: backend 2 covers 384 < H <= %d (got %d)",
                kWbMaxCta * kWbUnits, H);
  DANET_REQUIRE(aligned16(d_out) && aligned16(gates) && aligned16(cell_seq) && aligned16(host_Wh[0]) && aligned16(workspace),
                DANET_E_ALIGN, "lstm_seq_bwd: buffers must be 16-byte aligned");
  DANET_REQUIRE(workspace_bytes >= lstm_wide_bwd_workspace_bytes(n_dir, B, H), DANET_E_WORKSPACE,
                "lstm_seq_bwd: workspace %zu < %zu", workspace_bytes, lstm_wide_bwd_workspace_bytes(n_dir, B, H));
  const int ncta = wb_ncta(H);
  LstmWideBwdParams p;
  p.d_out = d_out; p.gates = gates; p.cell_seq = cell_seq;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.ldw = ldw; p.n_dir = n_dir; p.T = T; p.B = B; p.H = H;
  const int n_groups = (B + kWbNB - 1) / kWbNB;
  const size_t per_group = (size_t)2 * ncta * ncta * 256;             // LL words per (direction, group)
  p.prof = getenv("DANET_LSTM_PROFILE")                                 // the 256 spare bytes behind the exchange buffer
               ? reinterpret_cast<long long*>(reinterpret_cast<uint2*>(workspace) + (size_t)n_dir * n_groups * per_group)
               : nullptr;
  DANET_CUDA(cudaMemsetAsync(workspace, 0, (size_t)n_dir * n_groups * per_group * sizeof(uint2), stream));
  const size_t smem = 227 * 1024;        // whole SM: the step is latency-bound (lstm_tc.cu)
  DANET_CUDA(cudaFuncSetAttribute(lstm_wide_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_launch = num_sms() / (ncta * n_dir);
  DANET_REQUIRE(per_launch >= 1, DANET_E_SHAPE, "lstm_seq_bwd: one utterance group needs %d resident CTAs", ncta * n_dir);
  if (per_launch > n_groups) per_launch = n_groups;
  for (int g0 = 0; g0 < n_groups; g0 += per_launch) {
    const int ng = n_groups - g0 < per_launch ? n_groups - g0 : per_launch;
    p.group0 = g0;
    p.xch = reinterpret_cast<uint2*>(workspace) + (size_t)g0 * n_dir * per_group;
    void* args[] = {&p};
    DANET_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_wide_bwd_kernel, dim3(ncta, ng, n_dir), dim3(kWbThreads), args,
                                           smem, stream));
  }
  return DANET_OK;
}

}  // namespace danet
