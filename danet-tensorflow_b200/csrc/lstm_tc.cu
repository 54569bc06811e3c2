// K2b on tcgen05 (backend 1 of danet_lstm_seq_fwd): the (Bi)LSTM recurrence of
// Model.lyr_lstm (main.py:76-132) / ops.lyr_lstm_flat (app/ops.py:139-147) as a persistent
// thread-block-cluster kernel.
//
// One cluster = one direction x 16 utterances.  CTA r of the cluster owns hidden units
// [32r, 32r+32).  Its 128 gate rows (4 gates x 32 units, row m = 4*unit + gate) of Wh^T stay
// resident in TENSOR MEMORY for the whole sequence as packed bf16 hi/lo pairs (2 x K/2 = 320
// of the 512 TMEM columns at H = 300, K padded to 320), so the A operand never touches the
// shared-memory port (a shared-memory resident A costs 4 KB of reads per MMA: measured 65
// cycles per MMA, 3900 cycles per step).  Per step the CTA issues
//   D[128 gate rows, 16 utterances] = Wh^T[128, K] * h_{t-1}^T[K, 16]
// as 60 tcgen05.mma (A from TMEM, B from smem, M128 N16 K16; bf16x3: hi*hi + hi*lo + lo*hi,
// fp32 accumulate in TMEM).  The epilogue warps read the accumulator, add the hoisted input
// projection, apply  c = sig(i)*g + sig(f)*c ; h = sig(o)*tanh(c)  (candidate WITHOUT tanh),
// write h_t to the output, and broadcast their 32-unit slice of h_t (bf16 hi/lo, already in
// the UMMA K-major SWIZZLE_64B layout: K-block r of the B operand IS CTA r's slice) into every
// CTA's shared memory with one cp.async.bulk (DSMEM) per peer, completing on the peer's
// mbarrier.  No global-memory round trip and no cluster-wide barrier sits on the T-step
// critical path.
#include <stdlib.h>
#include <atomic>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "lstm_tc_common.cuh"

namespace danet {

using namespace tc;

constexpr int kUnits = 32;            // hidden units per CTA
constexpr int kRows = 128;            // gate rows per CTA = UMMA M
constexpr int kUmmaN = 16;            // UMMA N (minimum for M = 128); NB = 8 or 16 columns carry utterances
constexpr int kMaxCta = 12;           // cluster size limit (TMEM: 32 + 2*16*ncta <= 512 columns)
// One K-block (32 units) of the B operand h^T, 2 KB: [atom0: hi 512 B | lo 512 B][atom1: hi | lo], an atom
// being 8 utterance rows x 64 B in SWIZZLE_64B; hi and lo descriptors differ by 512 B and both use SBO = 1024.
// With NB = 8 only atom0 ever changes, so a CTA ships 1 KB per peer per step instead of 2 KB.
constexpr int kHBlock = 2048;
constexpr int kXchLd = 17;
constexpr int kEpiThreads = 128;
constexpr int kThreads = 160 + 32 * kMaxCta;   // 4 epilogue warps + 1 MMA warp + one sender warp per peer
constexpr int kTmemCols = 512;
constexpr int kAccCols = 32;          // accumulator: columns [0,16) of the first 32

struct LstmTcParams {
  const float* pre;        // [n_dir][T][B][4H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  const uint32_t* Wh_packed;   // nullable: pre-split rows in TMEM order (danet_lstm_pack_wh), read instead of Wh
  long long ldw;
  float* out;              // [B][T][n_dir*H]
  float* cell_seq;         // nullable [n_dir][T][B][H]
  float* gates_seq;        // nullable, indexed like pre: post-activation [g|i|f|o]; may alias pre
  __nv_bfloat16* out_split;  // nullable [2][B*T][out_kp]: the hidden sequence as bf16 hi rows then lo rows,
  int out_kp;                //   i.e. the next layer's tensor-core A operand, ready made
  int zero_pad;              // gen 2: the threads of units >= H write the K padding of out_split (no separate memset)
  long long pre_dir, pre_row;   // element strides of pre: address = dir*pre_dir + (t*B + b)*pre_row + gate*H + unit
  int n_dir, T, B, H;
  long long* prof;         // nullable: per-step phase timestamps of CTA (0,0,0) (DANET_LSTM_PROFILE=1)
  int prof_steps;          // 0: entry / exit stamps only (DANET_LSTM_PROFILE=2)
  // nullable: the input projections are still being PRODUCED while this kernel runs (danet_gemm_split_pipelined on another
  // stream): row r = b*T + t of the producer's A operand belongs to row tile r / 128, complete once
  // pre_flags[tile] >= flag_need.  A thread waits (acquire) the first time it needs a value of a tile.
  const int* pre_flags;
  int flag_need;
  // row order of the producer's A operand (and so of its row tiles): 0 = batch-major r = b*T + t, 1 = time-major r = t*B + b
  int flags_tm;
  // gen 2: emit out_split with time-major rows (t*B + b) -- the next layer's pipelined product then finishes the tiles
  // of the first / last frames of ALL utterances first, and its recurrence starts after 2 of its ~32 row tiles
  int split_tm;
  // programmatic dependent launch (gen 2): at this step every CTA lets the NEXT kernel of the stream be scheduled
  // (griddepcontrol.launch_dependents; -1 = only at exit).  The next layer's recurrence, launched with the programmatic
  // attribute, is then placed on the SMs this kernel frees the moment it frees them, runs its prologue, and waits for
  // this grid's completion (griddepcontrol.wait) before it touches global memory.
  int pdl_trigger_step;
};

constexpr int kProfSlots = 16;
#define DANET_PROF(slot)                                                                   \
  do {                                                                                     \
    if (prof_on) p.prof[(size_t)s * kProfSlots + (slot)] = clock64();                      \
  } while (0)

// sigma(x) and tanh(x) from ex2.approx / rcp.approx: ~3e-7 absolute error, far inside the bf16x3
// error of the recurrent product, and 4-5x shorter than expf/tanhf on the step's critical path
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_kernel(const LstmTcParams p) {
  constexpr int UPT = NB / 4;                      // hidden units per epilogue thread (4 or 2)
  constexpr uint32_t kSendBytes = NB * 128;        // per peer per step: NB rows x (64 B hi + 64 B lo)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;                      // cluster size = K-blocks
  const int rank = (int)cluster_ctarank();         // == blockIdx.x
  const int bt = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // shared memory carve-up (tile bases 1024-byte aligned)
  uint8_t* sH = smem;                                          // [2 buf][ncta][kHBlock]
  uint8_t* sStage = sH + 2 * ncta * kHBlock;                   // [2][kHBlock]
  float* sXch = reinterpret_cast<float*>(sStage + 2 * kHBlock);      // [128][17]
  uint64_t* h_full = reinterpret_cast<uint64_t*>(sXch + kRows * kXchLd + 2);   // [2 buf], 8-byte aligned
  uint64_t* acc_full = h_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int unit0 = rank * kUnits;
  const int b0 = bt * NB;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
                       (tid == 0 || warp == 4 || warp == 5);

  if (tid == 0) {
    // one arrival (+ expect_tx) per source CTA and phase, posted remotely by the sender itself
    mbar_init(h_full + 0, ncta);
    mbar_init(h_full + 1, ncta);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  // rows NB..15 of every K-block (and all of h_{-1}) are zero for the whole run
  for (int i = tid; i < (2 * ncta + 2) * kHBlock / 16; i += kThreads)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 4) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_a_hi = tmem_base + kAccCols;              // ncta*16 columns
  const uint32_t tmem_a_lo = tmem_a_hi + (uint32_t)ncta * 16;

  // ---- one-time: this CTA's rows of Wh^T -> packed bf16 hi/lo in TMEM (lane = gate row) ----
  if (warp < 4) {
    const int m = tid, u = m >> 2, g = m & 3, unit = unit0 + u;
    const float* wcol = p.Wh[dir] + (size_t)g * H + unit;      // W[k][g*H + unit], stride ldw over k
    const bool unit_ok = unit < H;
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    for (int k0 = 0; k0 < ncta * 32; k0 += 16) {
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = (unit_ok && k0 + i < H) ? __ldg(wcol + (size_t)(k0 + i) * p.ldw) : 0.f;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(w[2 * i], h0, l0);
        split_bf16(w[2 * i + 1], h1, l1);
        hi[i] = pack_bf16(h0, h1);
        lo[i] = pack_bf16(l0, l1);
      }
      tmem_st_32x8(tmem_a_hi + lane_sel + (k0 >> 1), hi);
      tmem_st_32x8(tmem_a_lo + lane_sel + (k0 >> 1), lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();                    // every CTA's barriers are initialised before any peer signals them

  if (warp == 4) {
    // ================= MMA issuer =================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kRows, kUmmaN);
      for (int s = 1; s < T; ++s) {
        const int buf = (s - 1) & 1;
        const uint64_t b0d = umma_desc_k_sw64(smem_u32(sH + (size_t)buf * ncta * kHBlock));
        DANET_PROF(0);
        mbar_wait(h_full + buf, ((s - 1) >> 1) & 1);              // every slice of h_{s-1} is in sH[buf]
        DANET_PROF(1);
        tc_fence_after();
#pragma unroll 2
        for (int j = 0; j < ncta; ++j) {
          const uint64_t b_hi = b0d + (uint64_t)((j * kHBlock) >> 4);
          const uint64_t b_lo = b_hi + (uint64_t)(512 >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ac = (uint32_t)(j * 16 + k * 8);        // A columns of this K16 step
            const uint64_t adv = (uint64_t)(k * 2);                 // 32 bytes inside the 64 B row
            umma_bf16_ts(tmem_acc, tmem_a_hi + ac, b_hi + adv, idesc, (j | k) != 0);
            umma_bf16_ts(tmem_acc, tmem_a_hi + ac, b_lo + adv, idesc, 1);
            umma_bf16_ts(tmem_acc, tmem_a_lo + ac, b_hi + adv, idesc, 1);
          }
        }
        umma_commit(acc_full);
        DANET_PROF(2);
      }
    }
  } else if (warp < 4) {
    // ================= epilogue warps: TMEM lane m = 32*warp + lane = 4*unit + gate =================
    const int m = tid;
    // cell-update ownership: utterance bl = lane % NB, units ub..ub+UPT-1 of this warp's 8 units
    const int bl = lane % NB, ub = 8 * warp + UPT * (lane / NB);
    const int b = b0 + bl, unit = unit0 + ub;
    const bool valid = b < B && unit < H;            // H % 4 == 0 and ub % UPT == 0: all-or-nothing
    float c[UPT];
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu) c[uu] = 0.f;
    // input projections are fetched TWO steps ahead: they do not depend on the recurrence, and under the
    // stream-group schedule other groups' GEMMs saturate L2, stretching global latency beyond one step
    float pre_q[2][4][UPT];
    auto load_pre = [&](int s, float (&dst)[4][UPT]) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) dst[g][uu] = 0.f;
      if (valid && s < T) {
        const int to = dir ? T - 1 - s : s;
        const float* q = p.pre + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (UPT == 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(q + g * H));
            dst[g][0] = v.x; dst[g][1] = v.y; dst[g][UPT - 2] = v.z; dst[g][UPT - 1] = v.w;
          } else {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(q + g * H));
            dst[g][0] = v.x; dst[g][1] = v.y;
          }
        }
      }
    };
    load_pre(0, pre_q[0]);
    load_pre(1, pre_q[1]);
    const int outw = p.n_dir * H;
    float* xw = sXch + m * kXchLd;
    const uint32_t stage_off = sw64_offset(bl, ub);      // UPT contiguous bf16: units ub.. of utterance bl
    for (int s = 0; s < T; ++s) {
      const int to = dir ? T - 1 - s : s;
      float a[UPT][4];                                   // [unit][gate]
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
        for (int g = 0; g < 4; ++g) a[uu][g] = pre_q[0][g][uu];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) pre_q[0][g][uu] = pre_q[1][g][uu];
      load_pre(s + 2, pre_q[1]);
      DANET_PROF(3);
      if (s > 0) {
        mbar_wait(acc_full, (s - 1) & 1);
        DANET_PROF(4);
        tc_fence_after();
        float v[NB];
        if (NB == 16) tmem_ld_32x16(tmem_acc + ((uint32_t)(32 * warp) << 16), *reinterpret_cast<float(*)[16]>(v));
        else tmem_ld_32x8(tmem_acc + ((uint32_t)(32 * warp) << 16), *reinterpret_cast<float(*)[8]>(v));
        DANET_PROF(5);
        tc_fence_before();
        __syncwarp();                                    // previous step's reads of sXch are done
#pragma unroll
        for (int j = 0; j < NB; ++j) xw[j] = v[j];
        __syncwarp();
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
          for (int g = 0; g < 4; ++g) a[uu][g] += sXch[(4 * (ub + uu) + g) * kXchLd + bl];
      }
      DANET_PROF(6);
      float h[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const float gg = a[uu][0];
        const float ig = fast_sigmoid(a[uu][1]), fg = fast_sigmoid(a[uu][2]), og = fast_sigmoid(a[uu][3]);
        c[uu] = ig * gg + fg * c[uu];
        h[uu] = valid ? og * fast_tanh(c[uu]) : 0.f;
        a[uu][1] = ig; a[uu][2] = fg; a[uu][3] = og;           // post-activation gates, kept for training
      }
      DANET_PROF(7);
      __nv_bfloat16 hi[UPT], lo[UPT];
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) split_bf16(h[uu], hi[uu], lo[uu]);
      if (s < T - 1) {
        // my units of h_s as bf16 hi/lo into the staging K-block (already UMMA layout)
        // staged once for the peers, and written straight into this CTA's own B operand (a bulk copy whose
        // destination is the issuing CTA is not a remote access; the own slice needs no transport at all)
        uint8_t* st = sStage + (s & 1) * kHBlock;
        uint8_t* own = sH + ((size_t)(s & 1) * ncta + rank) * kHBlock;
        if (UPT == 4) {
          const uint2 vh = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[UPT - 2], hi[UPT - 1]));
          const uint2 vl = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[UPT - 2], lo[UPT - 1]));
          *reinterpret_cast<uint2*>(st + stage_off) = vh;
          *reinterpret_cast<uint2*>(st + 512 + stage_off) = vl;
          *reinterpret_cast<uint2*>(own + stage_off) = vh;
          *reinterpret_cast<uint2*>(own + 512 + stage_off) = vl;
        } else {
          const uint32_t vh = pack_bf16(hi[0], hi[1]), vl = pack_bf16(lo[0], lo[1]);
          *reinterpret_cast<uint32_t*>(st + stage_off) = vh;
          *reinterpret_cast<uint32_t*>(st + 512 + stage_off) = vl;
          *reinterpret_cast<uint32_t*>(own + stage_off) = vh;
          *reinterpret_cast<uint32_t*>(own + 512 + stage_off) = vl;
        }
        fence_proxy_async_smem();
        // hand the staged slice to the sender warps (producer side of named barrier 1: no wait)
        asm volatile("bar.arrive 1, %0;" ::"r"(kEpiThreads + 32 * ncta) : "memory");
        DANET_PROF(8);
      }
      if (valid && p.out_split) {
        __nv_bfloat16* oh = p.out_split + ((size_t)b * T + to) * p.out_kp + dir * H + unit;
        __nv_bfloat16* ol = oh + (size_t)B * T * p.out_kp;
        if (UPT == 4) {
          *reinterpret_cast<uint2*>(oh) = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[UPT - 2], hi[UPT - 1]));
          *reinterpret_cast<uint2*>(ol) = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[UPT - 2], lo[UPT - 1]));
        } else {
          *reinterpret_cast<uint32_t*>(oh) = pack_bf16(hi[0], hi[1]);
          *reinterpret_cast<uint32_t*>(ol) = pack_bf16(lo[0], lo[1]);
        }
      }
      if (valid && p.gates_seq) {
        float* gs = p.gates_seq + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (UPT == 4) *reinterpret_cast<float4*>(gs + g * H) = make_float4(a[0][g], a[1][g], a[UPT - 2][g], a[UPT - 1][g]);
          else *reinterpret_cast<float2*>(gs + g * H) = make_float2(a[0][g], a[1][g]);
        }
      }
      if (valid) {
        float* o = p.out + ((size_t)b * T + to) * outw + dir * H + unit;
        float* cs = p.cell_seq ? p.cell_seq + (((size_t)dir * T + to) * B + b) * H + unit : nullptr;
        if (UPT == 4) {
          *reinterpret_cast<float4*>(o) = make_float4(h[0], h[1], h[UPT - 2], h[UPT - 1]);
          if (cs) *reinterpret_cast<float4*>(cs) = make_float4(c[0], c[1], c[UPT - 2], c[UPT - 1]);
        } else {
          *reinterpret_cast<float2*>(o) = make_float2(h[0], h[1]);
          if (cs) *reinterpret_cast<float2*>(cs) = make_float2(c[0], c[1]);
        }
      }
    }
  }
  if (warp >= 5 && warp - 5 < ncta) {
    // ================= sender warps: warp 5+i ships this CTA's slice of h_s to peer rank+i =================
    const int peer = (rank + (warp - 5)) % ncta;
    const uint32_t peer_dst = mapa(smem_u32(sH + (size_t)rank * kHBlock), peer);
    const uint32_t peer_bar = mapa(smem_u32(h_full), peer);
    for (int s = 0; s + 1 < T; ++s) {
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads + 32 * ncta) : "memory");
      if (elect_one_sync()) {
        if (peer == rank) {
          mbar_arrive(h_full + (s & 1));                      // own slice: already in place, just count it
        } else {
          const uint32_t boff = (uint32_t)(s & 1) * (uint32_t)ncta * kHBlock;
          const uint32_t bar = peer_bar + (uint32_t)(s & 1) * 8;
          mbar_arrive_expect_tx_cluster(bar, kSendBytes);     // one arrival + the byte count, posted by the sender
          dsmem_bulk_copy(peer_dst + boff, smem_u32(sStage + (s & 1) * kHBlock), kSendBytes, bar);
        }
        DANET_PROF(9);
      }
      __syncwarp();
    }
  }
  // nobody leaves while a peer may still read this CTA's staging tile or signal its barriers
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}


// ================================================================================================
// Second generation of the 8-utterances-per-cluster kernel (the default whenever every cluster is
// co-resident).  Same decomposition and transport as above; the T-step critical path is shorter.
// Measured anatomy of the first generation (profiles/r01_lstm_phase_cycles_v4.txt): the MMA phase is bound by
// streaming the A operand out of tensor memory (~14 cycles per fresh 128x16 bf16 tile, 5 when the tile repeats), the
// exchange of h by the ~17 B/cycle an SM can move through DSMEM while sending and receiving.  So:
//   * HF = 0 (backend 1, "bf16x3"): hi and lo of h share ONE B tile along N: a K-block is [lo atom | hi atom | zero
//     atom] (8 rows x 64 B each, SWIZZLE_64B, SBO = 512), A_hi x [lo | hi] puts hi*lo in accumulator columns 0-7 and
//     hi*hi in 8-15, A_lo x [hi | 0] adds lo*hi to columns 0-7: 40 MMAs per step instead of 60;
//   * HF = 1 (backend 2): h travels as ONE fp16 value (|h| < 1: 11 significant bits, 2^-12 relative rounding) and
//     multiplies the bf16 hi/lo pair of Wh: 40 MMAs, and HALF the DSMEM bytes (512 B per peer per step).  Measured
//     effect on the embedding (tools/precision_study.py): 1e-4 max-norm relative, against 2e-6 for HF = 0 and the
//     1e-3 gate;
//   * 8 epilogue warps (two per TMEM lane quadrant, 4 utterances each), one (unit, utterance) pair per thread;
//   * the gate-row -> (unit, utterance) regrouping is a 4x4 butterfly of warp shuffles among the 4 lanes that hold
//     one unit's gates (no shared-memory round trip);
//   * the two sigmoids feeding the cell share one reciprocal, as do the output gate and tanh: 6 MUFU ops per pair
//     instead of 8;
//   * Wh arrives pre-split (danet_lstm_pack_wh) through one bulk copy per CTA: prologue 45k -> 6k cycles.
constexpr int kEpi2Warps = 8;
#ifndef DANET_LSTM_PRE_DEPTH
#define DANET_LSTM_PRE_DEPTH 2      // measured 2 / 3 / 4: profiles/r02_lstm_experiment_skip_zero_slice.txt
#endif
constexpr int kPreDepth = DANET_LSTM_PRE_DEPTH;
constexpr int kEpi2Threads = 32 * kEpi2Warps;
constexpr int kThreads2 = kEpi2Threads + 32 + 32 * kMaxCta;

__device__ __forceinline__ void tmem_ld_32x4(uint32_t t0, float (&a)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t0) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __uint_as_float(r[i]);
}
// two 32x32b.x4 loads (columns c0.. and c1..) behind one wait
__device__ __forceinline__ void tmem_ld_2x4(uint32_t t0, uint32_t t1, float (&a)[4], float (&b)[4]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t0) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t1) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[4 + i]); }
}

template <int HF>
__global__ void __launch_bounds__(kThreads2, 1)
lstm_tc2_kernel(const LstmTcParams p) {
  constexpr int NB = 8;
  // one K-block (32 units) of the B operand: HF 0: [lo 512 | hi 512 | zero 512 | pad]; HF 1: [fp16 512 | zero 512]
  constexpr int kBlk = HF ? 1024 : 2048;
  constexpr uint32_t kSend = HF ? 512 : 1024;       // bytes per peer per step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;
  const int rank = (int)cluster_ctarank();
  const int bt = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kMmaWarp = kEpi2Warps, kSend0 = kEpi2Warps + 1;

  uint8_t* sH = smem;                                          // [2 buf][ncta][kBlk]
  uint8_t* sStage = sH + 2 * ncta * kBlk;                      // [2][kSend]
  uint64_t* h_full = reinterpret_cast<uint64_t*>(sStage + 2 * kSend);
  uint64_t* acc_full = h_full + 2;
  uint64_t* w_full = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint8_t* sW = sStage + 2 * kSend + 64;                       // packed Wh slice, prologue only (16-byte aligned)

  const int unit0 = rank * kUnits;
  const int b0 = bt * NB;
  // DANET_LSTM_PROFILE=2: only the entry / exit stamps of row 0 (no per-step stamps, no memset ahead of the launch: the
  // timing of the step is not disturbed)
  const bool prof_ee = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  const bool prof_on = p.prof != nullptr && p.prof_steps && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
                       (tid == 0 || warp == kMmaWarp || warp == kSend0);
  if (prof_ee) {
    p.prof[10] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[14] = (long long)gt;
  }

  if (tid == 0) {
    // ONE arrival per phase, made locally by this CTA's own-slice warp together with the bytes it expects from the
    // peers; the peers' bulk copies only complete transaction bytes here (a copy that lands before the local arrival
    // leaves the transaction count negative for a moment: the phase cannot complete while the arrival is pending).  One
    // remote mbarrier operation less per peer and step on the senders' path (first build: every sender posted
    // arrive.expect_tx remotely ahead of its copy); measured neutral on the step period (1687 against 1681 cycles).
    mbar_init(h_full + 0, 1);
    mbar_init(h_full + 1, 1);
    mbar_init(acc_full, 1);
    mbar_init(w_full, 1);
    fence_barrier_init();
    if (p.Wh_packed) {
      const uint32_t bytes = (uint32_t)kRows * (uint32_t)(ncta * 32 + 4) * 4u;
      // the packed buffer holds a bf16 image (backend 1) followed by an fp16 image (backend 2)
      const size_t image = (size_t)p.n_dir * ncta * kRows * (size_t)(ncta * 32 + 4);
      const uint32_t* src = p.Wh_packed + (HF ? image : 0) +
                            ((size_t)dir * ncta + rank) * (size_t)kRows * (size_t)(ncta * 32 + 4);
      mbar_arrive_expect_tx(w_full, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sW)), "l"(src), "r"(bytes), "r"(smem_u32(w_full)) : "memory");
    }
  }
  for (int i = tid; i < (2 * ncta * kBlk + 2 * (int)kSend) / 16; i += kThreads2)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_a_hi = tmem_base + kAccCols;
  const uint32_t tmem_a_lo = tmem_a_hi + (uint32_t)ncta * 16;

  // ---- one-time: this CTA's rows of Wh^T -> packed bf16 hi/lo in TMEM (lane = gate row 4*unit + gate) ----
  if (warp < 4) {
    // TMEM lane m = 32*q + 8*gate + u holds gate `gate` of unit 8*q + u (gate-major inside each 32-lane quadrant, so that
    // tcgen05.ld.16x128b hands one thread all four gates of a (unit, utterance) pair: see the epilogue)
    const int m = tid, u = 8 * (m >> 5) + (m & 7), g = (m >> 3) & 3, unit = unit0 + u;
    const bool unit_ok = unit < H;
    const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
    if (p.Wh_packed) {
      // pre-split rows (danet_lstm_pack_wh): [dir][rank][128 rows][ncta*16 hi words | ncta*16 lo words | 4 pad words].
      // The whole 160 KB slice was fetched into shared memory by ONE bulk copy (issued by thread 0 above); the 16-byte
      // row pad makes the per-lane row reads conflict-free.
      const int rw = ncta * 32 + 4;
      mbar_wait(w_full, 0);
      const uint4* row = reinterpret_cast<const uint4*>(sW + (size_t)m * rw * 4);
      for (int k0 = 0; k0 < ncta * 16; k0 += 8) {
        const uint4 h0 = row[k0 >> 2], h1 = row[(k0 >> 2) + 1];
        const uint4 l0 = row[(ncta * 16 + k0) >> 2], l1 = row[((ncta * 16 + k0) >> 2) + 1];
        const uint32_t hi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t lo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        tmem_st_32x8(tmem_a_hi + lane_sel + k0, hi);
        tmem_st_32x8(tmem_a_lo + lane_sel + k0, lo);
      }
    } else {
      const float* wcol = p.Wh[dir] + (size_t)g * H + unit;
      for (int k0 = 0; k0 < ncta * 32; k0 += 16) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = (unit_ok && k0 + i < H) ? __ldg(wcol + (size_t)(k0 + i) * p.ldw) : 0.f;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (HF) split2_f16(w[2 * i], w[2 * i + 1], hi[i], lo[i]);
          else split2_bf16(w[2 * i], w[2 * i + 1], hi[i], lo[i]);
        }
        tmem_st_32x8(tmem_a_hi + lane_sel + (k0 >> 1), hi);
        tmem_st_32x8(tmem_a_lo + lane_sel + (k0 >> 1), lo);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync();
  // launched as a programmatic dependent of the previous kernel in the stream (the previous layer's recurrence): everything
  // above ran while that grid was draining; from here on its results (and whatever else preceded it) are complete and
  // visible.  A no-op for a normally launched kernel.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (prof_ee) p.prof[11] = clock64();

  if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = HF ? umma_idesc_f16(kRows, kUmmaN) : umma_idesc_bf16(kRows, kUmmaN);
      for (int s = 1; s < T; ++s) {
        const int buf = (s - 1) & 1;
        const uint64_t b0d = umma_desc_k_sw64_sbo512(smem_u32(sH + (size_t)buf * ncta * kBlk));
        DANET_PROF(0);
        mbar_wait(h_full + buf, ((s - 1) >> 1) & 1);
        DANET_PROF(1);
        tc_fence_after();
#pragma unroll 2
        for (int j = 0; j < ncta; ++j) {
          // HF 0: rows 0-7 lo, rows 8-15 hi for A_hi; rows 0-7 hi, rows 8-15 zero for A_lo.  HF 1: [h | zero] for both.
          // (The last block's second K16 slice is all zeros at H = 300; NOT issuing it was measured: a conditional in this
          // loop costs 220 cycles per step, peeling the last block 60 -- the unrolled issue sequence is worth more than
          // two MMAs.  profiles/r02_lstm_experiment_skip_zero_slice.txt)
          const uint64_t b_first = b0d + (uint64_t)((j * kBlk) >> 4);
          const uint64_t b_second = HF ? b_first : b_first + (uint64_t)(512 >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ac = (uint32_t)(j * 16 + k * 8);
            const uint64_t adv = (uint64_t)(k * 2);
            umma_bf16_ts(tmem_acc, tmem_a_hi + ac, b_first + adv, idesc, (j | k) != 0);
            umma_bf16_ts(tmem_acc, tmem_a_lo + ac, b_second + adv, idesc, 1);
          }
        }
        umma_commit(acc_full);
        DANET_PROF(2);
      }
    }
  } else if (warp < kEpi2Warps) {
    // ================= epilogue: TMEM lane 32*q + lane = 4*unit + gate; this warp's utterances 4*hw .. 4*hw+3 =====
    const int q = warp & 3, hw = warp >> 2;
    const int u = lane >> 2, g = lane & 3;
    const int ul = 8 * q + u;                      // unit inside the CTA
    const int bl = 4 * hw + g;                     // utterance inside the tile (after the butterfly)
    const int b = b0 + bl, unit = unit0 + ul;
    const bool valid = b < B && unit < H;
    float c = 0.f;
    // input projections are fetched kPreDepth steps ahead (they do not depend on the recurrence): under the stream-group
    // schedule the other groups' dense products load L2, and a fetch that is late stalls all ten CTAs of the cluster
    float pre_q[kPreDepth][4];
    int tile_seen = -1;
    auto load_pre = [&](int s, float (&dst)[4]) {
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) dst[gg] = 0.f;
      if (valid && s < T) {
        const int to = dir ? T - 1 - s : s;
        if (p.pre_flags) {
          const int tile = (p.flags_tm ? to * B + b : b * T + to) >> 7;
          if (tile != tile_seen) {
            // bounded (~2 s): a producer that never comes must not hang the device -- the values of that tile are then
            // replaced by NaN, which the recurrence carries into every later output of the utterance (a visibly invalid
            // result instead of a silently wrong one)
            int got = 0;
            for (int spin = 0; spin < (1 << 24); ++spin) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(p.pre_flags + tile) : "memory");
              if (got >= p.flag_need) break;
              __nanosleep(100);
            }
            if (got < p.flag_need) {
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) dst[gg] = __int_as_float(0x7fc00000);
              return;
            }
            tile_seen = tile;
          }
        }
        const float* qp = p.pre + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) dst[gg] = __ldcg(qp + gg * H);
      }
    };
#pragma unroll
    for (int d = 0; d < kPreDepth; ++d) load_pre(d, pre_q[d]);
    const int outw = p.n_dir * H;
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    const uint32_t stage_off = sw64_offset(bl, ul & ~1);      // the (even, odd) unit pair of utterance bl: 4 bytes
    const uint32_t stage_addr = smem_u32(sStage) + stage_off;
    const uint32_t own_addr = smem_u32(sH + (size_t)rank * kBlk) + stage_off;
    const uint32_t stage16_addr = smem_u32(sStage) + sw64_offset(bl, ul);            // this thread's own 16-bit slot
    const uint32_t own16_addr = smem_u32(sH + (size_t)rank * kBlk) + sw64_offset(bl, ul);
    const bool odd = (u & 1) != 0;
    constexpr float kL2e = 1.4426950408889634f;
    for (int s = 0; s < T; ++s) {
      const int to = dir ? T - 1 - s : s;
      float a[4];
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        a[gg] = pre_q[0][gg];
#pragma unroll
        for (int d = 0; d + 1 < kPreDepth; ++d) pre_q[d][gg] = pre_q[d + 1][gg];
      }
      load_pre(s + kPreDepth, pre_q[kPreDepth - 1]);
      if (s == p.pdl_trigger_step) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      DANET_PROF(3);
      if (s > 0) {
        mbar_wait(acc_full, (s - 1) & 1);
        DANET_PROF(4);
        tc_fence_after();
        // tcgen05.ld.16x128b at lanes 32q (+16), columns 4hw..4hw+3: thread t receives (lane base + t/4, column t%4)
        // and (lane base + 8 + t/4, column t%4), i.e. with the gate-major row order gates 0,1 (2,3) of unit t/4 for
        // utterance 4hw + t%4 -- exactly this thread's pair, no cross-lane regrouping (layout probed on the B200:
        // tools/probes/tmem_ld_layout.cu)
        {
          const uint32_t t0 = tmem_acc + lane_sel + 4 * hw;
          uint32_t r01[2], r23[2];
          tmem_ld_16x128b(t0, r01);
          tmem_ld_16x128b(t0 + ((uint32_t)16 << 16), r23);
          if (HF) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            a[0] += __uint_as_float(r01[0]); a[1] += __uint_as_float(r01[1]);
            a[2] += __uint_as_float(r23[0]); a[3] += __uint_as_float(r23[1]);
          } else {                                           // columns j hold hi*lo + lo*hi, columns 8 + j hold hi*hi
            uint32_t s01[2], s23[2];
            tmem_ld_16x128b(t0 + 8, s01);
            tmem_ld_16x128b(t0 + 8 + ((uint32_t)16 << 16), s23);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            a[0] += __uint_as_float(r01[0]) + __uint_as_float(s01[0]); a[1] += __uint_as_float(r01[1]) + __uint_as_float(s01[1]);
            a[2] += __uint_as_float(r23[0]) + __uint_as_float(s23[0]); a[3] += __uint_as_float(r23[1]) + __uint_as_float(s23[1]);
          }
        }
        DANET_PROF(5);
        tc_fence_before();
      }
      DANET_PROF(6);
      // c = sig(i)*g + sig(f)*c ; h = sig(o)*tanh(c)   (candidate WITHOUT tanh, app/ops.py:141-147)
      const float ei = ex2_approx(-kL2e * fmaxf(a[1], -30.f));
      const float ef = ex2_approx(-kL2e * fmaxf(a[2], -30.f));
      const float eo = ex2_approx(-kL2e * fmaxf(a[3], -30.f));
      const float pi = 1.f + ei, pf = 1.f + ef, po = 1.f + eo;
      const float rif = rcp_approx(pi * pf);
      const float ig = rif * pf, fg = rif * pi;
      c = ig * a[0] + fg * c;
      const float ec = ex2_approx(-2.f * kL2e * fabsf(c));
      const float pc = 1.f + ec;
      const float roc = rcp_approx(po * pc);
      const float og = roc * pc;
      const float th = copysignf((1.f - ec) * roc * po, c);
      const float h = valid ? og * th : 0.f;
      DANET_PROF(7);
      uint32_t vh = 0, vl = 0;
      float he = 0.f, ho = 0.f;
      if (HF) {
        // every thread stores its own fp16 value (16-bit stores, neighbours share a word): the pairing shuffle is only
        // needed for the global stores below, behind the barrier
        if (s < T - 1) {
          const unsigned short hv = __half_as_ushort(__float2half_rn(h));
          st_shared_u16(stage16_addr + (uint32_t)(s & 1) * kSend, hv);
          st_shared_u16(own16_addr + (uint32_t)(s & 1) * (uint32_t)ncta * kBlk, hv);
          fence_proxy_async_smem();
          asm volatile("bar.arrive 1, %0;" ::"r"(kEpi2Threads + 32 * ncta) : "memory");
          DANET_PROF(8);
        }
        const float hn = __shfl_xor_sync(0xffffffffu, h, 4);
        he = odd ? hn : h; ho = odd ? h : hn;                  // units (ul & ~1), (ul | 1)
      } else {
        // bf16 hi / lo of my own value as two 16-bit stores each (stage + own copy): no pairing shuffle before the barrier
        __nv_bfloat16 bh, bl16;
        split_bf16(h, bh, bl16);
        if (s < T - 1) {
          const uint32_t st = stage16_addr + (uint32_t)(s & 1) * kSend;
          const uint32_t own = own16_addr + (uint32_t)(s & 1) * (uint32_t)ncta * kBlk;
          st_shared_u16(st, __bfloat16_as_ushort(bl16));
          st_shared_u16(st + 512, __bfloat16_as_ushort(bh));
          st_shared_u16(own, __bfloat16_as_ushort(bl16));
          st_shared_u16(own + 512, __bfloat16_as_ushort(bh));
          fence_proxy_async_smem();
          asm volatile("bar.arrive 1, %0;" ::"r"(kEpi2Threads + 32 * ncta) : "memory");
          DANET_PROF(8);
        }
        // neighbouring units (lane ^ 4) pair up for the 32-bit global stores of the next layer's operand
        const uint32_t mine = (uint32_t)__bfloat16_as_ushort(bh) | ((uint32_t)__bfloat16_as_ushort(bl16) << 16);
        const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 4);
        const uint32_t e = odd ? other : mine, o = odd ? mine : other;          // units (ul & ~1), (ul | 1)
        vh = (e & 0xffffu) | (o << 16);
        vl = (e >> 16) | (o & 0xffff0000u);
      }
      const size_t srow = p.split_tm ? (size_t)to * B + b : (size_t)b * T + to;      // row of the emitted operand
      if (p.zero_pad && odd && b < B && unit >= H) {
        // columns [n_dir*H, out_kp) of the operand: n_dir * (32*ncta - H) of them, one pair per idle unit pair (h = 0 here)
        __nv_bfloat16* oh = p.out_split + srow * p.out_kp + p.n_dir * H + dir * (ncta * kUnits - H) + (unit - 1 - H);
        *reinterpret_cast<uint32_t*>(oh) = 0u;
        *reinterpret_cast<uint32_t*>(oh + (size_t)B * T * p.out_kp) = 0u;
      }
      if (valid) {
        if (odd && p.out_split) {
          if (HF) split2_bf16(he, ho, vh, vl);
          __nv_bfloat16* oh = p.out_split + srow * p.out_kp + dir * H + (unit - 1);
          *reinterpret_cast<uint32_t*>(oh) = vh;
          *reinterpret_cast<uint32_t*>(oh + (size_t)B * T * p.out_kp) = vl;
        }
        p.out[((size_t)b * T + to) * outw + dir * H + unit] = h;
        if (p.cell_seq) p.cell_seq[(((size_t)dir * T + to) * B + b) * H + unit] = c;
        if (p.gates_seq) {
          float* gs = p.gates_seq + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
          gs[0] = a[0]; gs[H] = ig; gs[2 * H] = fg; gs[3 * H] = og;
        }
      }
    }
  } else if (warp >= kSend0 && warp - kSend0 < ncta) {
    // ================= sender warps: warp kSend0+i ships this CTA's slice of h_s to peer rank+i =================
    const int peer = (rank + (warp - kSend0)) % ncta;
    const uint32_t peer_dst = mapa(smem_u32(sH + (size_t)rank * kBlk), peer);
    const uint32_t peer_bar = mapa(smem_u32(h_full), peer);
    for (int s = 0; s + 1 < T; ++s) {
      asm volatile("bar.sync 1, %0;" ::"r"(kEpi2Threads + 32 * ncta) : "memory");
      if (elect_one_sync()) {
        if (peer == rank) {
          mbar_arrive_expect_tx(h_full + (s & 1), (uint32_t)(ncta - 1) * kSend);
        } else {
          const uint32_t boff = (uint32_t)(s & 1) * (uint32_t)ncta * kBlk;
          const uint32_t bar = peer_bar + (uint32_t)(s & 1) * 8;
          dsmem_bulk_copy(peer_dst + boff, smem_u32(sStage + (s & 1) * kSend), kSend, bar);
        }
        DANET_PROF(9);
      }
      __syncwarp();
    }
  }
  if (prof_ee) p.prof[12] = clock64();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
  if (prof_ee) {
    p.prof[13] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[15] = (long long)gt;
  }
}

// The kernel needs ~55 KB but asks for the whole SM's shared memory: the recurrence is latency-bound, and a
// co-resident CTA of another stream's kernel (attractor, mask, splits ...) would steal issue slots and
// shared-memory bandwidth from the step's critical path.
static size_t lstm_tc_smem_bytes(int ncta) {
  const size_t need = (size_t)(2 * ncta + 2) * kHBlock + (kRows * kXchLd + 2) * sizeof(float) + 64 + 1024;
  const size_t whole_sm = 227 * 1024;
  return need > whole_sm ? need : whole_sm;
}
static size_t lstm_tc2_packed_smem_bytes(int ncta) {
  return (size_t)2 * ncta * 2048 + 2 * 1024 + 64 + (size_t)kRows * (ncta * 32 + 4) * 4 + 1024;
}
static bool lstm_tc2_packed_fits(int ncta) { return lstm_tc2_packed_smem_bytes(ncta) <= 227 * 1024; }

static void cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, dim3 grid, int threads, size_t smem, int ncta,
                        cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

template <typename Kern>
static int launch_cluster(Kern kern, const LstmTcParams& p, int ncta, int nb, int threads, cudaStream_t stream,
                          bool programmatic = false) {
  const size_t smem = lstm_tc_smem_bytes(ncta);
  DANET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ncta > 8) DANET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  cluster_cfg(cfg, attr, dim3(ncta, (p.B + nb - 1) / nb, p.n_dir), threads, smem, ncta, stream);
  if (programmatic) {          // may be scheduled once the previous kernel of the stream has triggered (or finished)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  DANET_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return DANET_OK;
}

size_t lstm_tc_workspace_bytes(int, int, int) { return 256; }

bool lstm_tc_supported(int H) { return H % 4 == 0 && (H + kUnits - 1) / kUnits <= kMaxCta; }

// ---- Wh -> the recurrent kernel's TMEM image, once per weight update ------------------------------------------
// [image][dir][rank][128 rows m = 32*q + 8*gate + u, unit 8*q + u][ncta*16 hi words | ncta*16 lo words | 4 pad words]; word j of a row
// holds elements k = 2j (low half) and 2j+1 of Wh[k][gate*H + 32*rank + unit] as bf16 (image 0) or fp16 (image 1).
__global__ void lstm_pack_wh_kernel(const float* W0, const float* W1, long long ldw, int H, int ncta, uint32_t* out) {
  const int rank = blockIdx.x, dir = blockIdx.y, m = threadIdx.x;
  const bool f16 = blockIdx.z != 0;                 // image 0: bf16 pairs (backend 1), image 1: fp16 pairs (backend 2)
  out += (size_t)blockIdx.z * gridDim.y * ncta * kRows * (size_t)(ncta * 32 + 4);
  const int unit = rank * kUnits + 8 * (m >> 5) + (m & 7), g = (m >> 3) & 3;   // lane m = 32q + 8*gate + u (lstm_tc2_kernel)
  const int rw = ncta * 32 + 4;
  const float* W = dir ? W1 : W0;
  uint32_t* row = out + (((size_t)dir * ncta + rank) * kRows + m) * (size_t)rw;
  const bool unit_ok = unit < H;
  for (int j = 0; j < ncta * 16; ++j) {
    const int k = 2 * j;
    const float w0 = (unit_ok && k < H) ? __ldg(W + (size_t)k * ldw + (size_t)g * H + unit) : 0.f;
    const float w1 = (unit_ok && k + 1 < H) ? __ldg(W + (size_t)(k + 1) * ldw + (size_t)g * H + unit) : 0.f;
    uint32_t hi, lo;
    if (f16) split2_f16(w0, w1, hi, lo);
    else split2_bf16(w0, w1, hi, lo);
    row[j] = hi;
    row[ncta * 16 + j] = lo;
  }
  for (int j = 0; j < 4; ++j) row[ncta * 32 + j] = 0u;
}

size_t lstm_tc_pack_bytes(int n_dir, int H) {
  const int ncta = (H + kUnits - 1) / kUnits;
  return (size_t)2 * n_dir * ncta * kRows * (size_t)(ncta * 32 + 4) * 4;      // bf16 image + fp16 image
}

int lstm_tc_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed, cudaStream_t stream) {
  DANET_REQUIRE(lstm_tc_supported(H), DANET_E_SHAPE, "lstm_pack_wh: H %d is outside the tcgen05 backend's range", H);
  DANET_REQUIRE(aligned16(packed), DANET_E_ALIGN, "lstm_pack_wh: packed must be 16-byte aligned");
  const int ncta = (H + kUnits - 1) / kUnits;
  lstm_pack_wh_kernel<<<dim3(ncta, n_dir, 2), kRows, 0, stream>>>(host_Wh[0], n_dir > 1 ? host_Wh[1] : host_Wh[0], ldw, H, ncta,
                                                               reinterpret_cast<uint32_t*>(packed));
  DANET_CUDA(cudaGetLastError());
  return DANET_OK;
}

int lstm_tc_fwd(const float* pre, long long pre_dir, long long pre_row, const float* const* host_Wh, long long ldw,
                const void* wh_packed, float* out, float* cell_seq, float* gates_seq, void* out_split, int out_kp, int n_dir,
                int T, int B, int H, int h_fp16, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                const int* pre_flags, int flag_need, int flags_tm, int split_tm, int programmatic) {
  const int ncta = (H + kUnits - 1) / kUnits;
  DANET_REQUIRE(lstm_tc_supported(H), DANET_E_SHAPE,
                "lstm_seq: the tcgen05 backend keeps Wh resident in one cluster's tensor memory and needs "
                "H <= %d (got %d); use backend 0", kMaxCta * kUnits, H);
  DANET_REQUIRE(aligned16(pre) && aligned16(out) && (!cell_seq || aligned16(cell_seq)), DANET_E_ALIGN,
                "lstm_seq: pre/out/cell_seq must be 16-byte aligned");
  DANET_REQUIRE(!wh_packed || aligned16(wh_packed), DANET_E_ALIGN, "lstm_seq: wh_packed must be 16-byte aligned");
  long long* prof = nullptr;
  int prof_steps = 1;
  if (getenv("DANET_LSTM_PROFILE") && workspace && workspace_bytes >= (size_t)T * kProfSlots * sizeof(long long)) {
    prof = reinterpret_cast<long long*>(workspace);
    prof_steps = atoi(getenv("DANET_LSTM_PROFILE")) != 2;
    if (prof_steps) DANET_CUDA(cudaMemsetAsync(prof, 0, (size_t)T * kProfSlots * sizeof(long long), stream));
  }
  LstmTcParams p;
  p.pre = pre;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.Wh_packed = nullptr;
  p.ldw = ldw; p.out = out; p.cell_seq = cell_seq; p.gates_seq = gates_seq;
  p.pre_dir = pre_dir; p.pre_row = pre_row;
  p.out_split = reinterpret_cast<__nv_bfloat16*>(out_split);
  p.out_kp = out_kp;
  if (out_split) {
    DANET_REQUIRE(out_kp >= n_dir * H && out_kp % 64 == 0 && aligned16(out_split), DANET_E_SHAPE,
                  "lstm_seq: out_split needs a 16-byte aligned buffer with row length %d >= %d, multiple of 64", out_kp, n_dir * H);
  }
  p.zero_pad = 0;
  p.n_dir = n_dir; p.T = T; p.B = B; p.H = H; p.prof = prof; p.prof_steps = prof_steps;
  p.pre_flags = pre_flags; p.flag_need = flag_need;
  p.flags_tm = flags_tm ? 1 : 0; p.split_tm = split_tm ? 1 : 0;
  // ~11 us before the end: long enough to hide the launch of the dependent grid, short enough that its clusters, should
  // they find free SMs elsewhere, do not sit on them for long
  p.pdl_trigger_step = T > 16 ? T - 13 : -1;
  // The recurrence is latency-bound, so spread utterances thin: 8 per cluster (half the DSMEM bytes and
  // half the epilogue work per step) while all clusters are still co-resident, 16 per cluster otherwise.
  const int clusters8 = n_dir * ((B + 7) / 8);
  // per-cluster-size occupancy, queried once; the C-ABI is re-entrant, so the cache is atomic (two threads racing here
  // both compute the same value and store it)
  static std::atomic<int> resident_cache[kMaxCta + 1];
  int resident = resident_cache[ncta].load(std::memory_order_relaxed);
  if (resident == 0) {
    const size_t smem = lstm_tc_smem_bytes(ncta);
    DANET_CUDA(cudaFuncSetAttribute(lstm_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ncta > 8)
      DANET_CUDA(cudaFuncSetAttribute(lstm_tc2_kernel<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_cfg(cfg, attr, dim3(ncta, clusters8, 1), kThreads2, smem, ncta, nullptr);
    if (cudaOccupancyMaxActiveClusters(&resident, lstm_tc2_kernel<0>, &cfg) != cudaSuccess || resident <= 0) {
      cudaGetLastError();
      resident = num_sms() / (ncta + 2);
    }
    resident_cache[ncta].store(resident, std::memory_order_relaxed);
  }
  const char* force = getenv("DANET_LSTM_NB");
  const int nb = force ? atoi(force) : (clusters8 <= resident ? 8 : 16);
  // backend 2 (fp16 recurrent state) exists for the 8-per-cluster kernel only; a batch too large for that runs the
  // (more exact) bf16x3 kernel with 16 utterances per cluster instead
  bool pad_memset = false;
  const char* ver = getenv("DANET_LSTM_V");
  const bool gen1 = nb != 8 || (ver && atoi(ver) == 1 && !h_fp16);
  DANET_REQUIRE(!pre_flags || !gen1, DANET_E_SHAPE,
                "lstm_seq: the pipelined hand-over of the input projections needs the 8-utterances-per-cluster kernel "
                "(B = %d is too large for co-resident clusters)", B);
  DANET_REQUIRE(!split_tm || !gen1, DANET_E_SHAPE,
                "lstm_seq: a time-major out_split is written by the 8-utterances-per-cluster kernel only (B = %d)", B);
  if (gen1) {
    if (out_split && out_kp > n_dir * H)
      DANET_CUDA(cudaMemset2DAsync(p.out_split + n_dir * H, (size_t)out_kp * 2, 0, (size_t)(out_kp - n_dir * H) * 2,
                                   (size_t)2 * B * T, stream));
    return nb != 8 ? launch_cluster(lstm_tc_kernel<16>, p, ncta, 16, kThreads, stream)
                   : launch_cluster(lstm_tc_kernel<8>, p, ncta, 8, kThreads, stream);
  }
  if (wh_packed && lstm_tc2_packed_fits(ncta) && !getenv("DANET_LSTM_NOPACK"))
    p.Wh_packed = reinterpret_cast<const uint32_t*>(wh_packed);
  if (out_split && out_kp > n_dir * H) {
    // the K padding of the emitted operand must be zero: the generation-2 kernel writes it itself when the idle unit slots
    // of its last CTA cover it exactly (always for n_dir = 2), otherwise one memset
    if (out_kp - n_dir * H == n_dir * (ncta * kUnits - H)) p.zero_pad = 1;
    else pad_memset = true;
  }
  if (pad_memset)
    DANET_CUDA(cudaMemset2DAsync(p.out_split + n_dir * H, (size_t)out_kp * 2, 0, (size_t)(out_kp - n_dir * H) * 2,
                                 (size_t)2 * B * T, stream));
  return h_fp16 ? launch_cluster(lstm_tc2_kernel<1>, p, ncta, 8, kThreads2, stream, programmatic != 0)
                : launch_cluster(lstm_tc2_kernel<0>, p, ncta, 8, kThreads2, stream, programmatic != 0);
}

}  // namespace danet
