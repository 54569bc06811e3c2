// K2b for WIDE layers (384 < H <= 608: the `lstm-orig` encoder, app/modules.py:140-196, 4 x 600 unidirectional) on
// tcgen05 -- backend 2 of danet_lstm_seq_fwd (recurrent state carried as ONE fp16 value, Wh as an fp16 hi/lo pair).
//
// Why a second kernel: the cluster kernel of lstm_tc.cu keeps a CTA's 128 gate rows of Wh^T (hi AND lo, all K) in tensor
// memory and exchanges h through DSMEM; at H = 600 that is 19 CTAs (> the 16-CTA cluster limit) and 600 of the 512
// TMEM columns.  Here
//   * a GROUP of ncta = ceil(H/32) CTAs (19) owns 8 utterances of one direction; CTA r owns units [32r, 32r+32), i.e.
//     128 gate rows (TMEM lane 32q + 8*gate + u  <->  unit 8q + u, the layout tcgen05.ld.16x128b hands out gate-wise);
//   * A = Wh^T[128 rows, K]: the hi image (16*ncta columns) and as many K16 slices of the lo image as fit stay in
//     TENSOR MEMORY (24 of 38 slices at H = 600); the remaining lo slices sit in shared memory as a K-major SWIZZLE_128B
//     tile and feed SS-mode MMAs into the same accumulator.  Per step: 62 TS + 14 SS tcgen05.mma (M128 N16 K16);
//   * h travels through L2 with NCCL's "LL" idea: every 8-byte word carries two fp16 values and the step number, so
//     the data IS the flag -- a producer issues plain 8-byte stores (no fence, no separate flag), a consumer polls the
//     words it needs with 16-byte volatile loads and writes the payload straight into the UMMA B operand.  One L2 round
//     trip per step instead of store -> fence -> flag -> poll -> load.  Two buffers by step parity; a word of step s+2
//     can only be written after its producer has consumed every slice of step s+1, whose producers had consumed step s;
//   * the CTAs of a group spin on each other, so the grid is launched cooperatively (co-residency) and every spin is
//     bounded: a peer that never comes poisons the result with NaN instead of hanging the device.
// Everything else (activation math, output layouts, the emitted bf16 hi/lo operand of the next layer) is the epilogue
// of lstm_tc2_kernel<1>.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "lstm_tc_common.cuh"

namespace danet {

using namespace tc;

constexpr int kWUnits = 32;             // hidden units per CTA
constexpr int kWRows = 128;             // gate rows per CTA = UMMA M
constexpr int kWN = 16;                 // UMMA N; columns 0-7 carry utterances
constexpr int kWNB = 8;                 // utterances per group
constexpr int kWMinCta = 13;            // below that the cluster kernel of lstm_tc.cu is the better tool
constexpr int kWMaxCta = 19;            // H <= 608
constexpr int kWAccCols = 16;
constexpr int kWTmemCols = 512;
constexpr int kWEpiWarps = 8;
constexpr int kWEpiThreads = 32 * kWEpiWarps;
constexpr int kWThreads = kWEpiThreads + 32;     // + the MMA warp
constexpr int kWBlk = 1024;             // one K-block (32 units) of the B operand: [8 utterance rows x 64 B | 8 zero rows]
constexpr int kWTile = 16384;           // one 128-row x 64-element fp16 tile of the shared-memory lo image
constexpr int kWPreDepth = 2;
constexpr uint32_t kWPollLimit = 1u << 19;   // ~0.3 s of polling before a step is declared lost

__host__ __device__ constexpr int wide_lo_tmem_slices(int ncta) {
  return (kWTmemCols - kWAccCols - 16 * ncta) / 8 < 2 * ncta ? (kWTmemCols - kWAccCols - 16 * ncta) / 8 : 2 * ncta;
}
__host__ __device__ constexpr int wide_lo_smem_tiles(int ncta) {
  return (2 * ncta - wide_lo_tmem_slices(ncta) + 3) / 4;
}
// packed image of one (direction, CTA): [TMEM part: slice][128 rows][8 words] then the shared-memory lo tiles
__host__ __device__ constexpr size_t wide_image_words(int ncta) {
  return (size_t)(2 * ncta + wide_lo_tmem_slices(ncta)) * kWRows * 8 + (size_t)wide_lo_smem_tiles(ncta) * (kWTile / 4);
}

struct LstmWideParams {
  const float* pre;             // address = dir*pre_dir + (t*B + b)*pre_row + gate*H + unit
  const uint32_t* image;        // danet_lstm_pack_wh (wide layout)
  float* out;                   // [B][T][n_dir*H]
  float* cell_seq;              // nullable [n_dir][T][B][H]
  float* gates_seq;             // nullable, indexed like pre
  __nv_bfloat16* out_split;     // nullable [2][B*T][out_kp]
  int out_kp;
  long long pre_dir, pre_row;
  uint2* xch;                   // [n_dir][groups of this launch][2 parities][ncta][128] LL words, zeroed before the launch
  int n_dir, T, B, H;
  int group0;                   // first utterance group of this launch
  long long* prof;
};

constexpr int kWProfSlots = 8;
#define DANET_WPROF(slot)                                                                  \
  do {                                                                                     \
    if (prof_on) p.prof[(size_t)s * kWProfSlots + (slot)] = clock64();                     \
  } while (0)

__device__ __forceinline__ void ll_store(uint2* dst, uint32_t payload, uint32_t flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(payload), "r"(flag) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const uint2* src) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

__global__ void __launch_bounds__(kWThreads, 1)
lstm_wide_kernel(const LstmWideParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;
  const int rank = blockIdx.x;
  const int grp = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_lo_t = wide_lo_tmem_slices(ncta);
  const int n_tiles = wide_lo_smem_tiles(ncta);
  constexpr int kMmaWarp = kWEpiWarps;

  uint8_t* sAlo = smem;                                              // [n_tiles][16 KB], SWIZZLE_128B
  uint8_t* sH = sAlo + (size_t)n_tiles * kWTile;                     // [2 buf][ncta][kWBlk]
  uint64_t* h_full = reinterpret_cast<uint64_t*>(sH + 2 * ncta * kWBlk);
  uint64_t* acc_full = h_full + 2;
  uint64_t* w_full = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int unit0 = rank * kWUnits;
  const int b0 = (p.group0 + grp) * kWNB;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
                       (tid == 0 || warp == kMmaWarp);
  const uint32_t* image = p.image + ((size_t)dir * ncta + rank) * wide_image_words(ncta);
  uint2* xch = p.xch + ((size_t)dir * gridDim.y + grp) * 2 * (size_t)ncta * 128;
  if (prof_on && tid == 0) p.prof[(size_t)T * kWProfSlots + 0] = clock64();

  if (tid == 0) {
    mbar_init(h_full + 0, kWEpiThreads);
    mbar_init(h_full + 1, kWEpiThreads);
    mbar_init(acc_full, 1);
    mbar_init(w_full, 1);
    fence_barrier_init();
    if (n_tiles > 0) {
      const uint32_t bytes = (uint32_t)n_tiles * kWTile;
      const uint32_t* src = image + (size_t)(2 * ncta + n_lo_t) * kWRows * 8;
      mbar_arrive_expect_tx(w_full, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sAlo)), "l"(src), "r"(bytes), "r"(smem_u32(w_full)) : "memory");
    }
  }
  for (int i = tid; i < 2 * ncta * kWBlk / 16; i += kWThreads) reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kWTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_a_hi = tmem_base + kWAccCols;
  const uint32_t tmem_a_lo = tmem_a_hi + (uint32_t)ncta * 16;

  // ---- one-time: the TMEM part of the image.  Slice c of the image holds 8 words (16 K elements) of every row; the two
  // warps of a lane quadrant take alternate slices, four slices (8 x 16-byte loads) in flight per thread ----
  if (warp < kWEpiWarps) {
    const int q = warp & 3, half = warp >> 2;
    const int m = 32 * q + lane;
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    const int n_sl = 2 * ncta + n_lo_t;               // hi slices then lo slices: TMEM columns are contiguous as well
    for (int c0 = half; c0 < n_sl; c0 += 8) {
      uint4 v[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + 2 * i;
        if (c < n_sl) {
          const uint4* src = reinterpret_cast<const uint4*>(image + ((size_t)c * kWRows + m) * 8);
          v[i][0] = __ldg(src);
          v[i][1] = __ldg(src + 1);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + 2 * i;
        if (c < n_sl) {
          const uint32_t w[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
          tmem_st_32x8(tmem_a_hi + lane_sel + (uint32_t)c * 8, w);
        }
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (prof_on && tid == 0) p.prof[(size_t)T * kWProfSlots + 1] = clock64();

  if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_f16(kWRows, kWN);
      if (n_tiles > 0) mbar_wait(w_full, 0);
      const uint64_t a_s0 = umma_desc_k_sw128(smem_u32(sAlo));
      const int jt = n_lo_t >> 1;                     // K-blocks whose lo slices are both in tensor memory
      for (int s = 1; s < T; ++s) {
        const int buf = (s - 1) & 1;
        const uint64_t b0d = umma_desc_k_sw64_sbo512(smem_u32(sH + (size_t)buf * ncta * kWBlk));
        DANET_WPROF(0);
        mbar_wait(h_full + buf, ((s - 1) >> 1) & 1);
        DANET_WPROF(1);
        tc_fence_after();
#pragma unroll 2
        for (int j = 0; j < jt; ++j) {
          const uint64_t bj = b0d + (uint64_t)((j * kWBlk) >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ac = (uint32_t)(j * 16 + k * 8);
            umma_bf16_ts(tmem_acc, tmem_a_hi + ac, bj + (uint64_t)(k * 2), idesc, (j | k) != 0);
            umma_bf16_ts(tmem_acc, tmem_a_lo + ac, bj + (uint64_t)(k * 2), idesc, 1);
          }
        }
        for (int j = jt; j < ncta; ++j) {
          const uint64_t bj = b0d + (uint64_t)((j * kWBlk) >> 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int sl = 2 * j + k;
            const uint32_t ac = (uint32_t)(sl * 8);
            umma_bf16_ts(tmem_acc, tmem_a_hi + ac, bj + (uint64_t)(k * 2), idesc, sl != 0);
            if (sl < n_lo_t) {
              umma_bf16_ts(tmem_acc, tmem_a_lo + ac, bj + (uint64_t)(k * 2), idesc, 1);
            } else {
              const int js = sl - n_lo_t;
              umma_bf16(tmem_acc, a_s0 + (uint64_t)(((js >> 2) * kWTile + (js & 3) * 32) >> 4), bj + (uint64_t)(k * 2),
                        idesc, 1);
            }
          }
        }
        umma_commit(acc_full);
        DANET_WPROF(2);
      }
    }
  } else {
    // ================= epilogue warps: quadrant q = warp % 4, utterances 4*hw .. 4*hw+3 =================
    const int q = warp & 3, hw = warp >> 2;
    const int u = lane >> 2, g = lane & 3;
    const int ul = 8 * q + u;
    const int bl = 4 * hw + g;
    const int b = b0 + bl, unit = unit0 + ul;
    const bool valid = b < B && unit < H;
    const bool odd = (u & 1) != 0;
    float c = 0.f;
    float pre_q[kWPreDepth][4];
    auto load_pre = [&](int s, float (&dst)[4]) {
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) dst[gg] = 0.f;
      if (valid && s < T) {
        const int to = dir ? T - 1 - s : s;
        const float* qp = p.pre + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) dst[gg] = __ldcg(qp + gg * H);
      }
    };
#pragma unroll
    for (int d = 0; d < kWPreDepth; ++d) load_pre(d, pre_q[d]);
    const int outw = p.n_dir * H;
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    uint2* my_word = xch + (size_t)rank * 128 + bl * 16 + (ul >> 1);        // + parity * ncta * 128
    const int n_pairs = ncta * 64;                       // 16-byte units (two LL words) a CTA gathers per step
    bool dead = false;
    constexpr float kL2e = 1.4426950408889634f;
    for (int s = 0; s < T; ++s) {
      const int to = dir ? T - 1 - s : s;
      float a[4];
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        a[gg] = pre_q[0][gg];
#pragma unroll
        for (int d = 0; d + 1 < kWPreDepth; ++d) pre_q[d][gg] = pre_q[d + 1][gg];
      }
      load_pre(s + kWPreDepth, pre_q[kWPreDepth - 1]);
      if (s > 0) {
        mbar_wait(acc_full, (s - 1) & 1);
        DANET_WPROF(3);
        tc_fence_after();
        const uint32_t t0 = tmem_acc + lane_sel + 4 * hw;
        uint32_t r01[2], r23[2];
        tmem_ld_16x128b(t0, r01);
        tmem_ld_16x128b(t0 + ((uint32_t)16 << 16), r23);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        a[0] += __uint_as_float(r01[0]); a[1] += __uint_as_float(r01[1]);
        a[2] += __uint_as_float(r23[0]); a[3] += __uint_as_float(r23[1]);
        tc_fence_before();
      }
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
      const float hn = __shfl_xor_sync(0xffffffffu, h, 4);
      const float he = odd ? hn : h, ho = odd ? h : hn;            // units (ul & ~1), (ul | 1)
      if (s < T - 1 && !odd) {
        const __half2 hp = __floats2half2_rn(he, ho);
        ll_store(my_word + (size_t)(s & 1) * ncta * 128, *reinterpret_cast<const uint32_t*>(&hp), (uint32_t)(s + 1));
      }
      DANET_WPROF(4);
      if (valid) {
        if (odd && p.out_split) {
          uint32_t vh, vl;
          split2_bf16(he, ho, vh, vl);
          __nv_bfloat16* oh = p.out_split + ((size_t)b * T + to) * p.out_kp + dir * H + (unit - 1);
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
      if (s < T - 1) {
        // gather every CTA's slice of h_s (our own included) into the B operand of step s + 1
        const uint2* src = xch + (size_t)(s & 1) * ncta * 128;
        const uint32_t dst0 = smem_u32(sH + (size_t)(s & 1) * ncta * kWBlk);
        const uint32_t want = (uint32_t)(s + 1);
        constexpr int kPer = (kWMaxCta * 64 + kWEpiThreads - 1) / kWEpiThreads;     // 5
        uint4 v[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = tid + i * kWEpiThreads;
          if (idx < n_pairs) v[i] = ll_load2(src + 2 * idx);
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = tid + i * kWEpiThreads;
          if (idx < n_pairs) {
            uint32_t spins = 0;
            while ((v[i].y != want || v[i].w != want) && !dead) {
              if (++spins > kWPollLimit) { dead = true; break; }
              v[i] = ll_load2(src + 2 * idx);
            }
            if (dead) v[i].x = v[i].z = 0x7e007e00u;               // fp16 NaN pairs
            const int r = idx >> 6, ww = (idx & 63) * 2;             // producer, first word inside its slice
            st_shared_v2(dst0 + (uint32_t)r * kWBlk + sw64_offset(ww >> 4, (ww & 15) * 2), v[i].x, v[i].z);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(h_full + (s & 1));
        DANET_WPROF(5);
      }
    }
  }
  if (prof_on && tid == 0) p.prof[(size_t)T * kWProfSlots + 2] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kWTmemCols);
}

// ---- Wh -> the wide kernel's image, once per weight update --------------------------------------------------------
// grid (ncta, n_dir), 128 threads: thread m = TMEM lane m = 32q + 8*gate + u  <->  unit 32*rank + 8q + u
__global__ void lstm_wide_pack_kernel(const float* W0, const float* W1, long long ldw, int H, int ncta, uint32_t* out) {
  const int rank = blockIdx.x, dir = blockIdx.y, m = threadIdx.x;
  const int unit = rank * kWUnits + 8 * (m >> 5) + (m & 7), g = (m >> 3) & 3;
  const float* W = dir ? W1 : W0;
  const bool unit_ok = unit < H;
  const int n_lo_t = wide_lo_tmem_slices(ncta);
  uint32_t* img = out + ((size_t)dir * ncta + rank) * wide_image_words(ncta);
  uint8_t* tiles = reinterpret_cast<uint8_t*>(img + (size_t)(2 * ncta + n_lo_t) * kWRows * 8);
  for (int sl = 0; sl < 2 * ncta; ++sl) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 16 * sl + 2 * j;
      const float w0 = (unit_ok && k < H) ? __ldg(W + (size_t)k * ldw + (size_t)g * H + unit) : 0.f;
      const float w1 = (unit_ok && k + 1 < H) ? __ldg(W + (size_t)(k + 1) * ldw + (size_t)g * H + unit) : 0.f;
      split2_f16(w0, w1, hi[j], lo[j]);
    }
    uint4* dh = reinterpret_cast<uint4*>(img + ((size_t)sl * kWRows + m) * 8);
    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    if (sl < n_lo_t) {
      uint4* dl = reinterpret_cast<uint4*>(img + ((size_t)(2 * ncta + sl) * kWRows + m) * 8);
      dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    } else {
      // K-major SWIZZLE_128B tile: row m = 128 bytes (64 elements), 8-row atoms of 1024 bytes, 16-byte chunk c of row r
      // stored at chunk position c ^ (r & 7); slice js covers elements 16*(js & 3) .. +15 of tile js >> 2
      const int js = sl - n_lo_t;
      uint8_t* row = tiles + (size_t)(js >> 2) * kWTile + (size_t)(m >> 3) * 1024 + (size_t)(m & 7) * 128;
      const int c0 = 2 * (js & 3);
      *reinterpret_cast<uint4*>(row + (((c0) ^ (m & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(row + (((c0 + 1) ^ (m & 7)) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
  }
  // slices of the last tile beyond 2*ncta stay zero
  const int n_s = 2 * ncta - n_lo_t;
  for (int js = n_s; js < 4 * wide_lo_smem_tiles(ncta); ++js) {
    uint8_t* row = tiles + (size_t)(js >> 2) * kWTile + (size_t)(m >> 3) * 1024 + (size_t)(m & 7) * 128;
    const int c0 = 2 * (js & 3);
    *reinterpret_cast<uint4*>(row + (((c0) ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(row + (((c0 + 1) ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
}

static int wide_ncta(int H) { return (H + kWUnits - 1) / kWUnits; }

bool lstm_wide_supported(int H) { return H % 4 == 0 && wide_ncta(H) >= kWMinCta && wide_ncta(H) <= kWMaxCta; }

size_t lstm_wide_pack_bytes(int n_dir, int H) { return (size_t)n_dir * wide_ncta(H) * wide_image_words(wide_ncta(H)) * 4; }

static size_t wide_xch_bytes(int n_dir, int B, int H) {
  return (size_t)n_dir * ((B + kWNB - 1) / kWNB) * 2 * wide_ncta(H) * 128 * sizeof(uint2);
}
// exchange buffer + room for an image packed on the fly (callers that cache danet_lstm_pack_wh never touch that part)
size_t lstm_wide_workspace_bytes(int n_dir, int B, int H) {
  return ((wide_xch_bytes(n_dir, B, H) + 255) / 256) * 256 + lstm_wide_pack_bytes(n_dir, H) + 256;
}

int lstm_wide_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed, cudaStream_t stream) {
  DANET_REQUIRE(lstm_wide_supported(H), DANET_E_SHAPE, "lstm_pack_wh: H %d is outside the wide tcgen05 kernel's range", H);
  DANET_REQUIRE(aligned16(packed), DANET_E_ALIGN, "lstm_pack_wh: packed must be 16-byte aligned");
  const int ncta = wide_ncta(H);
  lstm_wide_pack_kernel<<<dim3(ncta, n_dir), kWRows, 0, stream>>>(host_Wh[0], n_dir > 1 ? host_Wh[1] : host_Wh[0], ldw, H, ncta,
                                                                 reinterpret_cast<uint32_t*>(packed));
  DANET_CUDA(cudaGetLastError());
  return DANET_OK;
}

static size_t wide_smem_bytes(int ncta) {
  const size_t need = (size_t)wide_lo_smem_tiles(ncta) * kWTile + (size_t)2 * ncta * kWBlk + 64 + 1024;
  const size_t whole_sm = 227 * 1024;       // keep other streams' CTAs off the SM: the step is latency-bound (lstm_tc.cu)
  return need > whole_sm ? need : whole_sm;
}

int lstm_wide_fwd(const float* pre, long long pre_dir, long long pre_row, const float* const* host_Wh, long long ldw,
                  const void* wh_packed, float* out, float* cell_seq, float* gates_seq, void* out_split, int out_kp,
                  int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  DANET_REQUIRE(lstm_wide_supported(H), DANET_E_SHAPE, "lstm_seq: H %d is outside the wide tcgen05 kernel's range", H);
  DANET_REQUIRE(aligned16(pre) && aligned16(out) && aligned16(workspace), DANET_E_ALIGN,
                "lstm_seq: pre/out/workspace must be 16-byte aligned");
  DANET_REQUIRE(!wh_packed || aligned16(wh_packed), DANET_E_ALIGN, "lstm_seq: wh_packed must be 16-byte aligned");
  DANET_REQUIRE(workspace_bytes >= lstm_wide_workspace_bytes(n_dir, B, H), DANET_E_WORKSPACE, "lstm_seq: workspace %zu < %zu",
                workspace_bytes, lstm_wide_workspace_bytes(n_dir, B, H));
  const int ncta = wide_ncta(H);
  const size_t xch_bytes = wide_xch_bytes(n_dir, B, H);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (!wh_packed) {
    void* img = ws + ((xch_bytes + 255) / 256) * 256;
    const int rc = lstm_wide_pack_wh(host_Wh, ldw, n_dir, H, img, stream);
    if (rc != DANET_OK) return rc;
    wh_packed = img;
  }
  LstmWideParams p;
  p.pre = pre;
  p.image = reinterpret_cast<const uint32_t*>(wh_packed);
  p.out = out; p.cell_seq = cell_seq; p.gates_seq = gates_seq;
  p.out_split = reinterpret_cast<__nv_bfloat16*>(out_split);
  p.out_kp = out_kp;
  p.pre_dir = pre_dir; p.pre_row = pre_row;
  p.n_dir = n_dir; p.T = T; p.B = B; p.H = H;
  p.prof = nullptr;
  if (out_split) {
    DANET_REQUIRE(out_kp >= n_dir * H && out_kp % 64 == 0 && aligned16(out_split), DANET_E_SHAPE,
                  "lstm_seq: out_split needs a 16-byte aligned buffer with row length %d >= %d, multiple of 64", out_kp, n_dir * H);
    if (out_kp > n_dir * H)
      DANET_CUDA(cudaMemset2DAsync(p.out_split + n_dir * H, (size_t)out_kp * 2, 0, (size_t)(out_kp - n_dir * H) * 2,
                                   (size_t)2 * B * T, stream));
  }
  const size_t smem = wide_smem_bytes(ncta);
  DANET_CUDA(cudaFuncSetAttribute(lstm_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_groups = (B + kWNB - 1) / kWNB;
  int per_launch = num_sms() / (ncta * n_dir);
  DANET_REQUIRE(per_launch >= 1, DANET_E_SHAPE, "lstm_seq: one utterance group needs %d resident CTAs", ncta * n_dir);
  if (per_launch > n_groups) per_launch = n_groups;
  DANET_CUDA(cudaMemsetAsync(ws, 0, xch_bytes, stream));
  const size_t prof_off = ((xch_bytes + 255) / 256) * 256 + lstm_wide_pack_bytes(n_dir, H);
  if (getenv("DANET_LSTM_PROFILE") && workspace_bytes >= prof_off + (size_t)(T + 1) * kWProfSlots * sizeof(long long)) {
    p.prof = reinterpret_cast<long long*>(ws + prof_off);
    DANET_CUDA(cudaMemsetAsync(p.prof, 0, (size_t)(T + 1) * kWProfSlots * sizeof(long long), stream));
  }
  for (int g0 = 0; g0 < n_groups; g0 += per_launch) {
    const int ng = n_groups - g0 < per_launch ? n_groups - g0 : per_launch;
    p.group0 = g0;
    // every launch gets its own slice of the exchange buffer ([dir][ng groups] inside the launch)
    p.xch = reinterpret_cast<uint2*>(ws) + (size_t)g0 * n_dir * 2 * ncta * 128;
    void* args[] = {&p};
    DANET_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_wide_kernel, dim3(ncta, ng, n_dir), dim3(kWThreads), args, smem,
                                           stream));
  }
  return DANET_OK;
}

}  // namespace danet
