// K2b for WIDE layers (384 < H <= 608: the `lstm-orig` encoder, app/modules.py:140-196, 4 x 600 unidirectional) on
// tcgen05 -- backend 2 of danet_lstm_seq_fwd (recurrent state carried as ONE fp16 value, Wh as an fp16 hi/lo pair).
//
// Why a second kernel: the cluster kernel of lstm_tc.cu keeps a CTA's 128 gate rows of Wh^T (hi AND lo, all K) in tensor
// memory and exchanges h through DSMEM; at H = 600 that is 19 CTAs (> the 16-CTA cluster limit) and 600 of the 512
// TMEM columns.  Here
//   * a GROUP of ncta = ceil(H/32) CTAs (19) owns 8 utterances of one direction; CTA r owns units [32r, 32r+32), i.e.
//     128 gate rows (TMEM lane 32q + 8*gate + u  <->  unit 8q + u, the layout tcgen05.ld.16x128b hands out gate-wise);
//   * A = Wh^T[128 rows, K] stays in TENSOR MEMORY for the whole sequence: the hi image as fp16 (16*ncta columns) and
//     the residual lo = W - fp16(W) as FP8 (e4m3, scaled per gate row by a power of two; 8*ncta columns).  lo only has
//     to carry the 4-5 bits that lift the weights from fp16's 2^-12 to ~2^-16 relative, so an fp8 image (and an fp8 copy
//     of h for that product) is enough: error 2^-3.5 * |lo||h| ~ 2^-15.5 |W||h| per term, an eighth of what rounding h
//     to fp16 contributes.  Per step 2*ncta kind::f16 MMAs (K16) into one accumulator and ncta kind::f8f6f4 MMAs (K32)
//     into a second one; the epilogue adds  acc_hi + 2^-e(row) * acc_lo.  57 TS-mode MMAs at H = 600.  (First build: lo
//     as fp16, 24 of its 38 slices in tensor memory and 14 in shared memory -- an SS-mode M128 K16 MMA costs ~65 cycles
//     whatever the swizzle, 910 cycles per step.)
//   * h travels through L2 with NCCL's "LL" idea: every 8-byte word carries two fp16 values and the step number, so
//     the data IS the flag -- a producer issues plain 8-byte stores (no fence, no separate flag), a consumer polls the
//     words it needs with 16-byte volatile loads and writes the payload straight into the UMMA B operand.  One L2 round
//     trip per step instead of store -> fence -> flag -> poll -> load.  Two buffers by step parity; a word of step s+2
//     can only be written after its producer has consumed every slice of step s+1, whose producers had consumed step s;
//     Measured and dropped: completing the gather in chunks of four producers with a barrier per chunk so that the MMAs
//     of a chunk start early (955 -> 1145 us per layer: five proxy fences + arrivals per thread cost more than the
//     overlap returns), delaying the first poll by 150-600 cycles (no change: the gather is bound by the ~64 B/cycle an
//     SM pulls from L2 and by the skew between the CTAs, not by a poll that leaves a moment too early), all fp16 MMAs
//     before all fp8 MMAs (+140 cycles against interleaving them);
//   * the CTAs of a group spin on each other, so the grid is launched cooperatively (co-residency) and every spin is
//     bounded: a peer that never comes poisons the result with NaN instead of hanging the device.
// Everything else (activation math, output layouts, the emitted bf16 hi/lo operand of the next layer) is the epilogue
// of lstm_tc2_kernel<1>.
#include <stdlib.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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
constexpr int kWAccCols = 32;            // two accumulators: fp16 hi product (columns 0-15), fp8 lo product (16-31)
constexpr int kWTmemCols = 512;
constexpr int kWEpiWarps = 8;
constexpr int kWEpiThreads = 32 * kWEpiWarps;
constexpr int kWThreads = kWEpiThreads + 32;     // + the MMA warp
constexpr int kWBlk = 1024;             // one K-block (32 units) of the fp16 B operand: [8 utterance rows x 64 B | 8 zero rows]
constexpr int kWBlk8 = 512;             // the same K-block as fp8: [8 rows x 32 B | 8 zero rows], SWIZZLE_32B
constexpr int kWPreDepth = 2;
constexpr uint32_t kWPollLimit = 1u << 19;   // ~0.3 s of polling before a step is declared lost

// packed image of one (direction, CTA): [slice][128 rows][8 words] for the 2*ncta fp16 hi slices (K16 each) and the ncta
// fp8 lo slices (K32 each) -- the order of the TMEM columns -- then the 128 per-row factors 2^-e that undo lo's scaling
__host__ __device__ constexpr size_t wide_image_words(int ncta) { return (size_t)3 * ncta * kWRows * 8 + kWRows; }

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

// One LL word = 64 bits written and read as ONE scalar access (single-copy atomic at its natural alignment: payload and
// flag can never be observed torn; a .v2.u32 store would formally be two 32-bit accesses).
__device__ __forceinline__ void ll_store(uint2* dst, uint32_t payload, uint32_t flag) {
  const unsigned long long v = ((unsigned long long)flag << 32) | payload;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
// two neighbouring LL words: (x, y) = (payload, flag) of the first, (z, w) of the second
__device__ __forceinline__ uint4 ll_load2(const uint2* src) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}
// K-major SWIZZLE_32B: rows of 32 bytes (16 fp16 = one K16 slice), 8-row atoms of 256 contiguous bytes
__device__ __forceinline__ uint64_t umma_desc_k_sw32(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// D[tmem] (+)= A[tmem] * B[smem], both e4m3, K = 32: A is [M lanes] x [8 columns] of four packed bytes (element 4j in the
// low byte of column j)
__device__ __forceinline__ void umma_f8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t f16x2_to_e4m3x2(uint32_t h2) {      // low half -> low byte
  unsigned short r;
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
  return (uint32_t)r;
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
  constexpr int kMmaWarp = kWEpiWarps;

  uint8_t* sH = smem;                                                // [2 buf][ncta][kWBlk]   fp16 copy of h
  uint8_t* sH8 = sH + 2 * ncta * kWBlk;                              // [2 buf][ncta][kWBlk8]  fp8 copy of h
  uint64_t* h_full = reinterpret_cast<uint64_t*>(sH8 + 2 * ncta * kWBlk8);
  uint64_t* acc_full = h_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

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
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * ncta * (kWBlk + kWBlk8) / 16; i += kWThreads)
    reinterpret_cast<uint4*>(sH)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kWTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_a_hi = tmem_base + kWAccCols;
  const uint32_t tmem_a_lo = tmem_a_hi + (uint32_t)ncta * 16;

  // ---- one-time: the image.  Slice c of the image holds 8 words (16 K elements) of every row; the two
  // warps of a lane quadrant take alternate slices, four slices (8 x 16-byte loads) in flight per thread ----
  if (warp < kWEpiWarps) {
    const int q = warp & 3, half = warp >> 2;
    const int m = 32 * q + lane;
    const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
    const int n_sl = 3 * ncta;                        // hi slices then lo slices: TMEM columns are contiguous as well
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
      constexpr uint32_t idesc = umma_idesc_f16(kWRows, kWN);      // e4m3 x e4m3 -> f32 has the same bit pattern
      for (int s = 1; s < T; ++s) {
        const int buf = (s - 1) & 1;
        const uint64_t b0d = umma_desc_k_sw64_sbo512(smem_u32(sH + (size_t)buf * ncta * kWBlk));
        const uint64_t b8d = umma_desc_k_sw32(smem_u32(sH8 + (size_t)buf * ncta * kWBlk8));
        DANET_WPROF(0);
        mbar_wait(h_full + buf, ((s - 1) >> 1) & 1);
        DANET_WPROF(1);
        tc_fence_after();
#pragma unroll 2
        for (int j = 0; j < ncta; ++j) {
          const uint64_t bj = b0d + (uint64_t)((j * kWBlk) >> 4);
          umma_bf16_ts(tmem_acc, tmem_a_hi + (uint32_t)(j * 16), bj, idesc, j != 0);
          umma_bf16_ts(tmem_acc, tmem_a_hi + (uint32_t)(j * 16 + 8), bj + 2, idesc, 1);
          umma_f8_ts(tmem_acc + 16, tmem_a_lo + (uint32_t)(j * 8), b8d + (uint64_t)((j * kWBlk8) >> 4), idesc, j != 0);
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
    float lo_scale[4];                                   // 2^-e of this thread's four gate rows (lane 32q + 8*gate + u)
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
      lo_scale[gg] = __uint_as_float(__ldg(image + (size_t)3 * ncta * kWRows * 8 + 32 * q + 8 * gg + u));
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
        uint32_t r01[2], r23[2], l01[2], l23[2];
        tmem_ld_16x128b(t0, r01);
        tmem_ld_16x128b(t0 + ((uint32_t)16 << 16), r23);
        tmem_ld_16x128b(t0 + 16, l01);
        tmem_ld_16x128b(t0 + 16 + ((uint32_t)16 << 16), l23);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        a[0] += fmaf(lo_scale[0], __uint_as_float(l01[0]), __uint_as_float(r01[0]));
        a[1] += fmaf(lo_scale[1], __uint_as_float(l01[1]), __uint_as_float(r01[1]));
        a[2] += fmaf(lo_scale[2], __uint_as_float(l23[0]), __uint_as_float(r23[0]));
        a[3] += fmaf(lo_scale[3], __uint_as_float(l23[1]), __uint_as_float(r23[1]));
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
        const uint32_t dst8 = smem_u32(sH8 + (size_t)(s & 1) * ncta * kWBlk8);
        const uint32_t want = (uint32_t)(s + 1);
        constexpr int kPer = (kWMaxCta * 64 + kWEpiThreads - 1) / kWEpiThreads;     // 5
        uint4 v[kPer];
        uint32_t pend = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = tid + i * kWEpiThreads;
          if (idx < n_pairs) {
            v[i] = ll_load2(src + 2 * idx);
            pend |= 1u << i;
          }
        }
        // rounds: check what has landed, then re-issue ALL loads that are still stale together (one L2 round trip per
        // round; checking them one after the other cost a round trip per word: 2600 cycles per step)
        uint32_t spins = 0;
        while (pend) {
#pragma unroll
          for (int i = 0; i < kPer; ++i) {
            if ((pend >> i) & 1u) {
              const bool ok = v[i].y == want && v[i].w == want;
              if (ok || dead) {
                const int idx = tid + i * kWEpiThreads;
                const int r = idx >> 6, ww = (idx & 63) * 2;         // producer, first word inside its slice
                const uint32_t nan2 = 0x7e007e00u;                   // fp16 NaN pair: a peer that never came
                const uint32_t x = ok ? v[i].x : nan2, z = ok ? v[i].z : nan2;
                const int row = ww >> 4, kk = (ww & 15) * 2;         // utterance, first of the four units
                st_shared_v2(dst0 + (uint32_t)r * kWBlk + sw64_offset(row, kk), x, z);
                // the same four values as e4m3 for the lo product: SWIZZLE_32B rows of 32 bytes, 16-byte chunk c at c ^ bit 2 of the row
                st_shared_u32(dst8 + (uint32_t)r * kWBlk8 + (uint32_t)(row * 32 + ((((kk >> 4) ^ (row >> 2)) & 1) << 4) + (kk & 15)),
                              f16x2_to_e4m3x2(x) | (f16x2_to_e4m3x2(z) << 16));
                pend &= ~(1u << i);
              }
            }
          }
          if (!pend) break;
          if (++spins > kWPollLimit) dead = true;
#pragma unroll
          for (int i = 0; i < kPer; ++i)
            if ((pend >> i) & 1u) v[i] = ll_load2(src + 2 * (tid + i * kWEpiThreads));
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
  uint32_t* img = out + ((size_t)dir * ncta + rank) * wide_image_words(ncta);
  const float* wcol = W + (size_t)g * H + unit;
  auto residual = [&](int k) {
    const float w = (unit_ok && k < H) ? __ldg(wcol + (size_t)k * ldw) : 0.f;
    return w - __half2float(__float2half_rn(w));
  };
  // pass 1: the row's largest residual fixes its power-of-two scale: max |lo| * 2^e in (128, 256] (e4m3 reaches 448)
  float mx = 0.f;
  for (int k = 0; k < 32 * ncta; ++k) mx = fmaxf(mx, fabsf(residual(k)));
  int e = 0;
  if (mx > 0.f) {
    int ex;
    frexpf(mx, &ex);                 // mx = f * 2^ex, f in [0.5, 1)
    e = 8 - ex;
    if (e > 126) e = 126;
  }
  const float scale = ldexpf(1.f, e);
  reinterpret_cast<float*>(img + (size_t)3 * ncta * kWRows * 8)[m] = ldexpf(1.f, -e);
  for (int sl = 0; sl < 2 * ncta; ++sl) {
    uint32_t hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 16 * sl + 2 * j;
      const float w0 = (unit_ok && k < H) ? __ldg(wcol + (size_t)k * ldw) : 0.f;
      const float w1 = (unit_ok && k + 1 < H) ? __ldg(wcol + (size_t)(k + 1) * ldw) : 0.f;
      const __half2 h = __floats2half2_rn(w0, w1);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    uint4* dh = reinterpret_cast<uint4*>(img + ((size_t)sl * kWRows + m) * 8);
    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  }
  for (int sl = 0; sl < ncta; ++sl) {                 // fp8 slices: 32 K elements, element 4j + i in byte i of word j
    uint32_t lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w |= (uint32_t)__nv_cvt_float_to_fp8(residual(32 * sl + 4 * j + i) * scale, __NV_SATFINITE, __NV_E4M3) << (8 * i);
      lo[j] = w;
    }
    uint4* dl = reinterpret_cast<uint4*>(img + ((size_t)(2 * ncta + sl) * kWRows + m) * 8);
    dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
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
  const size_t need = (size_t)2 * ncta * (kWBlk + kWBlk8) + 64 + 1024;
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
