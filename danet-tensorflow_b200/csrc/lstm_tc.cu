// K2b on tcgen05 (backend 1 of danet_lstm_seq_fwd): the (Bi)LSTM recurrence of
// Model.lyr_lstm (main.py:76-132) / ops.lyr_lstm_flat (app/ops.py:139-147) as a persistent
// thread-block-cluster kernel.
//
// One cluster = one direction x 16 utterances.  CTA r of the cluster owns hidden units
// [32r, 32r+32) and keeps the 128 gate rows (4 gates x 32 units, row m = 4*unit + gate) of
// Wh^T resident in shared memory for the whole sequence as bf16 hi/lo pairs (160 KB at
// H = 300, K padded to 320).  Per step the CTA issues  D[128 gate rows, 16 utterances] =
// Wh^T[128, K] * h_{t-1}^T[K, 16]  as 60 tcgen05.mma (M128 N16 K16; bf16x3: hi*hi + hi*lo +
// lo*hi, fp32 accumulate in TMEM), the epilogue warps read TMEM, add the hoisted input
// projection, apply  c = sig(i)*g + sig(f)*c ; h = sig(o)*tanh(c)  (candidate WITHOUT tanh),
// write h_t to the output, and broadcast their 32-unit slice of h_t (bf16 hi/lo, already in
// the UMMA K-major SWIZZLE_64B layout: K-block r of the B operand IS CTA r's slice) into every
// CTA's shared memory with one cp.async.bulk (DSMEM) per peer, completing on the peer's
// mbarrier.  No global-memory round trip and no cluster-wide barrier sits on the T-step
// critical path.
#include "common.cuh"
#include "tc_common.cuh"

namespace danet {

using namespace tc;

constexpr int kUnits = 32;            // hidden units per CTA
constexpr int kRows = 128;            // gate rows per CTA = UMMA M
constexpr int kNB = 16;               // utterances per cluster = UMMA N
constexpr int kMaxCta = 10;           // cluster size limit from shared memory (H <= 320)
constexpr int kATile = kRows * 64;    // bytes of one [128 x 32] bf16 K-block of A (SW64)
constexpr int kHTile = kNB * 64;      // bytes of one [16 x 32] bf16 K-block of h (SW64) = 1 KB
constexpr int kXchLd = 17;
constexpr int kEpiThreads = 128;
constexpr int kThreads = 160;         // 4 epilogue warps + 1 MMA warp

struct LstmTcParams {
  const float* pre;        // [n_dir][T][B][4H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  long long ldw;
  float* out;              // [B][T][n_dir*H]
  float* cell_seq;         // nullable [n_dir][T][B][H]
  int n_dir, T, B, H;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// local shared -> a peer CTA's shared memory, completion counted on the peer's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                uint32_t mbar_cluster) {
  asm volatile(
      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
      : "memory");
}
// K-major SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row atoms of 512 bytes (SBO = 512)
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         (4ull << 61);
}
// byte offset of element (row, kk) inside a SW64 K-block tile (kk in [0,32))
__device__ __forceinline__ uint32_t sw64_offset(int row, int kk) {
  const int r = row & 7;
  return (uint32_t)((row >> 3) * 512 + r * 64 + ((((kk >> 3) ^ (r >> 1)) & 3) << 4) + (kk & 7) * 2);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_kernel(const LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int ncta = gridDim.x;                      // cluster size = K-blocks
  const int rank = (int)cluster_ctarank();         // == blockIdx.x
  const int bt = blockIdx.y, dir = blockIdx.z;
  const int H = p.H, T = p.T, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // shared memory carve-up (all tile bases 1024-byte aligned)
  uint8_t* sA = smem;                                          // [hi|lo][ncta][kATile]
  uint8_t* sH = sA + 2 * ncta * kATile;                        // [2 buf][ncta][hi|lo][kHTile]
  uint8_t* sStage = sH + 2 * ncta * 2 * kHTile;                // [2][hi|lo][kHTile]
  float* sXch = reinterpret_cast<float*>(sStage + 2 * 2 * kHTile);   // [128][17]
  uint64_t* h_full = reinterpret_cast<uint64_t*>(sXch + kRows * kXchLd + 2);   // 8-byte aligned: 128*17+2 floats
  uint64_t* acc_full = h_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int unit0 = rank * kUnits;
  const int b0 = bt * kNB;

  // ---- one-time: Wh^T slice -> bf16 hi/lo, UMMA K-major SW64 layout ----
  {
    const float* Wg = p.Wh[dir];
    const int Kp = ncta * 32;
    for (int i = tid; i < 4 * Kp * kUnits; i += kThreads) {
      const int u = i % kUnits, g = (i / kUnits) & 3, k = i / (4 * kUnits);
      const int unit = unit0 + u;
      float w = 0.f;
      if (unit < H && k < H) w = __ldg(Wg + (size_t)k * p.ldw + g * H + unit);
      __nv_bfloat16 hi, lo;
      split_bf16(w, hi, lo);
      const int m = 4 * u + g;
      const uint32_t off = (uint32_t)(k >> 5) * kATile + sw64_offset(m, k & 31);
      *reinterpret_cast<__nv_bfloat16*>(sA + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(sA + ncta * kATile + off) = lo;
    }
  }
  if (tid == 0) {
    mbar_init(h_full + 0, 1);
    mbar_init(h_full + 1, 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 32);
  fence_proxy_async_smem();          // generic-proxy stores of sA -> visible to tcgen05.mma
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  cluster_sync();                    // every CTA's barriers are initialised before any peer signals them

  const uint32_t h_bytes = (uint32_t)ncta * 2 * kHTile;   // one full h_t (hi+lo, all K-blocks)

  if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      if (T >= 2) mbar_arrive_expect_tx(h_full + 0, h_bytes);     // h_0 lands in buffer 0
      if (T >= 3) mbar_arrive_expect_tx(h_full + 1, h_bytes);     // h_1 lands in buffer 1
      constexpr uint32_t idesc = umma_idesc_bf16(kRows, kNB);
      const uint32_t a_base = smem_u32(sA);
      for (int s = 1; s < T; ++s) {
        const int buf = (s - 1) & 1;
        mbar_wait(h_full + buf, ((s - 1) >> 1) & 1);              // h_{s-1} complete in sH[buf]
        if (s + 1 <= T - 2) mbar_arrive_expect_tx(h_full + buf, h_bytes);   // re-arm for h_{s+1}
        tc_fence_after();
        const uint32_t h_base = smem_u32(sH + (size_t)buf * ncta * 2 * kHTile);
        for (int j = 0; j < ncta; ++j) {
          const uint64_t a_hi = umma_desc_k_sw64(a_base + j * kATile);
          const uint64_t a_lo = umma_desc_k_sw64(a_base + (ncta + j) * kATile);
          const uint64_t b_hi = umma_desc_k_sw64(h_base + j * 2 * kHTile);
          const uint64_t b_lo = umma_desc_k_sw64(h_base + j * 2 * kHTile + kHTile);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);               // 32 bytes per K16 step
            umma_bf16(tmem_acc, a_hi + adv, b_hi + adv, idesc, (j | k) != 0);
            umma_bf16(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1);
            umma_bf16(tmem_acc, a_lo + adv, b_hi + adv, idesc, 1);
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ================= epilogue warps: TMEM lane m = 32*warp + lane = 4*unit + gate =================
    const int m = tid;
    // cell-update ownership: utterance bl = lane % 16, units ub..ub+3 (ub = 8*warp + 4*(lane/16))
    const int bl = lane & 15, ub = 8 * warp + 4 * (lane >> 4);
    const int b = b0 + bl, unit = unit0 + ub;
    const bool valid = b < B && unit < H;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    float4 pre_next[4];
    auto load_pre = [&](int s) {
#pragma unroll
      for (int g = 0; g < 4; ++g) pre_next[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && s < T) {
        const int to = dir ? T - 1 - s : s;
        const float* q = p.pre + (((size_t)dir * T + to) * B + b) * 4 * H + unit;
#pragma unroll
        for (int g = 0; g < 4; ++g) pre_next[g] = __ldg(reinterpret_cast<const float4*>(q + g * H));
      }
    };
    load_pre(0);
    const int outw = p.n_dir * H;
    float* xw = sXch + m * kXchLd;
    const uint32_t stage_off = sw64_offset(bl, ub);      // 8 contiguous bytes: units ub..ub+3 of utterance bl

    for (int s = 0; s < T; ++s) {
      const int to = dir ? T - 1 - s : s;
      float a[4][4];                                     // [unit][gate]
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        a[0][g] = (&pre_next[g].x)[0]; a[1][g] = (&pre_next[g].x)[1];
        a[2][g] = (&pre_next[g].x)[2]; a[3][g] = (&pre_next[g].x)[3];
      }
      load_pre(s + 1);
      if (s > 0) {
        mbar_wait(acc_full, (s - 1) & 1);
        tc_fence_after();
        float v[16];
        {
          float lo8[8], hi8[8];
          tmem_ld_32x8(tmem_acc + ((uint32_t)(32 * warp) << 16), lo8);
          tmem_ld_32x8(tmem_acc + ((uint32_t)(32 * warp) << 16) + 8, hi8);
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[i] = lo8[i]; v[8 + i] = hi8[i]; }
        }
        tc_fence_before();
        __syncwarp();                                    // previous step's reads of sXch are done
#pragma unroll
        for (int j = 0; j < 16; ++j) xw[j] = v[j];
        __syncwarp();
#pragma unroll
        for (int uu = 0; uu < 4; ++uu)
#pragma unroll
          for (int g = 0; g < 4; ++g) a[uu][g] += sXch[(4 * (ub + uu) + g) * kXchLd + bl];
      }
      float h[4];
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const float gg = a[uu][0];
        const float ig = sigmoidf_(a[uu][1]), fg = sigmoidf_(a[uu][2]), og = sigmoidf_(a[uu][3]);
        c[uu] = ig * gg + fg * c[uu];
        h[uu] = og * tanhf(c[uu]);
      }
      if (!valid) { h[0] = h[1] = h[2] = h[3] = 0.f; }
      if (s < T - 1) {
        // my 4 units of h_s as bf16 hi/lo into the staging K-block (already UMMA layout)
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) split_bf16(h[uu], hi[uu], lo[uu]);
        uint8_t* st = sStage + (s & 1) * 2 * kHTile;
        *reinterpret_cast<uint2*>(st + stage_off) = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]));
        *reinterpret_cast<uint2*>(st + kHTile + stage_off) =
            make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]));
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (tid == 0) {
          const uint32_t src = smem_u32(st);
          const uint32_t dst_local = smem_u32(sH + ((size_t)(s & 1) * ncta + rank) * 2 * kHTile);
          const uint32_t bar_local = smem_u32(h_full + (s & 1));
          for (int d = 0; d < ncta; ++d) {
            const int peer = (rank + d) % ncta;          // stagger the destinations
            dsmem_bulk_copy(mapa(dst_local, peer), src, 2 * kHTile, mapa(bar_local, peer));
          }
        }
      }
      if (valid) {
        *reinterpret_cast<float4*>(p.out + ((size_t)b * T + to) * outw + dir * H + unit) =
            make_float4(h[0], h[1], h[2], h[3]);
        if (p.cell_seq)
          *reinterpret_cast<float4*>(p.cell_seq + (((size_t)dir * T + to) * B + b) * H + unit) =
              make_float4(c[0], c[1], c[2], c[3]);
      }
    }
  }
  // nobody leaves while a peer may still read this CTA's staging tile or signal its barriers
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 4) tmem_dealloc(tmem_acc, 32);
}

static size_t lstm_tc_smem_bytes(int ncta) {
  return (size_t)2 * ncta * kATile + (size_t)2 * ncta * 2 * kHTile + 2 * 2 * kHTile +
         (kRows * kXchLd + 2) * sizeof(float) + 64 + 1024;
}

size_t lstm_tc_workspace_bytes(int, int, int) { return 256; }

bool lstm_tc_supported(int H) { return H % 4 == 0 && (H + kUnits - 1) / kUnits <= kMaxCta; }

int lstm_tc_fwd(const float* pre, const float* const* host_Wh, long long ldw, float* out, float* cell_seq,
                int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  (void)workspace; (void)workspace_bytes;
  const int ncta = (H + kUnits - 1) / kUnits;
  DANET_REQUIRE(lstm_tc_supported(H), DANET_E_SHAPE,
                "lstm_seq: the tcgen05 backend keeps Wh resident in one cluster's shared memory and needs "
                "H <= %d (got %d); use backend 0", kMaxCta * kUnits, H);
  DANET_REQUIRE(aligned16(pre) && aligned16(out) && (!cell_seq || aligned16(cell_seq)), DANET_E_ALIGN,
                "lstm_seq: pre/out/cell_seq must be 16-byte aligned");
  const size_t smem = lstm_tc_smem_bytes(ncta);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "lstm_seq: %zu B of shared memory needed", smem);
  DANET_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ncta > 8) DANET_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  LstmTcParams p;
  p.pre = pre;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.ldw = ldw; p.out = out; p.cell_seq = cell_seq;
  p.n_dir = n_dir; p.T = T; p.B = B; p.H = H;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncta, (B + kNB - 1) / kNB, n_dir);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DANET_CUDA(cudaLaunchKernelEx(&cfg, lstm_tc_kernel, p));
  return DANET_OK;
}

}  // namespace danet
