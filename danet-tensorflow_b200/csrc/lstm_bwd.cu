// K6 (part): backward-through-time of the (Bi)LSTM layer -- what TF derives for the tf.scan of
// main.py:125-131 over ops.lyr_lstm_flat (app/ops.py:139-147):
//   forward   a = pre + h_{t-1} Wh ; g = a_g (no tanh) ; i,f,o = sigmoid ; c = i g + f c_{t-1} ; h = o tanh(c)
//   backward  dh = d_out_t + da_{t+1} Wh^T
//             dc = dh o (1 - tanh(c)^2) + dc_{t+1} f_{t+1}
//             da_g = dc i ; da_i = dc g i(1-i) ; da_f = dc c_{t-1} f(1-f) ; da_o = dh tanh(c) o(1-o)
// The forward kernel leaves the post-activation gates [g|i|f|o] in the pre-activation buffer and the cell
// states in cell_seq; this kernel walks the sequence in reverse and overwrites the gates with da in place
// (the dW = X^T da / dX = da W^T products are danet_gemm calls on that buffer).
// Exact fp32, persistent cooperative kernel: a CTA owns 16 hidden units x 16 utterances of one direction,
// keeps its 16 rows of Wh (16 x 4H) in shared memory, and per step pulls the group's da_{t+1} [16 x 4H]
// through L2 (release/acquire counter, as lstm.cu).  backend 1 is the tcgen05 cluster kernel (lstm_bwd_tc.cu).
#include "common.cuh"

namespace danet {

int lstm_bwd_tc(const float* d_out, float* gates, const float* cell_seq, const float* const* host_Wh, long long ldw,
                int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream);
// wide layers (384 < H <= 608), backend 2: lstm_wide_bwd_tc.cu
bool lstm_wide_bwd_supported(int H);
size_t lstm_wide_bwd_workspace_bytes(int n_dir, int B, int H);
int lstm_wide_bwd(const float* d_out, float* gates, const float* cell_seq, const float* const* host_Wh, long long ldw,
                  int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// kKS = split of the 4H reduction across lanes: 4 for the 16 x 16 tile (H <= 320), 16 for the 8 x 8 tile of wide layers
// (H = 600: 256 threads per CTA instead of 64 -- with 64, pulling the 77 KB da tile of a step through L2 took ~19
// dependent load rounds per thread: 28 us per step, 240 ms per training step of `lstm-orig`)

__device__ __forceinline__ int ld_acquire_i(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct LstmBwdParams {
  const float* d_out;      // [B][T][n_dir*H]
  float* gates;            // [n_dir][T][B][4H]: in gates [g|i|f|o], out da
  const float* cell_seq;   // [n_dir][T][B][H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  long long ldw;
  int* counters;           // [n_dir][n_bt]
  int n_dir, T, B, H, bt0, n_bt_total;
  int dir0;                // first direction of this launch (wide layers: one direction per launch)
  int ldd;                 // row stride (floats) of the shared da tile, padded against bank conflicts
};

// kBU hidden units x kBB utterances per CTA (16 x 16 for H <= 320; 8 x 8 when 4H rows no longer fit)
template <int kBU, int kBB, int kKS>
__global__ void __launch_bounds__((kBU / 2) * (kBB / 2) * kKS)
lstm_bwd_kernel(LstmBwdParams p) {
  constexpr int NT = (kBU / 2) * (kBB / 2) * kKS;
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, T = p.T, B = p.B, G4 = 4 * H;
  float* sW = smem;                        // [kBU][4H]
  float* sD = sW + (size_t)kBU * G4;       // [kBB][ldd]  da_{t+1} of this batch tile
  const int ldd = p.ldd;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, bt = p.bt0 + blockIdx.y, dir = p.dir0 + blockIdx.z;
  const int n_chunks = gridDim.x;
  const int u0 = chunk * kBU, b0 = bt * kBB;

  const float* Wg = p.Wh[dir];
  for (int i = tid; i < kBU * (G4 / 4); i += NT) {
    const int u = i / (G4 / 4), q = i % (G4 / 4);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u0 + u < H) w = __ldg(reinterpret_cast<const float4*>(Wg + (size_t)(u0 + u) * p.ldw) + q);
    reinterpret_cast<float4*>(sW)[i] = w;
  }
  // thread -> 2 utterances x 2 units, one quarter of the 4H reduction; 64 tiles x 4 quarters
  const int ks = tid & (kKS - 1), tile = tid / kKS;
  const int tb = (tile % (kBB / 2)) * 2, tu = (tile / (kBB / 2)) * 2;   // local utterance / unit of the 2x2 tile
  const int kq = (G4 / 4 + kKS - 1) / kKS;                      // float4 per quarter (ceil)
  const int q_lo = ks * kq, q_hi = min(G4 / 4, q_lo + kq);
  int* counter = p.counters + dir * p.n_bt_total + bt;
  const int outw = p.n_dir * H;

  // each (utterance, unit) pair is finalised by the ks == 0 lane of its tile: it carries dc
  float dc_next[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  __syncthreads();

  for (int s = T - 1; s >= 0; --s) {       // processing index; original time of this step:
    const int to = dir ? T - 1 - s : s;
    const int tp = dir ? to + 1 : to - 1;  // original time of the previously processed state c_{t-1}
    // everything that does not depend on the recurrence is fetched BEFORE waiting for da_{s+1}
    float gg[2][2], ig[2][2], fg[2][2], og[2][2], cc[2][2], cp[2][2], dout[2][2];
    if (ks == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int b = b0 + tb + i, unit = u0 + tu + j;
          const bool ok = b < B && unit < H;
          const float* gt = p.gates + (((size_t)dir * T + to) * B + (ok ? b : 0)) * G4 + (ok ? unit : 0);
          gg[i][j] = __ldcg(gt); ig[i][j] = __ldcg(gt + H); fg[i][j] = __ldcg(gt + 2 * H); og[i][j] = __ldcg(gt + 3 * H);
          const size_t ci = ((size_t)dir * T * B + (ok ? b : 0)) * H + (ok ? unit : 0);
          cc[i][j] = __ldg(p.cell_seq + ci + (size_t)to * B * H);
          cp[i][j] = s > 0 ? __ldg(p.cell_seq + ci + (size_t)tp * B * H) : 0.f;
          dout[i][j] = __ldg(p.d_out + ((size_t)(ok ? b : 0) * T + to) * outw + dir * H + (ok ? unit : 0));
        }
    }
    float dh[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (s < T - 1) {
      // da of processing step s+1 (original time tn) from every CTA of the group
      const int tn = dir ? to - 1 : to + 1;
      if (tid == 0) {
        const int want = n_chunks * (T - 1 - s);
        while (ld_acquire_i(counter) < want) {}
      }
      __syncthreads();
      {
        constexpr int kUnroll = 8;
        const int nq = G4 / 4, total = kBB * nq;
        const float4* src = reinterpret_cast<const float4*>(p.gates + (((size_t)dir * T + tn) * B + b0) * G4);
        const int rows_ok = min(kBB, B - b0);
        for (int i0 = tid; i0 < total; i0 += NT * kUnroll) {
          float4 v[kUnroll];
#pragma unroll
          for (int k = 0; k < kUnroll; ++k) {
            const int i = i0 + k * NT;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total && i / nq < rows_ok) v[k] = __ldcg(src + i);     // rows of one time step are contiguous
          }
#pragma unroll
          for (int k = 0; k < kUnroll; ++k) {
            const int i = i0 + k * NT;
            if (i < total) *reinterpret_cast<float4*>(sD + (size_t)(i / nq) * ldd + 4 * (i % nq)) = v[k];
          }
        }
      }
      __syncthreads();
      const float4* d0 = reinterpret_cast<const float4*>(sD + (size_t)tb * ldd);
      const float4* d1 = reinterpret_cast<const float4*>(sD + (size_t)(tb + 1) * ldd);
      const float4* w0 = reinterpret_cast<const float4*>(sW + (size_t)tu * G4);
      const float4* w1 = reinterpret_cast<const float4*>(sW + (size_t)(tu + 1) * G4);
#pragma unroll 2
      for (int q = q_lo; q < q_hi; ++q) {
        const float4 a0 = d0[q], a1 = d1[q], x0 = w0[q], x1 = w1[q];
        dh[0][0] = fmaf(a0.x, x0.x, dh[0][0]); dh[0][0] = fmaf(a0.y, x0.y, dh[0][0]);
        dh[0][0] = fmaf(a0.z, x0.z, dh[0][0]); dh[0][0] = fmaf(a0.w, x0.w, dh[0][0]);
        dh[0][1] = fmaf(a0.x, x1.x, dh[0][1]); dh[0][1] = fmaf(a0.y, x1.y, dh[0][1]);
        dh[0][1] = fmaf(a0.z, x1.z, dh[0][1]); dh[0][1] = fmaf(a0.w, x1.w, dh[0][1]);
        dh[1][0] = fmaf(a1.x, x0.x, dh[1][0]); dh[1][0] = fmaf(a1.y, x0.y, dh[1][0]);
        dh[1][0] = fmaf(a1.z, x0.z, dh[1][0]); dh[1][0] = fmaf(a1.w, x0.w, dh[1][0]);
        dh[1][1] = fmaf(a1.x, x1.x, dh[1][1]); dh[1][1] = fmaf(a1.y, x1.y, dh[1][1]);
        dh[1][1] = fmaf(a1.z, x1.z, dh[1][1]); dh[1][1] = fmaf(a1.w, x1.w, dh[1][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int o = 1; o < kKS; o <<= 1) dh[i][j] += __shfl_xor_sync(0xffffffffu, dh[i][j], o);
        }
    }
    if (ks == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int b = b0 + tb + i, unit = u0 + tu + j;
          if (b < B && unit < H) {
            float* gt = p.gates + (((size_t)dir * T + to) * B + b) * G4 + unit;
            const float dht = dh[i][j] + dout[i][j];
            const float th = tanhf(cc[i][j]);
            const float dc = dht * og[i][j] * (1.f - th * th) + dc_next[i][j];
            dc_next[i][j] = dc * fg[i][j];
            __stcg(gt, dc * ig[i][j]);                                              // da_g (candidate has no tanh)
            __stcg(gt + H, dc * gg[i][j] * ig[i][j] * (1.f - ig[i][j]));            // da_i
            __stcg(gt + 2 * H, dc * cp[i][j] * fg[i][j] * (1.f - fg[i][j]));        // da_f
            __stcg(gt + 3 * H, dht * th * og[i][j] * (1.f - og[i][j]));             // da_o
          }
        }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(counter, 1);
    }
  }
}

template <int kBU, int kBB, int kKS>
static int launch_lstm_bwd(LstmBwdParams p, cudaStream_t st) {
  constexpr int NT = (kBU / 2) * (kBB / 2) * kKS;
  const int H = p.H, B = p.B, n_dir = p.n_dir;
  // pad the da rows so the 8 lanes of a quarter warp (4 reduction quarters x 2 utterance pairs) hit
  // distinct 16-byte bank groups: without it the 16 x 4H tile serialises 8-way on every LDS.128
  {
    const int q4 = 4 * H / 4, kq = (q4 + kKS - 1) / kKS;
    int best_pad = 0, best = -1;
    for (int pad = 0; pad < 8; ++pad) {
      unsigned seen = 0;
      for (int t = 0; t < 2; ++t)
        for (int ks = 0; ks < kKS; ++ks) seen |= 1u << ((ks * kq + t * 2 * (q4 + pad)) & 7);
      const int distinct = __builtin_popcount(seen);
      if (distinct > best) { best = distinct; best_pad = pad; }
    }
    p.ldd = 4 * (q4 + best_pad);
  }
  const size_t smem = ((size_t)kBU * 4 * H + (size_t)kBB * p.ldd) * sizeof(float);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "lstm_seq_bwd: H %d needs %zu B of shared memory", H, smem);
  DANET_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<kBU, kBB, kKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DANET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_bwd_kernel<kBU, kBB, kKS>, NT, smem));
  const int n_chunks = (H + kBU - 1) / kBU;
  const int n_bt = (B + kBB - 1) / kBB;
  const int resident = per_sm * num_sms();
  // the two directions are independent: when one batch tile of both does not fit the device (H = 600: 2 x 75 CTAs), they
  // run in separate launches
  const int dirs_per_launch = resident >= n_dir * n_chunks ? n_dir : 1;
  int bt_per_launch = resident / (dirs_per_launch * n_chunks);
  DANET_REQUIRE(bt_per_launch >= 1, DANET_E_SHAPE, "lstm_seq_bwd: one batch tile needs %d resident CTAs, device holds %d",
                n_chunks, resident);
  if (bt_per_launch > n_bt) bt_per_launch = n_bt;
  DANET_CUDA(cudaMemsetAsync(p.counters, 0, (size_t)n_dir * n_bt * sizeof(int), st));
  p.n_bt_total = n_bt;
  for (int dir0 = 0; dir0 < n_dir; dir0 += dirs_per_launch)
    for (int bt0 = 0; bt0 < n_bt; bt0 += bt_per_launch) {
      p.bt0 = bt0;
      p.dir0 = dir0;
      const int nb = (n_bt - bt0 < bt_per_launch) ? n_bt - bt0 : bt_per_launch;
      void* args[] = {&p};
      DANET_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_bwd_kernel<kBU, kBB, kKS>, dim3(n_chunks, nb, dirs_per_launch),
                                             dim3(NT), args, smem, st));
    }
  return DANET_OK;
}

}  // namespace danet

using namespace danet;

extern "C" size_t danet_lstm_seq_bwd_workspace_bytes(int n_dir, int B, int H) {
  if (n_dir < 1 || B < 1) return 256;
  const size_t counters = (((size_t)n_dir * ((B + 7) / 8) * sizeof(int)) + 255) / 256 * 256;
  if (H >= 4 && H % 4 == 0 && lstm_wide_bwd_supported(H)) {
    const size_t wide = lstm_wide_bwd_workspace_bytes(n_dir, B, H);
    return wide > counters ? wide : counters;
  }
  return counters;
}

extern "C" int danet_lstm_seq_bwd(const float* d_out, float* gates, const float* cell_seq,
                                  const float* const* host_Wh, long long ldw, int n_dir, int T, int B, int H,
                                  void* workspace, size_t workspace_bytes, int backend, void* stream) {
  DANET_REQUIRE(d_out && gates && cell_seq && host_Wh && workspace, DANET_E_ARG, "lstm_seq_bwd: null pointer");
  DANET_REQUIRE(n_dir == 1 || n_dir == 2, DANET_E_SHAPE, "lstm_seq_bwd: n_dir %d", n_dir);
  for (int d = 0; d < n_dir; ++d) DANET_REQUIRE(host_Wh[d], DANET_E_ARG, "lstm_seq_bwd: null Wh[%d]", d);
  DANET_REQUIRE(T >= 0 && B >= 0 && H >= 4 && H % 4 == 0 && ldw >= 4ll * H && ldw % 4 == 0, DANET_E_SHAPE,
                "lstm_seq_bwd: T %d B %d H %d ldw %lld", T, B, H, ldw);
  DANET_REQUIRE(aligned16(gates) && aligned16(host_Wh[0]), DANET_E_ALIGN, "lstm_seq_bwd: gates / Wh must be 16-byte aligned");
  DANET_REQUIRE(workspace_bytes >= danet_lstm_seq_bwd_workspace_bytes(n_dir, B, H), DANET_E_WORKSPACE,
                "lstm_seq_bwd: workspace too small");
  DANET_REQUIRE(backend >= 0 && backend <= 2, DANET_E_ARG, "lstm_seq_bwd: backend %d", backend);
  if (T == 0 || B == 0) return DANET_OK;
  cudaStream_t st = as_stream(stream);
  if (backend == 2)
    return lstm_wide_bwd(d_out, gates, cell_seq, host_Wh, ldw, n_dir, T, B, H, workspace, workspace_bytes, st);
  if (backend == 1) return lstm_bwd_tc(d_out, gates, cell_seq, host_Wh, ldw, n_dir, T, B, H, workspace, workspace_bytes, st);
  LstmBwdParams p;
  p.d_out = d_out; p.gates = gates; p.cell_seq = cell_seq;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.ldw = ldw;
  p.counters = reinterpret_cast<int*>(workspace);
  p.n_dir = n_dir; p.T = T; p.B = B; p.H = H; p.n_bt_total = 0; p.bt0 = 0; p.dir0 = 0;
  return H <= 320 ? launch_lstm_bwd<16, 16, 4>(p, st) : launch_lstm_bwd<8, 8, 16>(p, st);
}
