// K2b: (Bi)LSTM sequence kernel -- Model.lyr_lstm (main.py:76-132, tf.scan from zero
// state) over ops.lyr_lstm_flat (app/ops.py:139-147) and _lyr_bilstm (app/modules.py:120-137).
//   a = pre_t + h_{t-1} * Wh ; g = a[0:H] (NO tanh) ; i,f,o = sigmoid(a[H:4H])
//   c_t = i*g + f*c_{t-1} ; h_t = o*tanh(c_t) ; direction 1 walks t = T-1 .. 0.
// backend 0 (this file): persistent fp32 kernel.  One CTA owns 8 hidden units (their 4
// gate columns of Wh stay in shared memory for the whole sequence) of 16 utterances of one
// direction; the CTAs of one (direction, batch tile) group exchange h_t through L2 (the
// output tensor itself) and a release/acquire counter.  Launched cooperatively so every
// CTA of a group is resident.  backend 1: tcgen05 cluster kernel (lstm_tc.cu).
#include "common.cuh"

namespace danet {

int lstm_tc_fwd(const float* pre, long long pre_dir, long long pre_row, const float* const* host_Wh, long long ldw,
                const void* wh_packed,
                float* out, float* cell_seq, float* gates_seq, void* out_split, int out_kp, int n_dir, int T, int B,
                int H, int h_fp16, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                const int* pre_flags = nullptr, int flag_need = 0, int flags_tm = 0, int split_tm = 0,
                int programmatic = 0);
size_t lstm_tc_workspace_bytes(int n_dir, int B, int H);
size_t lstm_tc_pack_bytes(int n_dir, int H);
int lstm_tc_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed, cudaStream_t stream);
bool lstm_tc_supported(int H);
// 384 < H <= 608 (lstm_wide_tc.cu): groups of ceil(H/32) CTAs exchanging h through L2, backend 2 only
bool lstm_wide_supported(int H);
size_t lstm_wide_workspace_bytes(int n_dir, int B, int H);
size_t lstm_wide_pack_bytes(int n_dir, int H);
int lstm_wide_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed, cudaStream_t stream);
int lstm_wide_fwd(const float* pre, long long pre_dir, long long pre_row, const float* const* host_Wh, long long ldw,
                  const void* wh_packed, float* out, float* cell_seq, float* gates_seq, void* out_split, int out_kp,
                  int n_dir, int T, int B, int H, void* workspace, size_t workspace_bytes, cudaStream_t stream);

constexpr int kU = 8;      // hidden units per CTA
constexpr int kBt = 16;    // utterances per CTA

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct LstmSeqParams {
  const float* pre;        // [n_dir][T][B][4H]
  const float* Wh[2];      // recurrent rows [H][4H] (row stride ldw)
  long long ldw;
  float* out;              // [B][T][n_dir*H]
  float* cell_seq;         // nullable [n_dir][T][B][H]
  float* gates_seq;        // nullable, indexed like pre: post-activation [g|i|f|o]; may alias pre
  long long pre_dir, pre_row;   // element strides of pre: dir*pre_dir + (t*B + b)*pre_row + gate*H + unit
  int* counters;           // [n_dir][n_bt_total]
  int n_dir, T, B, H;
  int bt0;                 // first batch tile of this launch
  int dir0;                // first direction of this launch
  int n_bt_total;
};

__global__ void __launch_bounds__(256)
lstm_seq_kernel(LstmSeqParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, T = p.T, B = p.B;
  const int ldh = H + 4;
  float4* Ws = reinterpret_cast<float4*>(smem);            // [H][kU] float4 = 4 gates
  float* hs = smem + (size_t)H * kU * 4;                   // [kBt][ldh]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x, bt = p.bt0 + blockIdx.y, dir = p.dir0 + blockIdx.z;
  const int n_chunks = gridDim.x;
  const int u0 = chunk * kU, b0 = bt * kBt;

  // stage this CTA's slice of Wh: Ws[k][u] = (W[k][0H+u0+u], W[k][1H+..], W[k][2H+..], W[k][3H+..])
  const float* Wg = p.Wh[dir];
  for (int i = tid; i < H * kU; i += 256) {
    const int k = i / kU, u = i % kU, unit = u0 + u;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (unit < H) {
      const float* r = Wg + (size_t)k * p.ldw + unit;
      w = make_float4(__ldg(r), __ldg(r + H), __ldg(r + 2 * H), __ldg(r + 3 * H));
    }
    Ws[i] = w;
  }
  for (int i = tid; i < kBt * ldh; i += 256) hs[i] = 0.f;   // h_{-1} = 0 (main.py:112-121)

  const int pi = warp * 16 + (lane & 15);    // (utterance, unit) pair, two k-halves per pair
  const int khalf = lane >> 4;
  const int bl = pi / kU, u = pi % kU;
  const int b = b0 + bl, unit = u0 + u;
  const bool valid = b < B && unit < H;
  const int kmid = ((H / 2 + 3) / 4) * 4;
  const int kbeg = khalf ? kmid : 0, kend = khalf ? H : kmid;
  const int outw = p.n_dir * H;
  int* counter = p.counters + dir * p.n_bt_total + bt;

  float c = 0.f;
  float pre_next[4] = {0.f, 0.f, 0.f, 0.f};
  auto load_pre = [&](int s, float (&dst)[4]) {
    if (valid && khalf == 0 && s < T) {
      const int to = dir ? T - 1 - s : s;
      const float* q = p.pre + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
#pragma unroll
      for (int g = 0; g < 4; ++g) dst[g] = __ldcg(q + g * H);
    }
  };
  load_pre(0, pre_next);
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int to = dir ? T - 1 - s : s;
    float pre_cur[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) pre_cur[g] = pre_next[g];
    load_pre(s + 1, pre_next);
    if (s > 0) {
      // wait until every CTA of this (direction, batch tile) group has published h_{s-1}
      if (tid == 0) {
        const int want = n_chunks * s;
        while (ld_acquire(counter) < want) {}
      }
      __syncthreads();
      const int tp = dir ? to + 1 : to - 1;
      const int nq = H / 4;
      for (int i = tid; i < kBt * nq; i += 256) {
        const int r = i / nq, q = i % nq;
        if (b0 + r < B) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(
              p.out + ((size_t)(b0 + r) * T + tp) * outw + dir * H) + q);
          *reinterpret_cast<float4*>(hs + r * ldh + 4 * q) = v;
        }
      }
      __syncthreads();
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    {
      const float* hr = hs + bl * ldh;
#pragma unroll 2
      for (int k = kbeg; k < kend; k += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hr + k);
        const float4 w0 = Ws[(k + 0) * kU + u], w1 = Ws[(k + 1) * kU + u];
        const float4 w2 = Ws[(k + 2) * kU + u], w3 = Ws[(k + 3) * kU + u];
        a0 = fmaf(hv.x, w0.x, a0); a1 = fmaf(hv.x, w0.y, a1); a2 = fmaf(hv.x, w0.z, a2); a3 = fmaf(hv.x, w0.w, a3);
        a0 = fmaf(hv.y, w1.x, a0); a1 = fmaf(hv.y, w1.y, a1); a2 = fmaf(hv.y, w1.z, a2); a3 = fmaf(hv.y, w1.w, a3);
        a0 = fmaf(hv.z, w2.x, a0); a1 = fmaf(hv.z, w2.y, a1); a2 = fmaf(hv.z, w2.z, a2); a3 = fmaf(hv.z, w2.w, a3);
        a0 = fmaf(hv.w, w3.x, a0); a1 = fmaf(hv.w, w3.y, a1); a2 = fmaf(hv.w, w3.z, a2); a3 = fmaf(hv.w, w3.w, a3);
      }
    }
    a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
    a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
    a3 += __shfl_xor_sync(0xffffffffu, a3, 16);
    if (valid && khalf == 0) {
      const float g = a0 + pre_cur[0];
      const float ig = sigmoidf_(a1 + pre_cur[1]);
      const float fg = sigmoidf_(a2 + pre_cur[2]);
      const float og = sigmoidf_(a3 + pre_cur[3]);
      c = ig * g + fg * c;
      const float h = og * tanhf(c);
      __stcg(p.out + ((size_t)b * T + to) * outw + dir * H + unit, h);
      if (p.cell_seq) p.cell_seq[(((size_t)dir * T + to) * B + b) * H + unit] = c;
      if (p.gates_seq) {
        float* gs = p.gates_seq + (size_t)dir * p.pre_dir + ((size_t)to * B + b) * p.pre_row + unit;
        gs[0] = g; gs[H] = ig; gs[2 * H] = fg; gs[3 * H] = og;
      }
    }
    __syncthreads();   // all h_s stores of this CTA issued; hs free for the next step
    if (tid == 0) {
      __threadfence();
      atomicAdd(counter, 1);
    }
  }
}

static size_t lstm_smem_bytes(int H) { return ((size_t)H * kU * 4 + (size_t)kBt * (H + 4)) * sizeof(float); }

}  // namespace danet

using namespace danet;

extern "C" size_t danet_lstm_seq_workspace_bytes(int n_dir, int B, int H) {
  if (n_dir < 1 || B < 1 || H < 1) return 256;
  size_t simt = (((size_t)n_dir * ((B + kBt - 1) / kBt) * sizeof(int)) + 255) / 256 * 256;
  size_t tc = lstm_tc_workspace_bytes(n_dir, B, H);
  if (H % 4 == 0 && lstm_wide_supported(H)) tc = lstm_wide_workspace_bytes(n_dir, B, H);
  return simt > tc ? simt : tc;
}

extern "C" size_t danet_lstm_pack_wh_bytes(int n_dir, int H) {
  if (n_dir < 1 || H < 4) return 0;
  if (lstm_wide_supported(H)) return lstm_wide_pack_bytes(n_dir, H);
  if (!lstm_tc_supported(H)) return 0;
  return lstm_tc_pack_bytes(n_dir, H);
}

extern "C" int danet_lstm_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed,
                                  size_t packed_bytes, void* stream) {
  DANET_REQUIRE(host_Wh && packed, DANET_E_ARG, "lstm_pack_wh: null pointer");
  DANET_REQUIRE(n_dir == 1 || n_dir == 2, DANET_E_SHAPE, "lstm_pack_wh: n_dir %d", n_dir);
  for (int d = 0; d < n_dir; ++d) DANET_REQUIRE(host_Wh[d], DANET_E_ARG, "lstm_pack_wh: null Wh[%d]", d);
  DANET_REQUIRE(H >= 4 && H % 4 == 0 && ldw >= 4ll * H, DANET_E_SHAPE, "lstm_pack_wh: H %d ldw %lld", H, ldw);
  if (lstm_wide_supported(H)) {
    DANET_REQUIRE(packed_bytes >= lstm_wide_pack_bytes(n_dir, H), DANET_E_WORKSPACE, "lstm_pack_wh: buffer %zu < %zu",
                  packed_bytes, lstm_wide_pack_bytes(n_dir, H));
    return lstm_wide_pack_wh(host_Wh, ldw, n_dir, H, packed, as_stream(stream));
  }
  DANET_REQUIRE(lstm_tc_supported(H), DANET_E_SHAPE, "lstm_pack_wh: H %d is outside the tcgen05 backend's range", H);
  DANET_REQUIRE(packed_bytes >= lstm_tc_pack_bytes(n_dir, H), DANET_E_WORKSPACE, "lstm_pack_wh: buffer %zu < %zu",
                packed_bytes, lstm_tc_pack_bytes(n_dir, H));
  return lstm_tc_pack_wh(host_Wh, ldw, n_dir, H, packed, as_stream(stream));
}

static int lstm_seq_fwd_impl(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                             const float* const* host_Wh, long long ldw, const void* wh_packed, float* out,
                             float* cell_seq, float* gates_seq, void* out_split, int out_split_kp, int n_dir, int T,
                             int B, int H, void* workspace, size_t workspace_bytes, int backend, void* stream,
                             const int* pre_flags = nullptr, int flag_need = 0, int flags_tm = 0, int split_tm = 0,
                             int programmatic = 0);

extern "C" int danet_lstm_seq_fwd(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                                  const float* const* host_Wh, long long ldw, float* out, float* cell_seq,
                                  float* gates_seq, void* out_split, int out_split_kp, int n_dir, int T, int B, int H,
                                  void* workspace, size_t workspace_bytes, int backend, void* stream) {
  return lstm_seq_fwd_impl(pre, pre_dir_stride, pre_row_stride, host_Wh, ldw, nullptr, out, cell_seq, gates_seq,
                           out_split, out_split_kp, n_dir, T, B, H, workspace, workspace_bytes, backend, stream);
}

extern "C" int danet_lstm_seq_fwd_packed(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                                         const float* const* host_Wh, long long ldw, const void* wh_packed,
                                         float* out, float* cell_seq, float* gates_seq, void* out_split,
                                         int out_split_kp, int n_dir, int T, int B, int H, void* workspace,
                                         size_t workspace_bytes, int backend, void* stream) {
  DANET_REQUIRE(!wh_packed || backend >= 1, DANET_E_ARG, "lstm_seq: wh_packed is read by the tcgen05 backends only");
  return lstm_seq_fwd_impl(pre, pre_dir_stride, pre_row_stride, host_Wh, ldw, wh_packed, out, cell_seq, gates_seq,
                           out_split, out_split_kp, n_dir, T, B, H, workspace, workspace_bytes, backend, stream);
}

// The recurrence started while its input projections are still being written (danet_gemm_split_pipelined on another
// stream): pre_flags / flag_need are that call's tile_flags / *flag_need.  pre must be that product's C ([T][B][n_dir*4H],
// i.e. strides (4H, n_dir*4H)), B and T the same as there.  tcgen05 backends, B small enough for the 8-per-cluster kernel.
// pre_rows_time_major: the producer was called with rows_time_major (its row tiles cover rows t*B + b).
// out_split_time_major: emit out_split with rows t*B + b (for the NEXT layer's pipelined product) instead of b*T + t.
// programmatic_launch: launch as a programmatic dependent of the previous kernel in `stream` (the previous layer's
// recurrence, which triggers ~11 us before it ends): the clusters are placed on the SMs that kernel frees as it frees them
// and run their prologue while it drains; they wait for its completion before touching global memory.  Only meaningful
// when that kernel is the immediately preceding operation of the stream; otherwise an ordinary launch.
extern "C" int danet_lstm_seq_fwd_pipelined(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                                            const float* const* host_Wh, long long ldw, const void* wh_packed,
                                            float* out, void* out_split, int out_split_kp, int n_dir, int T, int B, int H,
                                            const int* pre_flags, int flag_need, int pre_rows_time_major,
                                            int out_split_time_major, int programmatic_launch, void* workspace,
                                            size_t workspace_bytes, int backend, void* stream) {
  DANET_REQUIRE(backend >= 1, DANET_E_ARG, "lstm_seq_pipelined: tcgen05 backends only (backend %d)", backend);
  DANET_REQUIRE(pre_flags && flag_need >= 1, DANET_E_ARG, "lstm_seq_pipelined: pre_flags / flag_need");
  return lstm_seq_fwd_impl(pre, pre_dir_stride, pre_row_stride, host_Wh, ldw, wh_packed, out, nullptr, nullptr, out_split,
                           out_split_kp, n_dir, T, B, H, workspace, workspace_bytes, backend, stream, pre_flags, flag_need,
                           pre_rows_time_major, out_split_time_major, programmatic_launch);
}

static int lstm_seq_fwd_impl(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                             const float* const* host_Wh, long long ldw, const void* wh_packed, float* out,
                             float* cell_seq, float* gates_seq, void* out_split, int out_split_kp, int n_dir, int T,
                             int B, int H, void* workspace, size_t workspace_bytes, int backend, void* stream,
                             const int* pre_flags, int flag_need, int flags_tm, int split_tm, int programmatic) {
  DANET_REQUIRE(pre && host_Wh && out && workspace, DANET_E_ARG, "lstm_seq: null pointer");
  DANET_REQUIRE(n_dir == 1 || n_dir == 2, DANET_E_SHAPE, "lstm_seq: n_dir %d", n_dir);
  for (int d = 0; d < n_dir; ++d) DANET_REQUIRE(host_Wh[d], DANET_E_ARG, "lstm_seq: null Wh[%d]", d);
  DANET_REQUIRE(T >= 0 && B >= 0 && H >= 4 && H % 4 == 0 && ldw >= 4ll * H, DANET_E_SHAPE,
                "lstm_seq: T %d B %d H %d (multiple of 4) ldw %lld", T, B, H, ldw);
  DANET_REQUIRE(backend >= 0 && backend <= 2, DANET_E_ARG, "lstm_seq: backend %d", backend);
  DANET_REQUIRE(aligned16(out), DANET_E_ALIGN, "lstm_seq: out must be 16-byte aligned");
  if (pre_dir_stride == 0 && pre_row_stride == 0) {       // default layout [n_dir][T][B][4H]
    pre_row_stride = 4ll * H;
    pre_dir_stride = (long long)T * B * 4 * H;
  }
  DANET_REQUIRE(pre_row_stride >= 4ll * H && pre_row_stride % 4 == 0 && pre_dir_stride % 4 == 0, DANET_E_SHAPE,
                "lstm_seq: pre strides %lld / %lld", pre_dir_stride, pre_row_stride);
  DANET_REQUIRE(backend >= 1 || !out_split, DANET_E_ARG, "lstm_seq: out_split is produced by the tcgen05 backends only");
  DANET_REQUIRE(workspace_bytes >= danet_lstm_seq_workspace_bytes(n_dir, B, H), DANET_E_WORKSPACE,
                "lstm_seq: workspace %zu < %zu", workspace_bytes,
                danet_lstm_seq_workspace_bytes(n_dir, B, H));
  if (T == 0 || B == 0) return DANET_OK;
  cudaStream_t st = as_stream(stream);
  if (backend == 2 && lstm_wide_supported(H)) {
    DANET_REQUIRE(!pre_flags && !split_tm, DANET_E_SHAPE, "lstm_seq: the pipelined hand-over needs H <= 384 (got %d)", H);
    return lstm_wide_fwd(pre, pre_dir_stride, pre_row_stride, host_Wh, ldw, wh_packed, out, cell_seq, gates_seq, out_split,
                         out_split_kp, n_dir, T, B, H, workspace, workspace_bytes, st);
  }
  if (backend >= 1)
    return lstm_tc_fwd(pre, pre_dir_stride, pre_row_stride, host_Wh, ldw, wh_packed, out, cell_seq, gates_seq, out_split, out_split_kp,
                       n_dir, T, B, H, backend == 2, workspace, workspace_bytes, st, pre_flags, flag_need, flags_tm, split_tm, programmatic);

  const size_t smem = lstm_smem_bytes(H);
  DANET_REQUIRE(smem <= 227 * 1024, DANET_E_SHAPE, "lstm_seq: H %d needs %zu B of shared memory", H, smem);
  DANET_CUDA(cudaFuncSetAttribute(lstm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  DANET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_kernel, 256, smem));
  const int n_chunks = (H + kU - 1) / kU;
  const int n_bt = (B + kBt - 1) / kBt;
  const int resident = per_sm * num_sms();
  // the directions are independent: when one batch tile of both does not fit the device (H = 600: 2 x 75 CTAs) they run
  // in separate launches
  const int dirs_per_launch = resident >= n_dir * n_chunks ? n_dir : 1;
  int bt_per_launch = resident / (dirs_per_launch * n_chunks);
  DANET_REQUIRE(bt_per_launch >= 1, DANET_E_SHAPE,
                "lstm_seq: one batch tile needs %d resident CTAs, device holds %d", n_chunks, resident);
  if (bt_per_launch > n_bt) bt_per_launch = n_bt;
  DANET_CUDA(cudaMemsetAsync(workspace, 0, (size_t)n_dir * n_bt * sizeof(int), st));
  LstmSeqParams p;
  p.pre = pre;
  p.Wh[0] = host_Wh[0];
  p.Wh[1] = n_dir > 1 ? host_Wh[1] : host_Wh[0];
  p.ldw = ldw;
  p.out = out;
  p.cell_seq = cell_seq;
  p.gates_seq = gates_seq;
  p.pre_dir = pre_dir_stride; p.pre_row = pre_row_stride;
  p.counters = reinterpret_cast<int*>(workspace);
  p.n_dir = n_dir; p.T = T; p.B = B; p.H = H;
  p.n_bt_total = n_bt;
  for (int dir0 = 0; dir0 < n_dir; dir0 += dirs_per_launch)
    for (int bt0 = 0; bt0 < n_bt; bt0 += bt_per_launch) {
      p.bt0 = bt0;
      p.dir0 = dir0;
      const int nb = (n_bt - bt0 < bt_per_launch) ? n_bt - bt0 : bt_per_launch;
      void* args[] = {&p};
      DANET_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_kernel, dim3(n_chunks, nb, dirs_per_launch),
                                             dim3(256), args, smem, st));
    }
  return DANET_OK;
}
