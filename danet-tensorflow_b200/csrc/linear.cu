// Dense layers: app/ops.py:37-90 (lyr_linear, last-axis branch :72-89).
//   C[M,N] = A[M,K] * W[K,N] (+ bias[N]), optional [B,T] -> [T,B] row remap so the
//   recurrent kernel reads its pre-activations time-major.
// backend 0: exact fp32 SIMT tiles (this file).  backend 1: tcgen05 bf16x3 (gemm_tc.cu).
#include "common.cuh"

namespace danet {

int linear_tc_fwd(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                  float* C, int M, int N, int K, int time_major_T, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
size_t linear_tc_workspace_bytes(int M, int N, int K);

constexpr int kBM = 128, kBN = 128, kBK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
             const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K, int T) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 8 x 8 outputs each
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 1, a_k = (tid & 1) * 8;   // A tile: 128 rows x 16 k
  const int b_row = tid >> 4, b_col = (tid & 15) * 8;  // W tile: 16 k x 128 cols
  for (int k0 = 0; k0 < K; k0 += kBK) {
    {
      const int gm = m0 + a_row;
      const float* ap = A + (size_t)gm * lda + k0 + a_k;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        As[a_k + i][a_row] = (gm < M && k0 + a_k + i < K) ? __ldg(ap + i) : 0.f;
      const int gk = k0 + b_row;
      const float* bp = W + (size_t)gk * ldw + n0 + b_col;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        Bs[b_row][b_col + i] = (gk < K && n0 + b_col + i < N) ? __ldg(bp + i) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[8];
      // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3 (conflict-free float4 reads)
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int nb = T > 0 ? M / T : 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (gm >= M) continue;
    const size_t orow = T > 0 ? (size_t)(gm % T) * nb + gm / T : (size_t)gm;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (gn < N) C[orow * N + gn] = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
    }
  }
}

}  // namespace danet

using namespace danet;

extern "C" size_t danet_linear_workspace_bytes(int M, int N, int K, int backend) {
  if (backend != 1 || M < 1 || N < 1 || K < 1) return 256;
  return linear_tc_workspace_bytes(M, N, K);
}

extern "C" int danet_linear_fwd(const float* A, long long lda, const float* W, long long ldw,
                                const float* bias, float* C, int M, int N, int K, int time_major_T,
                                void* workspace, size_t workspace_bytes, int backend, void* stream) {
  DANET_REQUIRE(A && W && C, DANET_E_ARG, "linear: null pointer");
  DANET_REQUIRE(M >= 0 && N >= 1 && K >= 1 && lda >= K && ldw >= N, DANET_E_SHAPE,
                "linear: M %d N %d K %d lda %lld ldw %lld", M, N, K, lda, ldw);
  DANET_REQUIRE(time_major_T >= 0 && (time_major_T == 0 || M % time_major_T == 0), DANET_E_SHAPE,
                "linear: M %d is not a multiple of T %d", M, time_major_T);
  DANET_REQUIRE(backend == 0 || backend == 1, DANET_E_ARG, "linear: backend %d", backend);
  if (M == 0) return DANET_OK;
  if (backend == 1)
    return linear_tc_fwd(A, lda, W, ldw, bias, C, M, N, K, time_major_T, workspace, workspace_bytes,
                         as_stream(stream));
  dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM);
  DANET_REQUIRE(grid.y <= 65535, DANET_E_SHAPE, "linear: M %d too large", M);
  sgemm_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, lda, W, ldw, bias, C, M, N, K, time_major_T);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
