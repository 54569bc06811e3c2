// Convolutional front / back end of the experimental CNN-LSTM encoder `conv-bilstm-v1` (app/modules.py:263-379):
//   tf.layers.conv2d(data_format='channels_first', padding='same', activation=leaky relu)   :289-298, 302-311, 342-353, 359-369
//   tf.layers.max_pooling2d((2,2),(2,2), channels_first)                                     :299-300, 312-313
//   s_mid3 += s_mid1                                                                           :335
// Direct fp32 convolution: a block owns a 32 x 8 pixel tile of one image and 16 output channels; input channels
// arrive in chunks of 8 (tile + halo in shared memory, the matching weight slab next to it), every thread keeps
// its pixel's 16 accumulators in registers (1 shared load of x + 4 broadcast float4 loads of w per 16 FMAs).
// Kernels are read in TensorFlow's own [kh][kw][Cin][Cout] layout, so checkpoints need no repacking.
#include "common.cuh"

namespace danet {

constexpr int kCvTW = 32, kCvTH = 8;       // output tile (pixels)
constexpr int kCvCo = 16;                  // output channels per block
constexpr int kCvCi = 8;                   // input channels per staged chunk
constexpr int kCvMaxK = 5;

// WT = true: the DATA GRADIENT of the same layer.  dX[ci][u][v] = sum_{co,kh,kw} dY[co][u-kh+R][v-kw+R] W[kh][kw][ci][co]
// is again a 'same' convolution -- of dY (Cout channels in) to Cin channels out, with the kernel flipped in both spatial
// axes and its two channel axes swapped -- so the kernel is shared and only the weight staging differs: it reads the
// forward layer's [kh][kw][Cin][Cout] array in place (here `Cin` / `Cout` are this launch's in / out channel counts, i.e.
// the forward layer's Cout / Cin).
template <int KS, bool WT = false>
__global__ void __launch_bounds__(kCvTW * kCvTH)
conv2d_same_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int Cin, int Cout, int H, int W, float leak) {
  constexpr int R = KS / 2;
  constexpr int SW = kCvTW + 2 * R, SH = kCvTH + 2 * R;
  __shared__ float s_x[kCvCi][SH][SW + 1];
  __shared__ __align__(16) float s_w[KS * KS * kCvCi][kCvCo];
  const int tx = threadIdx.x % kCvTW, ty = threadIdx.x / kCvTW;
  const int co_blocks = (Cout + kCvCo - 1) / kCvCo;
  const int b = blockIdx.z / co_blocks, co0 = (blockIdx.z % co_blocks) * kCvCo;
  const int x0 = blockIdx.x * kCvTW, y0 = blockIdx.y * kCvTH;
  const float* xb = x + (size_t)b * Cin * H * W;
  float acc[kCvCo];
#pragma unroll
  for (int i = 0; i < kCvCo; ++i) acc[i] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += kCvCi) {
    const int nci = min(kCvCi, Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < nci * SH * SW; i += kCvTW * kCvTH) {
      const int c = i / (SH * SW), r = (i / SW) % SH, q = i % SW;
      const int yy = y0 + r - R, xx = x0 + q - R;
      s_x[c][r][q] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xb + ((size_t)(ci0 + c) * H + yy) * W + xx) : 0.f;
    }
    for (int i = threadIdx.x; i < KS * KS * nci * kCvCo; i += kCvTW * kCvTH) {
      const int co = i % kCvCo, c = (i / kCvCo) % nci, kk = i / (kCvCo * nci);
      if (WT)
        s_w[kk * kCvCi + c][co] = co0 + co < Cout ? __ldg(w + ((size_t)(KS * KS - 1 - kk) * Cout + co0 + co) * Cin + ci0 + c) : 0.f;
      else
        s_w[kk * kCvCi + c][co] = co0 + co < Cout ? __ldg(w + ((size_t)kk * Cin + ci0 + c) * Cout + co0 + co) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < nci; ++c)
#pragma unroll
      for (int kh = 0; kh < KS; ++kh)
#pragma unroll
        for (int kw = 0; kw < KS; ++kw) {
          const float xv = s_x[c][ty + kh][tx + kw];
          const float4* wv = reinterpret_cast<const float4*>(s_w[(kh * KS + kw) * kCvCi + c]);
#pragma unroll
          for (int q = 0; q < kCvCo / 4; ++q) {
            const float4 ww = wv[q];
            acc[4 * q + 0] = fmaf(xv, ww.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(xv, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(xv, ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(xv, ww.w, acc[4 * q + 3]);
          }
        }
  }
  const int ox = x0 + tx, oy = y0 + ty;
  if (ox < W && oy < H) {
#pragma unroll
    for (int i = 0; i < kCvCo; ++i)
      if (co0 + i < Cout) {
        float v = acc[i] + (bias ? __ldg(bias + co0 + i) : 0.f);
        if (leak >= 0.f) v = fmaxf(v * leak, v);             // app/ops.py:103-106: max(alpha * x, x)
        y[(((size_t)b * Cout + co0 + i) * H + oy) * W + ox] = v;
      }
  }
}

__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, long long n_img, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;                          // 'valid': a trailing odd row / column is dropped
  const long long total = n_img * Ho * Wo;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
    const long long img = i / ((long long)Wo * Ho);
    const float* p = x + (img * H + 2 * yo) * W + 2 * xo;
    y[i] = fmaxf(fmaxf(__ldg(p), __ldg(p + 1)), fmaxf(__ldg(p + W), __ldg(p + W + 1)));
  }
}


// Weight and bias gradients of the same layer (TF autodiff of tf.layers.conv2d, main.py:357-358):
//   dW[kh][kw][ci][co] = sum_{b,y,x} X[b][ci][y+kh-R][x+kw-R] dY[b][co][y][x],   db[co] = sum_{b,y,x} dY[b][co][y][x]
// One block per (ci, co): every thread walks the output pixels with a fixed stride and keeps the KS x KS taps in
// registers; fixed-order block reduction (deterministic).  This encoder's training step is not a benchmarked path.
template <int KS>
__global__ void __launch_bounds__(256)
conv2d_bwd_weights_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                          float* __restrict__ db, int B, int Cin, int Cout, int H, int W) {
  constexpr int R = KS / 2;
  const int ci = blockIdx.x, co = blockIdx.y;
  float acc[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  const long long HW = (long long)H * W, total = (long long)B * HW;
  for (long long i = threadIdx.x; i < total; i += 256) {
    const int b = (int)(i / HW), yy = (int)((i % HW) / W), xx = (int)(i % W);
    const float g = __ldg(dy + ((size_t)b * Cout + co) * HW + (size_t)yy * W + xx);
    bsum += g;
    const float* xp = x + ((size_t)b * Cin + ci) * HW;
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const int sy = yy + kh - R;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const int sx = xx + kw - R;
        if (sx >= 0 && sx < W) acc[kh * KS + kw] = fmaf(__ldg(xp + (size_t)sy * W + sx), g, acc[kh * KS + kw]);
      }
    }
  }
  __shared__ float red[8][KS * KS + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  {
    const float v = warp_sum(bsum);
    if (lane == 0) red[warp][KS * KS] = v;
  }
  __syncthreads();
  if (threadIdx.x <= KS * KS) {
    float v = 0.f;
    for (int wq = 0; wq < 8; ++wq) v += red[wq][threadIdx.x];
    if (threadIdx.x < KS * KS) dw[((size_t)threadIdx.x * Cin + ci) * Cout + co] = v;
    else if (ci == 0 && db) db[co] = v;
  }
}

// gradient of the 2 x 2 / stride 2 max pooling: the whole window gradient goes to the window's FIRST maximum (row-major),
// as the argmax-based MaxPoolGrad does; a trailing odd row / column belongs to no window and gets zero
__global__ void __launch_bounds__(256)
maxpool2x2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n_img,
                      int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = n_img * H * W;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xx = (int)(i % W), yy = (int)((i / W) % H);
    const long long img = i / ((long long)W * H);
    const int yo = yy >> 1, xo = xx >> 1;
    float g = 0.f;
    if (yo < Ho && xo < Wo) {
      const float* p = x + (img * H + 2 * yo) * W + 2 * xo;
      const float v[4] = {__ldg(p), __ldg(p + 1), __ldg(p + W), __ldg(p + W + 1)};
      int best = 0;
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q] > v[best]) best = q;
      if (best == ((yy & 1) << 1 | (xx & 1))) g = __ldg(dy + (img * Ho + yo) * Wo + xo);
    }
    dx[i] = g;
  }
}

__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) out[i] = a[i] + b[i];
}

}  // namespace danet

using namespace danet;

extern "C" int danet_conv2d_fwd(const float* x, const float* w_hwio, const float* bias, float* y, int B, int Cin, int Cout,
                                int H, int W, int ksize, float leak, void* stream) {
  DANET_REQUIRE(B >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, DANET_E_SHAPE, "conv2d: B %d Cin %d Cout %d H %d W %d",
                B, Cin, Cout, H, W);
  DANET_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, DANET_E_SHAPE, "conv2d: kernel size %d (1, 3 or 5)", ksize);
  if (B == 0) return DANET_OK;
  DANET_REQUIRE(x && w_hwio && y, DANET_E_ARG, "conv2d: null pointer");
  const int co_blocks = (Cout + kCvCo - 1) / kCvCo;
  DANET_REQUIRE((long long)B * co_blocks <= 65535, DANET_E_SHAPE, "conv2d: B x ceil(Cout/16) = %lld > 65535",
                (long long)B * co_blocks);
  dim3 grid((W + kCvTW - 1) / kCvTW, (H + kCvTH - 1) / kCvTH, B * co_blocks);
  DANET_REQUIRE(grid.y <= 65535, DANET_E_SHAPE, "conv2d: H %d too large", H);
  cudaStream_t st = as_stream(stream);
  if (ksize == 5) conv2d_same_kernel<5><<<grid, kCvTW * kCvTH, 0, st>>>(x, w_hwio, bias, y, Cin, Cout, H, W, leak);
  else if (ksize == 3) conv2d_same_kernel<3><<<grid, kCvTW * kCvTH, 0, st>>>(x, w_hwio, bias, y, Cin, Cout, H, W, leak);
  else conv2d_same_kernel<1><<<grid, kCvTW * kCvTH, 0, st>>>(x, w_hwio, bias, y, Cin, Cout, H, W, leak);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_maxpool2x2_fwd(const float* x, float* y, long long n_img, int H, int W, void* stream) {
  DANET_REQUIRE(n_img >= 0 && H >= 2 && W >= 2, DANET_E_SHAPE, "maxpool2x2: n_img %lld H %d W %d", n_img, H, W);
  if (n_img == 0) return DANET_OK;
  DANET_REQUIRE(x && y, DANET_E_ARG, "maxpool2x2: null pointer");
  const long long total = n_img * (H / 2) * (W / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  maxpool2x2_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, y, n_img, H, W);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_add_fwd(const float* a, const float* b, float* out, long long n, void* stream) {
  DANET_REQUIRE(n >= 0, DANET_E_SHAPE, "add: n %lld", n);
  if (n == 0) return DANET_OK;
  DANET_REQUIRE(a && b && out, DANET_E_ARG, "add: null pointer");
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  add_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a, b, out, n);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_conv2d_bwd_data(const float* dy, const float* w_hwio, float* dx, int B, int Cin, int Cout, int H, int W,
                                     int ksize, void* stream) {
  DANET_REQUIRE(B >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, DANET_E_SHAPE, "conv2d_bwd_data: B %d Cin %d Cout %d H %d W %d",
                B, Cin, Cout, H, W);
  DANET_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, DANET_E_SHAPE, "conv2d_bwd_data: kernel size %d (1, 3 or 5)", ksize);
  if (B == 0) return DANET_OK;
  DANET_REQUIRE(dy && w_hwio && dx, DANET_E_ARG, "conv2d_bwd_data: null pointer");
  const int co_blocks = (Cin + kCvCo - 1) / kCvCo;                 // this launch's output channels = the layer's Cin
  DANET_REQUIRE((long long)B * co_blocks <= 65535, DANET_E_SHAPE, "conv2d_bwd_data: B x ceil(Cin/16) = %lld > 65535",
                (long long)B * co_blocks);
  dim3 grid((W + kCvTW - 1) / kCvTW, (H + kCvTH - 1) / kCvTH, B * co_blocks);
  DANET_REQUIRE(grid.y <= 65535, DANET_E_SHAPE, "conv2d_bwd_data: H %d too large", H);
  cudaStream_t st = as_stream(stream);
  if (ksize == 5) conv2d_same_kernel<5, true><<<grid, kCvTW * kCvTH, 0, st>>>(dy, w_hwio, nullptr, dx, Cout, Cin, H, W, -1.f);
  else if (ksize == 3) conv2d_same_kernel<3, true><<<grid, kCvTW * kCvTH, 0, st>>>(dy, w_hwio, nullptr, dx, Cout, Cin, H, W, -1.f);
  else conv2d_same_kernel<1, true><<<grid, kCvTW * kCvTH, 0, st>>>(dy, w_hwio, nullptr, dx, Cout, Cin, H, W, -1.f);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_conv2d_bwd_weights(const float* x, const float* dy, float* dw_hwio, float* dbias, int B, int Cin, int Cout,
                                        int H, int W, int ksize, void* stream) {
  DANET_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1 && Cout <= 65535, DANET_E_SHAPE,
                "conv2d_bwd_weights: B %d Cin %d Cout %d H %d W %d", B, Cin, Cout, H, W);
  DANET_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5, DANET_E_SHAPE, "conv2d_bwd_weights: kernel size %d (1, 3 or 5)", ksize);
  DANET_REQUIRE(x && dy && dw_hwio, DANET_E_ARG, "conv2d_bwd_weights: null pointer");
  dim3 grid(Cin, Cout);
  cudaStream_t st = as_stream(stream);
  if (ksize == 5) conv2d_bwd_weights_kernel<5><<<grid, 256, 0, st>>>(x, dy, dw_hwio, dbias, B, Cin, Cout, H, W);
  else if (ksize == 3) conv2d_bwd_weights_kernel<3><<<grid, 256, 0, st>>>(x, dy, dw_hwio, dbias, B, Cin, Cout, H, W);
  else conv2d_bwd_weights_kernel<1><<<grid, 256, 0, st>>>(x, dy, dw_hwio, dbias, B, Cin, Cout, H, W);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}

extern "C" int danet_maxpool2x2_bwd(const float* x, const float* dy, float* dx, long long n_img, int H, int W, void* stream) {
  DANET_REQUIRE(n_img >= 0 && H >= 2 && W >= 2, DANET_E_SHAPE, "maxpool2x2_bwd: n_img %lld H %d W %d", n_img, H, W);
  if (n_img == 0) return DANET_OK;
  DANET_REQUIRE(x && dy && dx, DANET_E_ARG, "maxpool2x2_bwd: null pointer");
  const long long total = n_img * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  maxpool2x2_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, dy, dx, n_img, H, W);
  DANET_LAUNCH_CHECK();
  return DANET_OK;
}
