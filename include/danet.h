/*
 * danet.h -- C ABI of libdanet_sm100.so: the B200 (sm_100a) kernels behind the
 * Encoder / Estimator / Separator plugin surface of khaotik/DaNet-Tensorflow.
 *
 * Conventions:
 *   - every entry point returns int: 0 = DANET_OK, negative = error; never throws,
 *     never allocates or frees device memory; the caller owns every buffer,
 *     including workspaces whose sizes come from the *_workspace_bytes queries;
 *   - all pointers are DEVICE pointers unless the name says `host_`;
 *   - kernels are asynchronous on the passed stream (a cudaStream_t cast to void*);
 *   - float tensors are contiguous fp32, complex tensors interleaved (re,im) fp32,
 *     index tensors int32; shapes follow the reference ([B,C,T,F], [B,T,F,E], ...);
 *   - F = 129 bins (FFT_SIZE 256, FFT_STRIDE 64) is the only front-end geometry.
 *
 * Each declaration cites the reference code it replaces (paths relative to the
 * reference repository root).
 */
#ifndef DANET_H_
#define DANET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DANET_OK          0
#define DANET_E_SHAPE    -1   /* unsupported / inconsistent sizes            */
#define DANET_E_ALIGN    -2   /* pointer not aligned as required             */
#define DANET_E_ARCH     -3   /* device is not sm_100                        */
#define DANET_E_CUDA     -4   /* a CUDA runtime call failed (see last error) */
#define DANET_E_ARG      -5   /* null pointer / bad enum                     */
#define DANET_E_WORKSPACE -6  /* workspace too small                         */

#define DANET_FFT_SIZE   256
#define DANET_FFT_STRIDE 64
#define DANET_FEATURE    129

/* library identity; danet_last_error_string() is thread-local, never NULL */
int         danet_version(void);
const char* danet_last_error_string(void);
/* 0 if the current device is compute capability 10.x, DANET_E_ARCH otherwise */
int         danet_check_device(void);
/* profiling aid: a 1-thread kernel on `stream` stores %globaltimer (ns) into *slot (device memory) */
int         danet_timestamp(unsigned long long* slot, void* stream);
/* host-side helper (no GPU work): CRC32C (Castagnoli) continued from `crc` (0 to start), the checksum TensorFlow's
 * checkpoint bundles carry per tensor and per table block (tf.train.Saver at main.py:192-206; tf_bundle.py) */
unsigned int danet_crc32c(const void* data, size_t n, unsigned int crc);

/* ---- K1  STFT front end -------------------------------------------------
 * replaces scipy.signal.stft as called at app/utils.py:117-122,
 * app/datasets/TIMIT/process.py:93-97 (window default.json:7): zero-pad 128 both
 * sides + tail to a hop multiple, frames of 256 @ hop 64, sqrt-hann, rFFT, 1/sum(w).
 * wav [n_sig, n_samples] -> spec [n_sig, T, 129] complex, T = ceil(n/64)+1.
 * logmag (nullable) [n_sig, T, 129] = log1p(|spec|)   (main.py:239-240).
 * n_samples < 256 -> DANET_E_SHAPE (scipy raises ValueError there). */
int danet_stft_num_frames(int n_samples);
int danet_stft_fwd(const float* wav, int n_sig, int n_samples,
                   float* spec_c64, float* logmag, void* stream);

/* ---- mixture features ---------------------------------------------------
 * replaces main.py:233-240: mix = sum_c src; src_pwr = |src|; mix_pwr = |mix|;
 * logmag = log1p(mix_pwr).  src [B,C,T,F] complex.  Any output may be NULL. */
int danet_mix_features_fwd(const float* src_c64, int B, int C, int TF,
                           float* mix_c64, float* src_pwr, float* mix_pwr,
                           float* logmag, void* stream);

/* ---- per-utterance mean centring -----------------------------------------
 * replaces app/modules.py:209-210 and :244-245 (x - mean over axes (1,2)).
 * x [B, n_per] -> y [B, n_per] (y may alias x).  workspace: B*64 floats. */
size_t danet_center_workspace_bytes(int B);
int danet_center_fwd(const float* x, int B, long long n_per, float* y,
                     float* workspace, void* stream);
/* the per-utterance means alone (same workspace): lets the centring be folded into the next dense layer's
 * epilogue, (x - mu) W = x W - mu colsum(W)  -- see danet_gemm_split(row_mu, col_s) */
int danet_mean_fwd(const float* x, int B, long long n_per, float* mean, float* workspace, void* stream);

/* leaky ReLU of the `toy` encoder: y = max(x*leak, x) (app/ops.py:93-107); y may alias x */
int danet_leaky_relu_fwd(const float* x, float* y, long long n, float leak, void* stream);
/* dx = dy * (y > 0 ? 1 : leak), from the activation's OUTPUT y; dx may alias dy */
int danet_leaky_relu_bwd(const float* y, const float* dy, float* dx, long long n, float leak, void* stream);

/* ---- dense layers (input projections of the LSTM, output projection) ------
 * replaces app/ops.py:37-90 (lyr_linear, last-axis branch :72-89).
 * C[M,N] = A[M,K] (row stride lda) * W[K,N] (row stride ldw) (+ bias[N]).
 * If time_major_T > 0 the logical row r = b*T + t of A is written to output row
 * t*(M/T) + b (the [T,B,N] layout the recurrent kernel consumes).
 * backend: 0 = exact fp32 SIMT (no workspace), 1 = tcgen05 bf16x3: operands split
 * x = hi + lo into bf16 pairs staged in the workspace, three tensor-core products
 * hi*hi + hi*lo + lo*hi accumulated in fp32 (error ~1e-5 of the output scale). */
size_t danet_linear_workspace_bytes(int M, int N, int K, int backend);
int danet_linear_fwd(const float* A, long long lda, const float* W, long long ldw,
                     const float* bias, float* C, int M, int N, int K,
                     int time_major_T, void* workspace, size_t workspace_bytes,
                     int backend, void* stream);

/* general tensor-core product used by the training step (dX = dY*W^T, dW = X^T*dY; TF autodiff of
 * tf.matmul at app/ops.py:77):  C[M,N] (row stride ldc) (+)= A'[M,K] * B'[K,N] (+ bias[N]).
 *   transA = 0: A stored [M,K] (lda);  1: A stored [K,M] (lda)
 *   transB = 0: B stored [K,N] (ldb);  1: B stored [N,K] (ldb)
 *   permA_T > 0 (needs transA): the reduction index k = t*(K/T) + b (time-major) reads A row
 *     b*T + t + shiftA of a batch-major [B,T,*] tensor, zero when t + shiftA leaves [0,T) --
 *     pairs the layer input X[b,t] (shift 0) or the previous hidden state h[b,t-1] / h[b,t+1]
 *     (shift -1 / +1) with the time-major gate gradients;
 *   out_perm_T: row permutation of C as in danet_linear_fwd;  accumulate != 0: C += .
 * Always tcgen05 bf16x3 (see danet_linear_fwd). */
size_t danet_gemm_workspace_bytes(int M, int N, int K);
int danet_gemm(const float* A, long long lda, int transA, int permA_T, int shiftA,
               const float* B, long long ldb, int transB, const float* bias,
               float* C, long long ldc, int M, int N, int K, int out_perm_T, int accumulate,
               void* workspace, size_t workspace_bytes, void* stream);

/* operands that do not change between calls (weights) or that a producer kernel can emit directly
 * (danet_lstm_seq_fwd out_split) skip the split pass of danet_gemm / danet_linear_fwd:
 *   danet_split_operand: X -> bf16 [2*rows_total, Kp] (hi rows [0,rows_total), lo rows after them), filling
 *     rows row0 .. row0+rows; Kp = K rounded up to 64, zero padded.  stored_k_major_rows = 0: X is
 *     [rows, K] (ld); 1: X is [K, rows] (ld), e.g. a [K,N] weight matrix.  Concatenating two weight
 *     matrices (row0) yields one product for both LSTM directions.
 *   danet_gemm_split: C[M,N] (ldc) (+)= A2 * B2^T (+ bias) on split operands A2 [2M,Kp], B2 [2N,Kp];
 *     row_mu (nullable [M / rows_per_mu]) and col_s [N] subtract row_mu[row / rows_per_mu] * col_s[n]
 *     in the epilogue (mean-centring of A, app/modules.py:244-245, folded into the projection). */
/*   danet_split_operand_paired: the transposing split (X is [K, rows]) for the weight-gradient products of the
 *     tf.scan LSTM (main.py:125-131): the reduction index is TIME-major (k = t*nb + b, nb = K / perm_T) while X is a
 *     batch-major [B*T, rows] activation; index k reads source row b*perm_T + t + shift, zero outside [0, perm_T)
 *     (shift -1 / +1 pairs h_{t-1} / h_{t+1} with the gate gradients of step t).  With row0 / rows_total the layer
 *     input and the shifted hidden sequence land in ONE operand [x ; h], i.e. one product for the stacked
 *     [I+H, 4H] weight gradient. */
size_t danet_split_operand_bytes(int rows, int K);
int danet_split_operand(const float* X, long long ld, int stored_k_major_rows, int rows, int K,
                        void* out_bf16, int row0, int rows_total, void* stream);
int danet_split_operand_paired(const float* X, long long ld, int rows, int K, int perm_T, int shift,
                               void* out_bf16, int row0, int rows_total, void* stream);
/*   danet_split_operand_time_major: danet_split_operand for a batch-major activation X [B*T, K] (main.py:76-132 feeds the
 *     scan time-major: `tf.transpose(s_x, [1, 0, 2])`) whose operand rows are wanted time-major: source row b*T + t is
 *     written to row t*B + b of out [2*rows, Kp] -- the A operand of danet_gemm_split_pipelined(rows_time_major = 1). */
int danet_split_operand_time_major(const float* X, long long ld, int rows, int K, int T, void* out_bf16, void* stream);
int danet_gemm_split(const void* A2, const void* B2, const float* bias, const float* row_mu,
                     const float* col_s, int rows_per_mu, float* C, long long ldc,
                     int M, int N, int K, int out_perm_T, int accumulate, void* stream);

/* ---- K2a -> K2b hand-over: the recurrence starts while its input projections are still being computed ----------
 * The reference runs `x W_x` inside every scan step (app/ops.py:139: a = [x_t, h] W); hoisted out of the scan it is one
 * product per layer, but a product the recurrence has to wait for.  These two calls overlap them:
 *   danet_gemm_split_pipelined  = danet_gemm_split(out_perm_T = T) with its 128-row tiles issued in the order a forward
 *     AND a backward scan consume them (tiles holding the first / last frames of an utterance first) and
 *     tile_flags[m] (int32[64], ZEROED BY THE CALLER before the call, on a stream the consumer is ordered after)
 *     counting the finished (column tile, epilogue warp) pairs of row tile m; *flag_need receives the count that means
 *     "rows 128 m .. 128 m + 127 of A, i.e. those (utterance, frame) pairs of C, are complete and visible".  M <= 8192.
 *   danet_lstm_seq_fwd_pipelined = danet_lstm_seq_fwd_packed (inference outputs only) launched on ANOTHER stream without
 *     waiting for the product: a thread spins (acquire) on the flag of the tile that holds the (utterance, frame) it is
 *     about to read.  The producer never waits for the consumer, so any interleaving completes. */
/* zero `bytes` (multiple of 4) at ptr on `stream`; small buffers by a one-block kernel (cheaper than a memset node inside a
 * captured graph): clears tile_flags */
int danet_zero_async(void* ptr, size_t bytes, void* stream);
/*   rows_time_major / pre_rows_time_major: A's rows are t*B + b instead of b*T + t (danet_split_operand_time_major, or the
 *     previous layer's recurrence called with out_split_time_major): the first and the last 128-row tile then hold the
 *     first / last 16 frames of ALL 8 utterances, the product issues its tiles alternately from the two ends of time, and
 *     both scans start after 2 of the ~32 row tiles instead of after every tile that holds some utterance's first or last
 *     frame (~9).  out_split_time_major: the recurrence emits its out_split in that row order.
 *   programmatic_launch: the call becomes a programmatic dependent (CUDA "PDL") of the previous kernel in `stream` -- meant
 *     for layer l+1's recurrence queued directly behind layer l's on the same stream: its clusters take over the SMs the
 *     previous kernel frees as it frees them and run their prologue while it drains (the previous kernel triggers ~11 us
 *     before its end), then wait for its completion before they touch global memory. */
int danet_gemm_split_pipelined(const void* A2, const void* B2, const float* bias, float* C, long long ldc, int M, int N,
                               int K, int T, int rows_time_major, int* tile_flags, int* flag_need, void* stream);
int danet_lstm_seq_fwd_pipelined(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                                 const float* const* host_Wh, long long ldw, const void* wh_packed, float* out,
                                 void* out_split, int out_split_kp, int n_dir, int T, int B, int H,
                                 const int* pre_flags, int flag_need, int pre_rows_time_major,
                                 int out_split_time_major, int programmatic_launch, void* workspace,
                                 size_t workspace_bytes, int backend, void* stream);

/* ---- K2c + K3 fused: output projection with the anchor estimator's sums in its epilogue ----
 * replaces, in ONE kernel + a per-utterance finalize, the mean-centred bias-free output layer of the encoder
 * (app/modules.py:244-259: V = (x - mean_b(x)) W, reshaped [B,T,F,E]) AND AnchoredEstimator for two sources
 * (app/modules.py:501-545 with C = 2: pair softmax eq.6, weighted means eq.7, similarity eq.8, argmin eq.9).
 * The embedding is still written (the separator reads it), but the estimator never reads it back: each 128 x 160
 * tcgen05 output tile (8 whole bins of E = 20) is turned into its contribution to sum S V / sum S while it sits in
 * the epilogue warps' shared memory, on the tensor cores (mma.sync 3xTF32).
 *   A2 [2*B*T, Kp], W2 [2*F*E, Kp]   split operands (danet_split_operand / danet_lstm_seq_fwd out_split)
 *   row_mu [B] (nullable) + col_s [F*E]   the centring term, as danet_gemm_split
 *   anchors [n_anchor, E]; embed [B, T*F, E] out; attractors [B,2,E] out; attractor_sets [B,P,2,E],
 *   similarities [B,P], choice [B] (int32) nullable outputs, P = n_anchor (n_anchor - 1) / 2
 * E must be 20 and n_anchor <= 6 (the configuration BASELINE.json benchmarks); anything else returns DANET_E_SHAPE
 * and the caller uses danet_gemm_split + danet_attractor_anchor_fwd.  Deterministic (partials indexed by tile). */
size_t danet_proj_anchor_workspace_bytes(int B, int T, int F, int E);
int danet_proj_anchor_fwd(const void* A2, const void* W2, const float* row_mu, const float* col_s,
                          const float* anchors, float* embed, float* attractors, float* attractor_sets,
                          float* similarities, int* choice, int B, int T, int F, int E, int K, int n_anchor,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2b  (Bi)LSTM sequence kernel ----------------------------------------
 * replaces Model.lyr_lstm (main.py:76-132: tf.scan from zero state) over
 * ops.lyr_lstm_flat (app/ops.py:139-147: gates [cand|i|f|o], candidate WITHOUT tanh,
 * c' = i*g + f*c, h' = o*tanh(c')) and _lyr_bilstm (app/modules.py:120-137).
 *   pre   x_t*Wx + bias for every step (danet_linear_fwd output), element (dir, t, b, j) at
 *         dir*pre_dir_stride + (t*B + b)*pre_row_stride + j; both strides 0 = the default
 *         [n_dir][T][B][4H]; (4H, 8H) = [T][B][n_dir][4H], the output of ONE product with the two
 *         directions' weights side by side.  Direction 1 is indexed by ORIGINAL time
 *         (it walks t = T-1 .. 0)
 *   Wh    n_dir pointers to the recurrent rows W[I:I+H, 0:4H] (row stride ldw)
 *   out   [B][T][n_dir*H]  hidden sequence, fwd in [0,H), bwd in [H,2H) (un-reversed)
 *   cell_seq (nullable) [n_dir][T][B][H] cell states kept for the backward pass
 *   gates_seq (nullable) indexed like pre: post-activation gates [g|i|f|o] kept for the
 *         backward pass; MAY ALIAS pre (each element is read once, then overwritten)
 *   out_split (nullable, backends 1-2) [2][B*T][out_split_kp] bf16: the hidden sequence again as hi
 *         rows then lo rows, K padded with zeros -- the next layer's danet_gemm_split operand
 * backend: 0 = fp32 SIMT cooperative kernel; 1 = tcgen05 cluster kernel, "bf16x3" (h and Wh as bf16 hi/lo pairs,
 * ~1e-5 of the fp64 result); 2 = the same kernel with h_{t-1} entering the recurrent product as ONE fp16 value
 * (Wh still hi/lo): half the bytes exchanged between the cluster's CTAs per step, ~1e-4 max-norm deviation of the
 * encoder output after 4 x 501 steps (tools/precision_study.py), inside the 1e-3 parity gate.  Batches too large
 * for 8 utterances per co-resident cluster run backend 1's 16-per-cluster kernel under either value. */
size_t danet_lstm_seq_workspace_bytes(int n_dir, int B, int H);
int danet_lstm_seq_fwd(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                       const float* const* host_Wh, long long ldw, float* out, float* cell_seq,
                       float* gates_seq, void* out_split, int out_split_kp, int n_dir, int T, int B, int H,
                       void* workspace, size_t workspace_bytes, int backend, void* stream);
/* Inference keeps weights fixed between calls, so the recurrent matrix can be handed to backend 1 already in
 * the layout its clusters hold in tensor memory (bf16 hi/lo pairs, one 128-row slice per CTA): the kernel's
 * prologue is then ONE bulk copy per CTA instead of strided fp32 reads + splitting (main.py:76-132 creates the
 * variable once; this is the same variable, repacked).  danet_lstm_pack_wh_bytes returns 0 when H is outside
 * backend 1's range.  danet_lstm_seq_fwd_packed == danet_lstm_seq_fwd with `wh_packed` (nullable) read instead
 * of host_Wh when the kernel variant supports it; host_Wh must still be valid. */
size_t danet_lstm_pack_wh_bytes(int n_dir, int H);
int danet_lstm_pack_wh(const float* const* host_Wh, long long ldw, int n_dir, int H, void* packed,
                       size_t packed_bytes, void* stream);
int danet_lstm_seq_fwd_packed(const float* pre, long long pre_dir_stride, long long pre_row_stride,
                              const float* const* host_Wh, long long ldw, const void* wh_packed, float* out,
                              float* cell_seq, float* gates_seq, void* out_split, int out_split_kp, int n_dir,
                              int T, int B, int H, void* workspace, size_t workspace_bytes, int backend,
                              void* stream);
/* backward through time (TF autodiff of the tf.scan at main.py:125-131): walks the sequence in
 * reverse, dh = d_out_t + da_{t+1} Wh^T, and overwrites `gates` ([g|i|f|o] from the forward) with
 * the pre-activation gradients da in place.  dWx / dWh / dX are danet_gemm calls on da;
 * the bias gradient is danet_colsum.  backend 0 = exact fp32 cooperative kernel, 1 = tcgen05
 * cluster kernel (Wh^T slices resident in TMEM, partial products reduce-scattered over DSMEM; H <= 384), 2 = tcgen05
 * kernel for wide layers (384 < H <= 608, the `lstm-orig` encoder of app/modules.py:140-196): groups of ceil(H/32) CTAs,
 * the weights as ONE fp16 value per element in TMEM (2^-12 relative), da as a scaled fp16 hi/lo pair, partial products
 * reduce-scattered through L2 with flag-in-data words. */
size_t danet_lstm_seq_bwd_workspace_bytes(int n_dir, int B, int H);
int danet_lstm_seq_bwd(const float* d_out, float* gates, const float* cell_seq,
                       const float* const* host_Wh, long long ldw, int n_dir, int T, int B, int H,
                       void* workspace, size_t workspace_bytes, int backend, void* stream);
/* out[n] (+)= sum over rows of x[rows, n] (row stride ld): bias gradients. workspace: 64*n floats */
size_t danet_colsum_workspace_bytes(int n);
int danet_colsum(const float* x, long long ld, long long rows, int n, float* out, int accumulate,
                 void* workspace, size_t workspace_bytes, void* stream);
/* a16: clip_by_value(+-clip) then TF1 AdamOptimizer (main.py:359-363, app/ozers.py:15-18):
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps).  grad_scale multiplies
 * the gradient first (1/world_size after an all-reduce sum).  clip <= 0 disables clipping. */
int danet_clip_adam(float* param, const float* grad, float* m, float* v, long long n, float grad_scale,
                    float clip, float lr, float beta1, float beta2, float eps, int step, void* stream);
/* clip_by_value then tf.train.GradientDescentOptimizer (app/ozers.py:9-12) */
int danet_clip_sgd(float* param, const float* grad, long long n, float grad_scale, float clip, float lr,
                   void* stream);

/* ---- K3  attractor estimation ---------------------------------------------
 * truth family replaces app/modules.py:390-412 (mode 0: plain, denominator n+1),
 * :425-450 (mode 1: weight = mix_pwr > 5, + EPS), :462-487 (mode 2: weight =
 * mix_pwr, + EPS).  class = first argmax_c src_pwr[b,:,t,f].
 *   embed [B,TF,E], src_pwr [B,C,TF], mix_pwr [B,TF] -> attractors [B,C,E] */
size_t danet_attractor_workspace_bytes(int B, int n_acc_rows, int E);
int danet_attractor_truth_fwd(const float* embed, const float* src_pwr,
                              const float* mix_pwr, float* attractors, float* den /* nullable [B,C] */,
                              int B, int C, int TF, int E, int mode,
                              void* workspace, size_t workspace_bytes, void* stream);
/* anchor estimator replaces app/modules.py:501-545 + app/ops.py:273-292:
 * P = C(n_anchor, C) subsets in itertools.combinations order; eq.6 softmax over the
 * subset's anchors, eq.7 weighted means, eq.8 max over the full CxC Gram (diagonal
 * included), eq.9 argmin.  Outputs: attractors [B,C,E]; optional attractor_sets
 * [B,P,C,E], similarities [B,P], choice [B] int32, den [B,P,C] (the eq.7 denominators,
 * kept for the backward pass; likewise den [B,C] = sum of weights in the truth family). */
int danet_anchor_num_subsets(int n_anchor, int C);
int danet_attractor_anchor_fwd(const float* embed, const float* anchors,
                               float* attractors, float* attractor_sets,
                               float* similarities, int* choice, float* den,
                               int B, int C, int TF, int E, int n_anchor,
                               void* workspace, size_t workspace_bytes, void* stream);
/* k-means estimator: NEW plugin without a reference twin (README.md:216-217 lists it
 * as unimplemented).  Lloyd iterations on embed[b] from the given initial centroids
 * (in/out [B,C,E]); an empty cluster keeps its previous centroid. */
int danet_attractor_kmeans_fwd(const float* embed, float* centroids,
                               int B, int C, int TF, int E, int n_iter,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4  mask x mixture (+ iSTFT) -------------------------------------------
 * replaces DotSeparatorSoftmax/Sigmoid (app/modules.py:577-603 / :548-574) and the
 * re-phasing at main.py:281-284 (complex(cos(phi)*p, sin(phi)*p) == mask * mix).
 *   embed [B,TF,E], attractors [B,C,E], mix [B,TF] complex and/or mix_pwr [B,TF] (the
 *   plugin interface hands the separator magnitudes only; |mix| is computed when
 *   mix_pwr is NULL; sep_c64 needs mix_c64) ->
 *   sep_pwr (nullable) [B,C,TF]; sep_c64 (nullable) [B,C,TF] complex;
 *   masks (nullable) [B,TF,C].   kind: 0 = softmax over C, 1 = sigmoid. */
int danet_mask_cmul_fwd(const float* embed, const float* attractors,
                        const float* mix_c64, const float* mix_pwr, float* sep_pwr, float* sep_c64,
                        float* masks, int B, int C, int TF, int E, int kind,
                        void* stream);
/* replaces utils.istft (app/utils.py:53-75) incl. its quirks: frames 0..T-5 only,
 * overlap-add of irfft(X[n])*w, division by sum(w^2) where non-zero, length 64*T.
 * spec [n_sig,T,129] complex -> wav [n_sig, 64*T] fp32. */
int danet_istft_fwd(const float* spec_c64, int n_sig, int T, float* wav, void* stream);
/* ---- conv-bilstm-v1 encoder (app/modules.py:263-379): its convolutional front and back end ----
 * danet_conv2d_fwd: tf.layers.conv2d(data_format='channels_first', padding='same', strides 1) + bias + leaky relu
 *   max(leak * v, v) (app/ops.py:103-106; leak < 0: linear).  x [B][Cin][H][W], y [B][Cout][H][W]; the kernel is read
 *   in TensorFlow's variable layout [k][k][Cin][Cout], k in {1, 3, 5}.
 * danet_maxpool2x2_fwd: tf.layers.max_pooling2d((2,2),(2,2)) on n_img = B*C images [H][W] -> [H/2][W/2] ('valid').
 * danet_add_fwd: out = a + b (the residual at app/modules.py:335). */
int danet_conv2d_fwd(const float* x, const float* w_hwio, const float* bias, float* y, int B, int Cin, int Cout,
                     int H, int W, int ksize, float leak, void* stream);
int danet_maxpool2x2_fwd(const float* x, float* y, long long n_img, int H, int W, void* stream);
/* backward of the three layers above (TF autodiff of tf.layers.conv2d / max_pooling2d at main.py:357-358; the leaky
 * ReLU's own derivative is danet_leaky_relu_bwd on the layer OUTPUT, applied to dy before these calls):
 *   danet_conv2d_bwd_data     dx [B,Cin,H,W]  = dy [B,Cout,H,W] convolved with the flipped, channel-swapped kernel
 *   danet_conv2d_bwd_weights  dw [k,k,Cin,Cout] = sum_{b,y,x} x[.., y+kh-r, x+kw-r] dy[.., y, x];  dbias [Cout] (nullable)
 *   danet_maxpool2x2_bwd      dx [n_img,H,W]: each window's gradient goes to its first maximum; x is the pooling INPUT */
int danet_conv2d_bwd_data(const float* dy, const float* w_hwio, float* dx, int B, int Cin, int Cout, int H, int W,
                          int ksize, void* stream);
int danet_conv2d_bwd_weights(const float* x, const float* dy, float* dw_hwio, float* dbias, int B, int Cin, int Cout,
                             int H, int W, int ksize, void* stream);
int danet_maxpool2x2_bwd(const float* x, const float* dy, float* dx, long long n_img, int H, int W, void* stream);

int danet_add_fwd(const float* a, const float* b, float* out, long long n, void* stream);

/* K4 fused (north star item 4): DotSeparatorSoftmax / DotSeparatorSigmoid (app/modules.py:548-603), the re-phasing
 * of main.py:281-284 and utils.istft (app/utils.py:53-75) for every source in ONE kernel:
 *   wav[b,c,:] = istft( mask_c(V[b], A[b]) * mix[b] ),  mask = softmax over c (kind 0) or sigmoid (kind 1).
 * embed [B][T*129][E], attractors [B][C][E], mix_c64 [B][T][129] complex64 -> wav [B][C][64*T] float32.
 * Same arithmetic as danet_mask_cmul_fwd followed by danet_istft_fwd; the separated spectra stay in shared memory.
 * C <= 4, E a multiple of 4 and <= 64. */
int danet_mask_cmul_istft_fwd(const float* embed, const float* attractors, const float* mix_c64, float* wav,
                              int B, int C, int T, int E, int kind, void* stream);
/* ---- K6  backward of the separation head ---------------------------------------------
 * What TF autodiff derives (main.py:357-358) for loss = pit_mse(src, mask*mix) with
 * mask = softmax/sigmoid(V.A) and A = estimator(V).  perm_idx comes from danet_pit_mse_fwd;
 * no gradient flows through argmax / argmin / the permutation choice.
 *   pass 1: d_attractors [B,C,E] (through the mask);
 *   pass 2: d_embed [B,TF,E] = mask path + estimator path, d_anchors [n_anchor,E].
 * est_mode 0..2 = truth / truth-threshold / truth-weighted (needs src_pwr, mix_pwr, den [B,C]),
 * 3 = anchor (needs anchors, choice [B], den [B,P,C], d_anchors). */
size_t danet_head_bwd_workspace_bytes(int B, int C, int E);
int danet_head_bwd_attractors(const float* embed, const float* attractors, const float* mix_c64,
                              const float* src_c64, const int* perm_idx, float* d_attractors,
                              int B, int C, int TF, int E, int kind,
                              void* workspace, size_t workspace_bytes, void* stream);
int danet_head_bwd_embed(const float* embed, const float* attractors, const float* mix_c64,
                         const float* src_c64, const int* perm_idx, const float* d_attractors,
                         int est_mode, const float* src_pwr, const float* mix_pwr,
                         const float* anchors, const int* choice, const float* den,
                         float* d_embed, float* d_anchors, int B, int C, int TF, int E, int kind,
                         int n_anchor, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K5  PIT-MSE + SNR ---------------------------------------------------------
 * replaces ops.pit_mse_loss (app/ops.py:374-431), the permutation gather at
 * main.py:293-306 and ops.batch_snr (app/ops.py:191-222).
 *   x, y [B,C,TF] complex (is_complex=1) or real;
 *   cross [B,C,C] mean |x_i - y_j|^2; perm_losses [B,C!] (itertools.permutations
 *   order); perm_idx [B] int32 first argmin; loss[1] = mean_b min; snr [B] (dB of
 *   mean|x|^2 over mean|x - y_aligned|^2, EPS 1e-7).  Outputs may be NULL except
 *   cross.  workspace: danet_pit_workspace_bytes. */
size_t danet_pit_workspace_bytes(int B, int C);
int danet_pit_mse_fwd(const float* x, const float* y, int B, int C, int TF,
                      int is_complex, float* cross, float* perm_losses, int* perm_idx,
                      float* loss, float* snr, void* workspace, size_t workspace_bytes,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DANET_H_ */
