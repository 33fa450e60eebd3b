/* picklebot_b200 -- C ABI of the B200 (sm_100a) kernels behind Picklebot's 3D mobile CNN hot path.
 *
 * The reference (hbfreed/Picklebot) has no FFI of its own: every op of the hot path is a torch.nn
 * module call (mobilenet.py, movinet.py).  Each entry point below therefore cites the reference
 * *operator* it replaces (file:line in /root/reference).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers.
 *   - Activations are NDHWC (channels-last-3d): row-major X[M][C], M = B*T*H*W, C % 8 == 0, base
 *     pointers 16-byte aligned.  `dtype` selects the activation storage type (PB_F32 / PB_BF16);
 *     parameters, statistics and parameter gradients are always fp32 (fp64 for raw sums).
 *   - The caller owns every buffer, including workspaces; the library allocates nothing on the
 *     device and keeps no pointers after a call returns.
 *   - Every call only enqueues work on `stream` (a cudaStream_t); no host synchronisation.
 *   - Return value: PB_OK or an error code; pb_last_error_string() describes the last failure on
 *     the calling thread.  Unsupported shapes are errors -- there is no CPU or library fallback.
 *   - Re-entrant: forward is typically called from the Python main thread and backward from the
 *     autograd engine's device thread.
 */
#ifndef PICKLEBOT_B200_H
#define PICKLEBOT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1

enum { PB_OK = 0, PB_ERR_BAD_ARG = 1, PB_ERR_UNSUPPORTED = 2, PB_ERR_CUDA = 3 };
enum { PB_F32 = 0, PB_BF16 = 1, PB_U8 = 2,
       PB_F32_RBF16 = 16 /* pb_cast_matrix only: fp32 storage, values rounded through bf16 */ };
enum { PB_ACT_NONE = 0, PB_ACT_RELU = 1, PB_ACT_HSWISH = 2, PB_ACT_LRELU = 3, PB_ACT_HSIGMOID = 4 };

typedef void* pb_stream_t; /* cudaStream_t */

int         pb_abi_version(void);
const char* pb_last_error_string(void);
long long   pb_launch_count(void);      /* kernels launched by this library since load (all threads) */
int         pb_device_check(void);      /* PB_OK iff the current device is sm_100 (B200) */

/* Which kernel family served a call: every entry point that can pick between a Blackwell fast path
 * (TMA-staged tiles / tcgen05) and a CUDA-core correctness path counts the choice here, so tests can assert
 * that the production path ran instead of inferring it from timings.  Counters are process-wide. */
enum { PB_PATH_DW_FWD_TMA = 0, PB_PATH_DW_FWD_GENERIC, PB_PATH_DW_DGRAD_TMA, PB_PATH_DW_DGRAD_GENERIC,
       PB_PATH_DW_WGRAD_TMA, PB_PATH_DW_WGRAD_GENERIC, PB_PATH_GEMM_TC, PB_PATH_GEMM_SIMT,
       PB_PATH_WGRAD_TC, PB_PATH_WGRAD_SIMT, PB_PATH_STEM_TC, PB_PATH_STEM_SIMT,
       PB_PATH_DW_BWD_FUSED_TMA, PB_PATH_DW_STREAM_TMA, PB_PATH_DW_STREAM_GENERIC,
       PB_PATH_STEM_TMA /* subset of STEM_TC: uint8 clip staged by TMA */, PB_PATH_COUNT };
long long   pb_path_count(int path);    /* calls served by `path` since load / the last reset; -1 if out of range */
void        pb_path_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Depthwise Conv3d, groups == C.  Replaces Bottleneck3D.depthwise_conv (mobilenet.py:67-75,86: kernel
 * (1,k,k) with SCALAR stride/padding, so time is padded and strided too) and MoviNetBottleneck.conv
 * (movinet.py:52-61,71: (kT,kH,kW), stride (1,s,s)), plus their autograd (train.py:269).
 *   x  [B][T][H][W][C], y/dy [B][To][Ho][Wo][C]; w_tc: fp32 weights repacked tap-major [kT*kH*kW][C]
 *   (pb_cast_matrix(..., transpose=1) of the (C,1,kT,kH,kW) parameter viewed as [C][taps]).  wgrad overwrites dw_tc (same layout).
 * ---------------------------------------------------------------------------------------------- */
int pb_dwconv3d_fwd(const void* x, const float* w_tc, void* y, int dtype,
                    int B, int C, int T, int H, int W, int kT, int kH, int kW,
                    int sT, int sH, int sW, int pT, int pH, int pW, int To, int Ho, int Wo,
                    pb_stream_t stream);
/* The same forward plus the global average pool of its output (SEBlock3D's AdaptiveAvgPool3d, mobilenet.py:15-16,84-88)
 * in the same pass: pool[B][C] fp32 = mean over (To,Ho,Wo) of the stored (rounded) outputs.  The TMA strip kernels
 * accumulate it while storing; shapes they decline are pooled by a second pass (pb_pool_fwd).  pool need not be zeroed. */
int pb_dwconv3d_fwd_pool(const void* x, const float* w_tc, void* y, float* pool, int dtype,
                    int B, int C, int T, int H, int W, int kT, int kH, int kW,
                    int sT, int sH, int sW, int pT, int pH, int pW, int To, int Ho, int Wo,
                    pb_stream_t stream);
int pb_dwconv3d_dgrad(const void* dy, const float* w_tc, void* dx, int dtype,
                      int B, int C, int T, int H, int W, int kT, int kH, int kW,
                      int sT, int sH, int sW, int pT, int pH, int pW, int To, int Ho, int Wo,
                      pb_stream_t stream);
int pb_dwconv3d_wgrad(const void* x, const void* dy, float* dw_tc, int dtype,
                      int B, int C, int T, int H, int W, int kT, int kH, int kW,
                      int sT, int sH, int sW, int pT, int pH, int pW, int To, int Ho, int Wo,
                      pb_stream_t stream);

/* Causal streaming variant (CausalConv3d semantics, movinet.py:23-39, applied to MoViNet's depthwise
 * convs): time is left-padded by kT-1 frames taken from `stream_buf` [B][kT-1][H][W][C] (the tail of
 * the previous chunk; zeros for the first chunk), temporal stride 1, To == T.  After the conv the last
 * kT-1 input frames are written back to `stream_buf_out` (may alias stream_buf only if kT-1 <= T). */
int pb_stream_dwconv3d_fwd(const void* x, const void* stream_buf, const float* w_tc, void* y,
                           void* stream_buf_out, int dtype,
                           int B, int C, int T, int H, int W, int kT, int kH, int kW,
                           int sH, int sW, int pH, int pW, int Ho, int Wo, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pointwise (1x1x1) convolutions and fully connected layers as GEMMs.  Replaces
 * Bottleneck3D.pointwise_conv1/2 (mobilenet.py:64,79,85,89), MoviNetBottleneck.expand/project
 * (movinet.py:47,63-64), block6 / block4 convs (mobilenet.py:179,245), `conv` (movinet.py:140) and the
 * classifier layers (mobilenet.py:185-190, movinet.py:146-154).
 *
 *   C[b][r][n] = ( sum_k A[b][r][k] * ascale[b][k] * W(n,k) + bias[n] ) * colscale[b][n] + coladd[b][n]
 *
 * A is [Bt][R][K], C is [Bt][R][N] (K % 8 == 0; N % 8 == 0 unless dtype is PB_F32).  bias, ascale,
 * colscale, coladd are optional fp32 vectors (NULL = absent).  With ascale = the squeeze-excite gate
 * this fuses `x * w` (mobilenet.py:25) into pointwise_conv2.
 *
 * _simt: fp32 W addressed as w[n*w_sn + k*w_sk] (so the same kernel serves fwd and dgrad); W is
 *        rounded to the activation dtype first, like autocast does.
 * _tc:   tcgen05/TMEM bf16 path; W is bf16 [Bw][N][K] (K contiguous) with Bw == 1 or Bw == Bt, e.g. the
 *        per-sample gate-folded weights written by pb_fold_gate_bf16.
 * ---------------------------------------------------------------------------------------------- */
int pb_pw_gemm_simt(const void* A, const float* W, long long w_sn, long long w_sk, const float* bias,
                    const float* ascale, const float* colscale, const float* coladd, void* C, int dtype,
                    int Bt, long long R, int K, int N, pb_stream_t stream);
/* stats (optional, NULL to skip; needs N <= 256 and no epilogue vectors): BatchNorm sums of the stored (rounded)
 * outputs, accumulated by the epilogue so that the following BatchNorm needs no statistics pass:
 * stats[r][0][c] = partial sum, stats[r][1][c] = partial sum of squares over PB_STAT_REPLICAS copies r,
 * c = column % stat_mod (stat_mod = real channel count of a row-folded problem, else N).  Zeroed here. */
int pb_pw_gemm_tc(const void* A, const void* W_bf16, int Bw, const float* bias,
                  const float* colscale, const float* coladd, void* C, double* stats, int stat_mod,
                  int Bt, long long R, int K, int N, pb_stream_t stream);
/* The same with an activation on the fp32 accumulators: C = act(A W^T + bias).  Inference form of
 * conv -> BatchNorm(eval) -> activation (mobilenet.py:89-91 in eval mode) once the caller has folded the BatchNorm
 * scale into W (pb_fold_scaled_bf16) and passes its shift as bias.  Needs the plain epilogue: no statistics, no
 * colscale, at most one of bias / coladd (either one pre-loads the TMEM accumulator). */
int pb_pw_gemm_tc_act(const void* A, const void* W_bf16, int Bw, const float* bias,
                  const float* colscale, const float* coladd, void* C, double* stats, int stat_mod,
                  int Bt, long long R, int K, int N, int act, float slope, pb_stream_t stream);
/* Weight gradient  dW[n][k] = sum_b ascale[b][k] * sum_r dC[b][r][n] * A[b][r][k]  (fp32, overwritten).
 * Optionally also dbias[n] = sum dC (NULL to skip). */
int pb_pw_wgrad_simt(const void* A, const void* dC, const float* ascale, float* dW, float* dbias,
                     int dtype, int Bt, long long R, int K, int N, pb_stream_t stream);
/* tcgen05 weight gradient for bf16 activations (both operands are MN-major shared-memory tiles fed by TMA).
 *   dW[n][k]    = sum_b gate[b][k] * P_b[n][k],   P_b[n][k] = sum_r dC[b][r][n] * A[b][r][k]
 *   dgate[b][k] = sum_n Wf32[n][k] * P_b[n][k]   (the squeeze-excite gate gradient, for free: no extra
 *                 pass over the expanded activations)
 * gate, Wf32 and dgate are optional (all NULL: plain wgrad; pass Bt = 1, R = total rows then).
 * workspace: pb_pw_wgrad_tc_workspace_bytes(Bt,R,K,N) bytes of caller-owned scratch. */
long long pb_pw_wgrad_tc_workspace_bytes(int Bt, long long R, int K, int N);
int pb_pw_wgrad_tc(const void* A, const void* dC, const float* gate, const float* Wf32, void* workspace,
                   float* dW, float* dgate, int Bt, long long R, int K, int N, pb_stream_t stream);

/* fp32 [rows][cols] -> dst (dtype) [cols][rows] if transpose else [rows][cols]. */
int pb_cast_matrix(const float* src, void* dst, int dst_dtype, int rows, int cols, int transpose,
                   pb_stream_t stream);
/* dst[b][n][k] = bf16( W[n][k] * gate[b][k] )  -- squeeze-excite gate folded into pointwise_conv2. */
int pb_fold_gate_bf16(const float* W, const float* gate, void* dst, int Bt, int N, int K, pb_stream_t stream);
/* transposed twin for the input gradient  dy2 = (dz W) * gate  of the same block (blocks.py se_pw2_backward):
 * dst[b][k][n] = bf16(W[n][k] * gate[b][k]), W fp32 [N][K], gate fp32 [Bt][K], dst bf16 [Bt][K][N], N % 8 == 0. */
/* inference: dst[b][n][k] = bf16(W[n][k] * gate[b][k] * rowscale[n]); gate (fp32 [Bt][K]) and rowscale (fp32 [N], the
 * eval-mode BatchNorm scale of the layer's output channels) are each optional (NULL); without a gate Bt must be 1. */
int pb_fold_scaled_bf16(const float* W, const float* gate, const float* rowscale, void* dst, int Bt, int N, int K,
                        pb_stream_t stream);
/* the same from a weight that is already transposed: dst[b][r][c] = bf16(Wt[r][c] * gate[b][r]), Wt fp32 [R][C] (e.g. the
 * cached pb_cast_matrix(..., PB_F32, transpose=1) of W2), gate fp32 [Bt][R], C % 8 == 0: all accesses coalesced. */
int pb_fold_rows_bf16(const float* Wt, const float* gate, void* dst, int Bt, int R, int C, pb_stream_t stream);
int pb_fold_gate_t_bf16(const float* W, const float* gate, void* dst, int Bt, int N, int K, pb_stream_t stream);
/* dst bf16 [F*N][F*K] = diag(W, ..., W) for W bf16 [N][K]: the weight of a row-folded GEMM.  A layer with
 * K <= 32 input channels is run as X'[rows/F][F*K] x dst^T = C'[rows/F][F*N], which is C[rows][N] in memory,
 * so that TMA moves 128-byte rows instead of 32-byte ones. */
int pb_block_diag_bf16(const void* W, void* dst, int F, int N, int K, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm3d/1d (+ activation + Dropout3d) -- mobilenet.py:80-82,90-92,142-143,180-181,247-248;
 * movinet.py:65,75-76,93,141-143,150-152.  Training mode uses batch statistics over the M rows.
 * ---------------------------------------------------------------------------------------------- */
/* Raw fp64 sums are accumulated into PB_STAT_REPLICAS interleaved copies (CTA i adds into copy
 * i % PB_STAT_REPLICAS, which spreads the same-address atomics over more L2 lines); the finalize calls
 * add the copies up.  Every `sums` argument below is a workspace of PB_STAT_REPLICAS*2*C doubles. */
#define PB_STAT_REPLICAS 16
/* sums[r][0][c] = partial sum_m x, sums[r][1][c] = partial sum_m x^2 (fp64, overwritten). */
int pb_colstats(const void* x, int dtype, long long M, int C, double* sums, pb_stream_t stream);
/* training: mean/var from sums, running stats updated (momentum, unbiased var), num_batches_tracked (int64
 * device scalar, may be NULL) incremented; else running stats.
 * Writes scale = gamma*invstd, shift = beta - mean*scale, mean, invstd (all [C]). */
int pb_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int training, float momentum, float eps,
                   float* scale, float* shift, float* mean, float* invstd, long long* num_batches_tracked,
                   int C, pb_stream_t stream);
/* out = act(z*scale + shift) * mask[b][c]   (mask NULL = no dropout; mask holds 0 or 1/(1-p)). */
int pb_bn_act_fwd(const void* z, const float* scale, const float* shift, const float* mask, void* out,
                  int dtype, int B, long long R, int C, int act, float slope, pb_stream_t stream);
/* Backward, pass 1: sums[0][c] = sum du, sums[1][c] = sum du*xhat with du = dout*mask*act'(u).
 * dout is [B][R][C] of `dtype`, or (dout_bcast != 0) fp32 [B][C] broadcast over the R rows (the
 * gradient of a global average pool, already divided by R). */
int pb_bn_act_bwd_reduce(const void* dout, int dout_bcast, const void* z, const float* scale,
                         const float* shift, const float* mean, const float* invstd, const float* mask,
                         double* sums, int dtype, int B, long long R, int C, int act, float slope,
                         pb_stream_t stream);
/* dgamma = sums[1], dbeta = sums[0]; coef[0][c] = sums[0]/M, coef[1][c] = sums[1]/M (zeros in eval). */
int pb_bn_bwd_finalize(const double* sums, long long M, int training, float* dgamma, float* dbeta,
                       float* coef, int C, pb_stream_t stream);
/* Pass 2: dz = scale * (du - coef0 - xhat*coef1). */
int pb_bn_act_bwd_apply(const void* dout, int dout_bcast, const void* z, const float* scale,
                        const float* shift, const float* mean, const float* invstd, const float* mask,
                        const float* coef, void* dz, int dtype, int B, long long R, int C, int act,
                        float slope, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Squeeze-excite and global pooling -- SEBlock3D (mobilenet.py:11-26), AdaptiveAvgPool3d in the
 * classifiers (mobilenet.py:186, 252; movinet.py:147).
 * ---------------------------------------------------------------------------------------------- */
int pb_pool_fwd(const void* x, int dtype, int B, long long R, int C, float* mean, pb_stream_t stream);
/* Streaming mode (MoViNet stream state, BASELINE config 4): global pools become cumulative means over all frames
 * seen so far.  sum[b][c] += chunk_mean[b][c] * R;  rows[0] += R;  mean_out = sum / rows.  `sum` (fp32 [B][C]) and
 * `rows` (int64 device scalar) are stream state resident in HBM between chunk calls (zero them to start a clip). */
int pb_stream_pool_update(const float* chunk_mean, long long R, float* sum, long long* rows, float* mean_out,
                          int B, int C, pb_stream_t stream);
/* Classifier-head linear layers (nn.Linear, mobilenet.py:184-190,250-256; movinet.py:146-154), fp32:
 * Y[b][n] = bias[n] + sum_k X[b][k] W[n][k]   and   dX[b][k] = scale * sum_n dY[b][n] W[n][k]. */
int pb_fc_fwd(const float* X, const float* W, const float* bias, float* Y, int B, int N, int K, pb_stream_t stream);
int pb_fc_dgrad(const float* dY, const float* W, float* dX, int B, int N, int K, float scale, pb_stream_t stream);
/* hidden = relu(W1 mean + b1) [B][Ch];  gate = hardsigmoid(W2 hidden + b2) [B][C]. */
int pb_se_fc_fwd(const float* mean, const float* W1, const float* b1, const float* W2, const float* b2,
                 float* hidden, float* gate, int B, int C, int Ch, pb_stream_t stream);
/* dgate [B][C] -> dmean [B][C] (already multiplied by inv_R) and fp32 parameter gradients (overwritten).
 * work: caller-provided scratch of B*(C+Ch) floats. */
int pb_se_fc_bwd(const float* dgate, const float* mean, const float* hidden, const float* gate,
                 const float* W1, const float* W2, float inv_R, float* dmean, float* work,
                 float* dW1, float* db1, float* dW2, float* db2, int B, int C, int Ch, pb_stream_t stream);
int pb_rowscale(const void* x, const float* gate, void* y, int dtype, int B, long long R, int C,
                pb_stream_t stream);                                      /* y = x * gate[b][c] */
int pb_rowdot(const void* g, const void* y, int dtype, int B, long long R, int C, float* out,
              pb_stream_t stream);                                        /* out[b][c] = sum_r g*y */
int pb_scale_add(void* g, const float* gate, const float* add, int dtype, int B, long long R, int C,
                 pb_stream_t stream);                                     /* g = g*gate[b][c] + add[b][c] */

/* ------------------------------------------------------------------------------------------------
 * Stem: dense Conv3d with tiny Cin.  Replaces block1.0 (mobilenet.py:141, 221: Conv3d(3,16,3,s2,p1)+bias;
 * movinet.py:92: Conv3d(3,16,(1,3,3),s(1,2,2),p(0,1,1)), no bias) and fuses extract_features_labels'
 * uint8 -> /255 conversion (train.py:106) when x_dtype == PB_U8 (in_div = 255; ignored otherwise).
 *   x: logical (B,Cin,T,H,W) addressed with element strides xs_*; w fp32 (Cout,Cin,kT,kH,kW); y NDHWC.
 * ---------------------------------------------------------------------------------------------- */
int pb_stem_conv_fwd(const void* x, int x_dtype, long long xs_b, long long xs_c, long long xs_t,
                     long long xs_h, long long xs_w, float in_div, const float* w, const float* bias,
                     void* y, int y_dtype, int B, int Cin, int T, int H, int W, int Cout,
                     int kT, int kH, int kW, int sT, int sH, int sW, int pT, int pH, int pW,
                     int To, int Ho, int Wo, pb_stream_t stream);
/* inference form: y = act(conv(x) + bias), the stem's eval-mode BatchNorm folded into w / bias by the caller; served by
 * the tensor-core kernels only (bf16 y, RGB channels-last clip), PB_ERR_UNSUPPORTED otherwise. */
int pb_stem_conv_fwd_act(const void* x, int x_dtype, long long xs_b, long long xs_c, long long xs_t,
                     long long xs_h, long long xs_w, float in_div, const float* w, const float* bias,
                     void* y, int y_dtype, int B, int Cin, int T, int H, int W, int Cout,
                     int kT, int kH, int kW, int sT, int sH, int sW, int pT, int pH, int pW,
                     int To, int Ho, int Wo, int act, float slope, pb_stream_t stream);
int pb_stem_conv_wgrad(const void* x, int x_dtype, long long xs_b, long long xs_c, long long xs_t,
                       long long xs_h, long long xs_w, float in_div, const void* dy, int y_dtype,
                       float* dw, float* dbias, int B, int Cin, int T, int H, int W, int Cout,
                       int kT, int kH, int kW, int sT, int sH, int sW, int pT, int pH, int pW,
                       int To, int Ho, int Wo, pb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step (train.py:208-212,283-289).  Multi-tensor AdamW with
 * torch.optim.AdamW's arithmetic (decoupled weight decay, bias-corrected moments, fp32 state), one launch for
 * all tensors.  Device tables: ptrs = [4][n_tensors] addresses (parameter, gradient, exp_avg, exp_avg_sq; all
 * fp32), sizes[n_tensors] element counts, and a chunk list: CTA c updates elements [chunk_start[c],
 * chunk_start[c] + pb_adamw_chunk_elems()) of tensor chunk_tensor[c].  bias_correction1 = 1 - beta1^t,
 * bias_correction2_sqrt = sqrt(1 - beta2^t); gradients are multiplied by grad_scale first (1/loss-scale).
 * ---------------------------------------------------------------------------------------------- */
int pb_adamw_chunk_elems(void);
int pb_adamw_step(const long long* ptrs, const long long* sizes, const int* chunk_tensor,
                  const long long* chunk_start, int n_tensors, int n_chunks, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float bias_correction1, float bias_correction2_sqrt,
                  float grad_scale, pb_stream_t stream);

/* Cross-entropy criterion + accuracy count (nn.CrossEntropyLoss, train.py:214,266-267; calculate_accuracy,
 * train.py:110-114), fp32 logits [B][NC], int64 labels [B]:
 *   loss[0]        = scale * mean_b (logsumexp(logits[b]) - logits[b][label_b])
 *   dlogits[b][c]  = scale / B * (softmax(logits[b])[c] - [c == label_b])        (NULL to skip)
 *   correct[0]     = #{b : argmax_c logits[b][c] == label_b}                     (int32 device scalar, NULL to skip)
 * Samples whose label lies outside [0, NC) (e.g. torch's ignore_index -100) are skipped and B above becomes the
 * number of remaining samples, as nn.CrossEntropyLoss(reduction="mean") does for ignored targets.
 * No host synchronisation: loss and count stay on the device. */
int pb_ce_loss(const float* logits, const long long* labels, float* loss, float* dlogits, int* correct,
               int B, int NC, float scale, pb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PICKLEBOT_B200_H */
