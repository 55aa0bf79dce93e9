/*
 * lasr.h -- C ABI of the B200-native lightning-asr hot path (liblasr_b200.so).
 *
 * The reference (kouyt5/lightning-asr) is pure Python/PyTorch and has NO plugin / operator / FFI
 * interface of its own: its hot path is the chain of torch / torchaudio library
 * calls made by
 *     data_module.py:150-174   AudioParser.parse_audio            (log-mel frontend)
 *     models/QuartNet.py:8-52  SeprationConv.forward              (dw conv, 1x1 conv, mask, BN, ReLU)
 *     models/QuartNet.py:55-78 QuartNetBlock.forward              (+ residual 1x1 conv + BN, add, ReLU)
 *     models/QuartNet.py:264-291 MyModel2.forward                 (decoder 1x1 conv, log_softmax)
 *     models/QuartNetContextSE.py:8-23 SELayer.forward            (squeeze-excitation)
 *     train.py:76-78,196       torch.nn.CTCLoss(blank=V, reduction='none')
 *     utils/asr_metrics.py:138-171 WER.ctc_decoder_predictions_tensor (greedy CTC decode)
 * Each entry point below replaces one of those library call sites (cited per function) and is what a
 * ctypes binding inside the reference's modules would call (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every function returns 0 (LASR_OK) or a negative LASR_ERR_* code; never throws, never
 *     synchronises, never allocates device memory; all launches go to `stream`;
 *   - all pointers are DEVICE pointers owned by the caller unless stated otherwise;
 *   - activations are channels-last: a [N, T, C] tensor is a row-major matrix of N*T rows ("frames")
 *     and C contiguous channels; `ld*` arguments are row pitches in ELEMENTS;
 *   - `dtype` selects the activation element type: LASR_F32 or LASR_BF16.  Statistics, weights
 *     gradients, losses and all accumulators are fp32 regardless;
 *   - `lengths[n]` = int(float32(T) * percents[n]) is computed by the caller exactly as
 *     models/QuartNet.py:311 does, and frames t >= lengths[n] of utterance n are "masked".
 */
#ifndef LASR_H_
#define LASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* lasr_stream_t; /* == cudaStream_t */

enum { LASR_F32 = 0, LASR_BF16 = 1 };

enum {
  LASR_OK = 0,
  LASR_ERR_BAD_SHAPE = -1,
  LASR_ERR_BAD_DTYPE = -2,
  LASR_ERR_WORKSPACE = -3,
  LASR_ERR_CUDA = -4,
  LASR_ERR_ALIGNMENT = -5,
  LASR_ERR_UNSUPPORTED = -6,
  LASR_ERR_DRIVER = -7
};

/* activation flags for lasr_bn_apply_act_* */
enum { LASR_ACT_NONE = 0, LASR_ACT_RELU = 1 };

const char* lasr_strerror(int code);
int lasr_abi_version(void);
/* compute capability check: returns LASR_OK only on an sm_100 device */
int lasr_check_device(void);
/* Caller's promise about PARAMETER tensors (conv weights / taps and their bf16 shadows) passed to the kernels: with
 * on != 0 they were last written before a full stream-order barrier (e.g. by the step prologue: memset + cast +
 * layout kernels), so kernels launched with programmatic dependent launch may prefetch them while the previous
 * kernel is still draining.  Default 0: parameters are read only after the previous kernel has completed, like every
 * other operand.  Returns the previous setting.  (The reference has no counterpart: torch's stream order covers it.) */
int lasr_set_early_param_loads(int on);
/* SMs the persistent kernels (GEMMs, depthwise, BatchNorm passes) size their grids for: default all 148.  A caller that
 * overlaps a collective with the compute stream (the gradient all-reduce, conf/conf.yaml:30) passes 148 minus the
 * collective's CTAs: a one-CTA-per-SM grid that finds SMs taken runs its last CTAs as a second wave. */
int lasr_set_sm_budget(int sms);

/* ------------------------------------------------------------------------------------------------
 * Layout conversion at the module boundary.
 * replaces: `input.squeeze(dim=1).contiguous()` models/QuartNet.py:154 (input [N,1,F,T] fp32, NCT)
 * x [N, C, T] fp32  ->  y [N, T, C] dtype
 * ---------------------------------------------------------------------------------------------- */
int lasr_nct_to_ntc(const float* x, void* y, int N, int C, int T, int dtype, lasr_stream_t stream);
/* y [N, T, C] dtype -> x [N, C, T] fp32 (used for gradients / debugging at the boundary) */
int lasr_ntc_to_nct(const void* y, float* x, int N, int C, int T, int dtype, lasr_stream_t stream);
/* fp32 master weights -> compute-dtype shadow (rows = 1: a whole flat parameter buffer in one launch), optionally
 * transposed: w [R, Ccols] -> out [R,Ccols] or [Ccols,R] */
int lasr_cast_weight(const float* w, void* out, int rows, int cols, int transpose, int dtype, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Depthwise Conv1d  (replaces nn.Conv1d(C, C, k, stride, padding=k//2, groups=C, bias=False),
 * models/QuartNet.py:14-21,30).  x [N, T_in, C], w fp32 in the reference's own layout [C, 1, K], y [N, T_out, C],
 * T_out = (T_in + 2*(K/2) - K)/stride + 1.
 *   flip = 0: y[n,t,c] = sum_j w[c,j] * x[n, t*stride + j - K/2, c]                (forward)
 *   flip = 1: same with w reversed along j (stride must be 1): the data gradient of the forward.
 *   addend (nullable, [N, T_out, C]): added to the result (used to fuse the residual-branch dgrad).
 * ---------------------------------------------------------------------------------------------- */
int lasr_dwconv1d_fwd(const void* x, const float* w, void* y, const void* addend, int N, int T_in, int T_out, int C,
                      int K, int stride, int flip, int dtype, lasr_stream_t stream);
/* weight gradient: dw[c,j] += sum_{n,t} dy[n,t,c] * x[n, t*stride + j - K/2, c]; dw fp32 [C, 1, K], ACCUMULATED
 * (caller zeroes).  Deterministic only up to fp32 atomic ordering. */
int lasr_dwconv1d_wgrad(const void* x, const void* dy, float* dw, int N, int T_in, int T_out, int C, int K, int stride,
                        int dtype, lasr_stream_t stream);
/* stride-1 backward of one layer in ONE launch: dx [N, T, C] = correlation of dy with the flipped taps (+ addend, the
 * residual-branch gradient, nullable) and dw [C, 1, K] += sum over frames of dy * shifted x  (autograd of
 * models/QuartNet.py:30) */
int lasr_dwconv1d_bwd(const void* x, const void* dy, const float* w, const void* addend, void* dx, float* dw, int N,
                      int T, int C, int K, int dtype, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pointwise (1x1) Conv1d as a GEMM over frames (replaces nn.Conv1d(Cin, Cout, 1), models/QuartNet.py:22-23,31,
 * :62-63, :145-146, :275).  bf16: tcgen05/TMEM tensor-core kernel; fp32: FFMA kernel (exact fp32 parity mode).
 * w is the weight in the reference's layout [Cout, Cin, 1], in the activation dtype (see lasr_cast_weight).
 *
 * fwd:   y[M, Cout] = x[M, Cin] * w[Cout, Cin]^T (+ bias[Cout])
 *        epilogue options (all nullable):
 *          lengths/T : zero rows whose frame index t = row % T is >= lengths[row / T]   (MaskCNN, :309-321)
 *          stats     : double [2, Cout]: column sum / sum of squares of the (masked) output ACCUMULATED with
 *                      RED.f64 (caller zeroes) -- the BatchNorm batch statistics, consumed by lasr_bn_*
 * dgrad: dx[M, Cin] = dy[M, Cout] * w[Cout, Cin]          (w read as an MN-major operand: no transposed copy)
 * wgrad: dw[Cout, Cin] (fp32) += dy[M, Cout]^T * x[M, Cin]   (split over M, fp32 RED; caller zeroes)
 * ---------------------------------------------------------------------------------------------- */
int lasr_pwconv_fwd(const void* x, const void* w, void* y, const float* bias, const int32_t* lengths, int T,
                    double* stats, int M, int Cin, int Cout, int ldx, int ldw, int ldy, int dtype,
                    lasr_stream_t stream);
/* Inference (eval-mode) block epilogue, SURVEY.md 8f-4: with BatchNorm's running statistics folded into the weight
 * rows and a bias (w' = diag(gamma / sqrt(var + eps)) w, bias = beta - mean * scale), the chain
 *   conv1x1 -> MaskCNN -> BatchNorm -> (+ residual branch) -> ReLU      models/QuartNet.py:31-37,71-78,145-148
 * is ONE GEMM:  y = act( mask(x w'^T) + bias [+ residual] ),  masked rows keep bias + residual (BN of a zeroed row).
 * residual [M, ld_res] dtype nullable, relu 0/1.  bf16 only (LASR_ERR_UNSUPPORTED otherwise); Cout % 32 == 0. */
int lasr_pwconv_fwd_fused(const void* x, const void* w, void* y, const float* bias, const void* residual,
                          const int32_t* lengths, int T, int relu, int M, int Cin, int Cout, int ldx, int ldw, int ldy,
                          int ld_res, int dtype, lasr_stream_t stream);
int lasr_pwconv_dgrad(const void* dy, const void* w, void* dx, int M, int Cin, int Cout, int lddy, int ldw, int lddx,
                      int dtype, lasr_stream_t stream);
int lasr_pwconv_wgrad(const void* dy, const void* x, float* dw, int M, int Cin, int Cout, int lddy, int ldx, int lddw,
                      int dtype, lasr_stream_t stream);
/* Grouped launches for a QuartNetBlock's two 1x1 convs of identical shape (the pointwise conv of its last
 * SeprationConv and the residual conv, models/QuartNet.py:31 and :62-63,75): dense operands (row pitch = channel
 * count), no bias; problem 1 / 2 carry their own MaskCNN lengths and BatchNorm statistics buffers (nullable).
 *   fwd2:   y1 = mask1(x1 w1^T), y2 = mask2(x2 w2^T)        dgrad2: dx1 = dy1 w1, dx2 = dy2 w2 */
int lasr_pwconv_fwd2(const void* x1, const void* w1, void* y1, const int32_t* lengths1, double* stats1, const void* x2,
                     const void* w2, void* y2, const int32_t* lengths2, double* stats2, int T, int M, int Cin, int Cout,
                     int dtype, lasr_stream_t stream);
int lasr_pwconv_dgrad2(const void* dy1, const void* w1, void* dx1, const void* dy2, const void* w2, void* dx2, int M,
                       int Cin, int Cout, int dtype, lasr_stream_t stream);
/* two weight gradients of identical shape in one launch (a block's pointwise conv, models/QuartNet.py:31, and its
 * residual conv, :62-63,75): dw1 += dy1^T x1, dw2 += dy2^T x2 */
int lasr_pwconv_wgrad2(const void* dy1, const void* x1, float* dw1, const void* dy2, const void* x2, float* dw2, int M,
                       int Cin, int Cout, int lddy, int ldx, int lddw, int dtype, lasr_stream_t stream);
/* out[c] += sum_m x[m, c], c < C (fp32, caller zeroes): the decoder bias gradient (models/QuartNet.py:275) */
int lasr_colsum(const void* x, float* out, int M, int C, int ld, int dtype, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm1d(eps=1e-3) (replaces nn.BatchNorm1d, models/QuartNet.py:24,35,64,147), fused with ReLU (:36-37, :77),
 * the residual add (:76) and the squeeze-excitation scale (models/QuartNetContextSE.py:23,55).
 * There is no separate "finalize" launch: the apply passes fold the raw statistics into per-channel coefficients
 * in their prologue.  The two descriptor structs are HOST structs of DEVICE pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct lasr_bn {
  const double* sums;           /* [2, C] batch sum / sum of squares of the BN input (training); NULL = eval mode */
  const float* gamma;           /* [C] */
  const float* beta;            /* [C] */
  float* running_mean;          /* [C] training: updated (momentum, nullable); eval: read */
  float* running_var;           /* [C] training: updated with the UNBIASED variance; eval: read */
  int64_t* num_batches_tracked; /* += 1 in training (nullable) */
  float* save_mean;             /* [C] out (training, nullable): batch mean, kept for the backward */
  float* save_invstd;           /* [C] out (training, nullable): 1/sqrt(biased var + eps) */
} lasr_bn_t;

typedef struct lasr_bn_bwd {
  const float* gamma;  /* [C] */
  const float* mean;   /* [C] save_mean of the forward */
  const float* invstd; /* [C] save_invstd of the forward */
  float* dgamma;       /* [C] += (nullable) */
  float* dbeta;        /* [C] += (nullable) */
} lasr_bn_bwd_t;

/* nn.Dropout(p) (models/QuartNet.py:27,38,149) fused into the BatchNorm passes.  HOST struct.  The keep mask is one byte
 * per element of the [M, C] activation (1 = keep); kept elements are scaled by 1/(1-p).  mode 0: no dropout (or pass
 * NULL); mode 1: the forward READS `mask` (a mask supplied by the caller: the parity hook, torch's Philox stream cannot
 * be reproduced); mode 2: the forward GENERATES the mask with Philox4x32-10 keyed by (seed, element index / 8) and
 * WRITES it to `mask`.  The backward passes always read `mask`.  Dropout acts on the normalised branch
 * BN1(y) [* gate] before the residual add and the final ReLU, exactly where SeprationConv.forward applies it. */
typedef struct lasr_dropout {
  uint8_t* mask; /* [M, C] bytes */
  int mode;
  float p;
  uint64_t seed;
  const uint64_t* seed_dev; /* nullable DEVICE pointer: a per-step counter added to `seed` inside the kernel, so that a
                               replayed CUDA graph (whose kernel arguments are frozen) draws a fresh mask every step */
} lasr_dropout_t;

/* scale = gamma*invstd, shift = beta - mean*scale ([C] each) from the descriptor; side_effects != 0 also performs the
 * training side effects (save_mean/save_invstd, running statistics, num_batches_tracked).  count = N*T rows. */
int lasr_bn_coeffs(const lasr_bn_t* bn, int C, int count, float eps, float momentum, float* scale, float* shift,
                   int side_effects, lasr_stream_t stream);
/* sums[n, c] = sum_t y[n, t, c] over ALL T frames (SE squeeze numerator, models/QuartNetContextSE.py:11,21) */
int lasr_sum_over_time(const void* y, float* sums, int N, int T, int C, int dtype, lasr_stream_t stream);

/* out = act( BN1(y) [* gate[n,c]] [* dropout] [+ BN2(r)] );  y, r, out [M, C]; gate [M/T, C] nullable; r / bn2 nullable
 * together; drop nullable.  One pass; side_effects != 0 performs the training side effects of bn1 and bn2 exactly once. */
/* relu_bits (nullable, every pass below): the sign of the forward output as one byte per (frame, 8 channels), bit i =
 * out[n, t, 8 cv + i] > 0, laid out [N][ceil(T / 8)][C / 8][8] (needs M == N * T).  The forward passes write it, the
 * backward passes then read it INSTEAD of the whole `out` tensor (which may be NULL there). */
int lasr_bn_apply_act_fwd(const void* y, const lasr_bn_t* bn1, const void* r, const lasr_bn_t* bn2, const float* gate,
                          void* out, int M, int C, int T, int count, float eps, float momentum, int act,
                          int side_effects, const lasr_dropout_t* drop, int dtype, uint8_t* relu_bits,
                          lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Channel-major series: the operand format of the TMA-fed depthwise kernels (bf16 only).
 *   xT [C][N][S]: frame t of utterance n, channel c at position off + t of row (c, n); zeros before and after;
 *   off = lasr_cm_offset(K) = K/2 rounded up to 8 (the conv's left padding), S = lasr_cm_pitch(T, K) (a multiple of 128);
 *   16-byte groups of 8 positions are stored at group index g ^ ((g >> 3) & 1) (the tensor cores' 32-byte swizzle
 *   applied in global memory).  K is the kernel size of the depthwise conv that will CONSUME the tensor.
 * lasr_bn_apply_act_fwd_cm = lasr_bn_apply_act_fwd (bf16, no dropout, C % 64 == 0) that also writes the series
 * companion outT of its output: the block that follows reads its depthwise input from it (models/QuartNet.py:30). */
int lasr_cm_offset(int K);
int lasr_cm_pitch(int T, int K);
/* Toeplitz factors (nullable extras of the three calls below): the depthwise kernels multiply the series with a banded
 * Toeplitz factor per channel built from the taps; (dw_w [C, 1, dw_K] fp32, toep, toep_flip) make the pass that writes
 * the series also build the factors of the conv that will read it -- toep for its forward, toep_flip (reversed taps) for
 * its data gradient, bf16 [C][lasr_cm_ks(K) * 16] each -- and lasr_dwconv1d_fwd_cm / _bwd_cm then fetch them with one
 * bulk copy instead of rebuilding them in every launch's prologue.  Valid as long as the taps are unchanged. */
int lasr_cm_ks(int K);
int lasr_bn_apply_act_fwd_cm(const void* y, const lasr_bn_t* bn1, const void* r, const lasr_bn_t* bn2, const float* gate,
                             void* out, void* outT, int N, int T, int C, int S, int off, float eps, float momentum,
                             int act, int side_effects, uint8_t* relu_bits, const float* dw_w, void* toep,
                             void* toep_flip, int dw_K, lasr_stream_t stream);
/* y [N, T, C] channels-last = depthwise conv of the series xT (flip = 1: reversed taps = the data gradient) + an
 * optional addend, given EITHER channels-last (addend [N, T, C]) OR as a series laid out like xT (addendT); same
 * arithmetic as lasr_dwconv1d_fwd(stride 1, bf16) */
int lasr_dwconv1d_fwd_cm(const void* xT, const float* w, void* y, const void* addend, const void* addendT,
                         const void* toep, int N, int T, int C, int K, int S, int flip, lasr_stream_t stream);
/* dw [C, 1, K] fp32 += sum_{n,t} dy[n,t,c] x[n, t + j - K/2, c] from the series of x and of dy (same K, same layout) */
int lasr_dwconv1d_wgrad_cm(const void* xT, const void* dyT, float* dw, int N, int T, int C, int K, int S,
                           lasr_stream_t stream);
/* stride-1 backward of one layer in ONE launch from series operands: dx [N, T, C] channels-last = correlation of dy with
 * the flipped taps (+ addend / addendT) and dw += (autograd of models/QuartNet.py:30) */
int lasr_dwconv1d_bwd_cm(const void* xT, const void* dyT, const float* w, const void* addend, const void* addendT,
                         const void* toep_flip, void* dx, float* dw, int N, int T, int C, int K, int S,
                         lasr_stream_t stream);
/* data gradient(s) of 1x1 convs written as channel-major series (bf16): dxT[c][n][off + t] = sum_k dy[n, t, k] w[k, c],
 * dy [N*T, Cout] channels-last, w [Cout, Cin]; the pads of dxT are written as zeros.  (dy2, w2, dxT2): optional second
 * problem of the same shape in the same launch (the block's residual conv), nullable together.  Returns
 * LASR_ERR_UNSUPPORTED for shapes the weight-stationary tensor-core kernels do not take (Cin % 128 != 0). */
int lasr_pwconv_dgrad_cm(const void* dy1, const void* w1, void* dxT1, const void* dy2, const void* w2, void* dxT2, int N,
                         int T, int Cin, int Cout, int S, int off, lasr_stream_t stream);

/* backward pass 1.  With g = dout * (act == RELU ? out > 0 : 1) and g1 = g * dropout factor (= g without dropout):
 *   totals[0][c] += sum g, totals[1][c] += sum g1*y, totals[2][c] += sum g*r     double [3, C], caller zeroes
 *   totals[3][c] += sum g1                                                        (only with dropout: double [4, C])
 *   per_n[n][0][c] += sum_t g1, per_n[n][1][c] += sum_t g1*y                      float [N, 3, C], nullable (SE) */
int lasr_bn_bwd_chunks(int N, int T);
int lasr_bn_act_bwd_reduce(const void* dout, const void* out, const void* y, const void* r, double* totals,
                           float* per_n, int N, int T, int C, int act, const lasr_dropout_t* drop, int dtype,
                           const uint8_t* relu_bits, lasr_stream_t stream);
/* standalone: coef [3, C] with d(BN input) = coef[0]*g + coef[1]*x + coef[2], from totals slots (0, slot_gx);
 * dgamma / dbeta += (nullable) */
int lasr_bn_bwd_coef(const double* totals, int C, int count, int slot_gx, const float* gamma, const float* mean,
                     const float* invstd, float* dgamma, float* dbeta, float* coef, lasr_stream_t stream);
/* pass 2: dy = mask_t<len( c1[0]*(g1*gate[n,c] + extra[n,c]) + c1[1]*y + c1[2] ),  dr = c2[0]*g + c2[1]*r + c2[2]
 * c1 / c2 are folded from `totals` and bn1 / bn2 in the prologue (dgamma / dbeta accumulated once), unless coef1 is
 * given (the gated SE branch).  gate/extra nullable together; r/dr/bn2 nullable together; lengths nullable = no
 * MaskCNN.  The mask zeroes the gradient that reaches the pointwise conv at padded frames exactly like
 * masked_fill's backward (models/QuartNet.py:320). */
int lasr_bn_act_bwd_apply(const void* dout, const void* out, const void* y, const void* r, const float* gate,
                          const float* extra, const double* totals, const float* coef1, const lasr_bn_bwd_t* bn1,
                          const lasr_bn_bwd_t* bn2, int count, const int32_t* lengths, int T, void* dy, void* dr,
                          int M, int C, int act, const lasr_dropout_t* drop, int dtype, const uint8_t* relu_bits,
                          lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Squeeze-excitation (models/QuartNetContextSE.py:8-23): gate[n,:] = sigmoid(W2 * relu(W1 * s[n,:])),
 * s[n,c] = mean_t BN(y)[n,t,c] = scale[c]*sums[n,c]/T + shift[c]  (sums from lasr_sum_over_time).
 * W1 [C/r, C], W2 [C, C/r] fp32 (nn.Linear layout, no bias).  s [N,C], hidden [N,C/r], gate [N,C] are written.
 * ---------------------------------------------------------------------------------------------- */
int lasr_se_excite_fwd(const float* sums, const float* scale, const float* shift, int T, const float* w1,
                       const float* w2, float* s, float* hidden, float* gate, int N, int C, int Cr,
                       lasr_stream_t stream);
/* backward of the excitation from lasr_bn_act_bwd_reduce's per_n sums (partials [N*chunks, 3, C] with chunks = 1):
 * dgate[n,c] = sum_t g*BN(y) -> through sigmoid, W2, ReLU, W1 -> extra[n,c] = d s[n,c] / T (the term every frame
 * of (n,c) receives); dW1 += , dW2 += (one writer per element).  ws: N*(C + C/r) floats of scratch (the per-utterance
 * gradients at the two linear layers, read by the weight-gradient kernel of the same call). */
int lasr_se_excite_bwd(const float* partials, int chunks, const float* scale, const float* shift, int T,
                       const float* w1, const float* w2, const float* s, const float* hidden, const float* gate,
                       float* extra, float* dw1, float* dw2, float* ws, int N, int C, int Cr, lasr_stream_t stream);
/* lasr_bn_bwd_finalize for the gated branch (its upstream gradient is g*gate + extra) */
int lasr_se_bn_bwd_finalize(const float* partials, int N, int chunks, int C, int T, const float* gate,
                            const float* extra, const float* sums_y, const float* gamma, const float* mean,
                            const float* invstd, float* dgamma, float* dbeta, float* coef, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Log-mel frontend (replaces AudioParser.parse_audio from the waveform tensor onward, data_module.py:155-172, with the
 * transforms built at :66-71: MelSpectrogram(sr=16000, n_fft=512, pad=32, win_length=320, hop_length=160, n_mels=64)
 * and AmplitudeToDB(stype="power")).  Batched: wave [N, S_max] fp32 zero padded, num_samples [N] int32;
 * T_n = 1 + (num_samples[n] + 64) / 160 frames per utterance, T_max = frames of the longest.
 *   prepare   : optional dither (y += 1e-5 * dither, :155), pre-emphasis (:157), zero pad 32, reflect pad 256 ->
 *               the sample stream the 320 non-zero window taps of each frame read, split into 3 bf16 terms
 *               parts [3, N, lasr_logmel_padded_len(T_max)] bf16
 *   fwd       : error-compensated (products = 6, or 3) bf16 tcgen05 windowed-DFT GEMM -> power -> mel -> dB.
 *               basis [3, 512, 320] bf16: the three bf16 terms of win[i]*cos / win[i]*sin(2 pi f i / 512); row 2f = re,
 *               2f+1 = im of bin f, except row 1 = the (real) Nyquist bin 256.  mel_idx / mel_w [257, 2]: the (at most
 *               two) triangular filters each bin feeds.  db [N, T_max, 64] fp32; stats [N, 2] double (sum, sum of
 *               squares over the utterance's own 64 * T_n values), ACCUMULATED (caller zeroes).
 *   normalize : (db - mean) / unbiased std (:171-172), zero for frames t >= T_n (the collate padding, :230,243);
 *               out_nct [N, 1, 64, T_max] fp32 (the reference layout) and / or out_ntc [N, T_max, 64] dtype, nullable.
 * ---------------------------------------------------------------------------------------------- */
int lasr_logmel_padded_len(int T_max);
int lasr_logmel_prepare(const float* wave, const float* dither, const int32_t* num_samples, void* parts, int N,
                        int S_max, int T_max, lasr_stream_t stream);
/* Train-time augmentation on the device, same arithmetic as the reference's host code:
 *   prepare_crop : sub_secquence (data_module.py:138-148, applied at :158-159 AFTER dither + pre-emphasis): utterance n
 *                  keeps samples [starts[n], starts[n] + num_samples[n]) of its pre-emphasised waveform
 *                  (starts == NULL: no crop).  The reference's slice x[:, loc:L] is starts = loc, num_samples = L - loc.
 *   spec_augment : spec_augment (data_module.py:97-122, applied to the dB spectrogram at :163-165, before the
 *                  normalisation): bands [N, 4] int32 = (f0, fw, t0, tw): mel bins [f0, f0+fw) of every frame and
 *                  frames [t0, t0+tw) of every bin are set to 0 and `stats` is corrected accordingly.  Call between
 *                  lasr_logmel_fwd and lasr_logmel_normalize.  The caller draws the band positions (the reference uses
 *                  an unseeded random.Random(), :65,112-116). */
int lasr_logmel_prepare_crop(const float* wave, const float* dither, const int32_t* starts,
                             const int32_t* num_samples, void* parts, int N, int S_max, int T_max,
                             lasr_stream_t stream);
int lasr_spec_augment(float* db, double* stats, const int32_t* num_samples, const int32_t* bands, int N, int T_max,
                      lasr_stream_t stream);
/* prepare with the waveform taken as it travels: wave_dtype LASR_WAVE_F32 ([-1, 1] floats) or LASR_WAVE_I16 (16-bit PCM,
 * x / 32768 like torchaudio.load(normalize=True), data_module.py:153 -- half the H2D bytes of fp32 samples).
 * Dither (data_module.py:155): `dither` [N, S_max] supplied by the caller, or drawn in the kernel (standard normal from
 * Philox keyed by dither_seed + *seed_dev, utterance, sample) when dither == NULL and dither_seed != 0.  seed_dev: a
 * device step counter, so a replayed CUDA graph draws fresh noise.  starts NULL = no crop. */
enum { LASR_WAVE_F32 = 0, LASR_WAVE_I16 = 1 };
int lasr_logmel_prepare_wave(const void* wave, int wave_dtype, const float* dither, uint64_t dither_seed,
                             const uint64_t* seed_dev, const int32_t* starts, const int32_t* num_samples, void* parts,
                             int N, int S_max, int T_max, lasr_stream_t stream);
/* The random draws of parse_audio(mask=True) on the device, in the reference's order and arithmetic (sub_secquence
 * data_module.py:138-148, spec_augment :97-122): num_samples [N] = whole-utterance sample counts ->
 * starts / kept [N] (the arguments of prepare_crop / prepare_wave), bands [N, 4] (the argument of lasr_spec_augment),
 * percents [N] = T_n / T_max (nullable; data_module.py:244 with the batch padded to T_max frames).
 * uniforms [N, 6] double (nullable) = the six random() values per utterance (parity hook); otherwise Philox keyed by
 * seed + *seed_dev.  crop / spec switch the two augmentations off individually (kept = num_samples, empty bands). */
int lasr_augment_draw(const int32_t* num_samples, const double* uniforms, uint64_t seed, const uint64_t* seed_dev,
                      int32_t* starts, int32_t* kept, int32_t* bands, float* percents, int N, int T_max, int crop,
                      int spec, lasr_stream_t stream);
int lasr_logmel_fwd(const void* parts, const void* basis, const int32_t* mel_idx, const float* mel_w,
                    const int32_t* num_samples, float* db, double* stats, int N, int T_max, int products,
                    lasr_stream_t stream);
int lasr_logmel_normalize(const float* db, const double* stats, const int32_t* num_samples, float* out_nct,
                          void* out_ntc, int N, int T_max, int dtype, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * log_softmax over classes (replaces nn.functional.log_softmax, models/QuartNet.py:290).
 * logits [M, ld] dtype (V valid columns) -> lse [M] fp32 and, if lp != NULL, log-probs lp [M, V] fp32 dense.
 * ---------------------------------------------------------------------------------------------- */
int lasr_log_softmax_fwd(const void* logits, float* lse, float* lp, int M, int V, int ld, int dtype,
                         lasr_stream_t stream);
/* dlogits[m, c] = dlp[m,c] - exp(lp[m,c]) * sum_c dlp[m,c]; dlogits [M, ld] dtype */
int lasr_log_softmax_bwd(const float* dlp, const float* lp, void* dlogits, int M, int V, int ld, int dtype,
                         lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * CTC loss (replaces torch.nn.CTCLoss(blank=V, reduction='none'), train.py:196, calls :76-78).
 * Input either log-probs (lse == NULL) or raw logits + their row log-sum-exp (fused log-softmax).
 *   x        [N, T, ldx] dtype, class c of frame t of utterance n at x[(n*T+t)*ldx + c]
 *   targets  [N, S_max] int64 (zero padded), input_lengths / target_lengths [N] int32
 *   alpha    workspace fp32 [N, T, 2*S_max+1]  (kept for the backward)
 *   beta     same shape, nullable: when given, the beta lattice is computed CONCURRENTLY with alpha
 *            (separate CTAs) so the backward is just the parallel combine pass
 *   nll      [N] fp32 out: -log p(target | x); +inf when no alignment exists
 *   scales / emis  workspaces of lasr_ctc_scales_bytes / lasr_ctc_emis_bytes bytes, nullable TOGETHER:
 *            given     -> scaled lattices: 8 warps per (utterance, direction), probabilities kept linear with an exact
 *                         power-of-two rescaling every 4th frame; alpha[n,t,s] = the stored value * 2^scales[0][n][t],
 *                         `beta` holds beta WITHOUT the frame's emission * 2^scales[1][n][t]; scales[2*N*T + 2n, +1] =
 *                         (float bits of P's mantissa part, its exponent).  emis [N, T, pad4(S_max+2)] fp32 receives
 *                         the emission probabilities the lattice reads (labels 0..S_n-1, blank at S_max).
 *            NULL      -> log-space lattices (round-1 kernels: one state per thread), alpha / beta are logarithms.
 *            The same `scales` (or NULL) must be passed to lasr_ctc_bwd.
 * bwd: grad [N, T, ldg] grad_dtype = (softmax - occupancy) * grad_out[n] for t < input_lengths[n], 0 after
 *      (this is also torch's "gradient w.r.t. log-probs").
 * ---------------------------------------------------------------------------------------------- */
size_t lasr_ctc_scales_bytes(int N, int T);
size_t lasr_ctc_emis_bytes(int N, int T, int S_max);
int lasr_ctc_fwd(const void* x, const float* lse, const int64_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, float* alpha, float* beta, float* nll, int32_t* scales, float* emis,
                 int N, int T, int V, int ldx, int S_max, int blank, int dtype, lasr_stream_t stream);
int lasr_ctc_bwd(const void* x, const float* lse, const int64_t* targets, const int32_t* input_lengths,
                 const int32_t* target_lengths, const float* alpha, const float* beta, const float* nll,
                 const int32_t* scales, const float* grad_out, void* grad, int N, int T, int V, int ldx, int ldg,
                 int S_max, int blank, int dtype, int grad_dtype, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Greedy CTC decode (replaces out.argmax(-1) train.py:80 + the collapse loop utils/asr_metrics.py:153-171).
 *   x [N, T, ldx] dtype scores (log-probs or logits; argmax ties -> lowest index, like torch.argmax)
 *   lengths [N] int32 or NULL (decode all T frames, predict.py:60)
 *   argmax  [N, T] int64 out: the raw per-frame argmax (torch.argmax semantics)
 *   tokens  [N, T] int32 out (nullable: argmax only): collapsed label ids, first counts[n] entries valid
 * ---------------------------------------------------------------------------------------------- */
int lasr_greedy_decode(const void* x, const int32_t* lengths, int64_t* argmax, int32_t* tokens, int32_t* counts, int N,
                       int T, int V, int ldx, int blank, int dtype, lasr_stream_t stream);
/* the collapse rule alone (utils/asr_metrics.py:159-167) on caller-supplied predictions [N, T] int64 */
int lasr_ctc_collapse(const int64_t* predictions, const int32_t* lengths, int32_t* tokens, int32_t* counts, int N,
                      int T, int blank, lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Novograd optimizer step + cosine-annealing-with-warm-restarts schedule over flat buffers.
 * Replaces scheduler/novograd.py:75-145 (Novograd.step, betas=(0.8, 0.5), weight decay added after the layer-wise
 * normalisation, train.py:46) and scheduler/cosine_annearing_with_warmup.py:53-89 (per-step LR schedule, train.py:53-55).
 *   params / grads / exp_avg   flat fp32 buffers of the same length (runtime.ParamBank layout), 16-byte aligned
 *   shadow_bf16                nullable: bf16 copy of params refreshed in the same pass
 *   chunk_off/len/param        [num_chunks] int32: consecutive pieces (<= a few thousand elements, offsets multiple of
 *                              4) that each lie inside ONE parameter tensor, and that tensor's index
 *   norms                      [num_params] fp64, ZEROED by the caller: receives |grad_p|^2
 *   exp_avg_sq, denom          [num_params] fp32: second-moment state (0 = uninitialised, like the reference) / scratch
 *   sched                      nullable device struct: when given, the step uses sched->lr and then advances the
 *                              schedule like CosineAnnealingWarmupRestarts.step(); when NULL the scalar `lr` is used
 *   lr_use                     [1] fp32 scratch (the learning rate this step applied; readable afterwards)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double base_max_lr, max_lr, min_lr, cycle_mult, gamma, lr;
  int32_t first_cycle_steps, cur_cycle_steps, warmup_steps, cycle, step_in_cycle, last_epoch;
} lasr_lr_sched_t;
int lasr_novograd_step(float* params, const float* grads, float* exp_avg, void* shadow_bf16, const int32_t* chunk_off,
                       const int32_t* chunk_len, const int32_t* chunk_param, int num_chunks, double* norms,
                       float* exp_avg_sq, float* denom, int num_params, lasr_lr_sched_t* sched, float* lr_use,
                       float lr, float beta1, float beta2, float eps, float weight_decay, int grad_averaging,
                       lasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Context BiLSTM recurrence.  Replaces pack_padded_sequence -> nn.LSTM(256, 40, bidirectional=True)
 * -> pad_packed_sequence at models/QuartNetContext.py:171-173,186-199 (and the `.cpu()` length sync at :171).  The
 * input projections are a pointwise-conv GEMM issued by the caller (lasr_pwconv_fwd with bias b_ih + b_hh):
 *   pre    [N, T, 320] dtype: column d*160 + g*40 + u = gate g (i, f, g, o) of unit u, direction d (0 fwd, 1 reverse)
 *   whh    [2, 160, 40] fp32 (weight_hh_l0, weight_hh_l0_reverse)
 *   lengths [N] int32 or NULL; frames >= lengths[n] give zeros (pad_packed_sequence) and no gradient
 *   out    [N, T, 80] dtype: [h_fwd | h_reverse]
 *   gates  [N, T, 2, 40, 4] fp32 (16-byte aligned) and cells [N, T, 2, 40] fp32: saved for the backward
 * bwd: dpre [N, T, 320] dtype out; dwhh [2, 160, 40] fp32 ACCUMULATED.  dW_ih / db / dx follow from dpre through
 *      lasr_pwconv_wgrad / lasr_colsum / lasr_pwconv_dgrad.   hidden must be 40 (the reference's only use).
 * ---------------------------------------------------------------------------------------------- */
int lasr_bilstm_fwd(const void* pre, const float* whh, const int32_t* lengths, void* out, float* gates, float* cells,
                    int N, int T, int hidden, int dtype, lasr_stream_t stream);
int lasr_bilstm_bwd(const void* dout, const void* out, const float* gates, const float* cells, const float* whh,
                    const int32_t* lengths, void* dpre, float* dwhh, int N, int T, int hidden, int dtype,
                    lasr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LASR_H_ */
