/*
 * dl4ds_b200 -- C ABI of the B200-native DL4DS convolutional super-resolution hot path.
 *
 * The reference (carlos-gg/dl4ds @ 232ae49) has no FFI: its hot path is whatever TensorFlow/Keras
 * executes for the graphs built in dl4ds/models/ and the step logic in dl4ds/training/.  This
 * library is the NEW lower boundary that replaces that runtime.  Every
 * entry point below names the reference call site(s) whose arithmetic it replaces.
 *
 * Conventions
 *  - plain C, no torch types; pointers are DEVICE pointers unless stated otherwise;
 *  - every function returns int: 0 = OK, negative = DL4DS_E_*; dl4ds_last_error() gives text;
 *  - no allocation inside: the caller owns every buffer, including workspaces;
 *  - every launch is asynchronous on the cudaStream_t passed as `void* stream`;
 *  - tensors are dense fp32 NHWC.  `*_ld` arguments are the channel pitch (elements between
 *    consecutive pixels) so a tensor may live inside a wider concat buffer; the pointer already
 *    includes the channel offset;
 *  - weights use the Keras layouts: Conv2D (kh,kw,Cin,Cout); Conv2DTranspose (kh,kw,Cout,Cin);
 *    ConvLSTM2D (kh,kw,Cin,4F)/(kh,kw,F,4F) gate order i,f,c,o; LocallyConnected2D 1x1
 *    W[H,W,Cin,F], b[H,W,F]; Dense (in,out);
 *  - parameter-gradient outputs ACCUMULATE (+=) into their destination (the flat gradient arena is
 *    zeroed once per step); this is what makes shared-weight layers (SubpixelConvolutionBlock
 *    conv2x, DeconvolutionBlock T2 -- blocks.py:415,421-422,528-531) sum over applications.
 */
#ifndef DL4DS_B200_H
#define DL4DS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DL4DS_OK             0
#define DL4DS_E_BADARG      -1
#define DL4DS_E_SHAPE       -2
#define DL4DS_E_UNSUPPORTED -3
#define DL4DS_E_CUDA        -4
#define DL4DS_E_NCCL        -5

/* activation codes (tf.keras.layers.Activation, blocks.py:75) */
#define DL4DS_ACT_NONE    0
#define DL4DS_ACT_RELU    1
#define DL4DS_ACT_SIGMOID 2
#define DL4DS_ACT_TANH    3

/* math modes for the convolution kernels */
#define DL4DS_MATH_FP32    0   /* CUDA-core fp32 FMA (exact fp32 semantics)                     */
#define DL4DS_MATH_TF32X3  1   /* tcgen05 kind::tf32, 3-term split (hi*hi+hi*lo+lo*hi), ~fp32   */
#define DL4DS_MATH_TF32    2   /* tcgen05 kind::tf32, single pass (10-bit mantissa operands)    */
#define DL4DS_MATH_F16X3   3   /* forward / dgrad: tcgen05 kind::f16, 3-term fp16 split with    */
                               /* power-of-two scales per tile and per weight tensor (~fp32,    */
                               /* half the MMAs of TF32X3); every other kernel as TF32X3        */

/* weight indexing modes of dl4ds_conv2d_fwd */
#define DL4DS_W_HWIO        0  /* B[(tap,c),n] = w[tap][c][n]            (Conv2D forward)          */
#define DL4DS_W_FLIP_T      1  /* B[(tap,c),n] = w[flip(tap)][n][c]      (Conv2D dgrad, ConvT fwd) */
#define DL4DS_W_PREPACKED   4  /* OR-ed into wmode: `ws` already holds dl4ds_conv2d_pack(w, wmode) */

const char* dl4ds_last_error(void);
int dl4ds_version(void);
/* 1 if the device behind the current context is sm_100 (tcgen05 paths usable). */
int dl4ds_device_is_sm100(void);
/* Number of tcgen05 (tensor-core) kernel launches issued by this process so far: lets callers and
 * tests verify that a tensor-core math mode did not silently take the CUDA-core path. */
int64_t dl4ds_tc_launch_count(void);
/* Developer aid: a device buffer of >= 1088 int64 that CTA (0,0) of the weight-gradient tensor-core kernel
 * fills with clock64() stamps of its pipeline stages (scratch/wg2_stamps.py); NULL (default) disables it. */
int dl4ds_debug_set_buffer(void* dev_i64);

/* ---------------------------------------------------------------------------------------------
 * Convolution family.  One generalized implicit-GEMM entry point covers:
 *   Conv2D forward                       blocks.py:49-61,91,97,208,299; sp_postups.py:134,156
 *   Conv2D input gradient (dgrad)        what TF's GradientTape derives for the same layers
 *   Conv2DTranspose forward              blocks.py:508-516 (DeconvolutionBlock)
 *   Conv2DTranspose input gradient       = strided Conv2D forward
 *
 * y[n,oy,ox,co] = act( sum_{kh,kw,c} x[n,(oy*stride+kh-pad_t)/up,(ox*stride+kw-pad_l)/up,c] * B[(kh,kw,c),co]
 *                      + bias[co] + res[n,oy,ox,co] )
 * where a tap only contributes if the numerator is divisible by `up` and the source pixel is in
 * range (zero padding).  `up`>1 is the fractional stride used by Conv2DTranspose forward / strided
 * dgrad.  If d2s_r>1 the result is stored through tf.nn.depth_to_space(., r) (NHWC DCR order,
 * blocks.py:427): y has shape (N, Ho*r, Wo*r, Cout/(r*r)) and `res` must be NULL.
 * bias and res may be NULL.  beta=1 accumulates into y (y += result; only with act NONE, no d2s).
 *
 * Tensor-core math modes (TF32 / TF32X3) run the tcgen05 implicit-GEMM kernel when the shape is in
 * its domain (stride 1, up 1, 'same' grid, Cin and Cout multiples of 8, Cout <= 256, W a power of two
 * in [8,128] or a multiple of 128, 16-byte aligned tensors); otherwise the call runs the CUDA-core
 * fp32 kernel.  They need `ws`: dl4ds_conv2d_fwd_workspace_bytes() bytes, 128-byte aligned, into which
 * the call packs the weights (tf32 hi/lo split, swizzled shared-memory image) -- or, with
 * DL4DS_W_PREPACKED in wmode, which already holds dl4ds_conv2d_pack() of the same (w, wmode), so one
 * pack per optimizer step serves every application of a layer.  ws may be NULL when the query is 0.
 * ------------------------------------------------------------------------------------------- */
int64_t dl4ds_conv2d_fwd_workspace_bytes(int N, int H, int W, int Cin, int Ho, int Wo, int Cout,
                                         int KH, int KW, int stride, int up, int d2s_r, int math_mode);
int dl4ds_conv2d_pack(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode,
                      void* ws, void* stream);
/* All weight images of a model in one launch: dl4ds_conv2d_pack_desc() writes one 64-byte descriptor (host memory)
 * for what dl4ds_conv2d_pack() would pack and returns its number of 16-byte units (< 0: error); the caller stores
 * the running sum of the units of the preceding descriptors into bytes [56, 64) of each record (int64), copies the
 * table to the device and calls dl4ds_conv2d_pack_multi(table, n, total units) once per optimizer step. */
int64_t dl4ds_conv2d_pack_desc(const float* w, int wmode, int KH, int KW, int Cin, int Cout, int math_mode,
                               void* ws, void* desc_out_host);
int dl4ds_conv2d_pack_multi(const void* descs_dev, int n, int64_t total_units, void* stream);
int dl4ds_conv2d_fwd(const float* x, int x_ld, const float* w, const float* bias,
                     const float* res, int res_ld, float* y, int y_ld,
                     int N, int H, int W, int Cin, int Ho, int Wo, int Cout,
                     int KH, int KW, int stride, int up, int pad_t, int pad_l,
                     int wmode, int act, int d2s_r, int beta, int math_mode, void* ws, void* stream);

/* Input gradient of a stride-1 'same' Conv2D with the epilogue-backward of the layer that PRODUCED the
 * convolution's input fused into the store (what GradientTape does between two Keras layers,
 * blocks.py:87-103,210-230: d(pre-activation) = d(activation) * act'(.), d(bias) = its sum over pixels):
 *   dz[n,y,x,c] = ( dgrad(dq, w)[n,y,x,c] (+ dz[n,y,x,c] if beta) ) * act'(y_prod[n,y,x,c])
 *   dbias[c]   += sum_{n,y,x} dz[n,y,x,c]
 * dq: (N,H,W,Cq) gradient w.r.t. this convolution's output; w: its HWIO kernel (KH,KW,Cp,Cq) read with
 * DL4DS_W_FLIP_T (or its packed image with DL4DS_W_PREPACKED); dz: (N,H,W,Cp).  y_prod (may be NULL: no
 * activation) is the producer's forward OUTPUT, act its activation code; dbias may be NULL.  Valid only when
 * this launch is the LAST contribution to dz.  Tensor-core math modes only, shapes in the halo-tile kernel's
 * domain (dl4ds_conv2d_dgrad_fused_supported() == 1: W % 8 == 0, H % 16 == 0, Cp and Cq multiples of 8,
 * Cp <= 256); otherwise DL4DS_E_UNSUPPORTED -- the caller then runs dl4ds_conv2d_fwd + dl4ds_bias_act_bwd.
 * Cq == Cp == 8 (the 8-channel HR tail): served by the warp-level fp16 3-term kernel instead (3x3, W % 32 == 0,
 * W <= 128 or W % 128 == 0, N*H*W >= 16384, math TF32X3 / F16X3); that kernel packs nothing -- ws may be NULL. */
int dl4ds_conv2d_dgrad_fused_supported(int N, int H, int W, int Cq, int Cp, int KH, int KW, int math_mode);
int dl4ds_conv2d_dgrad_fused(const float* dq, int dq_ld, const float* w, float* dz, int dz_ld,
                             const float* y_prod, int y_ld, int act, float* dbias,
                             int N, int H, int W, int Cq, int Cp, int KH, int KW, int pad_t, int pad_l,
                             int wmode, int beta, int math_mode, void* ws, void* stream);

/* Weight gradient of the same family (accumulating):
 *   dw[kh][kw][a][b] += sum_{n,oy,ox} P[n,oy*stride+kh-pad_t,ox*stride+kw-pad_l,a] * Q[n,oy,ox,b]
 * Conv2D: P = layer input (Ca=Cin), Q = dZ (Cb=Cout).  Conv2DTranspose (kernel (kh,kw,Cout,Cin)):
 * P = dY (Ca=Cout), Q = layer input (Cb=Cin).  (Hp,Wp) is P's grid, (Hq,Wq) is Q's grid.
 * `ws` is a caller-owned workspace of dl4ds_conv2d_wgrad_workspace_bytes() bytes (may be NULL if 0). */
int64_t dl4ds_conv2d_wgrad_workspace_bytes(int N, int Hq, int Wq, int Ca, int Cb, int KH, int KW,
                                           int math_mode);
int dl4ds_conv2d_wgrad(const float* P, int p_ld, const float* Q, int q_ld, float* dw,
                       int N, int Hp, int Wp, int Ca, int Hq, int Wq, int Cb,
                       int KH, int KW, int stride, int pad_t, int pad_l,
                       void* ws, int math_mode, void* stream);

/* SubpixelConvolutionBlock's last x2 stage (linear Conv2D + depth_to_space(r), blocks.py:421-427) composed with
 * the 1x1 convolution that consumes it (TransitionLast, sp_postups.py:205 / blocks.py:299):
 *   weff[row][d*Co+co] = sum_c w1[row][d*Cm+c] * w2[c][co],  beff[d*Co+co] = sum_c b1[d*Cm+c]*w2[c][co] + b2[co]
 * (row = (tap, ci) < rows, d < r*r).  conv2d_fwd(x, weff, beff, act, d2s_r=r) then equals
 * act(conv1x1(depth_to_space(conv(x, w1) + b1), w2) + b2) up to fp32 re-association, and the Cm-channel HR tensor is
 * never materialised.  _chain is the exact chain rule from (dweff, dbeff) to the original parameter gradients
 * (all four accumulate; gb1 / gb2 may be NULL). */
int dl4ds_spc_pointwise_compose(const float* w1, const float* b1, const float* w2, const float* b2,
                                float* weff, float* beff, int rows, int Cm, int Co, int r, void* stream);
int dl4ds_spc_pointwise_chain(const float* w1, const float* b1, const float* w2, const float* dweff, const float* dbeff,
                              float* gw1, float* gb1, float* gw2, float* gb2, int rows, int Cm, int Co, int r,
                              void* stream);

/* Backward of the fused bias+activation(+depth_to_space) epilogue:
 *   dz = dy * act'(y)   (act' expressed through the stored output y; NONE: dz = dy)
 *   dbias[c] += sum_pixels dz[.,c]          (dbias may be NULL)
 * With d2s_r>1 dy (and y, read only when act != NONE) are in the HR layout (N,Ho*r,Wo*r,C/(r*r)) and
 * dz is written un-shuffled (N,Ho,Wo,C) -- the space_to_depth adjoint of blocks.py:427.
 * dz may alias dy when d2s_r==1.  n_pix = N*Ho*Wo. */
int dl4ds_bias_act_bwd(const float* dy, int dy_ld, const float* y, int y_ld, float* dz, int dz_ld,
                       float* dbias, int N, int Ho, int Wo, int C, int act, int d2s_r, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Element-wise helpers (Add -- sp_postups.py:164; Concatenate -- blocks.py:276, sp_postups.py:186,201)
 * out[p,c] = a[p,c] + b[p,c]  (act applied after the sum);   copy/accumulate of channel slices.
 * ------------------------------------------------------------------------------------------- */
int dl4ds_add(const float* a, int a_ld, const float* b, int b_ld, float* out, int out_ld,
              int64_t n_pix, int C, int act, void* stream);
/* dst[p, 0:C] (= or +=) src[p, 0:C] with independent pitches. */
int dl4ds_copy_channels(const float* src, int src_ld, float* dst, int dst_ld,
                        int64_t n_pix, int C, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ChannelAttention2D -- blocks.py:537-593.  y = x * sigmoid(W2 relu(W1 mean_HW(x) + b1) + b2).
 *  fwd : pooled[N,C] (sum over H*W, workspace, zeroed by the call), hidden[N,Cr], scale[N,C] are
 *        saved for backward.  w1 (C,Cr), w2 (Cr,C) are the 1x1 Conv2D kernels.
 *  bwd : dx = dy*scale + dmean/(H*W); parameter gradients accumulate.  dsum[N,C] is workspace.
 *  n_groups attention vectors, each pooled over pix_per_group pixels.  Pixel p belongs to group
 *  (p / (pix_per_group*inner))*inner + p % inner.  4-D tensors: n_groups = N, pix_per_group = H*W,
 *  inner = 1.  5-D (B,T,H,W,C) tensors pooled over axes [1,2] = (T,H) as blocks.py:587 does when
 *  called from spt_postups.py:153-154: n_groups = B*W, pix_per_group = T*H, inner = W.
 * ------------------------------------------------------------------------------------------- */
int dl4ds_channel_attention_fwd(const float* x, int x_ld, float* y, int y_ld,
                                const float* w1, const float* b1, const float* w2, const float* b2,
                                float* pooled, float* hidden, float* scale,
                                int n_groups, int64_t pix_per_group, int inner, int C, int Cr,
                                void* stream);
int dl4ds_channel_attention_bwd(const float* x, int x_ld, const float* dy, int dy_ld,
                                float* dx, int dx_ld,
                                const float* w1, const float* w2,
                                const float* pooled, const float* hidden, const float* scale,
                                float* dsum, float* dw1, float* db1, float* dw2, float* db2,
                                int n_groups, int64_t pix_per_group, int inner, int C, int Cr,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pixel losses -- losses.py:5-20 (Keras MeanAbsoluteError / MeanSquaredError = global mean).
 * kind 0 = MAE, 1 = MSE.  loss_out[0] += scale * mean(...) (caller zeroes it); if dy != NULL,
 * dy = scale * d(mean)/d(y_pred) (dy += ... when DL4DS_LOSS_ACCUMULATE is OR-ed into kind: the second
 * pixel term of the weighted mixes, losses.py:71-84,143-151).  n = total element count.
 * ------------------------------------------------------------------------------------------- */
#define DL4DS_LOSS_ACCUMULATE 16
int dl4ds_pixel_loss(const float* y_pred, const float* y_true, float* loss_out, float* dy,
                     int64_t n, int kind, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SSIM-family losses -- losses.dssim / losses.msdssim, losses.py:27-59,96-131 (and the weighted mixes
 * :62-93,134-151, which add dl4ds_pixel_loss terms).  tf.image.ssim / tf.image.ssim_multiscale semantics:
 * 11x11 gaussian (sigma 1.5), k1 0.01, k2 0.03, VALID filtering, max_val = max(both) - min(both), each tensor
 * shifted by its own minimum when negative.  n_scales 1 = SSIM; > 1 = MS-SSIM with `power_factors` (HOST
 * array, n_scales entries; the reference passes 4: losses.py:128), 2x2 average pooling between scales (odd
 * sizes SYMMETRIC-padded by one first, as tf.image.ssim_multiscale does).  loss_out[0] += scale * mean_b((1 - ssim_b) / 2); if dy != NULL, dy (+)= scale * d/d y_pred,
 * including the gradient through the dynamic range and the shift (arg-max / arg-min elements of y_pred).
 * `ws`: dl4ds_ssim_loss_workspace_floats(...) floats of device scratch, 16-byte aligned (no allocation inside).
 * ------------------------------------------------------------------------------------------- */
int64_t dl4ds_ssim_loss_workspace_floats(int B, int H, int W, int C, int n_scales);
int dl4ds_ssim_loss(const float* y_pred, const float* y_true, int B, int H, int W, int C, int n_scales,
                    const float* power_factors, float scale, float* loss_out, float* dy, int accumulate,
                    float* ws, void* stream);

/* tf.image.ssim(img1, img2, max_val) per image as compute_metrics calls it (metrics.py:172-176): no shift, the given
 * dynamic range; out[b] = mean over channels and windows.  `ws`: dl4ds_ssim_loss_workspace_floats(B,H,W,C,1) floats. */
int dl4ds_ssim_index(const float* img1, const float* img2, int B, int H, int W, int C, float max_val, float* out,
                     float* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Verification metrics -- metrics.py:15-97,166-186 (compute_rmse / compute_correlation 'time' and 'space', PSNR,
 * MAE, dynamic range): raw fp64 moments of the (N, P) fp32 pair y / y_hat.
 *   pair_out  [N][11]: sum d^2, sum |d|, sum y, sum yh, sum y^2, sum yh^2, sum y*yh over the P values of sample n,
 *                      then min y, max y, min yh, max yh                                  (NULL: skipped)
 *   point_out [P][7] : the same seven sums over the N samples for every grid value        (NULL: skipped)
 * ------------------------------------------------------------------------------------------- */
int dl4ds_metrics_moments(const float* y, const float* y_hat, int N, int64_t P, double* pair_out, double* point_out,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Normalisation layers of the conv blocks (`normalization='bn' | 'ln'`) with the following activation fused --
 * blocks.py:63-71 (construction), :94-101 ConvBlock, :216-224 ResidualBlock, :263-272 DenseBlock, :298-305
 * TransitionBlock, :160-164,178 ConvNextBlock.  Keras defaults: axis -1, epsilon 1e-3 (1e-6 in ConvNext's LN),
 * BatchNormalization momentum 0.99.  NHWC tensors with a channel pitch (`*_ld`), C <= 256.
 *
 * dl4ds_batchnorm_stats: training-mode statistics over all n_pix pixels: mean[c], var[c] (biased, what the layer
 *   normalises with) and, if moving_mean/moving_var != NULL, the moving averages updated in place
 *   (moving = moving * momentum + batch * (1 - momentum); the moving variance takes the unbiased batch variance,
 *   as Keras' fused path does).  ws: 2*C floats of device scratch.
 * dl4ds_norm_apply: y = act(gamma * (x - mean) / sqrt(var + eps) + beta) -- batch statistics in training,
 *   the moving ones in inference (Predictor / validation).
 * dl4ds_batchnorm_bwd: dz = dy * act'(y); dgamma += sum dz*xhat; dbeta += sum dz (NULL = not wanted);
 *   dx = gamma/sqrt(var+eps) * (dz - mean(dz) - xhat * mean(dz*xhat)).  ws: 2*C floats.
 * dl4ds_layernorm_fwd / _bwd: the same per pixel over the channel axis (dx may be NULL).
 * ------------------------------------------------------------------------------------------- */
int dl4ds_batchnorm_stats(const float* x, int x_ld, int64_t n_pix, int C, float* mean, float* var,
                          float* moving_mean, float* moving_var, float momentum, float* ws, void* stream);
int dl4ds_norm_apply(const float* x, int x_ld, const float* mean, const float* var, const float* gamma,
                     const float* beta, float eps, float* y, int y_ld, int64_t n_pix, int C, int act,
                     void* stream);
int dl4ds_batchnorm_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* y, int y_ld,
                        const float* mean, const float* var, const float* gamma, float eps, float* dx, int dx_ld,
                        float* dgamma, float* dbeta, float* ws, int64_t n_pix, int C, int act, void* stream);
int dl4ds_layernorm_fwd(const float* x, int x_ld, const float* gamma, const float* beta, float eps, float* y,
                        int y_ld, int64_t n_pix, int C, int act, void* stream);
int dl4ds_layernorm_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* y, int y_ld,
                        const float* gamma, float eps, float* dx, int dx_ld, float* dgamma, float* dbeta,
                        int64_t n_pix, int C, int act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ConvNextBlock pieces -- blocks.py:131-184.  DepthwiseConv2D(kernel_size=7, padding='same', depth_multiplier=1)
 * (:147-148): w is the Keras depthwise kernel (k, k, C, 1) = [k][k][C], stride 1, odd k <= 7.
 *   fwd:   y (+)= bias + sum_ij w[i][j][c] x[h+i-r, w+j-r, c];  flip = 1 uses w[k-1-i][k-1-j]: the input gradient
 *          (call it with x := dy, bias := NULL).
 *   wgrad: dw[i][j][c] += sum dy[p, c] x[p + (i-r, j-r), c]   (the bias gradient is dl4ds_bias_act_bwd's dbias).
 * GELU, the block's default activation (:143,153), exact erf form: y = x Phi(x); dx = dy (Phi(x) + x phi(x)).
 * ------------------------------------------------------------------------------------------- */
int dl4ds_depthwise_conv_fwd(const float* x, int x_ld, const float* w, const float* bias, float* y, int y_ld,
                             int N, int H, int W, int C, int k, int flip, int accumulate, void* stream);
int dl4ds_depthwise_conv_wgrad(const float* x, int x_ld, const float* dy, int dy_ld, float* dw, int N, int H, int W,
                               int C, int k, void* stream);
/* ConvNextBlock's layer scale (blocks.py:166-179): y = gamma[c] * x; bwd: dx = gamma[c] * dy (dx may be NULL),
 * dgamma[c] += sum_pixels dy * x (dgamma may be NULL).  C <= 256 for the backward. */
int dl4ds_channel_scale_fwd(const float* x, int x_ld, const float* gamma, float* y, int y_ld, int64_t n_pix, int C,
                            void* stream);
int dl4ds_channel_scale_bwd(const float* dy, int dy_ld, const float* x, int x_ld, const float* gamma, float* dx,
                            int dx_ld, float* dgamma, int64_t n_pix, int C, void* stream);
int dl4ds_gelu_fwd(const float* x, float* y, int64_t n, void* stream);
int dl4ds_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dropout variants -- blocks.py:659-706 (`get_dropout_layer`) applied at blocks.py:89,95-96,211-214,219-220,
 * 265-266,271-272 and sp_postups.py:158.  variant 0 = Dropout (keep where u >= rate, scale 1/(1-rate)),
 * 1 = GaussianDropout (x * N(1, sqrt(rate/(1-rate)))), 2 = SpatialDropout2D / 3D (one draw per sample and channel;
 * sample = (pixel / pix_per_sample) % n_samples: n_samples = N for (N,H,W,C), = B for time-major frames (T*B,H,W,C)).
 * 3 = DropPath (blocks.py:106-129: one keep / drop draw per SAMPLE, kept samples scaled by 1/(1-rate)).
 * y = x * mask(seed, step, layer_id, element): Philox4x32-10, `rng_state` = DEVICE uint64[2] {seed, step}.  The
 * mask is a pure function of its arguments: the backward pass is the same call on dy.  dl4ds_rng_advance bumps
 * `step` (one kernel, captured with the step graph, so every replay draws new masks).  TensorFlow's own random
 * streams are not reproducible; parity tests obtain the mask by applying the call to a tensor of ones.
 * ------------------------------------------------------------------------------------------- */
int dl4ds_dropout(const float* x, int x_ld, float* y, int y_ld, int64_t n_pix, int64_t pix_per_sample, int n_samples,
                  int C, float rate, int variant, const uint64_t* rng_state, int layer_id, void* stream);
int dl4ds_rng_advance(uint64_t* rng_state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * tf.keras.optimizers.Adam step on a flat arena -- supervised.py:353, cgan.py:277-278.
 *   g = grad * grad_scale (grad_scale = 1/world_size folds Horovod's allreduce-average,
 *   supervised.py:365);  m,v updated;  theta -= lr_t * m / (sqrt(v) + eps)
 *   with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller-independent formula inside.
 * ------------------------------------------------------------------------------------------- */
int dl4ds_adam_step(float* theta, const float* grad, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, int t, float grad_scale,
                    void* stream);

/* Same update with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) read from DEVICE memory (one float), so that a
 * captured CUDA graph of the whole step can be replayed while the host advances t / the
 * PiecewiseConstantDecay schedule (supervised.py:336-353). */
int dl4ds_adam_step_dev(float* theta, const float* grad, float* m, float* v, int64_t n,
                        const float* lr_t_dev, float beta1, float beta2, float eps, float grad_scale,
                        void* stream);

/* Conv2DTranspose(k, strides=stride, padding='same', use_bias=False) (blocks.py:508-516) re-expressed as a stride-1
 * Kp x Kp convolution to stride^2*Co channels + depth_to_space(stride):  backward == 0 builds the HWIO weight image
 *   wp[my][mx][ci][(dy*stride+dx)*Co+co] = w[k-1-kh][k-1-kw][co][ci]  (w in the Keras layout (kh,kw,Co,Ci)), kh = stride*(my+off_min)-dy+pad
 * (0 where kh / kw fall outside the kernel); backward != 0 adds dwp back onto dw (= `w` argument) through the same map.
 * pad = k-1-pad_before of the transpose, off_min = smallest input offset any phase reads, Kp = number of offsets. */
int dl4ds_convt_rearrange(float* w, float* wp, int k, int stride, int pad, int off_min, int Kp, int Co, int Ci,
                          int backward, void* stream);

/* Device-resident batch assembly (create_batch_hr_lr, dataloader.py:297-360, without the per-sample host loop):
 * dst[b, y, x, dst_coff + c] = src[idx[b], y0[b] + y, x0[b] + x, c] for b < n, y < ph, x < pw, c < C.  src is the
 * whole (Ns, H, W, C) array resident in HBM; idx, y0, x0 are DEVICE int32 arrays (y0 / x0 NULL = no crop offset). */
int dl4ds_gather_crop(const float* src, const int* idx, const int* y0, const int* x0, float* dst,
                      int n, int H, int W, int C, int ph, int pw, int dst_ld, int dst_coff, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data path: HR -> LR coarsening by s x s block mean == cv2.resize(INTER_AREA) at an integer
 * factor -- utils.py:376-384 as called from dataloader.py:204,208.  (N,H,W,C) -> (N,H/s,W/s,C).
 * ------------------------------------------------------------------------------------------- */
int dl4ds_avgpool_coarsen(const float* x, float* y, int N, int H, int W, int C, int s, void* stream);

/* Data path: cv2.resize for every `interpolation` of utils.resize_array (utils.py:341-401: inter_area, nearest,
 * bilinear, bicubic, lanczos) as a separable resampling with tap tables: y[n, oy, ox, y_coff + c] =
 * sum_j wy[oy*Ky + j] * sum_k wx[ox*Kx + k] * x[n, iy[oy*Ky + j], ix[ox*Kx + k], c].  Tables are DEVICE arrays
 * (int32 indices, fp32 weights; zero weights are skipped); x dense (N,H,W,C), y a channel slice of pitch y_ld.
 * Used by the device-resident data path for `pin` pairs (coarsen + re-interpolate, dataloader.py:108-150) and for
 * the non-`inter_area` coarsenings (dataloader.py:157-222). */
int dl4ds_resample_taps(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo, const int* iy,
                        const float* wy, int Ky, const int* ix, const float* wx, int Kx, int y_ld, int y_coff,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Remaining graph ops for the dense / U-Net / recurrent / cGAN configurations.
 * ------------------------------------------------------------------------------------------- */
/* Resizing(h, w, 'bilinear') half-pixel centers -- blocks.py:489, discriminator.py:62.
 * bwd accumulates into dx (caller zeroes). */
int dl4ds_resize_bilinear_fwd(const float* x, int x_ld, float* y, int y_ld,
                              int N, int H, int W, int C, int Ho, int Wo, void* stream);
int dl4ds_resize_bilinear_bwd(const float* dy, int dy_ld, float* dx, int dx_ld,
                              int N, int H, int W, int C, int Ho, int Wo, void* stream);
/* Keras Resizing(interpolation=...) of ResizeConvolutionBlock -- blocks.py:457-491 (`rc_interpolation` of the
 * builders): tf.image.resize(method, antialias=False), half-pixel centres.  method 0 = bilinear (the two calls above),
 * 1 = nearest (in = min(floor((out+0.5)*scale), in-1)), 2 = bicubic (Keys cubic A = -0.5, offsets quantised to
 * 1/1024 as TF's coefficient table does, out-of-image taps zeroed and the weights renormalised).
 * bwd ACCUMULATES into dx (caller zeroes it), as dl4ds_resize_bilinear_bwd does. */
#define DL4DS_RESIZE_BILINEAR 0
#define DL4DS_RESIZE_NEAREST  1
#define DL4DS_RESIZE_BICUBIC  2
int dl4ds_resize_fwd(const float* x, int x_ld, float* y, int y_ld, int N, int H, int W, int C, int Ho, int Wo,
                     int method, void* stream);
int dl4ds_resize_bwd(const float* dy, int dy_ld, float* dx, int dx_ld, int N, int H, int W, int C, int Ho, int Wo,
                     int method, void* stream);
/* MaxPooling2D((2,2)) -- blocks.py:613.  bwd routes dy to the first max of each window and
 * WRITES dx fully (zeros elsewhere, including odd trailing rows/cols). */
int dl4ds_maxpool2_fwd(const float* x, int x_ld, float* y, int y_ld,
                       int N, int H, int W, int C, void* stream);
int dl4ds_maxpool2_bwd(const float* x, int x_ld, const float* dy, int dy_ld, float* dx, int dx_ld,
                       int N, int H, int W, int C, void* stream);
/* LocallyConnected2D(f,(1,1),implementation=3) -- blocks.py:322-328.  Parameter grads accumulate. */
int dl4ds_local_conv1x1_fwd(const float* x, int x_ld, const float* w, const float* b,
                            float* y, int y_ld, int N, int H, int W, int Cin, int F, void* stream);
int dl4ds_local_conv1x1_bwd(const float* x, int x_ld, const float* dy, int dy_ld, const float* w,
                            float* dx, int dx_ld, float* dw, float* db,
                            int N, int H, int W, int Cin, int F, void* stream);
/* ConvLSTM2D cell pointwise part -- blocks.py:350-355 (Keras 2.x: tanh / hard_sigmoid, gates
 * i,f,c,o).  z (n_pix,4F) = conv(x_t,Wx)+b+conv(h_{t-1},Wh) computed with dl4ds_conv2d_fwd.
 * fwd: c = f*c_prev + i*tanh(zc); h = o*tanh(c).  gates (n_pix,4F) saves i,f,g,o for backward.
 * bwd: given dh (total gradient wrt h_t) and dc_next (gradient flowing into c_t from t+1, may be
 * NULL), writes dz (n_pix,4F) and dc_prev. */
int dl4ds_convlstm_gates_fwd(const float* z, const float* c_prev, float* c, float* h, int h_ld,
                             float* gates, int64_t n_pix, int F, void* stream);
int dl4ds_convlstm_gates_bwd(const float* gates, const float* c_prev, const float* c,
                             const float* dh, int dh_ld, const float* dc_next,
                             float* dz, float* dc_prev, int64_t n_pix, int F, void* stream);
/* Standalone activation (RecurrentConvBlock act on h, blocks.py:391,397): y = act(x);
 * bwd: dx = dy*act'(y). */
int dl4ds_act_fwd(const float* x, int x_ld, float* y, int y_ld, int64_t n_pix, int C, int act,
                  void* stream);
/* Mean over `pix_per_group` pixels: GlobalAveragePooling2D -- discriminator.py:76.
 * out[g,c] = mean; bwd: dx[g,p,c] = dout[g,c]/pix_per_group (written, not accumulated). */
int dl4ds_group_mean_fwd(const float* x, int x_ld, float* out, int n_groups, int64_t pix_per_group,
                         int C, void* stream);
int dl4ds_group_mean_bwd(const float* dout, float* dx, int dx_ld, int n_groups,
                         int64_t pix_per_group, int C, void* stream);
/* out[p,c] = a[p,c] * b[p,c]  (Dropout mask multiply, discriminator.py:77). */
int dl4ds_mul(const float* a, const float* b, float* out, int64_t n, void* stream);
/* BinaryCrossentropy(from_logits=False) -- cgan.py:546-552,567-571 (Keras eps 1e-7 clipping).
 * loss_out[0] += scale * bce(target, p);  dp (+)= scale * d bce / d p  (accumulate flag). */
int dl4ds_bce_loss(const float* p, float target, float* loss_out, float* dp, int64_t n,
                   float scale, int accumulate, void* stream);
/* dst[b][a][:] = src[a][b][:] over frames of `frame_elems` floats: (B,T,...) <-> (T,B,...) layout
 * change around the ConvLSTM recurrence (spt_postups.py:96-163 keeps NTHWC; the recurrence here runs
 * time-major so each step is one dense tensor). */
int dl4ds_permute_frames(const float* src, float* dst, int A, int B, int64_t frame_elems, void* stream);
/* ZeroPadding2D(((0,dh),(0,dw))) of PadConcat -- blocks.py:639-655: dst (N,Hd,Wd,C) <- src (N,Hs,Ws,C)
 * zero-filled where src has no pixel.  With Hd<=Hs, Wd<=Ws it is the adjoint crop. */
int dl4ds_pad_bottom_right(const float* src, int src_ld, float* dst, int dst_ld,
                           int N, int Hs, int Ws, int Hd, int Wd, int C, void* stream);
/* y = a*x + b*y element-wise (gradient combination for the two-seed cGAN backward). */
int dl4ds_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel exchange (one process per GPU; replaces Horovod: hvd.DistributedOptimizer /
 * DistributedGradientTape average, training/supervised.py:363-365, training/cgan.py:608-611, and
 * hvd.broadcast_variables / BroadcastGlobalVariablesCallback(0), supervised.py:369, cgan.py:626-637).
 * NCCL over NVLink / NVSwitch, resolved with dlopen at the first call (libnccl.so.2; DL4DS_NCCL_LIB
 * overrides).  Rank 0 creates the 128-byte id, the caller's launcher distributes it (any channel), every
 * rank calls init_rank.  Collectives are asynchronous on `stream` and may be captured into a CUDA graph.
 * The gradient AVERAGE of Horovod is allreduce_sum here + grad_scale = 1/world in dl4ds_adam_step*.
 * ------------------------------------------------------------------------------------------- */
int dl4ds_comm_unique_id_bytes(void);                       /* 128 */
int dl4ds_comm_nccl_version(void);                          /* e.g. 22809; -1 if NCCL cannot be loaded */
int dl4ds_comm_get_unique_id(void* id_out_128);
int dl4ds_comm_init_rank(const void* id_128, int nranks, int rank);
int dl4ds_comm_size(void);                                  /* 0 without a communicator */
int dl4ds_comm_rank(void);                                  /* -1 without a communicator */
int dl4ds_comm_allreduce_sum(float* buf, int64_t n, void* stream);             /* in place */
int dl4ds_comm_broadcast(void* buf, int64_t nbytes, int root, void* stream);   /* in place */
int dl4ds_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* DL4DS_B200_H */
