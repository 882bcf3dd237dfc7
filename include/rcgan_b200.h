/* rcgan_b200 -- C ABI of the B200-native RCGAN training hot path.
 *
 * One shared object, `librcgan_b200.so`, `extern "C"`, plain pointers and sizes.
 * The reference (tkkiran/Robust-Conditional-GAN) is pure Python over TensorFlow 1.5
 * and has no FFI of its own: the "interface each entry replaces" is the TensorFlow
 * op the reference's wrapper calls, cited per entry as reference file:line.
 *
 * Conventions:
 *   - every entry returns 0 on success or a negative rcgan_status; the message of the
 *     last failure on the calling thread is available from rcgan_last_error();
 *   - never allocates device memory, never synchronises, never throws; all pointers are
 *     DEVICE pointers to caller-owned buffers unless the name says host;
 *   - kernels are enqueued on the caller's stream (`void* stream` is a cudaStream_t),
 *     so they can be captured into a CUDA graph by the caller;
 *   - activations are NHWC ("rows x channels", channels fastest) with an explicit
 *     channel stride `ld` (elements per pixel, >= channels) so padded concat buffers
 *     can be consumed in place;  `dtype` selects the activation storage type;
 *   - parameters, gradients of parameters, statistics and optimizer state are fp32.
 */
#ifndef RCGAN_B200_H
#define RCGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum rcgan_status { RCGAN_OK = 0, RCGAN_EBADSHAPE = -1, RCGAN_EUNSUPPORTED = -2, RCGAN_ECUDA = -3 };
enum rcgan_dtype { RCGAN_F32 = 0, RCGAN_BF16 = 1 };
enum rcgan_act { RCGAN_ACT_NONE = 0, RCGAN_ACT_RELU = 1, RCGAN_ACT_LRELU = 2, RCGAN_ACT_SIGMOID = 3, RCGAN_ACT_TANH = 4 };
/* GAN loss link functions: mnist/model.py:135-147, cifar10/gan_resnet.py:604-606,751 */
enum rcgan_loss_mode {
  RCGAN_HINGE_D_REAL = 0, /* relu(1-l) */ RCGAN_HINGE_D_FAKE = 1, /* relu(1+l) */ RCGAN_HINGE_G = 2, /* -l */
  RCGAN_CE_D_REAL = 3, /* sCE(l,1) */ RCGAN_CE_D_FAKE = 4, /* sCE(l,0) */ RCGAN_CE_G = 5 /* sCE(l,1) */
};

/* bumped on every signature change; robust_conditional_gan_b200/_C.py refuses a library whose version differs */
#define RCGAN_ABI_VERSION 11
const char* rcgan_last_error(void);
int rcgan_abi_version(void);
/* name of the kernel variant the last conv entry point (fprop / dgrad / wgrad / upconv) launched on the calling thread,
 * e.g. "conv_tc_persist<256,1,3,bf16,multi=0>", "conv_tc<128,3,im2col=1>", "wgrad_tc<128,im2col=1>", "conv_simt".
 * Test hook: the parity tests assert which instantiation a shape dispatches to. */
const char* rcgan_last_conv_variant(void);
/* every distinct variant launched by any thread since the last reset, ';'-separated (valid until the calling thread's next
 * call); reset != 0 clears the log afterwards.  Lets a test of a whole training step assert which kernels it ran. */
const char* rcgan_conv_variant_log(int reset);
/* number of kernels this library has launched in this process (captured graph replays are not re-counted) */
long rcgan_launch_count(void);
/* 1 when the library was built for sm_100a and the current device is compute capability 10.x */
int rcgan_device_ok(void);

/* ---------------------------------------------------------------- convolution family
 * Replaces tf.nn.conv2d (mnist/ops.py:62, cifar10/common/ops/conv2d.py:181-187) and its
 * autodiff dgrad/wgrad; tf.nn.conv2d_transpose (mnist/ops.py:78) is the dgrad entry used
 * as a forward op; tf.matmul linears (mnist/ops.py:114-116, cifar10/common/ops/linear.py:
 * 163-173) are the 1x1, h=w=1 case.  Filter layout HWIO [kh,kw,cin,cout] fp32. */
typedef struct {
  int n, h, w, cin;     /* conv input  x  [n,h,w,cin]   */
  int ho, wo, cout;     /* conv output y  [n,ho,wo,cout] */
  int kh, kw, stride;   /* square stride */
  int pad_t, pad_l;     /* TF SAME pad_before (top, left) */
  int ldx, ldy;         /* channel strides of the x-side and y-side buffers (elements) */
  int dtype;            /* rcgan_dtype of x/y/dx/dy */
} rcgan_conv_desc;

/* bytes of the bf16 tensor-core weight pack for this shape, 0 if the shape runs on the
 * CUDA-core path (tiny channel counts: cin or cout not a multiple of 8, K < 64 ...). */
size_t rcgan_conv_wpack_bytes(const rcgan_conv_desc* d);
/* 1 when the given direction (0 fprop, 1 dgrad, 2 wgrad) of this shape runs on the tcgen05 path when a pack is passed */
int rcgan_conv_uses_tensor_cores(const rcgan_conv_desc* d, int direction);
/* w_f32 (optionally scaled by *scale_dev, e.g. 1/sigma) -> bf16 pack (both GEMM layouts) */
int rcgan_conv_wpack(const rcgan_conv_desc* d, const float* w, const float* scale_dev, void* pack, void* stream);
/* the same for `count` weights in one launch (all packs of a training step are refreshed at its start) */
int rcgan_conv_wpack_batched(int count, const rcgan_conv_desc* const* descs, const float* const* w, void* const* packs, void* stream);

/* y = act(conv(x, w) + bias)        bias may be NULL.  out_dtype: storage type of y -- d->dtype, or RCGAN_F32
 * for a bf16 conv whose output feeds a batch norm (kept fp32: the norm's backward cancels catastrophically on
 * bf16-rounded inputs). */
int rcgan_conv2d_fprop(const rcgan_conv_desc* d, const void* x, const float* w, const void* wpack,
                       const float* bias, void* y, int out_dtype, int act, float leak, void* stream);
/* UpsampleConv forward (cifar10/gan_resnet.py:259-272: nearest-neighbour 2x upsample, then 3x3 SAME conv) WITHOUT the
 * upsampled tensor: each of the 4 output parity classes is a 2x2-tap conv of the small input with pre-summed filter taps
 * (4/9 of the flops).  d describes the 3x3 conv at the OUTPUT resolution (h, w even; cin, ldx multiples of 8); x_small is
 * [n, h/2, w/2, cin] with channel stride d->ldx.  rcgan_upconv2d_fold builds the folded fp32 filter (16*cin*cout floats)
 * and its bf16 pack (rcgan_upconv2d_pack_bytes; 0 = unsupported shape) -- once per weight update.  Forward only: the
 * reference needs no gradient through the generator in its discriminator step, where this runs. */
size_t rcgan_upconv2d_pack_bytes(const rcgan_conv_desc* d);
int rcgan_upconv2d_fold(const rcgan_conv_desc* d, const float* w, float* wfold, void* pack, void* stream);
int rcgan_upconv2d_fprop(const rcgan_conv_desc* d, const void* x_small, const void* pack, const float* bias, void* y,
                         int out_dtype, int act, float leak, void* stream);
/* y = act(conv2d(x, w) + bias) + res: the ResidualBlock's `shortcut + output` (cifar10/gan_resnet.py:328) fused into the
 * producing conv's epilogue; res has exactly y's layout and dtype.  (bf16: conv result rounded, then bf16 + bf16 in fp32,
 * rounded -- bit-identical to running rcgan_conv2d_fprop followed by rcgan_add.) */
int rcgan_conv2d_fprop_res(const rcgan_conv_desc* d, const void* x, const float* w, const void* wpack, const float* bias,
                           const void* res, void* y, int out_dtype, int act, float leak, void* stream);
/* dx (=|+=) act(conv_dgrad(dy, w) + bias): the backward-data op; with bias/act it is the
 * forward of deconv2d (mnist/ops.py:69-92).  bias may be NULL. */
int rcgan_conv2d_dgrad(const rcgan_conv_desc* d, const void* dy, const float* w, const void* wpack,
                       const float* bias, void* dx, int out_dtype, int act, float leak, int accumulate, void* stream);
/* Epilogue fusions of the tensor-core conv kernels: element-wise passes of the reference graph that touch exactly the tile
 * the conv is writing, applied in this order to the bf16 value act(conv + bias) (every pointer may be NULL):
 *   mask:  value *= act'(mask) -- the backward of a relu / lrelu whose forward OUTPUT is `mask` (laid out like the output):
 *          the `nonlinearity` in front of a conv (cifar10/gan_resnet.py:199-205, 318-325) differentiated inside the dgrad that
 *          produces its input gradient, instead of a separate pass over the gradient;
 *   res:   value += res -- ResidualBlock's `shortcut + output` (:328); res_up = 1 reads res [n, OH/2, OW/2, N] (channel stride
 *          ld_res) through a nearest-neighbour 2x upsampling, i.e. UpsampleConv_1x1's shortcut (:259-272, 305-309) is added
 *          without materialising the upsampled tensor.  Not combinable with accumulate;
 *   out2:  a second output relu(value) with the output's layout (the next block's `nonlinearity(inputs)`, :318).
 * Each step rounds to bf16 exactly where the separate kernels did, so results are bit-identical to the unfused sequence.
 * bf16 outputs on the tensor-core path only; the calls fail with RCGAN_EUNSUPPORTED otherwise (no fallback). */
typedef struct {
  const void* mask; int mask_act; float mask_leak;
  const void* res; int res_up; int ld_res;
  void* out2; int out2_act;
  float* colstats;
} rcgan_conv_epilogue;
/* colstats: the batch-norm statistics pass of the layer that FOLLOWS this conv (normalization.py:38-41 moments over [0,1,2]),
 * taken from the values as they are stored: every CTA of the persistent kernel writes its partial per-column sum and sum of
 * squares, [RCGAN_COLSTATS_PARTS][2][N] floats followed by [RCGAN_COLSTATS_PARTS] row counts (rcgan_colstats_floats(N) in
 * all); rcgan_bn_fwd_prestats merges them (Chan) instead of re-reading the tensor.  N in {64, 128, 256}, dense rows. */
#define RCGAN_COLSTATS_PARTS 148
size_t rcgan_colstats_floats(int n);
int rcgan_conv2d_fprop_ex(const rcgan_conv_desc* d, const void* x, const void* wpack, const float* bias, void* y, int out_dtype,
                          int act, float leak, const rcgan_conv_epilogue* ep, void* stream);
int rcgan_conv2d_dgrad_ex(const rcgan_conv_desc* d, const void* dy, const void* wpack, const float* bias, void* dx, int out_dtype,
                          int act, float leak, int accumulate, const rcgan_conv_epilogue* ep, void* stream);
/* dw (=|+=) x^T * dy ;  ws: caller workspace of rcgan_conv2d_wgrad_workspace() bytes */
size_t rcgan_conv2d_wgrad_workspace(const rcgan_conv_desc* d);
int rcgan_conv2d_wgrad(const rcgan_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate,
                       void* ws, size_t ws_bytes, void* stream);

/* Patch matrix of a conv whose input has very few channels (cin <= 4: d_h0_conv, D.Block.1.*, g_h3's transposed conv):
 * P[m, (ky*kw + kx)*cin + ci] = x[n, oy*s - pad_t + ky, ox*s - pad_l + kx, ci] (0 in the halo and for columns >= kh*kw*cin),
 * m = (n, oy, ox), row stride ldp.  The conv then runs as a dense GEMM on the tensor cores (1x1 desc over P) instead of a
 * K = 25..27 implicit GEMM on CUDA cores.  Same dtype as d->dtype. */
int rcgan_im2col(const rcgan_conv_desc* d, const void* x, void* patches, int ldp, void* stream);
/* Filter of the transposed conv: out[kh-1-ky][kw-1-kx][co][ci] (=|+=) w[ky][kx][ci][co], fp32.  The input gradient of a
 * stride-1 conv with very few OUTPUT channels (G.Output, cifar10/gan_resnet.py:405-407: 256 -> 3) is the conv of dL/dy with
 * this filter, so its backward runs as patch-matrix GEMMs like the few-input-channel convs above; applied to the GEMM's
 * filter gradient (cin/cout swapped) the same call maps it back. */
int rcgan_wflip(const float* w, float* out, int kh, int kw, int cin, int cout, int accumulate, void* stream);
/* `count` fp32 copies dst[i][0..numel[i]) = src[i][...] in one launch: the u <- u_new assignments of all spectral norms after a
 * step (mnist/sn.py:62-71, update_collection=None) */
int rcgan_copy_batched(int count, const float* const* src, float* const* dst, const long* numel, void* stream);
/* 3x3 filter folded with a 2x resampling into a 4x4 stride-2 filter (both maps linear; SURVEY section 7: "ConvMeanPool == 4x4-s2
 * conv and Upsample+3x3 == four 2x2 sub-pixel convs are algorithmic flop reductions, legal for results parity").
 *   mode 0: ConvMeanPool (cifar10/gan_resnet.py:231-241): meanpool2(conv3x3_SAME(x, w)) == conv4x4_stride2_SAME(x, w4), w4 HWIO
 *   mode 1: UpsampleConv (cifar10/gan_resnet.py:259-272): conv3x3_SAME(upsample2(x), w) == conv2d_transpose_4x4_stride2(x, w4),
 *           w4 in conv2d_transpose layout [4,4,cout,cin]
 * w, dw: [3,3,cin,cout] fp32; w4, dw4: 16*cin*cout fp32.  rcgan_wfold4_bwd is the adjoint: dw (=|+=) fold^T(dw4). */
int rcgan_wfold4(const float* w, float* w4, int cin, int cout, int mode, void* stream);
int rcgan_wfold4_bwd(const float* dw4, float* dw, int cin, int cout, int mode, int accumulate, void* stream);
/* Adjoint of rcgan_im2col: x[n, iy, ix, ci] (=|+=) act(bias[ci] + sum over taps (ky,kx) with (iy + pad_t - ky) % s == 0 ... of
 * T[(n, oy, ox), (ky*kw + kx)*cin + ci]), T fp32 with row stride ldt.  With T = dy[M, cout] * W^T[cout, kh*kw*cin] (one dense
 * GEMM that reads dy once) this is the input gradient of a conv with cin <= 4 -- i.e. the FORWARD of g_h3's
 * conv2d_transpose to the 1-channel image (mnist/model.py:726-728, ops.py:69-92) and the gradient d_h0_conv sends back into the
 * generated image (mnist/model.py:678) -- instead of a per-pixel gather that re-reads dy once per filter tap. */
int rcgan_col2im(const rcgan_conv_desc* d, const float* T, int ldt, const float* bias, void* x, int out_dtype, int act,
                 float leak, int accumulate, void* stream);

/* ---------------------------------------------------------------- rows x channels helpers */
/* db[c] (=|+=) sum_r dy[r,c]   (bias gradients of conv / deconv / linear) */
int rcgan_colsum(const void* dy, int rows, int c, int ld, int dtype, float* db, int accumulate, void* stream);
/* y = act(x + bias) and dx = dy * act'(y)  (standalone activations: lrelu/relu/sigmoid/tanh) */
int rcgan_bias_act_fwd(const void* x, const float* bias, void* y, long rows, int c, int ldx, int ldy, int dtype,
                       int act, float leak, void* stream);
int rcgan_act_bwd(const void* dy, const void* y, void* dx, long rows, int c, int ld_dy, int ld_y, int ld_dx,
                  int dtype, int act, float leak, int accumulate, void* stream);
/* rcgan_act_bwd in place followed by rcgan_colsum of the result, as ONE pass over dy: dy <- dy * act'(y); db (=|+=) column sums of
 * the new dy -- the backward of a conv with a fused activation and a bias (mnist/ops.py:62-66 + lrelu, gan_resnet.py:318-325)
 * starts with exactly these two steps.  Same roundings as the two kernels. */
int rcgan_act_bwd_colsum(void* dy, const void* y, int rows, int c, int ld_dy, int ld_y, int dtype, int act, float leak, float* db,
                         int accumulate, void* stream);
/* out[r, 0:c1] = a[r,:], out[r, c1:c1+c2] = yb[r / rows_per_sample, :], out[r, c1+c2:ldo] = 0
 * (ops.conv_cond_concat mnist/ops.py:46-51 and the z|y, h|y concats of mnist/model.py:714-728);
 * yb is fp32 [samples, c2].  Backward: da (=|+=) dout[:, 0:c1]. */
int rcgan_concat_label_fwd(const void* a, int lda, const float* yb, void* out, int ldo, long rows, int rows_per_sample,
                           int c1, int c2, int dtype, void* stream);
int rcgan_slice_bwd(const void* dout, int ldo, void* da, int lda, long rows, int c1, int dtype, int accumulate,
                    void* stream);
/* y[s,c] = mean over hw of x[s,hw,c]  (tf.reduce_mean(h3, axis=(1,2)) mnist/model.py:678; gan_resnet.py:407),
 * with optional relu applied to x first (gan_resnet.py:405).  Backward broadcasts dy/hw (times relu mask). */
int rcgan_meanhw_fwd(const void* x, void* y, int samples, int hw, int c, int dtype, int relu, void* stream);
int rcgan_meanhw_bwd(const void* dy, const void* x, void* dx, int samples, int hw, int c, int dtype, int relu,
                     int accumulate, void* stream);
/* 2x2 mean pool / nearest-neighbour 2x upsample and residual add (cifar10/gan_resnet.py:231-272, 328) */
int rcgan_avgpool2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream);
int rcgan_avgpool2_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dtype, int accumulate, void* stream);
int rcgan_upsample2_fwd(const void* x, void* y, int n, int h, int w, int c, int dtype, void* stream);
int rcgan_upsample2_bwd(const void* dy, void* dx, int n, int h, int w, int c, int dtype, int accumulate, void* stream);
int rcgan_add(const void* a, const void* b, void* out, long numel, int dtype, void* stream);
/* dst (=|+=) src */
int rcgan_copy_acc(const void* src, void* dst, long numel, int dtype, int accumulate, void* stream);
/* fp32 -> activation dtype cast (feeding inputs), and back */
int rcgan_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long numel, void* stream);

/* ---------------------------------------------------------------- (conditional) batch norm
 * tf.contrib.layers.batch_norm (mnist/ops.py:30-44) when labels == NULL (tables have one row);
 * cond_batchnorm (cifar10/common/ops/normalization.py:27-59) when labels != NULL.
 * x is [samples*hw, c] of type xdtype; y, dy and dx are of type ydtype (xdtype f32 with ydtype bf16 is the
 * bf16-mode layout); stats are over all rows, biased variance, eps inside rsqrt.
 * save[2*c] receives (mean, invstd).  moving_mean/var (may be NULL) are updated with `decay`
 * (Bessel-corrected variance).  train == 0 normalises with the moving statistics instead. */
size_t rcgan_bn_workspace(int samples, int hw, int c);
int rcgan_bn_fwd(const void* x, void* y, int samples, int hw, int c, int xdtype, int ydtype, const float* scale, const float* offset,
                 const int* labels, float eps, int act, float leak, int train, float decay, float* moving_mean,
                 float* moving_var, float* save, void* ws, size_t ws_bytes, void* stream);
/* rcgan_bn_fwd in training mode with the batch statistics taken from a producing conv's epilogue partials (see
 * rcgan_conv_epilogue.colstats): 1 read + 1 write of the activation instead of 2 reads + 1 write */
int rcgan_bn_fwd_prestats(const void* x, void* y, int samples, int hw, int c, int xdtype, int ydtype, const float* scale,
                          const float* offset, const int* labels, float eps, int act, float leak, float decay, float* moving_mean,
                          float* moving_var, float* save, const float* colstats, void* stream);
/* dx (=|+=) ; dscale/doffset [n_labels, c] (=|+=).  y is the activated forward output; with `offset` (the forward's table)
 * given and act relu / lrelu, the activation mask is re-derived from x as the sign of the forward pre-activation and y is not
 * read (may be NULL): 5 instead of 7 tensor passes. */
int rcgan_bn_bwd(const void* dy, const void* x, const void* y, void* dx, int samples, int hw, int c, int xdtype, int ydtype,
                 const float* scale, const int* labels, int n_labels, const float* save, int act, float leak,
                 float* dscale, float* doffset, int accumulate_dx, int accumulate_param, void* ws, size_t ws_bytes,
                 const float* offset, void* stream);
/* Batch norm with the label concat that follows it in the MNIST generator fused in (mnist/model.py:714-728: h = relu(bn(.)),
 * then concat([h, y]) / conv_cond_concat(h, yb)): y has row stride ldy >= c + c2 (a multiple of 8), its channels [c, c + c2)
 * receive yb[sample, :] (fp32 [samples, c2]), the rest of the padding is left untouched; the backward reads dy / y with the
 * same stride and ignores the label channels.  One pass less over the activation in each direction. */
int rcgan_bn_fwd_cat(const void* x, void* y, int ldy, const float* yb, int c2, int samples, int hw, int c, int xdtype, int ydtype,
                     const float* scale, const float* offset, const int* labels, float eps, int act, float leak, int train,
                     float decay, float* moving_mean, float* moving_var, float* save, void* ws, size_t ws_bytes, void* stream);
int rcgan_bn_bwd_cat(const void* dy, const void* x, const void* y, int ldy, void* dx, int samples, int hw, int c, int xdtype,
                     int ydtype, const float* scale, const int* labels, int n_labels, const float* save, int act, float leak,
                     float* dscale, float* doffset, int accumulate_dx, int accumulate_param, void* ws, size_t ws_bytes,
                     const float* offset, void* stream);

/* ---------------------------------------------------------------- spectral norm
 * spectral_normed_weight (mnist/sn.py:17-75 == cifar10/common/ops/sn.py), one power iteration.
 * W [m,c] fp32, u [c].  Outputs: w_bar [m,c] fp32 (may be NULL), u_new [c], and `save`
 * (rcgan_sn_save_floats(m,c) floats: sigma, 1/sigma, na, n, a[m], b[c], t[m] ...) for backward.
 * The gradient flows THROUGH the power iteration. */
size_t rcgan_sn_save_floats(int m, int c);
size_t rcgan_sn_workspace(int m, int c);
int rcgan_sn_fwd(const float* W, const float* u, int m, int c, float* w_bar, float* u_new, float* save, void* ws,
                 size_t ws_bytes, void* stream);
/* dW (=|+=) from G = dL/dW_bar */
int rcgan_sn_bwd(const float* W, const float* u, const float* G, int m, int c, const float* save, float* dW,
                 int accumulate, void* ws, size_t ws_bytes, void* stream);
/* Every spectrally-normalised weight of a step in one launch per pass (host arrays of `count` device pointers / sizes;
 * same per-item semantics as rcgan_sn_fwd / rcgan_sn_bwd).  The reference normalises each weight where its layer is built
 * (cifar10/common/ops/conv2d.py:181-216, linear.py:161-180); the results only depend on the parameters, so a step computes
 * them all up front: 16 weights x 6 small kernels become 6 launches. */
size_t rcgan_sn_workspace_batched(int count, const int* m, const int* c);
int rcgan_sn_fwd_batched(int count, const float* const* W, const float* const* u, const int* m, const int* c,
                         float* const* w_bar, float* const* u_new, float* const* save, void* ws, size_t ws_bytes, void* stream);
int rcgan_sn_bwd_batched(int count, const float* const* W, const float* const* u, const float* const* G, const int* m,
                         const int* c, float* const* save, float* const* dW, const int* accumulate, void* ws, size_t ws_bytes,
                         void* stream);

/* ---------------------------------------------------------------- losses
 * Unified noisy-channel projection loss (mnist/model.py:150-207,679-686;
 * cifar10/gan_resnet.py:588-606,649-685,751-760):
 *   l[b,j] = psi[b] + <h[b,:], V[j,:]>,   L = scale * sum_b sum_j wgt[b,j] * phi_mode(l[b,j])
 * h [B,d] (activation dtype, ld = d), psi [B] fp32, V [k,d] fp32, wgt [B,k] fp32.
 * Outputs (any may be NULL): loss_acc[0] += L; logits [B,k] fp32; dh (=|+=) [B,d] activation dtype;
 * dpsi [B] fp32 (=); dV [k,d] fp32 (+=); dwgt [B,k] fp32 (=). */
int rcgan_channel_loss(const void* h, const float* psi, const float* V, const float* wgt, int B, int d, int k,
                       int dtype, int mode, float scale, float* loss_acc, float* logits, void* dh, int accumulate_dh,
                       float* dpsi, float* dV, float* dwgt, void* stream);
/* mean sigmoid cross entropy (perm regulariser, mnist/model.py:214-224; gan_resnet.py:486-490,687-695):
 * loss_acc[0] += scale * sum sCE(logits, targets);  dlogits = scale * (sigmoid(l) - t).  fp32 [B,k]. */
int rcgan_sigmoid_ce(const float* logits, const float* targets, long numel, float scale, float* loss_acc,
                     float* dlogits, void* stream);
/* Scalar-logit GAN loss (vanilla discriminator, mnist/model.py:687-703 + :135-147):
 * loss_acc[0] += scale * sum_b phi_mode(l[b]);  dl[b] = scale * phi'(l[b]).  fp32 [B]. */
int rcgan_logit_loss(const float* logits, long B, int mode, float scale, float* loss_acc, float* dlogits, void* stream);
/* Row softmax of the confusion logits and its backward (mnist/model.py:106; gan_resnet.py:522):
 * C = softmax(L); dL = C * (dC - <dC, C>_row).  [k,k] fp32. */
int rcgan_softmax_rows_fwd(const float* logits, float* C, int rows, int k, void* stream);
int rcgan_softmax_rows_bwd(const float* C, const float* dC, float* dlogits, int rows, int k, int accumulate,
                           void* stream);
/* wgt[b,:] = C[y[b],:] (tensordot(onehot(y), C)); backward dC[y[b],:] += dwgt[b,:] (dC zeroed by the call when !accumulate) */
int rcgan_gather_rows_fwd(const float* C, const int* y, float* wgt, int B, int k, void* stream);
int rcgan_gather_rows_bwd(const float* dwgt, const int* y, float* dC, int B, int k, int rows, int accumulate, void* stream);

/* ---------------------------------------------------------------- optimiser
 * tf.train.AdamOptimizer (mnist/model.py:250-262; gan_resnet.py:802-817) over one flat arena:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t * m / (sqrt(v) + eps),
 * lr_t = lr * sqrt(1-b2^t)/(1-b1^t) is computed by the caller; when lr_t_dev != NULL it is read from that
 * device float instead (so a captured CUDA graph can be replayed with a new step count).  grad_scale multiplies g first
 * (1/world_size after an allreduce-sum).  Elements in [clip_lo[i], clip_hi[i]) (n_clip <= 8 ranges)
 * are clipped to [-1,1] after the update (max-norm variable constraint, mnist/ops.py:101-111). */
int rcgan_adam_tf(float* p, const float* g, float* m, float* v, long numel, float lr_t, const float* lr_t_dev, float b1,
                  float b2, float eps, float grad_scale, const long* clip_lo, const long* clip_hi, int n_clip,
                  void* stream);
/* ---------------------------------------------------------------- label recovery
 * DCGAN.recover_labels, mnist/model.py:494-640: gradient descent on z_recover [R*k, z_dim] and y_logit_recover [R, k] through
 * gen_sampler (the generator with batch norm in INFERENCE mode) against R real images.
 *
 * rcgan_bn_infer_bwd: backward of y = act(scale*(x - moving_mean)*rsqrt(moving_var + eps) + offset) with respect to x only
 *   (model.py:733-757 runs the norms with train=False): dx (=|+=) dy * act'(y) * scale[label] * save[c + ch].
 *   `save` is what rcgan_bn_fwd(train = 0) wrote; ws: 2*c floats; ldy: row stride of dy / y (c, or a fused concat's width).
 * rcgan_recover_mse (model.py:538-541): sq[r,j] = mean_p (actual[r,p] - sample[r*k + j, p])^2,
 *   loss_acc[0] += (1/R) sum_r sum_j sq[r,j] * y_rec[r,j];   dsample (=) (2/npix) (sample - actual) y_rec[r,j] / R;
 *   dyrec[r,j] (=) sq[r,j] / R.   fp32; any output may be NULL.
 * rcgan_sgd: tf.train.GradientDescentOptimizer (model.py:612-617): p -= lr * grad_scale * g. */
int rcgan_bn_infer_bwd(const void* dy, const void* y, int ldy, void* dx, int samples, int hw, int c, int dtype,
                       const float* scale, const int* labels, const float* save, int act, float leak, int accumulate_dx, void* ws,
                       size_t ws_bytes, void* stream);
int rcgan_recover_mse(const float* sample, const float* actual, const float* y_rec, int R, int k, int npix, float* loss_acc,
                      float* sq, float* dsample, float* dyrec, void* stream);
int rcgan_sgd(float* p, const float* g, long numel, float lr, float grad_scale, void* stream);

/* cudaMemsetAsync(ptr, 0, bytes) on the caller's stream (zeroing gradient arenas / loss slots inside a captured step) */
int rcgan_zero(void* ptr, size_t bytes, void* stream);

/* ---------------------------------------------------------------- label-noise sampler
 * numpy legacy RandomState stream on the device (mnist/model.py:795-834, :293-333;
 * cifar10/common/data/cifar10.py:29-38).  `state` is 625 uint32 (624 MT19937 words + position).
 * `table` (device, doubles) comes from rcgan_sampler_table_host(): per confusion-matrix row and
 * class index the binomial-inversion constants, computed on the HOST with libm exactly as numpy does. */
#define RCGAN_MT_STATE_WORDS 625
#define RCGAN_SAMPLER_TABLE_DOUBLES(k) ((k) * ((k) - 1) * 8)   /* (mode, qn, px1, p, T, T_safe, 0, 0) per (row, class) */
void rcgan_sampler_table_host(const double* C, int k, double* table);
int rcgan_mt_seed(uint32_t* state, uint32_t seed, void* stream);
/* Fisher-Yates permutation exactly as np.random.shuffle draws it; perm[n] int32 */
int rcgan_mt_shuffle_perm(uint32_t* state, int32_t* perm, int n, void* stream);
/* mnist: for each i: real[i] ~ C[y[i]], gen[i] = randint(k) (or real[i] if real_match), fake[i] ~ C[gen[i]] */
int rcgan_sample_labels_mnist(uint32_t* state, const double* table, int k, const int32_t* y, int n, int real_match,
                              int32_t* real, int32_t* gen, int32_t* fake, void* stream);
/* re-noising pass (mnist/model.py:329-333): real2[i] ~ C[real[i]], fake2[i] ~ C[fake[i]] interleaved */
int rcgan_sample_renoise_mnist(uint32_t* state, const double* table, int k, const int32_t* real, const int32_t* fake,
                               int n, int32_t* real2, int32_t* fake2, void* stream);
/* cifar: rnd = randint(k, size=n) first, then per i: labels[i] ~ C[labels[i]] (in place), biased[i] ~ C[rnd[i]] */
int rcgan_sample_labels_cifar(uint32_t* state, const double* table, int k, int32_t* labels, int n, int32_t* rnd,
                              int32_t* biased, void* stream);
/* out[i] = lo + (hi-lo)*random_double()  (np.random.uniform; batch_z mnist/model.py:342), stored as fp32 */
int rcgan_mt_uniform(uint32_t* state, float* out, long n, double lo, double hi, void* stream);
/* CIFAR real-data preprocessing (gan_resnet.py:548-552): int32 CHW [n,3072] in [0,255] ->
 * 2*(v/256 - .5) + noise[n,3072 NHWC-ordered after transpose] (noise may be NULL), NHWC, activation dtype */
int rcgan_preprocess_cifar(const int32_t* chw, const float* noise, void* out, int n, int dtype, void* stream);
/* the same from uint8 pixels: a quarter of the host -> device bytes of the int32 placeholder */
int rcgan_preprocess_cifar_u8(const uint8_t* chw, const float* noise, void* out, int n, int dtype, void* stream);
/* In-graph random inputs: out[i] = a + b * N(0,1) (normal != 0: tf.random_normal([n, 128]), gan_resnet.py:363-364) or U[a, b)
 * (tf.random_uniform(0, 1/128) dequantisation noise, :550).  Philox4x32-10 keyed by `seed`, counter = (element block, *step_dev,
 * stream_id): the step is read from device memory, so a captured CUDA graph draws new numbers on every replay. */
int rcgan_random_fill(float* out, long n, int normal, float a, float b, unsigned long long seed, const long* step_dev,
                      unsigned stream_id, void* stream);

#ifdef __cplusplus
}
#endif
#endif
