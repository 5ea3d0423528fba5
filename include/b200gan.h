/*
 * b200gan.h -- C ABI of libb200gan.so, the sm_100a StyleGAN2 operator library
 * behind gan-control's generator / discriminator hot path.
 *
 * The reference (amazon-science/gan-control) has no native boundary of its own:
 * `models/gan_model.py:19-50` is a source-level switch (`FUSED`) whose True
 * branch is a stub that would import `FusedLeakyReLU, fused_leaky_relu,
 * upfirdn2d` (and, upstream, `conv2d_gradfix`) from a CUDA-op package located at
 * `gan_control.models.op` (`trainers/non_leaking.py:6`).  This header is the
 * C-level contract that package binds: every entry point names the reference
 * computation it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - plain C types only; tensors are raw device pointers + explicit sizes.
 *  - activations are NHWC ("channels last"): x[n][h][w][c], c contiguous.
 *    An NCHW-contiguous tensor (N,C,H,W) is the NHWC tensor (N*C,H,W,1).
 *  - `dtype`: B200GAN_F32, B200GAN_BF16 or B200GAN_F16 storage; all arithmetic accumulates
 *    in fp32.  Reductions / weight gradients are always written as fp32.
 *  - every call is asynchronous on `stream` (a cudaStream_t), never allocates
 *    device memory, never synchronises, and is re-entrant.
 *  - return value: 0 on success, otherwise a negative B200GAN_E* code or a
 *    positive cudaError_t; `b200gan_last_error()` gives a thread-local message.
 */
#ifndef B200GAN_H_
#define B200GAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GAN_F32 0
#define B200GAN_BF16 1
#define B200GAN_F16 2    /* IEEE half storage: the tensor-core PARITY mode (same tcgen05 kind::f16 instruction as bf16, operand
                          * round-off 2^-11 instead of 2^-9; fp32 accumulation) */

#define B200GAN_EINVAL (-1)   /* bad argument (shape / dtype / alignment)   */
#define B200GAN_ENOSUP (-2)   /* configuration not supported by this build   */

int b200gan_version(void);
const char* b200gan_last_error(void);
/* number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`); monotonically increasing. */
uint64_t b200gan_launch_count(void);

/* ---- upfirdn2d ---------------------------------------------------------
 * Replaces `upfirdn2d(input, kernel, up, down, pad)` gm.py:45-50 ->
 * `upfirdn2d_native` pytorch_upfirdn2d.py:9-51.   Per channel:
 *   y[oy][ox] = sum_{ky,kx} z[oy*down + ky - pad0][ox*down + kx - pad0] * f[ky][kx]
 * where z is x zero-upsampled by `up` (sample i at z[i*up]) and zero outside,
 * f = kernel flipped in both axes when `flip_kernel` != 0 (the reference's
 * true convolution; the adjoint op passes flip_kernel = 0).  The output extent
 * (out_h, out_w) is explicit, so a trailing pad/crop is implicit.
 * kernel: fp32 [kh][kw] on the device, kh,kw <= 16.                        */
int b200gan_upfirdn2d(const void* x, void* y, const float* kernel, int dtype,
                      int n, int in_h, int in_w, int c, int out_h, int out_w,
                      int kh, int kw, int up, int down, int pad0_y, int pad0_x,
                      int flip_kernel, float gain, void* stream);

/* upfirdn2d with the StyledConv tail fused on the filter output (the upsampling
 * StyledConv: `Blur` gm.py:307, demodulation scale gm.py:288-289, NoiseInjection
 * gm.py:340-345, FusedLeakyReLU gm.py:32-35 -- four extra passes in the reference):
 *   y = act_gain * lrelu(fir(x) * rowscale[n][c] + noise_w * noise[n][oy][ox] + bias[c]).
 * bias / rowscale / noise may be NULL; `gain` still scales the taps.            */
int b200gan_upfirdn2d_act(const void* x, void* y, const float* kernel, int dtype,
                          int n, int in_h, int in_w, int c, int out_h, int out_w,
                          int kh, int kw, int up, int down, int pad0_y, int pad0_x,
                          int flip_kernel, float gain, const float* bias,
                          const float* rowscale, const void* noise, const float* noise_w,
                          float slope, float act_gain, void* stream);

/* ---- bias + noise + scaled leaky-ReLU -----------------------------------
 * Replaces `fused_leaky_relu` / `FusedLeakyReLU` gm.py:25-41, the noise add of
 * `NoiseInjection` gm.py:340-345 and the demodulation scale of gm.py:288-289
 * when the activation-scaled form is used:
 *   y = gain * lrelu_slope( x * rowscale[n][c] + noise_w * noise[n][hw] + bias[c] )
 * Element e of x has channel (e / inner) % c and sample e / (c*inner*outer?) --
 * layout is described by (n, hw, c, inner): NHWC -> inner = 1 (x[n][hw][c]);
 * NCHW -> inner = hw with `hw` passed as 1 (x[n][c][inner]).
 * Any of bias / rowscale / noise may be NULL.  slope = 1, gain = 1 is linear. */
int b200gan_bias_act_fwd(const void* x, void* y, const float* bias, const float* rowscale,
                         const void* noise, const float* noise_w, int dtype,
                         int64_t n, int64_t hw, int64_t c, int64_t inner,
                         float slope, float gain, void* stream);
/* gx = gy * gain * (y > 0 ? 1 : slope) [* rowscale]; `y` is the saved OUTPUT
 * (sign-preserving).  Linear in gy => it is its own
 * double-backward. */
int b200gan_bias_act_bwd(const void* gy, const void* y, void* gx, const float* rowscale, int dtype,
                         int64_t n, int64_t hw, int64_t c, int64_t inner,
                         float slope, float gain, void* stream);
/* Fused backward of the epilogue  y = gain*lrelu(z*rowscale + noise_w*noise + bias)  from the saved
 * OUTPUT y (NHWC): writes gconv = gy*gain*act'(y)*rowscale (the gradient entering the conv backward)
 * and accumulates gb[c] (bias), gd[n][c] (rowscale; z recovered from y), gnw[1] (noise strength).
 * rowscale / noise / bias / gd / gb / gnw may be NULL; fp32 outputs zero-initialised by the caller. */
int b200gan_epilogue_bwd(const void* gy, const void* y, void* gconv, const float* rowscale, const void* noise,
                         const float* noise_w, const float* bias, float* gd, float* gb, float* gnw, int dtype,
                         int64_t n, int64_t hw, int64_t c, float slope, float gain, void* stream);
/* Pixel reductions of  a[n][p][c] * b[n][p][c] * pixw[n][p]  (b, pixw may be NULL):
 * out_c[c] = sum over n,p (bias / noise-strength grads), out_nc[n][c] = sum over p
 * (demod-scale grads); either may be NULL.  NHWC; fp32 outputs must be zero-initialised by
 * the caller (atomic accumulation). */
int b200gan_reduce_nhwc(const void* a, const void* b, const void* pixw, float* out_c, float* out_nc,
                        int dtype, int64_t n, int64_t hw, int64_t c, void* stream);

/* ---- convolution family ----------------------------------------------------
 * One gather form covers every convolution on the path and every data
 * gradient of them:
 *   y[b][oy][ox][o] = sum_{ky,kx,i} z[b][oy*down + ky - pad0][ox*down + kx - pad0][i]
 *                                   * w[wb][ky][kx][o][i]
 * z = x zero-upsampled by `up` (extent (H-1)*up+1), zero outside; wb = b when
 * `w_per_sample`, else 0.
 *   up=1,down=1 : F.conv2d stride 1 (EqualConv2d gm.py:152-160, ModulatedConv2d
 *                 plain branch gm.py:326-329 with groups=batch => w_per_sample)
 *   up=1,down=2 : stride-2 conv after Blur (ConvLayer downsample gm.py:857-881)
 *   up=2,down=1 : F.conv_transpose2d stride 2 (gm.py:304) with the kernel
 *                 flipped/transposed by the caller, pad0 = k-1
 * weights: [wb][kh][kw][oc][ic] in `dtype`, ic contiguous ("K-major").
 * Optional fused epilogue (NULL = off), applied in this order:
 *   v = acc * rowscale[b][o] + noise_w * noise[b][oy][ox] + bias[o];
 *   y = gain * lrelu_slope(v)                                               */
int b200gan_conv_fwd(const void* x, const void* w, void* y, int dtype,
                     int b, int in_h, int in_w, int ic, int out_h, int out_w, int oc,
                     int kh, int kw, int up, int down, int pad0, int w_per_sample,
                     const float* bias, const float* rowscale, const void* noise, const float* noise_w,
                     float slope, float gain, void* stream);
/* The same convolution with the full epilogue description.  Beyond b200gan_conv_fwd's operands:
 *   addend: a tensor of the OUTPUT's shape and dtype added to the accumulator first -- the residual sum of ResBlock
 *           `(out + skip) / sqrt(2)` gm.py:920, ToRGB's `out + skip` gm.py:433, and in the backward pass the gradient that
 *           another consumer of the same tensor contributes (what autograd would add in a separate pass);
 *   gate:   a tensor of the OUTPUT's shape and dtype: BACKWARD mode.  The convolution is then a data gradient whose
 *           result is the gradient w.r.t. the output y of a bias + leaky-ReLU layer, and `gate` is that saved y:
 *               out = (acc + addend) * rowscale[b][o] * gain * (gate > 0 ? 1 : slope)
 *           i.e. the activation backward of the PRODUCER layer (`fused_leaky_relu` gm.py:39-41 from its output, SURVEY
 *           App. A.4) is applied here instead of in a pass of its own.  bias / noise are ignored in gate mode.
 * pack_in / pack_out as in b200gan_conv_fwd_packed (0 / 0 = plain). */
typedef struct {
    const float* bias;       /* [oc]      or NULL */
    const float* rowscale;   /* [b][oc]   or NULL */
    const void* noise;       /* [b][out pixels], `dtype`, or NULL */
    const float* noise_w;    /* [1]       or NULL */
    float slope, gain;
    const void* addend;      /* output-shaped, `dtype`, or NULL */
    const void* gate;        /* output-shaped, `dtype`, or NULL */
} b200gan_conv_epilogue;
int b200gan_conv_fwd_ex(const void* x, const void* w, void* y, int dtype,
                        int b, int in_h, int in_w, int ic, int out_h, int out_w, int oc,
                        int kh, int kw, int up, int down, int pad0, int w_per_sample, int pack_in, int pack_out,
                        const b200gan_conv_epilogue* epilogue, void* stream);
/* Engine selection for the convolution family: 0 (default) = tcgen05/TMEM implicit GEMM whenever the
 * shape qualifies (bf16, IC = 32 or a multiple of 8 >= 64, OC a multiple of 16, k <= 3x3), CUDA-core
 * gather kernel otherwise; 1 = CUDA-core kernel only; 2 = automatic but without the halo-reuse variant
 * (tests, A/B timing).  Returns the previous value. */
int b200gan_set_conv_engine(int engine);
/* Which engine served the convolution calls of this process: per-engine launch counters and the engine of the most
 * recent call (-1 before the first).  FWD_* count b200gan_conv_fwd / _packed (forward passes AND data gradients, which
 * are the same gather form), WGRAD_* count b200gan_conv_wgrad / _packed.  UMMA / HALO are the tcgen05 + TMEM + TMA
 * kernels; POINTWISE the streaming 1x1 kernels for a <= 4-channel side (ToRGB, from_rgb); SIMT the CUDA-core kernel
 * (fp32 parity engine, odd shapes).  Tests assert on these so that a silent fallback cannot pass as the tensor-core path. */
#define B200GAN_ENGINE_FWD_SIMT 0
#define B200GAN_ENGINE_FWD_POINTWISE 1
#define B200GAN_ENGINE_FWD_UMMA 2
#define B200GAN_ENGINE_FWD_HALO 3
#define B200GAN_ENGINE_WGRAD_SIMT 4
#define B200GAN_ENGINE_WGRAD_POINTWISE 5
#define B200GAN_ENGINE_WGRAD_UMMA 6
#define B200GAN_ENGINE_WGRAD_HALO 7
#define B200GAN_ENGINE_COUNT 8
uint64_t b200gan_engine_launches(int engine);
int b200gan_last_conv_engine(void);
/* Weight gradient of the form above:
 *   gw[wb][ky][kx][o][i] += sum_{b,oy,ox} gy[b][oy][ox][o] * z[b][oy*down+ky-pad0][..][i]
 * gw is fp32, same [wb][kh][kw][oc][ic] layout, must be zero-initialised
 * (split-K atomic accumulation). */
int b200gan_conv_wgrad(const void* x, const void* gy, float* gw, int dtype,
                       int b, int in_h, int in_w, int ic, int out_h, int out_w, int oc,
                       int kh, int kw, int up, int down, int pad0, int w_per_sample,
                       void* stream);

/* Convolution between space-to-depth VIEWS (up = down = 1).  Replaces, in ONE pass each,
 *   `conv_transpose2d(stride 2)` -> `Blur`            (upsampling ModulatedConv2d, gm.py:295-307): pack_out
 *   `Blur` -> `conv2d(stride 2)`                      (ConvLayer(downsample=True), gm.py:857-872): pack_in
 * and their data / weight gradients, with composite (FIR (*) conv) weights prepared by the caller:
 * the blurred / zero-inserted intermediate of the reference never exists.
 * pack_in : x is the plain NHWC tensor (b, 2*in_h, 2*in_w, ic/4); logical channel (py*2+px)*ic/4 + c
 *           of logical pixel (iy, ix) is physical pixel (2*iy+py, 2*ix+px), channel c.
 * pack_out: y is the plain NHWC tensor (b, 2*out_h, 2*out_w, oc/4), same correspondence; bias / rowscale are
 *           indexed by the physical channel, noise by the physical pixel.
 * in_h .. oc describe the LOGICAL (view) geometry; w is [wb][kh][kw][oc][ic] over logical channels. */
int b200gan_conv_fwd_packed(const void* x, const void* w, void* y, int dtype,
                            int b, int in_h, int in_w, int ic, int out_h, int out_w, int oc,
                            int kh, int kw, int pad0, int w_per_sample, int pack_in, int pack_out,
                            const float* bias, const float* rowscale, const void* noise, const float* noise_w,
                            float slope, float gain, void* stream);
/* Weight gradient of the packed form: pack_x / pack_gy say which operand is read through the view. */
int b200gan_conv_wgrad_packed(const void* x, const void* gy, float* gw, int dtype,
                              int b, int in_h, int in_w, int ic, int out_h, int out_w, int oc,
                              int kh, int kw, int pad0, int w_per_sample, int pack_x, int pack_gy,
                              void* stream);

/* ---- weight (de)modulation -------------------------------------------------
 * Replaces the weight path of `ModulatedConv2d.forward` gm.py:284-289 (`scale * weight * style`, `rsqrt(sum w^2 +
 * 1e-8)`, `weight * demod`: five broadcast / reduction passes over a (B,OC,IC,k,k) tensor) by ONE kernel that writes
 * the per-sample weights straight into the K-major operand layout of b200gan_conv_fwd:
 *   wk[b][tap'][o][i] = scale * weight[o][i][tap] * s[b][i] * d[b][o],   d[b][o] = rsqrt(sum_{i,tap} (scale*weight*s)^2 + 1e-8)
 * (d = 1 when !demodulate; tap' = kh*kw-1-tap when `flip`: the gather form of `conv_transpose2d`, gm.py:301-306).
 * weight: fp32 [oc][ic][kh][kw] (the parameter); s: fp32 [b][ic] (the modulation EqualLinear's output, gm.py:284);
 * d (may be NULL): fp32 [b][oc] written; wk: `dtype` [b][kh*kw][oc][ic]; wk_adjoint (may be NULL): `dtype`
 * [b][kh*kw][ic][oc] with the taps reversed relative to wk = the operand of the data-gradient convolution.  kh*kw <= 9. */
int b200gan_modweight_fwd(const float* weight, const float* s, float* d, void* wk, void* wk_adjoint, int dtype,
                          int b, int oc, int ic, int kh, int kw, float scale, int demodulate, int flip, void* stream);
/* First-order backward of the above given g = dL/dwk (fp32, wk's layout: the output of b200gan_conv_wgrad with
 * w_per_sample): gs[b][ic] (zero-initialised by the caller, atomic accumulation) and / or gweight[oc][ic][kh][kw]
 * (written).  d: the forward's output; e: fp32 [b][oc] scratch (demodulation only; written by the style pass and read by
 * the weight pass, so gweight with demodulation needs gs too). */
int b200gan_modweight_bwd(const float* g, const float* weight, const float* s, const float* d, float* gs, float* e,
                          float* gweight, int b, int oc, int ic, int kh, int kw, float scale, int demodulate, int flip,
                          void* stream);

/* ---- dense layers ------------------------------------------------------------
 * Replaces `EqualLinear.forward` gm.py:189-197 (`F.linear` + bias*lr_mul
 * [+ fused_leaky_relu]):  y[m][n] = act( scale * sum_k x[m][k] w[n][k] + bias[n]*bias_mul )
 * x, y in `dtype`; w, bias fp32 (parameters). act = 1 -> sqrt(2)*lrelu_0.2.   */
int b200gan_linear_fwd(const void* x, const float* w, const float* bias, void* y, int dtype,
                       int m, int n, int k, float scale, float bias_mul, int act, void* stream);
/* Generic small fp32-accumulating GEMM used by the linear backward passes:
 *   c[m][n] (+)= alpha * sum_k A(m,k) * B(k,n);  A(m,k) = a[m*lda + k] or a[k*lda + m] (trans_a),
 *   B(k,n) = b[n*ldb + k] (trans_b = 1, "NT") or b[k*ldb + n].  All fp32.       */
int b200gan_gemm_f32(const float* a, const float* b, float* c, int m, int n, int k,
                     int lda, int ldb, int ldc, int trans_a, int trans_b, float alpha, float beta,
                     void* stream);

/* Whole mapping network in one persistent (cooperative) kernel:
 * `Generator.style` gm.py:633-642 (vanilla) and `MultiFcStack` gm.py:489-502
 * (block-diagonal split-FC; PixelNorm gm.py:52-57 per slice), and the controller
 * `FcStack` controller_model.py:24-43 (normalize = 0).
 * `layers` is a DEVICE array of n_layers * n_groups descriptors, layer-major
 * (layers[l * n_groups + g]).  `acts` is a caller-provided fp32 buffer
 * [n_layers + 1][batch][row_width]: row 0 receives the (normalised) input, row l
 * the output of layer l; the last row is the w latent.  It doubles as the saved
 * activations of the backward pass.
 * `normalize` is a flag word: bit 0 = PixelNorm the input slices, bit 1 = LINEAR layers (no fused leaky-ReLU): with
 * bit 1 set and n_layers = 1 the call is a batch of independent EqualLinear layers without activation -- the style
 * modulations `ModulatedConv2d.modulation` of ALL generator layers (gm.py:245, 284) in one launch, each group reading its
 * own 512-wide slice of the (B, n_latent * 512) W+ tensor.                                                        */
typedef struct {
    const float* w;      /* [out_dim][in_dim] fp32 */
    const float* bias;   /* [out_dim] fp32 */
    int in_dim, out_dim;
    int in_off, out_off; /* column offsets into the input / output activation rows */
    float scale;         /* lr_mul / sqrt(in_dim) */
    float bias_mul;      /* lr_mul */
} b200gan_fc_layer;
int b200gan_mapping_fwd(const float* z, float* acts, const b200gan_fc_layer* layers,
                        int n_groups, int n_layers, int batch, int z_dim, int row_width,
                        int normalize, void* stream);
/* First-order backward of the whole mapping network in one cooperative kernel (autograd through gm.py:633-642 / 489-502
 * issues ~6 launches per EqualLinear).  acts: the forward's buffer.  gbuf: fp32 scratch [2][batch][row_width]; the caller
 * stores dL/d(acts[n_layers]) in gbuf[0] (it is overwritten).  grads: DEVICE array parallel to `layers` naming where each
 * layer's weight / bias gradient is WRITTEN (NULL = skip).  dz (may be NULL): [batch][z_dim] gradient of the input. */
typedef struct {
    float* gw;           /* [out_dim][in_dim] */
    float* gb;           /* [out_dim] */
} b200gan_fc_layer_grad;
int b200gan_mapping_bwd(const float* z, const float* acts, float* gbuf, const b200gan_fc_layer* layers,
                        const b200gan_fc_layer_grad* grads, float* dz, int n_groups, int n_layers, int batch,
                        int z_dim, int row_width, int normalize, void* stream);

/* ---- ADA augmentation: geometric warp + colour transform -------------------------
 * Replaces, in `trainers/non_leaking.py`, the sampling-grid chain `make_grid` / `affine_grid` / rescale and
 * `F.grid_sample(img_2x, grid, mode="bilinear", align_corners=False, padding_mode="zeros")` of `random_apply_affine`
 * (:341-357) together with `apply_color` (:373-383), in one pass and without a grid tensor:
 *   (sx, sy) = (m0*ox + m1*oy + m2,  m3*ox + m4*oy + m5)          mat: DEVICE double [n][6], source PIXEL coordinates
 *   s[ch]    = bilinear sample of x[b][ch] at (sx, sy), zero outside the image
 *   y[b][o][oy][ox] = sum_ch color[b][o][ch] * s[ch] + color[b][o][c]     color: DEVICE float [n][c][c+1] or NULL (y = s)
 * x (n, c, in_h, in_w) and y (n, c, out_h, out_w), c <= 4, are addressed through HOST arrays of four element strides
 * (n, c, h, w): planar NCHW and channels-last tensors are both taken as they are.
 * `_bwd` is the adjoint w.r.t. the image (the matrices carry no gradient, generator_trainer.py:421-422): it ADDS
 * into gx, fp32 planar (n, c, in_h, in_w), which the caller zeroes.                                                */
int b200gan_affine_color_fwd(const void* x, void* y, const double* mat, const float* color, int dtype, int n, int c,
                             int in_h, int in_w, int out_h, int out_w, const int64_t* x_strides,
                             const int64_t* y_strides, void* stream);
int b200gan_affine_color_bwd(const void* gy, float* gx, const double* mat, const float* color, int dtype, int n, int c,
                             int in_h, int in_w, int out_h, int out_w, const int64_t* gy_strides, void* stream);

/* ---- optimiser ------------------------------------------------------------------
 * Adam step as `torch.optim.Adam` (gt.py:161-173; eps added after the bias-corrected sqrt) fused with
 * the generator EMA `accumulate` (trainers/utils.py:8-12; ema may be NULL).  fp32, in place.
 * bias_corr: DEVICE float[2] = {1 - beta1^t, 1 - beta2^t} (on the device so that a captured CUDA
 * graph replays with the current step count).                                          */
int b200gan_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t numel,
                     float lr, float beta1, float beta2, float eps, const float* bias_corr,
                     float ema_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GAN_H_ */
