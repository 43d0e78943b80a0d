/* csmri_dc.h - C ABI of libcsmri_dc.so, the B200 (sm_100a) drop-in for the
 * k-space data-consistency (DC) hot path of mseitzer/csmri-refinement.
 *
 * The reference has no FFI of its own for this path (it is pure Python on top
 * of the third-party CUDA package pytorch-fft 0.14); each entry point below
 * names the reference interface it replaces (paths relative to the reference
 * root).  Tensors are fp32, NCHW contiguous, C = 2 planar (channel 0 = Re,
 * channel 1 = Im) exactly as data/reconstruction/deep_med_lib/utils/dnn_io.py:4-22
 * produces them: element (b,c,h,w) lives at ((b*2+c)*H+h)*W+w.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's caching
 *     allocator); the library allocates nothing except one immutable twiddle
 *     table per device (constant memory, uploaded by csmri_init / first use);
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL is
 *     the legacy default stream) and never synchronise;
 *   - return value 0 = ok, negative = CSMRI_E_*; csmri_last_error() returns a
 *     thread-local message.  Unsupported sizes are an error, never a fallback;
 *   - H and W must each be a power of two in [32, 1024] or 320 (= 2^6 * 5, radix-5
 *     path); outputs must not alias inputs unless stated.
 */
#ifndef CSMRI_DC_H_
#define CSMRI_DC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSMRI_OK 0
#define CSMRI_E_SHAPE (-1)    /* unsupported B/H/W                     */
#define CSMRI_E_NULLPTR (-2)  /* required pointer is NULL              */
#define CSMRI_E_ALIGN (-3)    /* pointer not 4-byte aligned            */
#define CSMRI_E_CUDA (-4)     /* CUDA runtime error, see last_error    */
#define CSMRI_E_ARG (-5)      /* bad scalar argument                   */

int csmri_version(void);
const char* csmri_last_error(void);

/* Upload the twiddle table to the current device.  Optional (done lazily by
 * every call below), but must have happened before CUDA-graph capture. */
int csmri_init(void);

/* Bytes of scratch the *_general / prepare / undersample entry points need
 * for a (B,2,H,W) problem (one hybrid-space tensor). */
size_t csmri_dc_workspace_bytes(int B, int H, int W);

/* ---- once per batch ------------------------------------------------------
 * Replaces nothing in the reference one-to-one: k0 and mask are constant over
 * the cascade (models/recnet.py:139-151 passes the same kspace, mask to every
 * dc_layers[idx].perform), so everything that depends only on them is hoisted.
 *
 *   dtab   (B,H)      out: D[b,h]/H with D = 1-m (noiseless) or (1-m)+m/(1+v),
 *                          m = mask[b,0,h,0]   (autograd of myfft.py:139,141)
 *   addend (B,2,H,W)  out: iFFT_W(c*k0)/sqrt(H*W), c = 1 (noiseless) or m*v/(1+v):
 *                          the k0 term of myfft.py:139,141 in hybrid (k_H, w)
 *                          space, which the strip kernel adds between its
 *                          forward and inverse column passes; may be NULL to skip
 *   row_constant      out: device int, set to 1 iff every mask row is constant
 *                          along W and mask[:,0]==mask[:,1] (the property of
 *                          compressed_sensing.py:115-116 that the Cartesian
 *                          strip kernel relies on), else 0
 *   scratch           unused by prepare (kept for ABI symmetry), may be NULL
 * noise_lvl == 0 selects the noiseless branch; every other value, negative
 * ones included, takes the noisy formula - the truthiness test `if v:` of
 * myfft.py:137-138 (Python callers map None to 0).
 */
int csmri_dc_prepare(const float* k0, const float* mask, int B, int H, int W,
                     float noise_lvl, float* dtab, float* addend,
                     int* row_constant, void* scratch, void* stream);

/* ---- once per batch, from what a Cartesian loader actually holds -----------
 * Same outputs as csmri_dc_prepare for a Cartesian acquisition described
 * compactly: the sampled-line table and only the sampled lines of k0.  A host
 * caller ships N bytes + 8*L*W bytes per slice instead of the two dense
 * (2,H,W) tensors the reference's loader builds (dnn_io.py:47-61 expands the
 * mask to a dense 2-channel array; compressed_sensing.py:510 leaves k0 zero off
 * the sampled lines), e.g. 0.27 instead of 2 tensors at 4x acceleration.
 *   k0_lines (B,2,L,W)  k0[b,c,h_j,:] for the L sampled rows h_0 < h_1 < ... of
 *                       slice b (= kspace[b][:, rows[b] != 0, :]); k0 is taken
 *                       to be zero off the sampled lines (k0 = mask * k0)
 *   rows     (B,H)      uint8, 1 = sampled line, as for csmri_undersample
 *   lines_ok            out: device int, 0 if some slice does not have exactly L
 *                       sampled rows (the plan is then invalid), else 1
 */
int csmri_dc_prepare_lines(const float* k0_lines, const unsigned char* rows,
                           int B, int H, int W, int L, float noise_lvl,
                           float* dtab, float* addend, int* lines_ok,
                           void* stream);

/* ---- DC forward, Cartesian (row-constant) masks ---------------------------
 * Replaces DataConsistencyInKspace.perform (myfft.py:153-163) =
 * Fft2d.forward (:83-90) + data_consistency (:131-142) + Ifft2d.forward
 * (:110-117) + both torch.cat (:159,161), and optionally the residual add of
 * models/recnet.py:147-148 (residual may be NULL).
 *   out = iFFT_H( dtab * FFT_H(x [+ residual]) + addend )        per column
 * One kernel, one pass over HBM: reads x (+residual) and addend, writes out.
 */
int csmri_dc_forward_cartesian(const float* x, const float* residual,
                               const float* dtab, const float* addend,
                               float* out, int B, int H, int W, void* stream);

/* ---- DC adjoint (backward), Cartesian masks -------------------------------
 * Replaces Ifft2d.backward (myfft.py:119-128) + autograd of the blend +
 * Fft2d.backward (myfft.py:92-102):  gx = iFFT_H( dtab * FFT_H(g) ).
 * The operator is self-adjoint, so this is the forward without the addend.
 */
int csmri_dc_adjoint_cartesian(const float* grad_out, const float* dtab,
                               float* grad_x, int B, int H, int W,
                               void* stream);

/* ---- DC forward / adjoint, arbitrary masks ---------------------------------
 * Same reference functions as above, evaluated as written for any mask:
 * row FFT -> (column FFT, blend with k0 under the dense mask, column iFFT)
 * -> row iFFT; three kernels, the hybrid tensor lives in `scratch`.
 */
int csmri_dc_forward_general(const float* x, const float* residual,
                             const float* k0, const float* mask, float* out,
                             int B, int H, int W, float noise_lvl,
                             void* scratch, void* stream);
int csmri_dc_adjoint_general(const float* grad_out, const float* mask,
                             float* grad_x, int B, int H, int W,
                             float noise_lvl, void* scratch, void* stream);

/* ---- unified entry point ---------------------------------------
 * mask_is_row_constant != 0: dtab/addend from csmri_dc_prepare are used and
 * k0/mask may be NULL; == 0: the general path is taken and dtab/addend may be
 * NULL.  adjoint: pass k0 = addend = NULL.
 */
int csmri_dc_forward(const float* x, const float* residual, const float* k0,
                     const float* mask, const float* dtab, const float* addend,
                     float* out, int B, int H, int W, float noise_lvl,
                     int mask_is_row_constant, void* scratch, void* stream);
int csmri_dc_adjoint(const float* grad_out, const float* mask,
                     const float* dtab, float* grad_x, int B, int H, int W,
                     float noise_lvl, int mask_is_row_constant, void* scratch,
                     void* stream);

/* ---- undersampling (loader side) -------------------------------------------
 * Replaces cs.undersample (deep_med_lib/utils/compressed_sensing.py:460-512,
 * centred=False, norm='ortho', noise=0) + dnn_io.to_tensor_format
 * (dnn_io.py:47-61) x4 + the channel split of scar_segmentation.py:212-218.
 *   img   (B,H,W)  real image in [0,1]
 *   rows  (B,H)    uint8, 1 = sampled phase-encode line (already ifftshift-ed;
 *                  chosen on the HOST by compressed_sensing.py:82-123 because
 *                  numpy's legacy RandomState cannot be reproduced on device)
 *   inp, kspace, mask, target: (B,2,H,W) outputs
 *   dtab (B,H), addend (B,2,H,W): optional (both or neither).  The noiseless DC
 *                  plan csmri_dc_prepare would compute for (kspace, mask): the
 *                  masked hybrid-space rows are an intermediate anyway, so a
 *                  loader that asks for it saves the prepare pass and the
 *                  row-constancy read (the mask is row-constant by construction).
 * Two passes (the mask is constant along W, so x_u = F_H^-1(r . F_H x)): a column
 * kernel writes target, the masked hybrid rows and inp; a row kernel transforms
 * the sampled rows only into kspace and writes the dense mask.  Masked k-space
 * entries are +0.0 (the reference's `mask * x_f` leaves signed zeros; no
 * consumer can tell them apart).  scratch: one (B,2,H,W) tensor, unused when
 * addend is given.
 */
int csmri_undersample(const float* img, const unsigned char* rows, float* inp,
                      float* kspace, float* mask, float* target, float* dtab,
                      float* addend, int B, int H, int W, void* scratch,
                      void* stream);

/* Plain ortho FFT2 / iFFT2 of a planar complex batch (Fft2d / Ifft2d forward,
 * myfft.py:78-128); inverse != 0 selects the inverse.  Used by the tests to
 * check the conventions pinned at myfft.py:225,241-242. */
int csmri_fft2(const float* x, float* out, int B, int H, int W, int inverse,
               void* scratch, void* stream);

/* ---- reporting-side pointwise ops on the DC output ---------------------------
 * csmri_magnitude_clamp replaces complex_abs (utils/tensor_transforms.py:62-75)
 * followed by torch.clamp(., lo, hi) as in output_transform
 * (data/reconstruction/rec_transforms.py:79-85):
 *   x (B,2,H,W) -> out (B,1,H,W) = clamp(sqrt(re^2 + im^2), lo, hi), bit-identical
 *   to the torch expression (separately rounded squares, sum, sqrt).
 * csmri_psnr_sum fuses that transform of prediction and target with the squared
 * error reduction of compute_psnr (metrics/image_metrics.py:7-19):
 *   *sum_sq (device double) = sum over B*H*W pixels of (|pred|_c - |target|_c)^2;
 *   PSNR = 10*log10(1 / (sum_sq / (B*H*W))).  Any H, W > 0. */
int csmri_magnitude_clamp(const float* x, float* out, int B, int H, int W,
                          float lo, float hi, void* stream);
int csmri_psnr_sum(const float* pred, const float* target, double* sum_sq,
                   int B, int H, int W, float lo, float hi, void* stream);

/* ---- loader tail (data/reconstruction/rec_transforms.py:40-47,62-65) -----------
 * csmri_shift_crop is the index plumbing of CenterCropInKspace
 * (myImageTransformations.py:935-954: fft2c -> crop_image_at (:105-117) -> ifft2c
 * -> abs; fft2c = fftshift . fft2 . ifftshift, mymath.py:18-29) around the two
 * csmri_fft2 calls: one gather pass per stage instead of roll / slice / pad / roll
 * copies.  Per axis (y with IH/OH, x with IW/OW):
 *   out[i] = in[(v + in_roll) mod I] with v = ((i + out_roll) mod O) + off,
 *   0 where v is outside [0, I) (the zero padding of crop_image_at).
 * in (B,in_ch,IH,IW), out (B,out_ch,OH,OW); in_ch = 1 reads a real image (imaginary
 * plane 0); out_ch = 1 writes sqrt(re^2 + im^2) rounded like csmri_magnitude_clamp
 * and, if absmax != NULL, the maximum per slice into absmax[B].  Rolls must lie in
 * [0, axis length); off may be negative.
 * csmri_plane_absmax / csmri_plane_divide are `x / np.max(np.abs(x))`
 * (rec_transforms.py:47,65) per plane of n contiguous floats: absmax[p] = max |x|,
 * out = x / denom[p] (IEEE division, bit-identical to the torch expression). */
int csmri_shift_crop(const float* in, float* out, int B, int in_ch, int IH, int IW,
                     int out_ch, int OH, int OW, int in_roll_y, int in_roll_x,
                     int off_y, int off_x, int out_roll_y, int out_roll_x,
                     float* absmax, void* stream);
int csmri_plane_absmax(const float* x, float* absmax, int planes, int n, void* stream);
int csmri_plane_divide(const float* x, const float* denom, float* out, int planes, int n,
                       void* stream);

/* ---- refinement-path pointwise ops --------------------------------------------
 * A "plane" is n contiguous floats; plane p starts at x + p*pitch (floats), so the
 * real channel of a (B,2,H,W) tensor is addressed with n = H*W, pitch = 2*H*W.
 *
 * csmri_plane_minmax: minimum[p] = min x[p,:], maximum[p] = max(x[p,:] - minimum[p])
 *   - the two reductions of _scale (models/refinement_wrapper.py:66-69) and of
 *   magnitude_image (utils/tensor_transforms.py:95-97) in one pass.
 * csmri_plane_scale, mode 0: out = (x - min) / max            (magnitude_image, :96-98)
 *                    mode 1: out = (x - min) / max * 2 - 1    (_scale, :68-71)
 *                    mode 2: out = ((x + 1) / 2) * max + min  (_unscale, :89-91)
 * csmri_refine_real_penalty_add: RefinementWrapper._refinement_real_penalty_add
 *   (models/refinement_wrapper.py:173-197) given the learnable model's output:
 *   pred[:,0] = _unscale(_scale(pretrained[:,0]) + scale * learnable), pred[:,1] =
 *   pretrained[:,1]; minimum / maximum (B) are returned for the backward.
 *   pretrained, pred (B,2,H,W); learnable (B,1,H,W); scale: device scalar (the
 *   nn.Parameter).  The pretrained output is treated as detached (:211-212,226-227).
 * csmri_refine_real_penalty_add_backward: grad_learnable (B,1,H,W) and
 *   csmri_refine_partials() partial sums per slice of d/d scale
 *   (grad_scale_partial, B * partials floats; their total is the gradient).
 * All are rounded op by op like the reference's tensor expressions: bit-identical
 * to the torch evaluation for finite inputs.  Any H, W > 0. */
int csmri_plane_minmax(const float* x, float* minimum, float* maximum, int planes,
                       int n, long long pitch, void* stream);
int csmri_plane_scale(const float* x, const float* minimum, const float* maximum,
                      float* out, int planes, int n, long long pitch_in,
                      long long pitch_out, int mode, void* stream);
int csmri_refine_real_penalty_add(const float* pretrained, const float* learnable,
                                  const float* scale, float* pred, float* minimum,
                                  float* maximum, int B, int H, int W, void* stream);
int csmri_refine_partials(void);
int csmri_refine_real_penalty_add_backward(const float* grad_pred, const float* learnable,
                                           const float* scale, const float* maximum,
                                           float* grad_learnable, float* grad_scale_partial,
                                           int B, int H, int W, void* stream);

/* ---- training step: weight gradient of RecNet's 3x3 convolutions --------------
 * The backward-weight half of torch.nn.Conv2d(CI, CO, 3, stride=1) as RecNet
 * builds it (models/recnet.py:37-48), i.e. what autograd computes for `weight`
 * in Runner._train_step (training/runner.py:154-178):
 *   dw[co][ci][ky][kx] = sum_{n,y,x} dy[n][co][y][x] * x[n][ci][y+ky-pad][x+kx-pad]
 *   x  (N, CI, H + 2 - 2*pad, W + 2 - 2*pad)   pad = 0: input already padded by the
 *                                              reference's ZeroPad2d layer; pad = 1:
 *                                              zero padding inside the convolution
 *   dy (N, CO, H, W);  dw (CO, CI, 3, 3), overwritten;  fp32 throughout, summed in a
 *   fixed order (deterministic).
 * CI and CO multiples of 32 with H a multiple of 4, or the thin layers 2 -> 32 and
 * 32 -> 2 with H a multiple of 16; W a multiple of 32.  Anything else is
 * CSMRI_E_SHAPE (the caller keeps its own convolution backend for those).
 * workspace: csmri_conv3x3_wgrad_workspace_bytes(CI, CO) bytes of device memory. */
size_t csmri_conv3x3_wgrad_workspace_bytes(int CI, int CO);
int csmri_conv3x3_wgrad(const float* x, const float* dy, float* dw, void* workspace,
                        int N, int CI, int CO, int H, int W, int pad, void* stream);
/* 32 -> 32 channels, zero padding 1, H % 16 == 0, W % 64 == 0 (the tensor-core path of
 * the call above): the same dw plus the bias gradient of the layer,
 *   db[co] = sum_{n,y,x} dy[n][co][y][x]        (what autograd computes for `bias`),
 * a by-product of the kernel reading every dy value once; db (32), overwritten.  Same
 * workspace.  Other shapes are CSMRI_E_SHAPE. */
int csmri_conv3x3_wgrad_bias(const float* x, const float* dy, float* dw, float* db, void* workspace,
                             int N, int H, int W, void* stream);
/* The same for RecNet's 2 -> 32 layer (x (N,2,H,W), dy (N,32,H,W), zero padding 1, H % 8 == 0,
 * W % 32 == 0, x 16-byte aligned): dw (32,2,3,3) and db (32). */
int csmri_conv3x3_wgrad_thin_bias(const float* x, const float* dy, float* dw, float* db,
                                  void* workspace, int N, int H, int W, void* stream);
/* And for the 32 -> 2 layer that closes a block (models/recnet.py:48; x (N,32,H,W), dy (N,2,H,W),
 * same shape rules): dw (2,32,3,3) and db (2).  Workspace of csmri_conv3x3_wgrad_workspace_bytes(32, 2). */
int csmri_conv3x3_wgrad_thin_in_bias(const float* x, const float* dy, float* dw, float* db,
                                     void* workspace, int N, int H, int W, void* stream);

/* RecNet's thin 3x3 convolutions (first / last layer of a block, models/recnet.py:
 * 45-48), stride 1, zero padding 1:  y = act(conv(x, w) + bias)
 *   x (N,A,H,W), w (B,A,3,3), bias (B) or NULL, y (N,B,H,W);
 *   (A,B) = (2,32) or (32,2); slope > 0 applies LeakyReLU, slope = 0 none.
 * The data gradient of such a layer is the same call with w flipped in both
 * spatial axes and its first two axes transposed.  H % 16 == 0, W % 32 == 0. */
int csmri_conv3x3_thin(const float* x, const float* w, const float* bias, float* y,
                       int N, int A, int B, int H, int W, float slope, void* stream);
/* The 2 -> 32 form without bias / activation, its result multiplied by the derivative of a
 * LeakyReLU whose output signs csmri_conv3x3_tc_signs recorded:
 *   y[n][c] = conv(x, w)[n][c] * (bit c of signs[n] ? 1 : act_slope)
 * With w = the 32 -> 2 layer's weights flipped and transposed this is that layer's data
 * gradient followed by the backward of the nn.LeakyReLU in front of it (models/recnet.py:
 * 45-48), in one pass.  x (N,2,H,W), w (32,2,3,3), signs (N,H,W) uint32, y (N,32,H,W). */
int csmri_conv3x3_thin_masked(const float* x, const float* w, const unsigned* signs, float* y,
                              int N, int H, int W, float act_slope, void* stream);
/* Data gradient of a thin layer with weights w (CO,CI,3,3), zero padding 1 - what autograd's
 * convolution backward returns for the layer's input: dy (N,CO,H,W) -> dx (N,CI,H,W).  The
 * kernel of the opposite shape reads w through the transposed, mirrored index map (no flipped
 * copy of the weights is made).  signs (optional, CI = 32 only): as csmri_conv3x3_thin_masked,
 * dx is also multiplied by (bit c ? 1 : act_slope) - the derivative of the LeakyReLU that
 * produced the layer's input. */
int csmri_conv3x3_thin_dgrad(const float* dy, const float* w, const unsigned* signs, float* dx,
                             int N, int CI, int CO, int H, int W, float act_slope, void* stream);

/* RecNet's 32 -> 32 channel 3x3 convolutions (the inner layers of every ConvBlock,
 * models/recnet.py:37-44), stride 1, zero padding 1, on the tcgen05 tensor cores
 * with an error-compensated TF32 split (fp32-level accuracy, fp32 accumulation in
 * tensor memory):  y = act(conv(x, wq) + bias)
 *   x (N,32,H,W), w (32,32,3,3), bias (32) or NULL, y (N,32,H,W), y must not alias x;
 *   slope > 0 applies LeakyReLU, slope = 0 none;
 *   transpose_flip = 0: wq = w (what nn.Conv2d.forward computes);
 *   transpose_flip = 1: wq[co][ci][ky][kx] = w[ci][co][2-ky][2-kx], i.e. the data
 *   gradient of the same layer (what autograd computes for its input).
 * H % 8 == 0, W % 128 == 0; anything else is CSMRI_E_SHAPE (the caller keeps its
 * own convolution backend for those shapes). */
int csmri_conv3x3_tc(const float* x, const float* w, const float* bias, float* y,
                     int N, int C, int H, int W, float slope, int transpose_flip, void* stream);
/* The forward form of the call above that also records the sign of every output:
 *   signs (N,H,W) uint32, bit c of a pixel = (y[n][c][h][w] > 0)   (4 bytes per pixel),
 * and optionally (in_signs != NULL) the same word for the INPUT x - when x is the output of
 * a LeakyReLU, that is what the data gradient of this layer needs to carry its derivative. */
int csmri_conv3x3_tc_signs(const float* x, const float* w, const float* bias, float* y, unsigned* signs,
                           unsigned* in_signs, int N, int C, int H, int W, float slope, void* stream);
/* The same convolution without bias / activation, its result multiplied by the derivative
 * of the LeakyReLU that produced its consumer's input:
 *   y[n][c] = conv(x, wq)[n][c] * (bit c of signs[n] ? 1 : act_slope)
 * With transpose_flip = 1 this is "data gradient of layer L+1, then backward of the
 * nn.LeakyReLU(relu_leakiness) between layer L and L+1" (models/recnet.py:45-47) in one
 * pass: signs is what csmri_conv3x3_tc_signs wrote in layer L's forward pass. */
int csmri_conv3x3_tc_masked(const float* x, const float* w, const unsigned* signs, float* y,
                            int N, int C, int H, int W, float act_slope, int transpose_flip,
                            void* stream);

/* Bias + LeakyReLU after a convolution (models/recnet.py:45-48: Conv2d(bias=True)
 * followed by nn.LeakyReLU(relu_leakiness, inplace=True)), fused into one pass:
 *   csmri_bias_lrelu:           z (N,C,H,W) <- lrelu(z + bias[c]), in place
 *   csmri_bias_lrelu_backward:  grad_z = grad_y * (y > 0 ? 1 : slope),
 *                               grad_bias[c] = sum_{n,h,w} grad_z   (fixed order)
 * y is the forward OUTPUT; slope > 0.  H*W must be a multiple of 4, pointers
 * 16-byte aligned.  partial: N*C*8 floats of scratch. */
int csmri_bias_lrelu(float* z, const float* bias, int N, int C, int H, int W, float slope,
                     void* stream);
int csmri_bias_lrelu_backward(const float* grad_y, const float* y, float* grad_z,
                              float* grad_bias, float* partial, int N, int C, int H, int W,
                              float slope, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSMRI_DC_H_ */
