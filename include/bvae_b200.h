/* bvae_b200.h -- C ABI of the B200 (sm_100a) Beta-VAE-SE training hot path.
 *
 * The reference (jtang25/Beta-VAE) has no FFI: its hot path is a Python class surface
 * (src/models/beta_vae.py, src/models/se_blocks.py) that dispatches to torch.nn modules.
 * Every entry point below replaces the ATen/cuDNN/cuBLAS call(s) one of those modules
 * issues; the reference line it stands in for is cited next to it.  The Python side
 * (beta-vae_b200/src/models, beta-vae_b200/bvae_rt) binds these with ctypes and wraps them in
 * torch.autograd.Function so `loss.backward()` in the reference's train loop keeps working.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated;
 *   - no allocation inside: outputs and workspaces are caller-provided (size queries below);
 *   - `stream` is a cudaStream_t (NULL = legacy default stream);
 *   - activations are NHWC ("pixel-major"): [B][H][W][C], dtype BVAE_F32 or BVAE_BF16;
 *     parameters and gradients cross the boundary as fp32 in the reference's own layout
 *     ([Cout][Cin][3][3], fc in NCHW-flatten feature order) so state_dict / optimizer /
 *     checkpoint code is untouched;
 *   - return 0 on success, negative BVAE_ERR_* otherwise; bvae_last_error() gives the text
 *     (thread-local).  No entry point ever falls back to a CPU computation.
 */
#ifndef BVAE_B200_H
#define BVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* bvae_stream_t; /* cudaStream_t */

enum { BVAE_F32 = 0, BVAE_BF16 = 1 };
enum { BVAE_LOSS_MSE = 0, BVAE_LOSS_L1 = 1, BVAE_LOSS_BCE = 2 };
enum {
  BVAE_OK = 0,
  BVAE_ERR_ARG = -1,         /* null pointer / bad size */
  BVAE_ERR_UNSUPPORTED = -2, /* shape or dtype outside what the kernels cover */
  BVAE_ERR_CUDA = -3,        /* a CUDA runtime call or launch failed */
  BVAE_ERR_WORKSPACE = -4    /* workspace too small */
};

int bvae_version(void);
const char* bvae_last_error(void);
/* 1 when the device behind the current context is compute capability 10.x */
int bvae_device_ok(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
unsigned long long bvae_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Weight packing: fp32 master parameters (reference layout) -> compute copies.
 * ---------------------------------------------------------------------------------------- */
/* nn.Conv2d weight [Cout][Cin][3][3] (beta_vae.py:37,54).
 *   layout 0 ("kn", CUDA-core kernels):   wp[t][ci][co]   t = r*3+s
 *   layout 1 ("nk", tcgen05 B operand):   wp[t][co][ci]
 * `transpose` != 0 swaps the roles of ci/co (dgrad: contraction runs over Cout). */
int bvae_pack_conv_weight(const float* w, void* wp, int Cout, int Cin, int layout, int transpose,
                          int dtype, bvae_stream_t stream);
/* The same for n <= 32 weights in one launch (host arrays of device pointers / dims): the packs are re-made from
 * the fp32 masters every training step. */
int bvae_pack_conv_weights(int n, const void* const* w, void* const* wp, const int* Cout, const int* Cin,
                           const int* layout, const int* transpose, int dtype, bvae_stream_t stream);
/* fc_mu / fc_logvar weights [L][C*P] in NCHW-flatten order (beta_vae.py:117-118,138-140)
 * -> one matrix wh[2L][P*C] in NHWC-flatten order (rows 0..L-1 = mu, L..2L-1 = logvar). */
int bvae_pack_heads_weight(const float* w_mu, const float* w_logvar, void* wh, int L, int C, int P,
                           int dtype, bvae_stream_t stream);
/* fc_dec weight [C*P][L], bias [C*P] (beta_vae.py:119,155) -> wd[L][P*C], bd[P*C] (fp32). */
int bvae_pack_fcdec_weight(const float* w, const float* b, void* wd, float* bd, int L, int C, int P,
                           int dtype, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3x3 convolution, padding 1, stride 1 or 2  (nn.Conv2d at beta_vae.py:37/42 and :54;
 * backward = what autograd's convolution_backward computes for them).
 *   x  [B][H][W][Cin], y [B][Ho][Wo][Cout], Ho = (H-1)/stride + 1.
 * ---------------------------------------------------------------------------------------- */
/* Which weight pack the kernels want for this shape: 1 = [t][n][k] (tcgen05 implicit GEMM, bf16),
 * 0 = [t][k][n] (CUDA-core kernels: fp32, odd shapes).  kind: 0 = fprop, 1 = dgrad. */
int bvae_conv3x3_weight_layout(int H, int W, int Cin, int Cout, int stride, int dtype, int kind);
int bvae_conv3x3_fprop(const void* x, const void* wp, const float* bias, void* y, int B, int H, int W,
                       int Cin, int Cout, int stride, int x_dtype, int y_dtype, bvae_stream_t stream);
/* fprop that also emits, from its epilogue, per-sample partial (sum, sum of squares) of the stored y -- the
 * GroupNorm(1,C) statistics of the block (beta_vae.py:43,63) without a pass over y:
 *   stat_part [B][n][2] doubles, n = bvae_conv3x3_fprop_stats_parts(...) (0: not available for this shape). */
int bvae_conv3x3_fprop_stats_parts(int H, int W, int Cin, int Cout, int stride, int x_dtype, int y_dtype);
int bvae_conv3x3_fprop_stats(const void* x, const void* wp, const float* bias, void* y, double* stat_part, int B, int H,
                             int W, int Cin, int Cout, int stride, int x_dtype, int y_dtype, bvae_stream_t stream);
/* dx = conv_transpose(dy, w); wp_t is the layout-0, transpose=1 pack. */
int bvae_conv3x3_dgrad(const void* dy, const void* wp_t, void* dx, int B, int H, int W, int Cin, int Cout,
                       int stride, int dtype, bvae_stream_t stream);
size_t bvae_conv3x3_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int stride);
/* dw [Cout][Cin][3][3] fp32, dbias [Cout] fp32 (may be NULL). Deterministic (two-stage split-K). */
int bvae_conv3x3_wgrad(const void* x, const void* dy, float* dw, float* dbias, void* ws, size_t ws_bytes,
                       int B, int H, int W, int Cin, int Cout, int stride, int x_dtype, int dy_dtype,
                       bvae_stream_t stream);

/* First encoder layer (Cin = 1, stride 2, Cout = 64; nn.Conv2d at beta_vae.py:37 with in_ch = 1) on tcgen05:
 * xcol [B][H/2][W/2][16] bf16 is the 9-tap im2col of the fp32 image x [B][H][W] plus a ones column, built once per
 * forward and reused by the weight gradient.  y / dy [B][H/2][W/2][64] bf16; dw [64][1][3][3], db [64] fp32. */
int bvae_conv_first_supported(int H, int W, int Cin, int Cout, int stride);
int bvae_conv_first_im2col(const float* x, void* xcol, int B, int H, int W, bvae_stream_t stream);
size_t bvae_conv_first_workspace(void);
int bvae_conv_first_fprop(const void* xcol, const float* w, const float* bias, void* y, void* ws, size_t ws_bytes, int B,
                          int Ho, int Wo, bvae_stream_t stream);
int bvae_conv_first_wgrad(const void* xcol, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int B, int Ho,
                          int Wo, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Bilinear x2 upsample, align_corners=False (nn.Upsample at beta_vae.py:53).
 *   x [B][H][W][C] -> y [B][2H][2W][C];  bwd: dy [B][2H][2W][C] -> dx [B][H][W][C].
 * ---------------------------------------------------------------------------------------- */
int bvae_upsample2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, bvae_stream_t stream);
int bvae_upsample2x_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm(1,C) + ReLU + Squeeze-and-Excitation, fused
 * (beta_vae.py:20-21,43-45,63-65 and se_blocks.py:15-19).
 *   y   [B][P][C]  conv output (pre-norm), P = H*W
 *   out [B][P][C]  relu(gn(y)) * gate
 * SE weights are fp32 in reference layout: w1 [R][C], b1 [R], w2 [C][R], b2 [C];
 * pass w1 == NULL for a block without SE (use_decoder_se: false) -> gate = 1.
 * Saved for backward (fp32): stats [B][2] = (mean, rstd); cpos [B][C] = #positive pixels;
 * spos [B][C] = sum of xhat over positive pixels; hidden [B][R]; gate [B][C].
 * ---------------------------------------------------------------------------------------- */
size_t bvae_gn_relu_se_workspace(int B, int P, int C);
/* stat_part (optional, may be NULL): [B][n_stat_part][2] doubles = per-sample partial (sum, sum of squares) of y that
 * the convolution's epilogue produced (bvae_conv3x3_fprop_stats); without it a statistics pass over y runs first. */
int bvae_gn_relu_se_fwd(const void* y, void* out, const float* gamma, const float* beta, const float* w1,
                        const float* b1, const float* w2, const float* b2, int B, int P, int C, int R,
                        float* stats, float* cpos, float* spos, float* hidden, float* gate,
                        const double* stat_part, int n_stat_part, void* ws, size_t ws_bytes, int dtype,
                        bvae_stream_t stream);
/* dy may alias dout.  Parameter gradients are fp32, reference layout, overwritten (not accumulated). */
int bvae_gn_relu_se_bwd(const void* dout, const void* y, void* dy, const float* gamma, const float* beta,
                        const float* w1, const float* b1, const float* w2, const float* b2, int B, int P,
                        int C, int R, const float* stats, const float* cpos, const float* spos,
                        const float* hidden, const float* gate, float* dgamma, float* dbeta, float* dw1,
                        float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int dtype,
                        bvae_stream_t stream);
/* Deferred form of the two backward entry points (up != 0: the fused-upsample one, dout = [B][2H][2W][C]): dy is
 * produced now; the reduction of the parameter gradients over the batch -- a few microseconds of work behind a launch,
 * needed only by the optimizer -- is described in `job_out` (HOST memory, bvae_gn_params_job_bytes()) instead of being
 * launched.  bvae_gn_bwd_params_batched then runs the jobs of all blocks in ONE launch (`jobs`: host array of records,
 * back to back) once backward has been enqueued.  `ws` and every pointer of a job must stay valid until then. */
size_t bvae_gn_params_job_bytes(void);
int bvae_gn_relu_se_bwd_deferred(const void* dout, const void* y, void* dy, const float* gamma, const float* beta,
                                 const float* w1, const float* w2, int B, int H, int W, int C, int R, int up,
                                 const float* stats, const float* cpos, const float* spos, const float* hidden,
                                 const float* gate, float* dgamma, float* dbeta, float* dw1, float* db1, float* dw2,
                                 float* db2, void* ws, size_t ws_bytes, int dtype, void* job_out, bvae_stream_t stream);
int bvae_gn_bwd_params_batched(const void* jobs, int njobs, bvae_stream_t stream);
/* The same block followed by the NEXT decoder block's nn.Upsample(x2, bilinear, align_corners=False)
 * (beta_vae.py:52-53), fused: the activation is never written at its own resolution.
 *   fwd: y [B][H][W][C] -> out_up [B][2H][2W][C] = upsample2x(relu(gn(y)) * gate)
 *   bwd: d_up [B][2H][2W][C] (gradient w.r.t. out_up) -> dy [B][H][W][C]; upsample2x_bwd(d_up) lives in the workspace.
 * Bit-identical to bvae_gn_relu_se_* composed with bvae_upsample2x_*. */
size_t bvae_gn_relu_se_up_workspace(int B, int H, int W, int C);
/* 1: this shape is served by the single-launch cluster-resident kernel (keep the upsample as its own pass); 0: the
 * streaming passes run, and the fused entry points below save one write + one read of the activation. */
int bvae_gn_relu_se_prefers_unfused(int B, int P, int C, int R, int dtype);
int bvae_gn_relu_se_up_fwd(const void* y, void* out_up, const float* gamma, const float* beta, const float* w1,
                           const float* b1, const float* w2, const float* b2, int B, int H, int W, int C, int R,
                           float* stats, float* cpos, float* spos, float* hidden, float* gate,
                           const double* stat_part, int n_stat_part, void* ws, size_t ws_bytes, int dtype,
                           bvae_stream_t stream);
int bvae_gn_relu_se_up_bwd(const void* d_up, const void* y, void* dy, const float* gamma, const float* beta,
                           const float* w1, const float* b1, const float* w2, const float* b2, int B, int H,
                           int W, int C, int R, const float* stats, const float* cpos, const float* spos,
                           const float* hidden, const float* gate, float* dgamma, float* dbeta, float* dw1,
                           float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int dtype,
                           bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * mu/logvar heads + clamp + reparameterisation + KL + KL-side objective, fused
 * (beta_vae.py:137-150 and :219-253).
 *   h [B][F] activations (NHWC flatten, F = P*C), wh [2L][F] packed, eps [B][L] fp32
 *   (eps == NULL or deterministic != 0 -> z = mu and all KL terms are zero).
 *   scalars[0] = kl_mean, [1] = kl_effective, [2] = KL-side objective
 *   (beta*kl_effective  |  capacity_weight*|kl_mean - capacity|)  + latent_reg, [3] = latent_reg.
 *   use_capacity != 0 selects the capacity objective.
 * ---------------------------------------------------------------------------------------- */
size_t bvae_heads_workspace(int B, int L, int F);
int bvae_heads_reparam_kl_fwd(const void* h, const void* wh, const float* b_mu, const float* b_logvar,
                              const float* eps, float lv_lo, float lv_hi, float beta, int use_capacity,
                              float capacity, float capacity_weight, float free_bits, float reg_lambda,
                              int deterministic, float* mu, float* logvar, float* z, float* kl_per_dim,
                              float* scalars, uint8_t* lv_pass, void* ws, size_t ws_bytes, int B, int L,
                              int F, int dtype, bvae_stream_t stream);
/* g_* are upstream gradients (device, fp32): g_mu/g_logvar/g_z [B][L], g_scalars [4]
 * (only [0] kl_mean and [2] objective are differentiable); any may be NULL (= zero).
 * Outputs: dh [B][F] (act dtype); dw_mu/dw_logvar [L][C*P] and db_* [L] fp32 reference layout. */
int bvae_heads_reparam_kl_bwd(const void* h, const void* wh, const float* eps, const float* mu,
                              const float* logvar, const float* kl_per_dim, const float* scalars,
                              const uint8_t* lv_pass, const float* g_mu, const float* g_logvar,
                              const float* g_z, const float* g_scalars, float beta, int use_capacity,
                              float capacity, float capacity_weight, float free_bits, float reg_lambda,
                              int deterministic, void* dh, float* dw_mu, float* dw_logvar, float* db_mu,
                              float* db_logvar, void* ws, size_t ws_bytes, int B, int L, int C, int P,
                              int dtype, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fc_dec (beta_vae.py:153-161): optional latent clamp, Linear(L -> C*P), output in NHWC.
 * ---------------------------------------------------------------------------------------- */
int bvae_fc_dec_fwd(const float* z, const void* wd, const float* bd, void* h0, float z_clamp, int B, int L,
                    int F, int dtype, bvae_stream_t stream);
size_t bvae_fc_dec_workspace(int B, int L, int F);
/* dz [B][L] fp32; dw [C*P][L], db [C*P] fp32 reference layout.  z_clamp <= 0 means no clamp. */
int bvae_fc_dec_bwd(const void* dh0, const float* z, const void* wd, float z_clamp, float* dz, float* dw,
                    float* db, void* ws, size_t ws_bytes, int B, int L, int C, int P, int dtype,
                    bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * final_conv (C -> Co<=4, 3x3, pad 1) + sigmoid + reconstruction loss, fused
 * (beta_vae.py:163-164 and :185-192).
 *   h [B][H][W][C] act dtype; w [Co][C][3][3], b [Co] fp32 (reference layout, read directly);
 *   target/recon [B][Co][H][W] fp32 (NCHW, as the caller of BetaVAE.loss provides / expects);
 *   loss_out[0] = sum(elementwise loss) / B   (target == NULL -> decode only, loss untouched).
 * ---------------------------------------------------------------------------------------- */
size_t bvae_finalconv_workspace(int B, int H, int W, int C, int Co);
int bvae_finalconv_sigmoid_recon_fwd(const void* h, const float* w, const float* b, const float* target,
                                     float* recon, float* loss_out, int loss_type, void* ws,
                                     size_t ws_bytes, int B, int H, int W, int C, int Co, int dtype,
                                     bvae_stream_t stream);
/* g_recon [B][Co][H][W] (may be NULL) and g_loss (device scalar, may be NULL) are the upstream
 * gradients of the two forward outputs.  dh may be NULL (no input gradient wanted). */
int bvae_finalconv_sigmoid_recon_bwd(const void* h, const float* w, const float* recon, const float* target,
                                     const float* g_recon, const float* g_loss, void* dh, float* dw,
                                     float* db, int loss_type, void* ws, size_t ws_bytes, int B, int H,
                                     int W, int C, int Co, int dtype, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser step over a flat fp32 buffer: global-norm clip + Adam, one pass
 * (callbacks.py:18-30 clip_grad_norm_ + torch.optim.Adam at callbacks.py:96).
 *   sumsq_out[0] receives ||g||^2 (before clipping).  max_norm <= 0 disables clipping.
 * ---------------------------------------------------------------------------------------- */
size_t bvae_adam_workspace(size_t n);
int bvae_grad_sumsq(const float* g, size_t n, float* sumsq_out, void* ws, size_t ws_bytes,
                    bvae_stream_t stream);
int bvae_adam_clip_step(float* p, const float* g, float* m, float* v, size_t n, const float* sumsq,
                        float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int decoupled_wd, int step, bvae_stream_t stream);
/* Same step with the per-step scalars read from DEVICE memory, so one captured CUDA graph can be
 * replayed for every step: hyper[0] = lr, hyper[1] = 1 - beta1^t, hyper[2] = sqrt(1 - beta2^t). */
int bvae_adam_clip_step_dev(float* p, const float* g, float* m, float* v, size_t n, const float* sumsq,
                            float grad_scale, float max_norm, const float* hyper, float beta1, float beta2,
                            float eps, float weight_decay, int decoupled_wd, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Focal-frequency loss, spectrum side (FocalFrequencyLoss.forward, src/training/losses.py:23-33, without the
 * FFT itself): spec = one-sided rfft2 of (pred - target), [B][C][H][Wh = W/2+1] interleaved (re, im) fp32.
 *   loss[0] = mean(w * |F|^2),  w = clamp_min(|F|^2 / scale_c, eps)^alpha,  scale_c = mean_{n,h,w} |F|^2 + eps
 * scale [C] and U [C] are saved for the backward, which returns the gradient w.r.t. the complex spectrum
 * (PyTorch convention: d/dre + i d/dim), g_loss = upstream gradient of the scalar (device).
 * ---------------------------------------------------------------------------------------- */
size_t bvae_ffl_workspace(int C);
int bvae_ffl_fwd(const float* spec, int B, int C, int H, int Wh, int W, float alpha, float eps, float* loss,
                 float* scale, float* U, void* ws, size_t ws_bytes, bvae_stream_t stream);
int bvae_ffl_bwd(const float* spec, const float* scale, const float* U, const float* g_loss, float* g_spec, int B,
                 int C, int H, int Wh, int W, float alpha, float eps, bvae_stream_t stream);
/* The whole term on images, 1-D line FFTs in shared memory, rows then columns (power-of-two sides <= 128):
 *   loss = weight * FFL(pred, target);  pred / target / d_pred [B][C][H][W] fp32 (NCHW);
 *   spec [B][C][H][W/2+1][2] fp32 is the saved one-sided spectrum of pred - target (ortho), scale / U [C] as above. */
int bvae_ffl_image_supported(int H, int W);
size_t bvae_ffl_image_workspace(int B, int C, int H, int W);
int bvae_ffl_image_fwd(const float* pred, const float* target, int B, int C, int H, int W, float alpha, float eps,
                       float weight, float* spec, float* loss, float* scale, float* U, void* ws, size_t ws_bytes,
                       bvae_stream_t stream);
int bvae_ffl_image_bwd(const float* spec, const float* scale, const float* U, const float* g_loss, float weight,
                       float* d_pred, void* ws, size_t ws_bytes, int B, int C, int H, int W, float alpha, float eps,
                       bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Scalar tail of BetaVAE.loss (beta_vae.py:216-217, 249-253) in one launch each way:
 *   recon = recon_base + recon_lpips + recon_ffl;  total = recon + kl_scalars[2]
 *   out6 = (total, recon, kl_mean, recon_ffl, recon_base, recon_lpips); recon_lpips / recon_ffl may be NULL (= 0).
 *   bwd: g5 = (d recon_base [= d lpips = d ffl], d kl_scalars[0..3]) from g_total / g_recon (either may be NULL).
 * ---------------------------------------------------------------------------------------- */
int bvae_loss_combine_fwd(const float* recon_base, const float* recon_lpips, const float* recon_ffl,
                          const float* kl_scalars, float* out6, bvae_stream_t stream);
int bvae_loss_combine_bwd(const float* g_total, const float* g_recon, float* g5, bvae_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange: in-place sum of a symmetric fp32 buffer over the ranks of one node
 * through the NVSwitch multicast object it is bound to (multimem.ld_reduce / multimem.st; one kernel per rank).
 *   mc_ptr     multicast address of the first float of the range (16-byte aligned);  n floats, a multiple of 4
 *   flag_ptrs  DEVICE array of `world` pointers: rank q's symmetric flag buffer (bvae_allreduce_nvls_flag_count uint32,
 *              zeroed once) as mapped on this device;  state: bvae_allreduce_nvls_state_bytes() of zeroed device memory
 * Every rank launches it once per exchange (same range, same nblocks, same order of exchanges); CUDA-graph replayable.
 * No shared memory, 512-thread CTAs: it co-resides with the persistent convolution kernels, so the engine runs it per
 * gradient bucket on a side stream under the rest of backward.
 * ---------------------------------------------------------------------------------------- */
size_t bvae_allreduce_nvls_state_bytes(void);
size_t bvae_allreduce_nvls_flag_count(int nblocks, int world);
int bvae_allreduce_nvls(void* mc_ptr, size_t n, void* flag_ptrs, void* state, int rank, int world, int nblocks,
                        bvae_stream_t stream);
/* Same sum through plain peer mappings (NVLink loads / stores, no multicast): rank r loads the `world` copies of its
 * slice through `buf_ptrs` (DEVICE array of `world` pointers to element 0 of every rank's buffer as mapped on this
 * device), adds them in rank order and stores the sum into every copy.  Range = floats [off, off + n), multiples of 4;
 * world in {2, 4, 8}.  Flags / state / nblocks as above. */
int bvae_allreduce_p2p(void* buf_ptrs, size_t off, size_t n, void* flag_ptrs, void* state, int rank, int world,
                       int nblocks, bvae_stream_t stream);

/* Input pipeline: uint8 image bytes -> fp32 in [0, 1], dst[i] = float(src[i]) / 255 with IEEE division -- the values
 * torchvision's ToTensor produces from the dataset's PNGs (datautils.py:59-65), so shipping bytes over PCIe and
 * converting on the device is bit-identical to shipping ToTensor's floats.  n = element count; 16-byte aligned. */
int bvae_u8_to_unit_f32(const void* src, float* dst, size_t n, bvae_stream_t stream);

/* Train-time augmentation of a uint8 batch [B][C][H][W] on the device: horizontal flip -> nearest-neighbour rotation
 * about the image centre (fill 0) -> brightness, per sample -- what the reference's train transform does to each PIL
 * image inside its DataLoader workers (src/data_processing/augmentations.py:5-18: torchvision RandomHorizontalFlip,
 * RandomRotation(degrees), ColorJitter(brightness), followed by ToTensor), bit for bit.
 *   params    DEVICE array [B][8] of 32-bit words per sample: {flags, a0, a1, a2, a3, a4, a5, factor(fp32 bits)};
 *             flags bit 4 = mirror, bit 5 = apply brightness `factor`; a0..a5 = Pillow's 16.16 fixed-point inverse
 *             affine coefficients (identity: a0 = a4 = 65536, others 0).  bvae_rt/augment.py builds them.
 *   dst_u8    augmented bytes, or null;  dst_unit: the same pixels as fp32 v / 255 (ToTensor), or null.  Not in place.
 * W must be a multiple of 4. */
int bvae_augment_u8(const void* src, void* dst_u8, float* dst_unit, const void* params, int B, int C, int H, int W,
                    bvae_stream_t stream);

/* NCHW fp32 -> NHWC act dtype (RGB input path, data.grayscale: false) and back. */
int bvae_nchw_to_nhwc(const float* x, void* y, int B, int C, int H, int W, int dtype, bvae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BVAE_B200_H */
