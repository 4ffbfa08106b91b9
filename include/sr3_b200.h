/* sr3_b200 -- C ABI of the B200-native SR3 hot path (libsr3_b200.so).
 *
 * The reference (Janspiry/Image-Super-Resolution-via-Iterative-Refinement) is pure Python/PyTorch and has no
 * FFI; its boundary for this path is the Python factory model/networks.py:83-116 `define_G(opt)` and the methods of
 * the module it returns.  Each entry point below replaces the reference function cited next to it, with the same
 * argument meaning, on plain device pointers (fp32, NCHW contiguous, exactly the tensors the reference passes).
 * No torch types appear here; `stream` is a cudaStream_t passed as void*.  All functions return 0 on success and a
 * non-zero code otherwise, with a message available from sr3_last_error().  There is no CPU fallback.
 *
 * Binding shown in INTEGRATION.md (ctypes, what a maintainer of the reference would add to model/networks.py).
 */
#ifndef SR3_B200_H
#define SR3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR3_MAX_LEVELS 8

/* opt['model']['unet'] + opt['model']['diffusion'] as consumed by define_G (model/networks.py:83-109) */
typedef struct sr3_unet_config {
    int in_channel;                    /* 6 conditional, 3 unconditional */
    int out_channel;                   /* 3 */
    int inner_channel;                 /* 64 (must be a multiple of 64) */
    int norm_groups;                   /* 32 (define_G default, networks.py:89-90) or 16 */
    int n_mults;
    int channel_mults[SR3_MAX_LEVELS]; /* channel_multiplier */
    int n_attn_res;
    int attn_res[SR3_MAX_LEVELS];
    int res_blocks;
    int image_size;
    int channels;                      /* diffusion.channels (3) */
    int conditional;                   /* diffusion.conditional */
    int precision;                     /* 0: bf16 operands, fp32 accumulate / stream (north_star tolerance 1e-2 rel);
                                          1: "precise" -- every tensor-core operand is a (hi, lo) bf16 pair and each product is
                                             hi*hi + hi*lo + lo*hi (fp32-level accuracy, north_star tolerance 1e-3 rel; ~3x the MMA work).
                                          The reference arithmetic is fp32 nn.Conv2d / nn.Linear (unet.py:87). */
} sr3_unet_config;

typedef struct sr3_engine sr3_engine;

/* Message of the last failure on the calling thread ("" if none). */
const char* sr3_last_error(void);
/* ABI version of this header. */
int sr3_abi_version(void);

/* UNet.__init__ (model/sr3_modules/unet.py:161-233) + GaussianDiffusion.__init__ (diffusion.py:64-82):
 * builds the layer plan, allocates activations / packed weights on `device` for a fixed batch size.
 * Threading: calls on one engine must be serialised by the caller.  Kernels whose CTAs wait for partners are safe next to other work on the
 * device: the split-K partners of a tile are one thread-block cluster (gang-scheduled by the hardware), the persistent step kernel (SR3_MEGA=1)
 * is a cooperative launch -- engines driven concurrently from different streams of one device cannot deadlock each other. */
int sr3_engine_create(const sr3_unet_config* cfg, int batch, int device, sr3_engine** out);
void sr3_engine_destroy(sr3_engine* e);

/* Parameter table in the reference's state_dict order and naming ("downs.1.res_block.block1.block.3.weight", ...;
 * keys are relative to denoise_fn).  shape has up to 4 entries (OIHW for convs). */
int sr3_engine_num_params(const sr3_engine* e);
int sr3_engine_param_info(const sr3_engine* e, int index, char* name, int name_cap, int64_t shape[4], int* ndim);
/* load_state_dict for one tensor (model/model.py:146-160): `src` is a DEVICE fp32 pointer in the reference layout.
 * Weights are re-packed (OIHW -> K-major bf16) by a kernel on `stream`. */
int sr3_engine_load_param(sr3_engine* e, const char* name, const float* src, int64_t numel, void* stream);
/* The whole state_dict in one call (srcs[i] = DEVICE fp32 pointer of parameter i in sr3_engine_param_info order) including the finalisation;
 * asynchronous on `stream`.  What the training loop calls after every optimizer step (model/model.py:58). */
int sr3_engine_load_all_params(sr3_engine* e, const float* const* srcs, int n, void* stream);
/* Must be called after the last load_param of a batch of updates (fuses bias vectors). */
int sr3_engine_finalize_params(sr3_engine* e, void* stream);

/* set_new_noise_schedule (model/sr3_modules/diffusion.py:92-139): HOST pointers to the fp32 buffers
 * sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
 * posterior_log_variance_clipped (each [T]) and the float64 sqrt_alphas_cumprod_prev ([T+1]). */
int sr3_engine_set_schedule(sr3_engine* e, int T, const float* sqrt_recip_ac, const float* sqrt_recipm1_ac, const float* post_coef1,
                            const float* post_coef2, const float* post_logvar, const double* sqrt_ac_prev, void* stream);

/* UNet.forward (model/sr3_modules/unet.py:235-259): x [B,in_channel,H,W], noise_level [B] (the reference's [B,1]) -> eps
 * [B,out_channel,H,W].  All DEVICE fp32. */
int sr3_unet_forward(sr3_engine* e, const float* x, const float* noise_level, float* eps, void* stream);

/* p_mean_variance (diffusion.py:151-167): x [B,3,H,W], condition_x [B,3,H,W] or NULL, integer t -> posterior mean
 * [B,3,H,W] (DEVICE) and the clipped log-variance (HOST scalar, may be NULL). */
int sr3_p_mean_variance(sr3_engine* e, const float* x, const float* condition_x, int t, int clip_denoised, float* mean, float* log_variance,
                        void* stream);

/* p_sample (diffusion.py:169-174): x_{t-1} = mean + exp(0.5 logvar) * z for t > 0.  `noise` (DEVICE [B,3,H,W]) replaces
 * torch.randn_like when non-NULL; otherwise z comes from Philox4x32-10 keyed by (seed, first_sample_index + b, pixel, t). */
int sr3_p_sample(sr3_engine* e, const float* x, const float* condition_x, int t, const float* noise, uint64_t seed,
                 uint64_t first_sample_index, float* x_prev, void* stream);

/* p_losses forward (diffusion.py:221-246) with the random draws injected: hr = x_in['HR'] [B,3,H,W], sr = x_in['SR'] (NULL when
 * unconditional), gamma [B] = continuous_sqrt_alpha_cumprod, noise [B,3,H,W] (all DEVICE fp32).  q_sample (diffusion.py:212-219),
 * the UNet and the summed L1 (loss_type 1) / L2 (2) loss (diffusion.py:84-90) run natively; *loss_host receives the scalar.
 * Forward value only: the backward pass is not implemented. */
int sr3_p_losses(sr3_engine* e, const float* hr, const float* sr, const float* gamma, const float* noise, int loss_type, double* loss_host,
                 void* stream);

/* ---- training step: DDPM.optimize_parameters (model/model.py:48-58) = zero_grad -> GaussianDiffusion.forward / p_losses
 * (diffusion.py:221-249) in train() mode -> l_pix.sum() / (b c h w) -> backward -> Adam step (model.py:39-40).
 *
 * sr3_engine_create_train: as sr3_engine_create, for a fixed batch, with a plan that keeps every intermediate of the forward and has the
 * backward recorded next to it.  `dropout` = opt['model']['unet']['dropout'] (nn.Dropout in block2 of every ResnetBlock, unet.py:86,100-101);
 * bf16 precision only.  Inference entry points keep working on such an engine (eval-mode semantics are NOT applied: use a plain engine). */
int sr3_engine_create_train(const sr3_unet_config* cfg, int batch, int device, float dropout, sr3_engine** out);
/* p_losses forward in training mode with the random draws injected (as sr3_p_losses): q_sample, UNet with Dropout (Philox keyed by
 * dropout_seed, or the masks given to sr3_train_set_dropout_mask), summed L1 / L2 loss -> *loss_host (may be NULL: no synchronisation). */
int sr3_train_forward(sr3_engine* e, const float* hr, const float* sr, const float* gamma, const float* noise, int loss_type, uint64_t dropout_seed,
                      double* loss_host, void* stream);
/* loss.backward() for the forward that just ran: grads[i] (DEVICE fp32, reference layout, one pointer per parameter in
 * sr3_engine_param_info order, n_grads == sr3_engine_num_params) is OVERWRITTEN with grad_scale * d(summed loss)/d(parameter i)
 * (grad_scale = 1 / (b c h w) reproduces model.py:50-53).  One backward per forward. */
int sr3_train_backward(sr3_engine* e, float grad_scale, float* const* grads, int n_grads, void* stream);
/* The same backward layer by layer, so that the caller can overlap the gradient all-reduce (training row) with the layers still to
 * come: begin -> block n-1, n-2, ..., 0 -> finish (FiLM projections + noise-level MLP).  sr3_train_block_params lists the parameters whose
 * gradient is final once `block` has run AND sr3_train_backward_flush has been called (conv weight gradients leave the tensor-core kernel as
 * per-slice partial tiles; one table-driven launch per flush sums them) -- the rest (noise_func / block1 conv bias / noise_level_mlp) are
 * final after finish. */
int sr3_train_num_backward_blocks(const sr3_engine* e);
int sr3_train_backward_begin(sr3_engine* e, float grad_scale, float* const* grads, int n_grads);
int sr3_train_backward_block(sr3_engine* e, int block, void* stream);
int sr3_train_backward_flush(sr3_engine* e, void* stream);    /* reduce the weight-gradient partial tiles produced since the last flush (one launch) */
int sr3_train_backward_finish(sr3_engine* e, void* stream);
int sr3_train_block_params(const sr3_engine* e, int block, int* indices, int cap, int* n);
/* Profiling: the whole backward with CUDA events around every op; ms_by_kind[8]: device time per op kind (0 data-gradient tile kernel,
 * 1 GroupNorm / elementwise, 4 other, 6 weight gradient + slice reduction, 7 attention GEMMs). */
int sr3_train_backward_profile(sr3_engine* e, float grad_scale, float* const* grads, int n_grads, float* ms_by_kind, void* stream);
/* Tests: replace the Philox dropout mask of ResnetBlock `block_name` ("downs.1.res_block.block2", the module that owns the nn.Dropout) by a
 * keep-mask, uint8 DEVICE [B][C][H][W] (1 = keep), e.g. the one the reference drew; NULL restores Philox. */
int sr3_train_set_dropout_mask(sr3_engine* e, const char* block_name, const unsigned char* mask_nchw);
int sr3_train_num_dropout_layers(const sr3_engine* e);
int sr3_train_dropout_layer_name(const sr3_engine* e, int index, char* name, int name_cap);
/* torch.optim.Adam(lr, betas, eps, weight_decay 0) step `step` (1-based) over a DEVICE table of n_tensors records
 * {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64 numel} (40 bytes each) in ONE launch; gradients are multiplied by
 * grad_scale first (1 / world_size after a summing all-reduce). */
int sr3_adam_step(const void* table_dev, int n_tensors, float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* p_sample_loop / super_resolution / sample (diffusion.py:176-210): runs t = T-1 .. 0 as T launches of one captured CUDA
 * graph.  x_T: DEVICE [B,3,H,W] initial noise (the reference's torch.randn(shape)).  noises: optional DEVICE
 * [T][B,3,H,W], noises[i] used at step i.  Every (i % (1|T/10) == 0) the image is appended to `snapshots`
 * (DEVICE [n][B,3,H,W], capacity snapshot_cap images-batches; may be NULL) -- the `continous=True` return value without
 * its first B rows.  final: DEVICE [B,3,H,W] = x_0.  *n_snapshots receives the count. */
int sr3_p_sample_loop(sr3_engine* e, const float* condition_x, const float* x_T, const float* noises, uint64_t seed,
                      uint64_t first_sample_index, float* final, float* snapshots, int snapshot_cap, int* n_snapshots, void* stream);

/* Same loop on HOST buffers (H2D of condition_x / x_T, D2H of final inside): the end-to-end entry point. */
int sr3_super_resolution_host(sr3_engine* e, const float* condition_x_host, const float* x_T_host, uint64_t seed,
                              uint64_t first_sample_index, float* final_host, void* stream);

/* Run `steps` reverse steps starting at timestep t_start on the engine's resident state (benchmark / profiling hook:
 * no copies, no snapshots).  State must have been initialised by sr3_p_sample_loop_begin. */
int sr3_p_sample_loop_begin(sr3_engine* e, const float* condition_x, const float* x_T, uint64_t seed, uint64_t first_sample_index, void* stream);
int sr3_p_sample_steps(sr3_engine* e, int t_start, int steps, void* stream);
int sr3_read_state(sr3_engine* e, float* x_out, void* stream);

/* core/metrics.py:8-34 `tensor2img` on the device: src fp32 DEVICE [n][C][H][W] -> clamp to [min_v, max_v] -> [0, 1] -> * 255, round half
 * to even -> uint8 DEVICE, HWC.  n == 1: dst [H][W][C].  n > 1: the images are tiled like torchvision.utils.make_grid(nrow, padding 2,
 * pad_value 0), which is what the reference does for 4-D input: dst [rows*(H+2)+2][cols*(W+2)+2][C], cols = min(nrow, n).
 * Saves the D2H of fp32 snapshots (model.py:98-110 get_current_visuals + sr.py): a quarter of the bytes cross PCIe. */
int sr3_tensor2img(const float* src, unsigned char* dst_u8, int n, int C, int H, int W, int nrow, float min_v, float max_v, void* stream);
/* core/metrics.py:42-50 `calculate_psnr`: exact integer sum of squared differences of two uint8 DEVICE images (n elements) -> *ssd_host;
 * PSNR = 20 log10(255 / sqrt(ssd / n)) is formed by the caller in float64 as the reference does. */
int sr3_ssd_u8(const unsigned char* a_u8, const unsigned char* b_u8, int64_t n, unsigned long long* ssd_host, void* stream);

/* Entrance of the path: the conditioning image.  data/prepare_data.py:17-40 `trans_fn.resize(img, size, Image.BICUBIC)` (Pillow's two-pass
 * fixed-point bicubic resampler on uint8) + data/util.py:74-83 `transform_augment` (ToTensor, optional horizontal flip, range mapping).
 * src uint8 DEVICE [B][h][w][C] (HWC); dst_u8 (optional) uint8 DEVICE [B][H][W][C]; dst_f32 (optional) fp32 DEVICE [B][C][H][W] =
 * (resized / 255) * (max_v - min_v) + min_v, mirrored along W when flip != 0.  Integer-exact against Pillow 12. */
/* Host-only helper: Pillow's integer coefficient tables of one bicubic pass in_size -> out_size (bounds [out][2] = first tap, tap count;
 * coef [out][ksize], 22 fractional bits). */
int sr3_pil_bicubic_tables(int in_size, int out_size, int* bounds, int* coef, int coef_cap, int* ksize);
int sr3_resize_bicubic_u8(const unsigned char* src_u8, unsigned char* dst_u8, float* dst_f32, int B, int h, int w, int C, int H, int W, int flip,
                          float min_v, float max_v, void* stream);

/* Introspection for tests / bench. */
int sr3_engine_num_launches_per_step(const sr3_engine* e);   /* kernel launches per reverse step: 1 with the persistent step kernel */
int sr3_engine_num_ops_per_step(const sr3_engine* e);        /* launches of the per-layer path (the default; also what sr3_engine_profile_step times) */
/* 1 when one reverse step (reference p_sample: diffusion.py:151-174, UNet.forward unet.py:235-259) runs as ONE persistent
 * cooperative launch (csrc/step_megakernel.cuh), 0 when it runs as a CUDA graph of per-layer launches (the default; SR3_MEGA=1 selects the step kernel). */
int sr3_engine_uses_step_kernel(const sr3_engine* e);
/* Per-op device time (us) of the most recent step-kernel launch: globaltimer stamps taken by CTA 0 after each grid barrier.
 * types: 0 tensor-core tile loop, 1 GroupNorm apply, 2 fused attention, 3 row softmax, 4 embedding + FiLM, 5 statistics clear. */
int sr3_engine_step_kernel_profile(sr3_engine* e, int cap, int* types, double* us, double* phases_or_null, int* n_ops, void* stream);
/* phases (optional, [cap][4] us): per op, time CTA 0 spent in set-up (barrier arrive + parameter / stage-table copy), waiting at the grid
 * barrier, in the op body, and in the end-of-op fence. */
int64_t sr3_engine_workspace_bytes(const sr3_engine* e);
/* Per-kernel timing of one eager (non-graph) reverse step at timestep t, averaged over `reps` repetitions after one warm-up,
 * CUDA events on `stream` around every launch.  kinds: 0 tensor-core tile kernel, 1 GroupNorm apply, 2 cast/upsample,
 * 3 softmax, 4 other; flops / bytes are the executed work of each launch.  Does not modify the sampler state. */
int sr3_engine_profile_step(sr3_engine* e, int t, int reps, int cap, int* kinds, float* ms, double* flops, double* bytes, int* n_ops,
                            void* stream);
/* Debug tap: copy the fp32 NHWC output of top-level layer `name` ("downs.3", "mid.0", ...) of the last forward to dst
 * (DEVICE, [B,H,W,C]); returns C*H*W*B through *numel. */
int sr3_engine_read_activation(sr3_engine* e, const char* name, float* dst, int64_t cap, int64_t* numel, int shape_bhwc[4], void* stream);

/* Stand-alone tile GEMM for unit tests: D[M,N] = A[M,K] * B[N,K]^T (bf16 row-major DEVICE inputs, fp32 output), M%128==0,
 * K%64==0, N%block_n==0. */
int sr3_test_gemm(const void* a_bf16, const void* b_bf16, float* d, int M, int N, int K, int block_n, void* stream);
/* Test hook for the fused attention core (S = q k^T / sqrt(C), softmax per image, O = P v; unet.py:129-139): qk bf16 [nz*Lt][2C]
 * (q | k), vT bf16 [nz*C][Lt], out bf16 [nz*Lt][C]; Lt in {128, 256} keys per attention batch, HW tokens per image (Lt % HW == 0). */
int sr3_test_attention(const void* qk_bf16, const void* vT_bf16, void* out_bf16, int nz, int Lt, int HW, int C, void* stream);
/* Stand-alone NHWC conv for unit tests: x bf16 [B,H,W,Cin], w fp32 OIHW [Cout,Cin,k,k] (k in {1,3}), stride in {1,2},
 * y fp32 [B,OH,OW,Cout]; stats (optional) fp64 [B,Cout,2] (sum, sum of squares per image and channel: the GroupNorm statistics of
 * unet.py:84, accumulated with order-independent fp64 atomics) must be zeroed by the caller. */
int sr3_test_conv(const void* x_bf16, const float* w_oihw, const float* bias, float* y, double* stats, int B, int H, int W, int Cin,
                  int Cout, int ksize, int stride, void* stream);

/* Test hook for the GroupNorm path of a Block (unet.py:80-91): y = conv(x) + bias with the statistics taken in the conv epilogue, then
 * a = [silu](GroupNorm(y; groups, gamma, beta, eps 1e-5)) as bf16 NHWC.  Shapes as sr3_test_conv (stride 1). */
int sr3_test_conv_groupnorm(const void* x_bf16, const float* w_oihw, const float* bias, const float* gamma, const float* beta, int groups,
                            int silu, float* y, void* a_bf16, int B, int H, int W, int Cin, int Cout, int ksize, void* stream);

/* Timing harness for one conv shape on zero-filled buffers (kernel-tuning experiments): average ms over `reps` launches. */
int sr3_bench_conv(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int with_resid, int with_stats, int reps, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SR3_B200_H */
