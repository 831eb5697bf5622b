/* i2v_b200.h -- C-ABI of the B200-native image->video sampling path.
 *
 * The reference (CompVis/image2video-synthesis-using-cINNs) is pure Python/PyTorch and has no
 * FFI; the boundary it offers is the Python class get_model.Model (get_model.py:10-103) and the
 * nn.Module seams below it.  Each entry point here replaces the arithmetic of one of those seams
 * and is what a reference-side binding (ctypes) would call:
 *
 *   i2v_flow_reverse / i2v_flow_forward   <- ConditionalFlow.forward(x, embedding, reverse)
 *                                            stage2_cINN/modules/flow_blocks.py:31-57
 *                                            (called from SupervisedTransformer.forward, INN.py:59-73)
 *   i2v_embedder_forward                  <- ResnetEncoder.encode(x).mode()
 *                                            stage2_cINN/AE/modules/AE.py:126-166, distributions.py:41
 *   i2v_decoder_forward                   <- Generator.forward(img, motion)
 *                                            stage1_VAE/modules/decoder.py:97-120
 *   i2v_encoder3d_forward                 <- Encoder.forward(x) -> (mu | logvar); get_model.py:87 consumes
 *                                            mu only.  stage1_VAE/modules/resnet3D.py:202-219
 *   i2v_op_*                              <- the individual kernels, exported for parity tests
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer named dev_* is a CUDA
 * device pointer on the current device; `stream` is a cudaStream_t passed as void*; no entry point
 * allocates device memory or synchronises -- the caller provides a workspace of at least
 * i2v_*_workspace_bytes() bytes (256-byte aligned) that must stay alive until the stream has
 * drained.  Weights are registered by name with i2v_*_set_tensor (layouts in DESIGN.md section 3;
 * packed from the reference's state-dicts by loader.py) and are NOT copied: the caller keeps them
 * alive for the life of the handle.  All functions returning int return 0 on success and a negative
 * code on failure, with a message available from i2v_last_error() (thread-local).  Nothing throws
 * across the ABI.  Handles are not thread-safe; use one per host thread / stream.
 */
#ifndef I2V_B200_H_
#define I2V_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I2V_ABI_VERSION 1

int i2v_abi_version(void);
const char* i2v_last_error(void);

/* Launch accounting.  i2v_launch_count(): kernels launched by this library since load.
 * i2v_prof_enable(1) brackets every subsequent launch with CUDA events on its stream;
 * i2v_prof_collect() waits for them and returns per kernel family (0 halo tensor-core conv, 1 stats,
 * 2 modulate, 3 flow, 4 other, 5 per-tap tensor-core conv, 6 fp32 SIMT conv; arrays of 7) the summed
 * device time [ms], algorithmic FLOPs, algorithmic bytes and launch counts since the last collect. */
long long i2v_launch_count(void);
void i2v_prof_enable(int on);
int i2v_prof_is_enabled(void);
int i2v_prof_collect(double* ms, double* flops, double* bytes, long long* launches);
/* when set (non-NULL, non-empty) i2v_prof_collect also writes one CSV row per launch to this path */
void i2v_prof_dump_path(const char* path);
/* Process-wide tuning switches for A/B measurements (defaults = the measured optima): "pdl" (programmatic dependent
 * launch), "tc_flags", "tc_persist", "tc_min_stages", "tc_pair" (2-CTA conv tiles), "linear_bfly", "flow_cluster".
 * None changes a result beyond summation order.  Returns -2 for an unknown name. */
int i2v_set_option(const char* name, double value);

/* ---------------------------------------------------------------- conditional INN (stage 2) */
typedef struct i2v_flow i2v_flow;
/* d: latent width (64); zc: conditioning width padded to a multiple of 4; hidden: MLP width (<=512);
 * depth: hidden->hidden layers per MLP (2); cond_mode[n_flows]: 1 where the block's subnets see only
 * the conditioning (flow_blocks.py:24, control=True and fl % 4 != 0), may be NULL (= all 0). */
i2v_flow* i2v_flow_create(int n_flows, int d, int zc, int hidden, int depth, const unsigned char* cond_mode);
int i2v_flow_set_tensor(i2v_flow* h, const char* name, const void* dev_ptr, size_t nbytes);
size_t i2v_flow_workspace_bytes(const i2v_flow* h, int batch);
/* z[B,d] = flow^-1(residual[B,d] | cond[B,zc])            (Model.forward, get_model.py:65) */
int i2v_flow_reverse(i2v_flow* h, const float* dev_residual, const float* dev_cond, float* dev_z, int batch,
                     void* dev_ws, size_t ws_bytes, void* stream);
/* out[B,d], logdet[B] = flow(z[B,d] | cond[B,zc])         (Model.transfer, get_model.py:90) */
int i2v_flow_forward(i2v_flow* h, const float* dev_z, const float* dev_cond, float* dev_out, float* dev_logdet,
                     int batch, void* dev_ws, size_t ws_bytes, void* stream);
void i2v_flow_destroy(i2v_flow* h);

/* ---------------------------------------------------------------- start-frame embedder */
typedef struct i2v_embedder i2v_embedder;
/* norm_mode: 0 = InstanceNorm2d (norm "in"), 1 = BatchNorm folded into conv weights+bias at load */
i2v_embedder* i2v_embedder_create(int zc, int norm_mode);
int i2v_embedder_set_tensor(i2v_embedder* h, const char* name, const void* dev_ptr, size_t nbytes);
/* options: "tc_mode" = 0 fp32 SIMT convs only, 1 (default) tensor-core convs (error-compensated fp16 split, fp32-grade)
 * for the stride-1 convs of the InstanceNorm variant whose GEMM fills the machine, 2 wherever the shape is supported;
 * "tc_min_ctas" = fewest 128x128 output tiles for which mode 1 picks the tensor-core engine (default 8: measured, B = 64) */
int i2v_embedder_set_scalar(i2v_embedder* h, const char* name, double value);
size_t i2v_embedder_workspace_bytes(const i2v_embedder* h, int batch, int height, int width);
/* x0: [B,3,H,W] fp32 NCHW in [-1,1]  ->  embed: [B, zc] (the posterior mean) */
int i2v_embedder_forward(i2v_embedder* h, const float* dev_x0, float* dev_embed, int batch, int height, int width,
                         void* dev_ws, size_t ws_bytes, void* stream);
void i2v_embedder_destroy(i2v_embedder* h);

/* ---------------------------------------------------------------- 3-D conv decoder (stage 1) */
typedef struct i2v_decoder i2v_decoder;
/* nf: Decoder.channel_factor; upsample_s/t: the two config lists (decoder.py:65-66).
 * conv_engine: 0 = fp32 SIMT implicit GEMM everywhere (exact-arithmetic engine);
 *              1 = tcgen05 tensor-core engine, error-compensated fp16 split (hi*hi+hi*lo+lo*hi): fp32-grade parity;
 *              2 = tcgen05 tensor-core engine, single fp16 product (fast mode, ~1e-3). */
i2v_decoder* i2v_decoder_create(int nf, int z_dim, const int upsample_s[2], const int upsample_t[2], int conv_engine);
int i2v_decoder_set_tensor(i2v_decoder* h, const char* name, const void* dev_ptr, size_t nbytes);
/* host-side per-layer constants of the tensor-core engine ("<block>.spade.sa": split scale of SPADE's hidden map) */
int i2v_decoder_set_scalar(i2v_decoder* h, const char* name, double value);
size_t i2v_decoder_workspace_bytes(const i2v_decoder* h, int batch, int height, int width);
/* img: [B,3,H,W] NCHW; z: [B,z_dim]; frames: [B,16,3,H,W] (contiguous; the reference returns the
 * same values as a transposed view, decoder.py:120) */
int i2v_decoder_forward(i2v_decoder* h, const float* dev_img, const float* dev_z, float* dev_frames, int batch,
                        int height, int width, void* dev_ws, size_t ws_bytes, void* stream);
void i2v_decoder_destroy(i2v_decoder* h);

/* ---------------------------------------------------------------- 3-D video encoder (transfer) */
typedef struct i2v_encoder3d i2v_encoder3d;
i2v_encoder3d* i2v_encoder3d_create(const int channels[5], const int stride_s[4], const int stride_t[4], int z_dim);
int i2v_encoder3d_set_tensor(i2v_encoder3d* h, const char* name, const void* dev_ptr, size_t nbytes);
/* "tc_mode": 0 fp32 SIMT convs only, 1 (default) tensor-core engine for the stride-1 3x3x3 convs (resnet3D.py:101-135) whose GEMM
 * has >= "tc_min_ctas" (default 8) 128x128 tiles, 2 wherever the shape is supported; needs the "<conv>.wh/.wl/.ws" tensors */
int i2v_encoder3d_set_scalar(i2v_encoder3d* h, const char* name, double value);
size_t i2v_encoder3d_workspace_bytes(const i2v_encoder3d* h, int batch, int frames, int height, int width);
/* seq: [B,T,3,H,W] fp32 (the query clip without its first frame, get_model.py:87)
 * -> mu_logvar [B, 2*z_dim] = (conv_mu | conv_var) outputs, resnet3D.py:202-203 */
int i2v_encoder3d_forward(i2v_encoder3d* h, const float* dev_seq, float* dev_mu_logvar, int batch, int frames, int height,
                          int width, void* dev_ws, size_t ws_bytes, void* stream);
void i2v_encoder3d_destroy(i2v_encoder3d* h);

/* ---------------------------------------------------------------- single kernels (parity tests) */
/* channels-last convolution, weights [taps,Cout,Cin]; res may be NULL; act: 0 none 1 relu 2 lrelu(0.2)
 * 3 tanh; out_mode 0: [B,To,Ho,Wo,Cout], 1: [B,To,Cout,Ho,Wo]; engine 0 = fp32 SIMT */
int i2v_op_conv(const float* dev_x, const float* dev_w, const float* dev_bias, const float* dev_res, float* dev_y,
                int B, int Ti, int Hi, int Wi, int Cin, int Cout, int kt, int kh, int kw, int st, int sh, int sw,
                int pt, int ph, int pw, int res_ut, int res_uh, int res_uw, int act, int out_mode, int engine,
                void* stream);
/* SPADE's Conv2d(3 -> 128, k3, p1) + act (normalization_layer.py:13,21) on img [B,H,W,3], weights [9,128,3]; result as the
 * fp16 split hi = fp16(s*v), lo = fp16(s*v - hi), each [B,H,W,128] */
int i2v_op_spade_conv3(const float* dev_img, const float* dev_w, const float* dev_bias, void* dev_y_hi, void* dev_y_lo,
                       float split_scale, int B, int H, int W, int act, void* stream);
/* tensor-core conv on fp32 inputs: splits x (scale_a) and w (scale_w; [taps,cout_pad,Cin], rows >= Cout zero) into
 * fp16 (hi, lo) inside the workspace (>= 4*(|x|+|w|)+2048 bytes), then runs the tcgen05 engine; terms 3 or 1;
 * variant 0 = auto, 1 = per-tap box kernel, 2 = 256-row H-halo kernel without kw stacking, 3 = H-halo kernel with the three
 * kw taps stacked along N (Cout <= 64), 4 = CTA-pair kernel (cta_group::2 tiles, two TMEM accumulator sets); 2, 3 and 4 fail
 * if the shape is not eligible */
int i2v_op_conv_tc(const float* dev_x, const float* dev_w, const float* dev_bias, const float* dev_res, float* dev_y,
                   int B, int T, int H, int W, int Cin, int Cout, int cout_pad, int kt, int kh, int kw, int res_ut,
                   int res_uh, int res_uw, int act, int out_mode, int terms, int variant, float scale_a, float scale_w,
                   void* dev_ws, size_t ws_bytes, void* stream);
/* Temporal phase form (conv_0 of a GeneratorBlock behind a x2 temporal nearest upsample, decoder.py:102-111): x is the
 * PRE-upsample tensor [B,T/2,H,W,Cin], w the phase-combined 3x3x3 weights [2 phases][2 taps][3][3][cout_pad][Cin]
 * (loader.phase_weights); y [B,T,H,W,Cout] equals conv3d(repeat_interleave(x, 2, dim=T), w_original) + bias. */
int i2v_op_conv_tc_phase(const float* dev_x, const float* dev_w, const float* dev_bias, float* dev_y, int B, int T, int H, int W,
                         int Cin, int Cout, int cout_pad, int terms, int variant, float scale_a, float scale_w, void* dev_ws,
                         size_t ws_bytes, void* stream);
/* 3x3x3 tensor-core conv with a side input: y = conv3x3x3(x, w) + conv1x1x1(x2, w2) + bias as ONE implicit GEMM (the
 * GeneratorBlock's learned shortcut fused into conv_1, decoder.py:44-50).  w2 is [3 (kw)][cout_pad][Cin2] with only
 * the kw = 1 slab non-zero. */
int i2v_op_conv_tc_side(const float* dev_x, const float* dev_w, const float* dev_x2, const float* dev_w2, const float* dev_bias,
                        float* dev_y, int B, int T, int H, int W, int Cin, int Cin2, int Cout, int cout_pad, int act,
                        int out_mode, int terms, int variant, float scale_a, float scale_w, void* ws, size_t ws_bytes,
                        void* stream);
/* profiling aid: when dev_buf != NULL the halo conv kernel writes 8 x uint64 %globaltimer stamps per CTA (first `ctas`
 * CTAs of grid row 0): 0 start, 1 prologue done, 2 first stage landed, 3 last MMA issued, 4 accumulators complete,
 * 5 epilogue stores issued, 6 CTA end */
int i2v_debug_conv_tc_timestamps(void* dev_buf, int ctas);
/* profiling aid: 2 x 16 uint64 %globaltimer stamps of the flow kernel's 5th coupling (CTA 0 and CTA 100): 0 start, 1/2 around
 * the barrier after layer 1, 3/4 and 5/6 around the barriers after the hidden layers, 7/8 after the last layer, 9 end */
int i2v_debug_flow_timestamps(void* dev_buf);
/* sums[B,C,2] (double) of x[B,V,C] */
int i2v_op_channel_stats(const float* dev_x, double* dev_sums, int B, int64_t V, int C, void* stream);
int i2v_op_norm_coeffs(const double* dev_sums, float* dev_coef, int B, int C, int64_t V, int groups, float eps,
                       const float* dev_gamma, const float* dev_beta, const float* dev_mod, void* stream);
int i2v_op_modulate(const float* dev_x, const float* dev_coef, const float* dev_gb, const float* dev_r,
                    const float* dev_coef2, float* dev_out, int B, int T, int H, int W, int C, int ut, int uh, int uw,
                    int act, void* stream);
/* The same pass writing the conv-ready fp16 pair of the tensor-core engine: hi = fp16(s*v), lo = fp16(s*v - hi).
 * Optional second result of the same read (no upsampling): outb = split(s * (coef_b.A * x + coef_b.B)). */
int i2v_op_modulate_split(const float* dev_x, const float* dev_coef, const float* dev_gb, void* dev_out_hi, void* dev_out_lo,
                          int B, int T, int H, int W, int C, int ut, int uh, int uw, int act, float split_scale,
                          const float* dev_coef_b, void* dev_outb_hi, void* dev_outb_lo, void* stream);
int i2v_op_linear(const float* dev_x, const float* dev_w, const float* dev_bias, float* dev_y, int B, int K, int N,
                  int act, void* stream);
int i2v_op_resize_bilinear(const float* dev_img, float* dev_out, int B, int C, int H0, int W0, int H, int W,
                           void* stream);
int i2v_op_maxpool3x3s2(const float* dev_x, float* dev_y, int B, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------- CLI pre/post-processing */
/* Start frame as generate_samples.py:36-41 prepares it: uint8 HWC image (bgr != 0: cv2.imread channel order) -> RGB ->
 * /255 -> Normalize(0.5, 0.5) -> bilinear resize to HxW (align_corners=False) -> fp32 [3,H,W] (one slot of x_0). */
int i2v_op_preprocess_u8(const unsigned char* dev_img_hwc, float* dev_out_chw, int H0, int W0, int H, int W, int bgr, void* stream);
/* *dev_max = max over the n floats of denorm(x) = clamp((x+1)/2, 0, 1)   (utils/auxiliaries.py:21,53-55) */
int i2v_op_frames_max(const float* dev_frames, float* dev_max, int64_t n, void* stream);
/* frames [N,T,3,H,W] -> uint8 RGB at out[n*stride_n + t*stride_t + h*stride_h + w*3 + c] = trunc(255*denorm(x) / *dev_max)
 * (utils/auxiliaries.py:15-22 + .astype(uint8), generate_samples.py:61).  GIF canvas [T,H,N*W,3]: stride_n = 3W,
 * stride_h = 3NW, stride_t = 3HNW; per-video [N,T,H,W,3]: stride_n = 3THW, stride_t = 3HW, stride_h = 3W. */
int i2v_op_frames_to_u8(const float* dev_frames, const float* dev_max, unsigned char* dev_out, int N, int T, int H, int W,
                        int64_t stride_n, int64_t stride_t, int64_t stride_h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* I2V_B200_H_ */
