// Internal kernel launchers (C++); the public C-ABI lives in include/i2v_b200.h.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstddef>

namespace i2v {

// Process-wide tuning switches (i2v_set_option; defaults are the measured optima, profiles/*_bench_ab*.txt).  They
// exist for A/B measurements and never change results beyond summation order.
struct TuneOptions {
    int pdl = 1;             // programmatic dependent launch attribute on every kernel
    int tc_flags = 1;        // halo conv kernel: bit 0 = residual L2 prefetch from the idle epilogue warps
    int tc_persist = 1;      // halo conv kernel: persistent tile loop (0: one CTA per tile)
    int tc_min_stages = 2;   // halo conv kernel: pipeline depth aimed for when picking the channel chunk (2..6)
    int tc_pair = 1;         // conv engine: CTA-pair (cta_group::2) tiles where the layer is eligible
    int tc_pair_stack = 1;   // kw-stacked narrow layers on the CTA-pair kernel: 0 never, 1 frame-layout output only (conv_img), 2 all
    int tc_pair_stages = 0;  // CTA-pair kernel: cap on the ring depth (0 = as many as fit, up to 8)
    int linear_bfly = 1;     // linear kernel: 9-shuffle transpose-reduce (0: 8 x warp_sum)
    int linear_k64 = 1;      // Linear: thread-per-feature kernel for K = 64, N >= 8192 (0: warp-per-feature kernel everywhere)
    int flow_cluster = 1;    // flow: cluster-resident kernel where eligible (0: cooperative grid-barrier kernel)
    int tc_t2_split = 1;     // per-tap conv kernel: two-frame clips under a temporal 3-tap kernel run as one-frame tiles (skip the padded tap)
    int mod_spade = 1;       // SPADE modulate passes: pipelined fine-grained kernel (0: generic T-walking kernel)
};
TuneOptions& tune();

// ----------------------------------------------------------------------------- convolution
struct ConvArgs {
    const float* x;      // [B, Ti, Hi, Wi, Cin]
    const float* w;      // [taps, Cout, Cin]
    const float* bias;   // [Cout] or nullptr
    const float* res;    // residual [B, To/res_ut, Ho/res_uh, Wo/res_uw, Cout] or nullptr
    float* y;
    int B, Ti, Hi, Wi, Cin;
    int To, Ho, Wo, Cout;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int res_ut, res_uh, res_uw;
    int act;        // i2v::Act
    int out_mode;   // 0: [B,To,Ho,Wo,Cout]   1: [B,To,Cout,Ho,Wo] (video frames)
    // optional fp16 split of the result for the tensor-core engine: hi = fp16(s*v), lo = fp16(s*v - hi)
    __half* y_hi = nullptr; __half* y_lo = nullptr; float split_scale = 1.f;
    // optional split-K scratch (partials [tiles][ksplit][128*64] fp32 + one zero-initialised counter per tile)
    float* splitk_ws = nullptr; size_t splitk_ws_bytes = 0; unsigned* splitk_counters = nullptr; int splitk_max_tiles = 0;
};
int launch_conv_simt(const ConvArgs& a, cudaStream_t stream);

// Tensor-core engine (conv_tc.cu): stride-1 'same' convolution on pre-split fp16 operands.
struct ConvTcArgs {
    const __half* x_hi; const __half* x_lo;   // [B,T,H,W,Cin]   activations * s_a, split
    const __half* w_hi; const __half* w_lo;   // [taps,cout_pad,Cin] weights * s_w, split (rows >= Cout are zero)
    const float* scale_ptr;                   // device scalar 1/(s_a*s_w)
    const float* bias; const float* res; float* y;
    double* stats = nullptr;                  // optional fused output statistics [B, Cout, 2] (must be zeroed by the caller)
    int B, T, H, W, Cin, Cout, cout_pad;
    int kt, kh, kw;
    int res_ut, res_uh, res_uw, act, out_mode;
    int terms;                                // 3: hi*hi+hi*lo+lo*hi (fp32-grade)   1: hi*hi only
    int variant = 0;                          // 0: auto  1: per-tap kernel  2: halo kernel, never kw-stacked  3: halo kernel, kw-stacked
    // 1: the logical input is x nearest-upsampled x2 in time; x_hi/x_lo hold it at T/2 planes and w_hi/w_lo hold the
    //    phase-combined weights [2 phases][2 taps][kh][kw][cout_pad][Cin] (halo kernel only)
    int t_phase = 0;
    // optional side input (fused 1x1x1 conv through the centre tap, halo kernel only; see ConvTcHArgs::cc2):
    // x2 [B,T,H,W,Cin2] split like x, w2 [3 (kw)][cout_pad][Cin2] split with the SAME scales (only the kw = 1 slab non-zero)
    const __half* x2_hi = nullptr; const __half* x2_lo = nullptr;
    const __half* w2_hi = nullptr; const __half* w2_lo = nullptr;
    int Cin2 = 0;
    // 1: run every eligibility check of the launch (tile / pipeline configuration, fused-statistics requirements) and return
    // without launching: lets a caller ask "can this conv fuse the statistics?" with the launcher's own predicate
    int dry_run = 0;
};
bool conv_tc_fuses_stats(int T, int H, int W);
bool conv_tc_halo_eligible(int H, int W, int kh);
bool conv_tc_side_eligible(int H, int W, int Cin, int Cin2, int cout_pad, int terms);   // fused 1x1x1 side input possible?
bool conv_tc_supported(int B, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw);
int launch_conv_tc(const ConvTcArgs& a, cudaStream_t stream);
int conv_tc_set_debug(unsigned long long* buf, int ctas);   // phase timestamps of the halo kernel (profiling aid)
int launch_split_fp16(const float* x, __half* hi, __half* lo, float scale, long long n, cudaStream_t stream);

// ----------------------------------------------------------------------------- normalisation
// Per-(sample, channel) sums of a channels-last tensor x[B, V, C] -> sums[B, C, 2] (double):
// the single statistics pass feeding InstanceNorm (per channel), GroupNorm (per 16-group) and the
// global average pool.
int launch_channel_stats(const float* x, double* sums, int B, long long V, int C, cudaStream_t stream);

// Turn sums into per-(b, c) affine coefficients  y = A*x + Bc  of the normalisation layer:
//   groups == 0 : instance norm (one group per channel);  groups > 0 : GroupNorm(groups)
//   gamma/beta  : optional per-channel affine [C] (GroupNorm affine=True)
//   mod         : optional per-sample modulation [B, 2C] = (gamma | beta) (AdaIN, gamma*x_hat+beta)
int launch_norm_coeffs(const double* sums, float* coef /*[B,C,2]*/, int B, int C, long long V, int groups,
                       float eps, const float* gamma, const float* beta, const float* mod, cudaStream_t stream);

// Fused element-wise pass producing a conv-ready tensor (channels-last):
//   v   = A[b,c] * x[b, t/ut, h/uh, w/uw, c] + Bc[b,c]              (coef may be null: v = x)
//   v   = v * (1 + g[b,h,w,c]) + bt[b,h,w,c]                         (SPADE maps gb[B,H,W,2C], optional)
//   v  += A2[b,c] * r[b,t,h,w,c] + B2[b,c]                           (second normalised branch, optional)
//   out = act(v)
struct ModArgs {
    const float* x; const float* coef;        // coef [B,C,2] or nullptr
    const float* gb;                          // [B,H,W,2C] or nullptr
    const float* r; const float* coef2;       // residual branch (same shape as out) + its coef or nullptr
    float* out;
    int B, T, H, W, C;                        // OUTPUT dims
    int ut, uh, uw;                           // nearest-upsample factors from x to out
    int act;
    // when out_hi != nullptr the result is written as an fp16 (hi, lo) split of split_scale*v instead of fp32
    __half* out_hi = nullptr; __half* out_lo = nullptr; float split_scale = 1.f;
    // with out_hi set: ALSO store the fp32 result here (a tensor that is both a conv operand and a later residual)
    float* out_f32 = nullptr;
    // optional second result of the same read of x (split output, no upsampling only): outb = coef_b.A * x + coef_b.B,
    // no maps, no activation -- the GroupNorm-affine input of a block's learned shortcut (decoder.py:44-50)
    const float* coef_b = nullptr; __half* outb_hi = nullptr; __half* outb_lo = nullptr;
};
int launch_modulate(const ModArgs& a, cudaStream_t stream);

// ----------------------------------------------------------------------------- small ops
// y[b, n] = act(sum_k W[n, k] x[b, k] + bias[n])      (nn.Linear layout)
int launch_linear(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act,
                  cudaStream_t stream);
// img [B,3,H0,W0] (NCHW) -> out [B,H,W,3] bilinear, align_corners=True (normalization_layer.py:20);
// H==H0 && W==W0 degenerates to an exact NCHW->NHWC repack.
int launch_resize_bilinear_nchw_to_nhwc(const float* img, float* out, int B, int C, int H0, int W0, int H, int W,
                                        cudaStream_t stream);
// SPADE's Conv2d(3 -> 128, k3, p1) + activation on img [B,H,W,3], result as the fp16 split of split_scale * v
// (y_hi / y_lo [B,H,W,128]); same bits as launch_conv_simt with y_hi set.  H*W must tile into 64-voxel patches.
bool spade_conv3_tiles(int H, int W);     // does the plane tile into the kernel's patches?
int launch_spade_conv3(const float* img, const float* w, const float* bias, __half* y_hi, __half* y_lo, float split_scale, int B,
                       int H, int W, int act, cudaStream_t stream);
// CLI pre/post-processing on the device (generate_samples.py:36-41,57-62; utils/auxiliaries.py:15-22,53-55)
int launch_preprocess_u8(const unsigned char* img_hwc, float* out_chw, int H0, int W0, int H, int W, int bgr, cudaStream_t stream);
int launch_frames_max(const float* frames, float* mx, long long n, cudaStream_t stream);
int launch_frames_to_u8(const float* frames, const float* mx, unsigned char* out, int N, int T, int H, int W, long long sn,
                        long long st, long long sh, cudaStream_t stream);
// 3x3 stride-2 pad-1 max pool, channels-last [B,H,W,C] -> [B,Ho,Wo,C]
int launch_maxpool3x3s2(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream);
// mean over V from channel sums: y[b,c] = sums[b,c,0] / V
int launch_mean_from_sums(const double* sums, float* y, int B, int C, long long V, cudaStream_t stream);

// ----------------------------------------------------------------------------- flow
struct FlowWeights {
    int n_flows, d, half, zc, hidden, depth;  // depth == 2 hidden->hidden layers
    const unsigned char* cond_mode;           // host array [n_flows]: 1 = 'cond' block (no x input)
    // packed device tensors (see loader.py::pack_flow)
    const float* w1x;    // [n_flows, 2, 2*hidden, half]   x-part of the first Linear (s rows, then t rows)
    const float* w1c;    // [n_flows*2*2*hidden, zc]       cond-part of the first Linear
    const float* b1;     // [n_flows*2*2*hidden]
    const float* wh;     // [n_flows, 2, depth, 2, hidden, hidden]   hidden Linears (s then t)
    const float* bh;     // [n_flows, 2, depth, 2*hidden]
    const float* wo;     // [n_flows, 2, 2*half, hidden]   last Linear (s rows then t rows)
    const float* bo;     // [n_flows, 2, 2*half]
    const float* loc;    // [n_flows, d]
    const float* scale;  // [n_flows, d]
    const int* perm_fwd; // [n_flows, d]
    const int* perm_bwd; // [n_flows, d]
    // optional: the x-part of the first Linear, the hidden Linears and the last Linear once more, cut into the per-CTA,
    // k-major chunks the cluster kernel streams (flow_cluster.cu; loader.pack_flow "wpack"); nullptr = cooperative kernel only
    const float* wpack = nullptr;
};
bool flow_cluster_eligible(const FlowWeights& fw);
int flow_cluster_set_debug(unsigned long long* buf);
// cluster-resident kernel; c1 = hoisted conditioning GEMM output.  Returns 1 if the device cannot schedule the cluster.
int launch_flow_cluster(const FlowWeights& fw, const float* in, const float* c1, float* out, float* logdet, int B, bool reverse,
                        cudaStream_t stream);
size_t flow_workspace_bytes(const FlowWeights& fw, int B);
int flow_set_debug(unsigned long long* buf);   // phase timestamps of coupling #4 (profiling aid)
// reverse: z = flow^-1(residual | cond);  forward: (out, logdet) = flow(z | cond)
int launch_flow(const FlowWeights& fw, const float* in, const float* cond, float* out, float* logdet, int B,
                bool reverse, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace i2v
