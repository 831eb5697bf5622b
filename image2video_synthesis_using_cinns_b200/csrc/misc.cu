// Small ops of the path: nn.Linear, bilinear start-frame resize, max-pool.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

namespace {

// nn.Linear y[b, n] = act(sum_k W[n, k] x[b, k] + bias[n]), K % 4 == 0: decoder.fc (decoder.py:99), AdaIN's Linear
// (normalization_layer.py:44,49), the conditioning half of every flow subnet's first Linear (flow_blocks.py:33-41), the
// embedder's 1x1 "fc" conv on the pooled feature (AE.py:121-124) and conv_mu on the flattened 4x4 map
// (resnet3D.py:179,203).  One warp per (output feature, group of 8 batch rows), lanes stride over k.  (One warp per feature
// walking all rows left the embedder fc -- K = 2048, N = 64 -- on 8 CTAs: 200 us, now 32.  A thread-per-feature variant with
// the batch in shared memory was measured slower for the shallow-K shapes: 91 us vs 27 for AdaIN's Linear.)
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, int B,
                                                     int K, int N, int act, int bfly) {
    pdl_launch_dependents();
    pdl_wait();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int groups = (B + 7) >> 3;
    const int n = gw / groups, b0 = (gw - n * groups) * 8;
    if (n >= N) return;
    const int K4 = K >> 2;
    const float4* wr = reinterpret_cast<const float4*>(w + (long long)n * K);
    const float bn = bias != nullptr ? __ldg(bias + n) : 0.f;
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int k = lane; k < K4; k += 32) {
        const float4 wv = __ldg(wr + k);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (b0 + r < B) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(b0 + r) * K) + k);
                acc[r] = fmaf(wv.x, xv.x, acc[r]); acc[r] = fmaf(wv.y, xv.y, acc[r]);
                acc[r] = fmaf(wv.z, xv.z, acc[r]); acc[r] = fmaf(wv.w, xv.w, acc[r]);
            }
        }
    }
    if (bfly) {
        // transpose-reduce: 9 shuffles instead of 8 x 5.  After the three exchange steps a lane holds ONE row's partial sum
        // (row = bits 4,3,2 of the lane index), the last two steps fold the lanes that share a row.
        const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
        float t[4], u[2];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float recv = __shfl_xor_sync(0xffffffffu, u16 ? acc[r] : acc[r + 4], 16);
            t[r] = (u16 ? acc[r + 4] : acc[r]) + recv;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float recv = __shfl_xor_sync(0xffffffffu, u8 ? t[r] : t[r + 2], 8);
            u[r] = (u8 ? t[r + 2] : t[r]) + recv;
        }
        float v = (u4 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, u4 ? u[0] : u[1], 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int r = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
        if ((lane & 3) == 0 && b0 + r < B) y[(long long)(b0 + r) * N + n] = apply_act(v + bn, act);
        return;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float s = warp_sum(acc[r]);
        if (lane == 0 && b0 + r < B) y[(long long)(b0 + r) * N + n] = apply_act(s + bn, act);
    }
}

// Shallow, wide Linears (K = 64, N >= 8192): the flow's hoisted conditioning GEMM (N = n_flows * 2 * 2 * hidden = 40 960 for
// BAIR) and the decoder's fc (N = 16 * 16 nf).  The warp-per-feature kernel above leaves half its lanes idle at K = 64 and lives
// for one dependent load -> FMA -> shuffle round per warp: 0.17 ms for 21 MB of traffic.  Here a thread owns a feature, keeps its
// 64 weights in registers and walks the batch rows, which sit in shared memory and are read as broadcasts: FMA-bound.
constexpr int LSK_K = 64, LSK_ROWS = 64, LSK_THREADS = 128;
__global__ void __launch_bounds__(LSK_THREADS) linear_k64_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y, int B, int N,
                                                                 int act) {
    __shared__ float4 xs[LSK_ROWS * (LSK_K / 4)];
    pdl_launch_dependents();
    pdl_wait();
    const int b0 = blockIdx.y * LSK_ROWS;
    const int nb = B - b0 < LSK_ROWS ? B - b0 : LSK_ROWS;
    const int nb4 = (nb + 3) & ~3;
    for (int i = threadIdx.x; i < nb4 * (LSK_K / 4); i += LSK_THREADS)
        xs[i] = i < nb * (LSK_K / 4) ? __ldg(reinterpret_cast<const float4*>(x + (long long)b0 * LSK_K) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = blockIdx.x * LSK_THREADS + threadIdx.x;
    const int nc = n < N ? n : N - 1;          // spare threads of the last CTA repeat its last feature (no divergent barrier)
    float4 wr[LSK_K / 4];
#pragma unroll
    for (int k = 0; k < LSK_K / 4; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w + (long long)nc * LSK_K) + k);
    const float bn = bias != nullptr ? __ldg(bias + nc) : 0.f;
    __syncthreads();
    for (int b = 0; b < nb4; b += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < LSK_K / 4; ++k) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 xv = xs[(b + r) * (LSK_K / 4) + k];
                acc[r] = fmaf(wr[k].x, xv.x, acc[r]); acc[r] = fmaf(wr[k].y, xv.y, acc[r]);
                acc[r] = fmaf(wr[k].z, xv.z, acc[r]); acc[r] = fmaf(wr[k].w, xv.w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (n < N && b + r < nb) y[(long long)(b0 + b + r) * N + n] = apply_act(acc[r] + bn, act);
    }
}

// SPADE's first conv (normalization_layer.py:13,21): Conv2d(3 -> 128, k3, p1) + LeakyReLU(0.2) on the resized start frame,
// written as the fp16 (hi, lo) split the gamma|beta conv consumes.  K = 27: as an implicit GEMM it is all overhead (the
// SIMT engine ran it at 6 TFLOP/s, 220 us per 64x64 block at B = 64).  Here a lane keeps the 27 x 4 weights of its 4
// output channels in registers, a warp owns a voxel (128 channels = one 256-byte row of each output word), the CTA's
// input patch sits in shared memory and is read as broadcasts: 108 FMAs per 27 shared loads, FMA-bound.
// Summation order = the SIMT engine's (taps outer, input channels inner, one fp32 FMA chain), so both produce the same bits.
// A CTA takes 256 voxels on planes that have them (64 below): the 108 weight loads per lane and the patch staging are paid once
// per 32 voxels of a warp instead of once per 8 (the kernel holds one CTA per SM and was latency-bound at 64 voxels: 0.20-0.25 ms
// per 64x64 block at B = 64 for 134 MB of output).
constexpr int SP_TILE = 256;      // most voxels per CTA
constexpr int SP_PATCH = 2400;    // floats: the largest (rows + 2) x (cols + 2) x 3 patch of a 256-voxel tile (1 x 256 voxels: 3 x 258 x 3)
static int spade_tile_voxels(int HW) { return HW >= SP_TILE ? SP_TILE : (HW < 64 ? HW : 64); }
__global__ void __launch_bounds__(256) spade_conv3_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                          const float* __restrict__ bias, __half* __restrict__ y_hi,
                                                          __half* __restrict__ y_lo, float scale, int H, int W, int act, int tile_v) {
    __shared__ float in_s[SP_PATCH];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int HW = H * W;
    const int v0 = blockIdx.x * tile_v;
    // patch: a segment of one row (W >= tile_v) or tile_v / W whole rows
    const int cols = W >= tile_v ? tile_v : W, rows = tile_v / cols;
    const int h0 = v0 / W, w0 = v0 - h0 * W;
    const int pc = cols + 2, pr = rows + 2;
    for (int i = threadIdx.x; i < pr * pc * 3; i += blockDim.x) {
        const int ch = i % 3, c = (i / 3) % pc, r = i / (3 * pc);
        const int hh = h0 - 1 + r, ww = w0 - 1 + c;
        in_s[i] = ((unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) ? __ldg(img + (((long long)b * H + hh) * W + ww) * 3 + ch) : 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n = lane * 4;
    float wreg[9][4][3];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) wreg[tap][q][c] = __ldg(w + ((long long)tap * 128 + n + q) * 3 + c);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
    __syncthreads();
    for (int j = warp; j < tile_v; j += 8) {
        const int r = j / cols, c = j - r * cols;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const float* xp = in_s + ((r + dh) * pc + c + dw) * 3;
                const float x0 = xp[0], x1 = xp[1], x2 = xp[2];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[q] = fmaf(x0, wreg[dh * 3 + dw][q][0], acc[q]);
                    acc[q] = fmaf(x1, wreg[dh * 3 + dw][q][1], acc[q]);
                    acc[q] = fmaf(x2, wreg[dh * 3 + dw][q][2], acc[q]);
                }
            }
        const float f[4] = {apply_act(acc[0] + b4.x, act) * scale, apply_act(acc[1] + b4.y, act) * scale,
                            apply_act(acc[2] + b4.z, act) * scale, apply_act(acc[3] + b4.w, act) * scale};
        __half hh[4], ll[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_f16(f[q], hh[q], ll[q]);
        const long long o = ((long long)b * HW + v0 + j) * 128 + n;
        *reinterpret_cast<uint2*>(y_hi + o) = *reinterpret_cast<const uint2*>(hh);
        *reinterpret_cast<uint2*>(y_lo + o) = *reinterpret_cast<const uint2*>(ll);
    }
}

__global__ void resize_bilinear_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int C, int H0,
                                       int W0, int H, int W, float sh, float sw) {
    pdl_launch_dependents();
    pdl_wait();
    const long long n = (long long)B * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        // align_corners=True: src = dst * (in-1)/(out-1)   (ATen area_pixel_compute_source_index)
        const float fy = sh * h, fx = sw * w;
        int y0 = (int)fy, x0 = (int)fx;
        y0 = min(y0, H0 - 1); x0 = min(x0, W0 - 1);
        const int y1 = min(y0 + 1, H0 - 1), x1 = min(x0 + 1, W0 - 1);
        const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
        const float* pc = img + ((long long)b * C + c) * H0 * W0;
        const float v = hy * (hx * __ldg(pc + (long long)y0 * W0 + x0) + lx * __ldg(pc + (long long)y0 * W0 + x1)) +
                        ly * (hx * __ldg(pc + (long long)y1 * W0 + x0) + lx * __ldg(pc + (long long)y1 * W0 + x1));
        out[i] = v;
    }
}

__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C,
                                    int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();
    const int C4 = C >> 2;
    const long long n = (long long)B * Ho * Wo * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long p = i / C4;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int dh = 0; dh < 3; ++dh) {
            const int hi = ho * 2 - 1 + dh;
            if ((unsigned)hi >= (unsigned)H) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int wi = wo * 2 - 1 + dw;
                if ((unsigned)wi >= (unsigned)W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(x) + (((long long)b * H + hi) * W + wi) * C4 + c4);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        reinterpret_cast<float4*>(y)[i] = m;
    }
}

// ---- CLI pre/post-processing on the device --------------------------------------------------------------
// Start frame: cv2.imread's uint8 HWC image -> RGB -> /255 -> Normalize(0.5, 0.5) -> Resize((S, S)) (bilinear,
// align_corners=False, no antialias: kornia 0.5's Resize is F.interpolate) -> one fp32 CHW slot of the batch tensor
// (generate_samples.py:36-41).  Source index and weights follow ATen's area_pixel_compute_source_index.
__global__ void preprocess_u8_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int H0, int W0, int H, int W,
                                     float sh, float sw, int bgr) {
    pdl_launch_dependents();
    pdl_wait();
    const int n = 3 * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int w = i % W, h = (i / W) % H, c = i / (W * H);
        const int cs = bgr ? 2 - c : c;
        float fy = sh * ((float)h + 0.5f) - 0.5f, fx = sw * ((float)w + 0.5f) - 0.5f;
        fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
        int y0 = (int)fy, x0 = (int)fx;
        y0 = min(y0, H0 - 1); x0 = min(x0, W0 - 1);
        const int y1 = y0 + (y0 < H0 - 1 ? 1 : 0), x1 = x0 + (x0 < W0 - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        auto px = [&](int y, int x) {
            const float t = __fdiv_rn((float)img[((long long)y * W0 + x) * 3 + cs], 255.f);
            return __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
        };
        out[i] = hy * (hx * px(y0, x0) + lx * px(y0, x1)) + ly * (hx * px(y1, x0) + lx * px(y1, x1));
    }
}

// max over the clip of denorm(x) = clamp((x + 1) / 2, 0, 1) (utils/auxiliaries.py:53-55): the normaliser of
// convert_seq2gif (utils/auxiliaries.py:21).  Values are >= 0, so their float bits order like ints.
__global__ void __launch_bounds__(256) frames_max_kernel(const float* __restrict__ x, float* __restrict__ mx, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __fadd_rn(__ldg(x + i), 1.f) * 0.5f;
        m = fmaxf(m, fminf(v, 1.f));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(mx), __float_as_int(m));
}

// frames [N,T,3,H,W] in [-1,1] -> uint8 RGB pixels at out[n*sn + t*st + h*sh + w*3 + c] = trunc(255 * denorm(x) / max)
// (utils/auxiliaries.py:15-22 + the .astype(np.uint8) of generate_samples.py:61; same fp32 operation order as numpy).
// Strides select the layout: the GIF canvas [T, H, N*W, 3] (videos side by side) or per-video [N, T, H, W, 3].
__global__ void __launch_bounds__(256) frames_to_u8_kernel(const float* __restrict__ x, const float* __restrict__ mx,
                                                           unsigned char* __restrict__ out, int N, int T, int H, int W, long long sn,
                                                           long long st, long long sh) {
    pdl_launch_dependents();
    pdl_wait();
    const float m = __ldg(mx);
    const long long HW = (long long)H * W, n_px = (long long)N * T * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        long long p = i / W;
        const int h = (int)(p % H); p /= H;
        const int t = (int)(p % T);
        const int n = (int)(p / T);
        const float* src = x + (((long long)n * T + t) * 3) * HW + (long long)h * W + w;
        unsigned char* dst = out + n * sn + t * st + h * sh + (long long)w * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fadd_rn(__ldg(src + c * HW), 1.f) * 0.5f;
            v = fminf(fmaxf(v, 0.f), 1.f);
            dst[c] = (unsigned char)(int)__fdiv_rn(__fmul_rn(255.f, v), m);
        }
    }
}

}  // namespace

int launch_preprocess_u8(const unsigned char* img, float* out, int H0, int W0, int H, int W, int bgr, cudaStream_t stream) {
    I2V_REQUIRE(H0 > 0 && W0 > 0 && H > 0 && W > 0, "preprocess_u8: empty image");
    // ATen area_pixel_compute_scale (align_corners=False, no explicit scale factor): in / out
    const float sh = (float)H0 / (float)H, sw = (float)W0 / (float)W;
    const int n = 3 * H * W;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    I2V_CHECK_CUDA(launch_k(preprocess_u8_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, stream, img, out, H0, W0, H, W, sh, sw, bgr));
    return 0;
}

int launch_frames_max(const float* frames, float* mx, long long n, cudaStream_t stream) {
    I2V_CHECK_CUDA(cudaMemsetAsync(mx, 0, sizeof(float), stream));
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    ProfScope ps(PROF_OTHER, 0, 4.0 * (double)n, stream);
    I2V_CHECK_CUDA(launch_k(frames_max_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, frames, mx, n));
    return 0;
}

int launch_frames_to_u8(const float* frames, const float* mx, unsigned char* out, int N, int T, int H, int W, long long sn,
                        long long st, long long sh, cudaStream_t stream) {
    const long long n_px = (long long)N * T * H * W;
    long long blocks = (n_px + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope ps(PROF_OTHER, 0, 15.0 * (double)n_px, stream);
    I2V_CHECK_CUDA(launch_k(frames_to_u8_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, frames, mx, out, N, T, H, W, sn, st, sh));
    return 0;
}

int launch_linear(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act,
                  cudaStream_t stream) {
    I2V_REQUIRE(K % 4 == 0, "linear: K=%d must be a multiple of 4", K);
    ProfScope ps(PROF_OTHER, 2.0 * (double)B * K * N, 4.0 * ((double)K * N + (double)B * (K + N)), stream);
    if (K == LSK_K && N >= 8192 && tune().linear_k64) {
        I2V_CHECK_CUDA(launch_k(linear_k64_kernel, dim3(ceil_div(N, LSK_THREADS), ceil_div(B, LSK_ROWS)), dim3(LSK_THREADS), 0, stream, x, w,
                                bias, y, B, N, act));
        return 0;
    }
    const long long warps = (long long)N * ((B + 7) / 8);
    const int bfly = tune().linear_bfly;   // A/B switch (0: 8 x warp_sum)
    I2V_CHECK_CUDA(launch_k(linear_kernel, dim3(ceil_div(warps * 32, 256)), dim3(256), 0, stream, x, w, bias, y, B, K, N, act, bfly));
    return 0;
}

// planes the dedicated kernel takes: whole tiles that are a segment of one row or a stack of whole rows, patch within SP_PATCH
bool spade_conv3_tiles(int H, int W) {
    const int HW = H * W, tv = spade_tile_voxels(HW);
    if (tv <= 0 || HW % tv != 0 || !(W >= tv ? W % tv == 0 : tv % W == 0)) return false;
    const int cols = W >= tv ? tv : W, rows = tv / cols;
    return (rows + 2) * (cols + 2) * 3 <= SP_PATCH;
}

int launch_spade_conv3(const float* img, const float* w, const float* bias, __half* y_hi, __half* y_lo, float split_scale, int B,
                       int H, int W, int act, cudaStream_t stream) {
    const int HW = H * W, tile_v = spade_tile_voxels(HW);
    I2V_REQUIRE(spade_conv3_tiles(H, W) && B < 65536, "spade_conv3: plane %dx%d does not tile into %d-voxel patches", H, W, tile_v);
    const double M = (double)B * HW;
    ProfScope ps(PROF_CONV_SIMT, 2.0 * M * 128 * 27, 4.0 * (M * 3 + M * 128 + 27.0 * 128), stream);
    I2V_CHECK_CUDA(launch_k(spade_conv3_kernel, dim3(HW / tile_v, B), dim3(256), 0, stream, img, w, bias, y_hi, y_lo, split_scale, H, W, act,
                            tile_v));
    return 0;
}

int launch_resize_bilinear_nchw_to_nhwc(const float* img, float* out, int B, int C, int H0, int W0, int H, int W,
                                        cudaStream_t stream) {
    const float sh = H > 1 ? (float)(H0 - 1) / (float)(H - 1) : 0.f;
    const float sw = W > 1 ? (float)(W0 - 1) / (float)(W - 1) : 0.f;
    const long long n = (long long)B * H * W * C;
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    I2V_CHECK_CUDA(launch_k(resize_bilinear_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, img, out, B, C, H0, W0, H, W, sh, sw));
    return 0;
}

int launch_maxpool3x3s2(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream) {
    I2V_REQUIRE(C % 4 == 0, "maxpool: C=%d must be a multiple of 4", C);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long n = (long long)B * Ho * Wo * (C / 4);
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    I2V_CHECK_CUDA(launch_k(maxpool3x3s2_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, y, B, H, W, C, Ho, Wo));
    return 0;
}

}  // namespace i2v
