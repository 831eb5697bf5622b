// Small ops of the path: nn.Linear, bilinear start-frame resize, max-pool.
#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

namespace {

// One warp per output feature n, all batch rows; K % 4 == 0.  Used for decoder.fc (decoder.py:99),
// AdaIN's Linear (normalization_layer.py:44,49), the embedder's 1x1 "fc" conv on the pooled feature
// (AE.py:121-124) and conv_mu on the flattened 4x4 map (resnet3D.py:179,203).
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y, int B,
                                                     int K, int N, int act) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const int K4 = K >> 2;
    const float4* wr = reinterpret_cast<const float4*>(w + (long long)warp * K);
    const float bn = bias != nullptr ? __ldg(bias + warp) : 0.f;
    for (int b0 = 0; b0 < B; b0 += 8) {
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
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float s = warp_sum(acc[r]);
            if (lane == 0 && b0 + r < B) y[(long long)(b0 + r) * N + warp] = apply_act(s + bn, act);
        }
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

}  // namespace

int launch_linear(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act,
                  cudaStream_t stream) {
    I2V_REQUIRE(K % 4 == 0, "linear: K=%d must be a multiple of 4", K);
    ProfScope ps(PROF_OTHER, 2.0 * (double)B * K * N, 4.0 * ((double)K * N + (double)B * (K + N)), stream);
    I2V_CHECK_CUDA(launch_k(linear_kernel, dim3(ceil_div((long long)N * 32, 256)), dim3(256), 0, stream, x, w, bias, y, B, K, N, act));
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
