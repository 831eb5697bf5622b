// Statistics + fused modulation passes of the decoder / encoders (HBM-bound, channels-last fp32).
//
//   channel_stats : one read of x -> per-(sample, channel) sum and sum of squares (double)
//   norm_coeffs   : sums -> per-(sample, channel) affine (A, Bc) of GroupNorm(16) / InstanceNorm /
//                   AdaIN / GroupNorm-affine  (normalization_layer.py:11,31,42 ; eps 1e-5, biased var)
//   modulate      : y = act( (A x + Bc) * (1 + gamma) + beta  [+ A2 r + B2] ), reading x through a
//                   nearest-upsample index map so the upsampled tensor of decoder.py:102-114 is never
//                   materialised, SPADE's gamma/beta maps are read as 2-D maps (not repeated over T as
//                   normalization_layer.py:22-23 does).
#include <cuda_fp16.h>

#include <cmath>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

namespace {

constexpr int STATS_THREADS = 256;
constexpr int STATS_ELEMS_PER_THREAD = 256;   // most voxels summed by one thread per channel group (fewer on small tensors)

__global__ void __launch_bounds__(STATS_THREADS) channel_stats_kernel(const float* __restrict__ x,
                                                                      double* __restrict__ sums, long long V,
                                                                      int C, int lanes_c, int rows, int ept) {
    extern __shared__ double sh[];   // [C][2]
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int C4 = C >> 2;
    for (int i = threadIdx.x; i < 2 * C; i += STATS_THREADS) sh[i] = 0.0;
    __syncthreads();
    const int cl = threadIdx.x % lanes_c, rl = threadIdx.x / lanes_c;
    const long long chunk = (long long)rows * ept;
    const long long v0 = (long long)blockIdx.x * chunk;
    const long long v1 = min(V, v0 + chunk);
    const float4* xb = reinterpret_cast<const float4*>(x + (long long)b * V * C);
    if (rl < rows) {
        for (int cg = cl; cg < C4; cg += lanes_c) {
            double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
            long long v = v0 + rl;
            while (v < v1) {
                float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
#pragma unroll 4
                for (int it = 0; it < 16 && v < v1; ++it, v += rows) {
                    const float4 t = __ldg(xb + v * C4 + cg);
                    fs[0] += t.x; fs[1] += t.y; fs[2] += t.z; fs[3] += t.w;
                    fq[0] = fmaf(t.x, t.x, fq[0]); fq[1] = fmaf(t.y, t.y, fq[1]);
                    fq[2] = fmaf(t.z, t.z, fq[2]); fq[3] = fmaf(t.w, t.w, fq[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { s[j] += fs[j]; q[j] += fq[j]; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&sh[2 * (cg * 4 + j)], s[j]);
                atomicAdd(&sh[2 * (cg * 4 + j) + 1], q[j]);
            }
        }
    }
    __syncthreads();
    double* out = sums + (long long)b * C * 2;
    for (int i = threadIdx.x; i < 2 * C; i += STATS_THREADS) atomicAdd(out + i, sh[i]);
}

__global__ void norm_coeffs_kernel(const double* __restrict__ sums, float* __restrict__ coef, int B, int C,
                                   double count_per_channel, int groups, float eps,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mod) {
    pdl_launch_dependents();
    pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * C) return;
    const int b = idx / C, c = idx - b * C;
    double s = 0, q = 0, n;
    if (groups > 0) {
        const int cpg = C / groups, g = c / cpg;
        const double* p = sums + ((long long)b * C + g * cpg) * 2;
        for (int i = 0; i < cpg; ++i) { s += p[2 * i]; q += p[2 * i + 1]; }
        n = count_per_channel * cpg;
    } else {
        s = sums[(long long)idx * 2]; q = sums[(long long)idx * 2 + 1];
        n = count_per_channel;
    }
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float A = rstd, Bc = (float)(-mean) * rstd;
    if (gamma != nullptr) { A *= gamma[c]; Bc = Bc * gamma[c] + beta[c]; }
    if (mod != nullptr) {
        const float g = mod[(long long)b * 2 * C + c], bt = mod[(long long)b * 2 * C + C + c];
        A *= g; Bc = Bc * g + bt;
    }
    coef[(long long)idx * 2] = A;
    coef[(long long)idx * 2 + 1] = Bc;
}

__global__ void __launch_bounds__(256) modulate_kernel(const ModArgs a, long long total4) {
    pdl_launch_dependents();
    pdl_wait();
    const int C4 = a.C >> 2;
    const int Ts = a.T / a.ut, Hs = a.H / a.uh, Ws = a.W / a.uw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long vox = i / C4;
        const int w = (int)(vox % a.W); vox /= a.W;
        const int h = (int)(vox % a.H); vox /= a.H;
        const int t = (int)(vox % a.T);
        const int b = (int)(vox / a.T);
        const long long src = ((((long long)b * Ts + t / a.ut) * Hs + h / a.uh) * Ws + w / a.uw) * C4 + c4;
        float4 xv = __ldg(reinterpret_cast<const float4*>(a.x) + src);
        float v[4] = {xv.x, xv.y, xv.z, xv.w};
        if (a.coef != nullptr) {
            const float4* cf = reinterpret_cast<const float4*>(a.coef + ((long long)b * a.C + c4 * 4) * 2);
            const float4 c0 = __ldg(cf), c1 = __ldg(cf + 1);
            v[0] = fmaf(c0.x, v[0], c0.y); v[1] = fmaf(c0.z, v[1], c0.w);
            v[2] = fmaf(c1.x, v[2], c1.y); v[3] = fmaf(c1.z, v[3], c1.w);
        }
        if (a.gb != nullptr) {
            const float4* gp = reinterpret_cast<const float4*>(a.gb + (((long long)b * a.H + h) * a.W + w) * 2 * a.C);
            const float4 g = __ldg(gp + c4), bt = __ldg(gp + C4 + c4);
            v[0] = fmaf(v[0], 1.f + g.x, bt.x); v[1] = fmaf(v[1], 1.f + g.y, bt.y);
            v[2] = fmaf(v[2], 1.f + g.z, bt.z); v[3] = fmaf(v[3], 1.f + g.w, bt.w);
        }
        if (a.r != nullptr) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(a.r) + i);
            float r[4] = {rv.x, rv.y, rv.z, rv.w};
            if (a.coef2 != nullptr) {
                const float4* cf = reinterpret_cast<const float4*>(a.coef2 + ((long long)b * a.C + c4 * 4) * 2);
                const float4 c0 = __ldg(cf), c1 = __ldg(cf + 1);
                r[0] = fmaf(c0.x, r[0], c0.y); r[1] = fmaf(c0.z, r[1], c0.w);
                r[2] = fmaf(c1.x, r[2], c1.y); r[3] = fmaf(c1.z, r[3], c1.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += r[j];
        }
        float4 o;
        o.x = apply_act(v[0], a.act); o.y = apply_act(v[1], a.act);
        o.z = apply_act(v[2], a.act); o.w = apply_act(v[3], a.act);
        if (a.out_hi != nullptr) {
            const float s = a.split_scale;
            const float f[4] = {o.x * s, o.y * s, o.z * s, o.w * s};
            __half hh[4], ll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_f16(f[j], hh[j], ll[j]);
            reinterpret_cast<uint2*>(a.out_hi)[i] = *reinterpret_cast<const uint2*>(hh);
            reinterpret_cast<uint2*>(a.out_lo)[i] = *reinterpret_cast<const uint2*>(ll);
            if (a.out_f32 != nullptr) reinterpret_cast<float4*>(a.out_f32)[i] = o;
        } else {
            reinterpret_cast<float4*>(a.out)[i] = o;
        }
    }
}

// Fast path of the tensor-core engine's prologue pass: 8 channels per thread (16-byte stores of the fp16 hi and lo
// words), shift/mask index decode (W and C/8 are powers of two in the decoder).  No second branch, output always
// the fp16 split.  A thread owns one (h, w, 8-channel) position of one sample and walks ALL T planes: the
// normalisation coefficients and SPADE's gamma/beta (which do not depend on t, normalization_layer.py:22-23) are
// loaded once instead of once per plane, and a source plane feeds its `ut` output planes from registers
// (ncu before: the gamma|beta maps were re-read T times and the pass ran at 3.0 TB/s of DRAM traffic against
// 5.3 TB/s for the map-free AdaIN pass).
__global__ void __launch_bounds__(256) modulate8_split_kernel(const ModArgs a, int c8_shift, int w_shift) {
    pdl_launch_dependents();
    pdl_wait();
    const int C8 = a.C >> 3;
    const int b = blockIdx.y;
    const int Ts = a.T / a.ut, Hs = a.H / a.uh, Ws = a.W / a.uw;
    const int per_plane = a.H * a.W * C8;
    const long long src_plane = (long long)Hs * Ws * (C8 * 2);      // float4 per source plane
    const float4* xb = reinterpret_cast<const float4*>(a.x) + (long long)b * Ts * src_plane;
    const float4* cf = a.coef ? reinterpret_cast<const float4*>(a.coef + (long long)b * a.C * 2) : nullptr;
    const float4* gbp = a.gb ? reinterpret_cast<const float4*>(a.gb + (long long)b * a.H * a.W * 2 * a.C) : nullptr;
    uint4* oh = reinterpret_cast<uint4*>(a.out_hi) + (long long)b * a.T * per_plane;
    uint4* ol = reinterpret_cast<uint4*>(a.out_lo) + (long long)b * a.T * per_plane;
    // second result (host guarantees ut = uh = uw = 1 when it is requested)
    const float4* cf2 = a.outb_hi ? reinterpret_cast<const float4*>(a.coef_b + (long long)b * a.C * 2) : nullptr;
    uint4* o2h = reinterpret_cast<uint4*>(a.outb_hi) + (long long)b * a.T * per_plane;
    uint4* o2l = reinterpret_cast<uint4*>(a.outb_lo) + (long long)b * a.T * per_plane;
    const float s = a.split_scale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_plane; i += gridDim.x * blockDim.x) {
        const int c8 = i & (C8 - 1);
        const int hw = i >> c8_shift;
        const int w = hw & (a.W - 1), h = hw >> w_shift;
        const int src = ((h / a.uh) * Ws + (w / a.uw)) * (C8 * 2) + c8 * 2;
        float c2a[8], c2b[8];
        if (cf2 != nullptr) {
            const float4* c = cf2 + c8 * 4;
            const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
            c2a[0] = c0.x; c2b[0] = c0.y; c2a[1] = c0.z; c2b[1] = c0.w; c2a[2] = c1.x; c2b[2] = c1.y; c2a[3] = c1.z; c2b[3] = c1.w;
            c2a[4] = c2.x; c2b[4] = c2.y; c2a[5] = c2.z; c2b[5] = c2.w; c2a[6] = c3.x; c2b[6] = c3.y; c2a[7] = c3.z; c2b[7] = c3.w;
        }
        float ca[8], cb[8], ga[8], gbv[8];      // v = (ca * x + cb) * ga + gbv
#pragma unroll
        for (int j = 0; j < 8; ++j) { ca[j] = 1.f; cb[j] = 0.f; ga[j] = 1.f; gbv[j] = 0.f; }
        if (cf != nullptr) {
            const float4* c = cf + c8 * 4;
            const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
            ca[0] = c0.x; cb[0] = c0.y; ca[1] = c0.z; cb[1] = c0.w; ca[2] = c1.x; cb[2] = c1.y; ca[3] = c1.z; cb[3] = c1.w;
            ca[4] = c2.x; cb[4] = c2.y; ca[5] = c2.z; cb[5] = c2.w; ca[6] = c3.x; cb[6] = c3.y; ca[7] = c3.z; cb[7] = c3.w;
        }
        if (gbp != nullptr) {
            const float4* g = gbp + (long long)hw * (C8 * 4) + c8 * 2;      // row of 2C floats = C8*4 float4: gamma | beta
            const float4 g0 = __ldg(g), g1 = __ldg(g + 1), b0 = __ldg(g + C8 * 2), b1 = __ldg(g + C8 * 2 + 1);
            ga[0] = 1.f + g0.x; ga[1] = 1.f + g0.y; ga[2] = 1.f + g0.z; ga[3] = 1.f + g0.w;
            ga[4] = 1.f + g1.x; ga[5] = 1.f + g1.y; ga[6] = 1.f + g1.z; ga[7] = 1.f + g1.w;
            gbv[0] = b0.x; gbv[1] = b0.y; gbv[2] = b0.z; gbv[3] = b0.w; gbv[4] = b1.x; gbv[5] = b1.y; gbv[6] = b1.z; gbv[7] = b1.w;
        }
#pragma unroll 4
        for (int ts = 0; ts < Ts; ++ts) {
            const float4 x0 = __ldg(xb + ts * src_plane + src), x1 = __ldg(xb + ts * src_plane + src + 1);
            float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            __half2 hh[4], ll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // same operation order as the generic kernel: fma(A, x, B), then fma(v, 1 + gamma, beta)
                float f0 = v[2 * j], f1 = v[2 * j + 1];
                if (cf != nullptr) { f0 = fmaf(ca[2 * j], f0, cb[2 * j]); f1 = fmaf(ca[2 * j + 1], f1, cb[2 * j + 1]); }
                if (gbp != nullptr) { f0 = fmaf(f0, ga[2 * j], gbv[2 * j]); f1 = fmaf(f1, ga[2 * j + 1], gbv[2 * j + 1]); }
                f0 = apply_act(f0, a.act) * s; f1 = apply_act(f1, a.act) * s;
                __half h0, h1, l0, l1;
                split_f16(f0, h0, l0); split_f16(f1, h1, l1);
                hh[j] = __halves2half2(h0, h1);
                ll[j] = __halves2half2(l0, l1);
            }
            for (int r = 0; r < a.ut; ++r) {
                const long long o = (long long)(ts * a.ut + r) * per_plane + i;
                oh[o] = *reinterpret_cast<const uint4*>(hh);
                ol[o] = *reinterpret_cast<const uint4*>(ll);
            }
            if (cf2 != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float f0 = fmaf(c2a[2 * j], v[2 * j], c2b[2 * j]) * s, f1 = fmaf(c2a[2 * j + 1], v[2 * j + 1], c2b[2 * j + 1]) * s;
                    __half h0, h1, l0, l1;
                    split_f16(f0, h0, l0); split_f16(f1, h1, l1);
                    hh[j] = __halves2half2(h0, h1);
                    ll[j] = __halves2half2(l0, l1);
                }
                const long long o = (long long)ts * per_plane + i;
                o2h[o] = *reinterpret_cast<const uint4*>(hh);
                o2l[o] = *reinterpret_cast<const uint4*>(ll);
            }
        }
    }
}

// NaN-propagating clamp to the fp16 range (same values as sat_f16: NaN stays NaN) in two FMNMX instead of compare + select + FMNMX
__device__ __forceinline__ float sat_f16_fast(float f) {
    float r;
    asm("{\n\t.reg .f32 t;\n\tmax.NaN.f32 t, %1, 0fC77FE000;\n\tmin.NaN.f32 %0, t, 0f477FE000;\n\t}" : "=f"(r) : "f"(f));
    return r;
}
// split of a PAIR of (already scaled) values: one packed conversion for the two high words
__device__ __forceinline__ void split_f16x2(float f0, float f1, __half2& hi, __half2& lo) {
    f0 = sat_f16_fast(f0); f1 = sat_f16_fast(f1);
    hi = __floats2half2_rn(f0, f1);
    lo = __floats2half2_rn(f0 - __low2float(hi), f1 - __high2float(hi));
}

// SPADE passes of the decoder (coefficients + gamma|beta maps + LeakyReLU(0.2), the only form decoder_run issues for the
// T-walking kernel): the same arithmetic as modulate8_split_kernel, restructured around the memory system.  ncu on the generic
// kernel (profiles/r02_ncu_modulate_summary.txt): 128 registers, 16 warps per SM, and -- because the activation switch and the
// runtime `ut` loop keep branches between a plane's loads and its stores -- only ONE plane (32 bytes) in flight per thread:
// long-scoreboard-bound at 3.3 TB/s.  Here the activation and the temporal factor are compile-time, a thread owns one
// (h, w, 8-channel) position (fine-grained grid: no 8.2-wave tail), and a rotating window keeps TCH source planes (128 bytes)
// in flight per thread: plane ts + TCH is requested the moment plane ts has been written out.  Instruction diet (the generic
// kernel spends ~210 instructions per plane and thread, enough to be issue-bound at 5 TB/s): power-of-two split scale folded
// into gamma / beta, lrelu as max(v, 0.2 v), NaN-propagating two-instruction clamp, packed conversions, pointer increments.
template <bool SECOND, int UT, int TCH>
__global__ void __launch_bounds__(256, 2) modulate8_spade_kernel(const ModArgs a, int c8_shift, int w_shift) {
    pdl_launch_dependents();
    pdl_wait();
    // TCH source planes in flight (4; 2 for the two-plane tensors of g_0 / g_1); host guarantees Ts % TCH == 0 and a power-of-two scale
    const int C8 = a.C >> 3;
    const int b = blockIdx.y;
    const int Ts = a.T / UT, Hs = a.H / a.uh, Ws = a.W / a.uw;
    const int per_plane = a.H * a.W * C8;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= per_plane) return;
    const int c8 = i & (C8 - 1);
    const int hw = i >> c8_shift;
    const int w = hw & (a.W - 1), h = hw >> w_shift;
    const long long src_plane = (long long)Hs * Ws * (C8 * 2);      // float4 per source plane
    const float4* xp = reinterpret_cast<const float4*>(a.x) + (long long)b * Ts * src_plane + ((h / a.uh) * Ws + (w / a.uw)) * (C8 * 2) + c8 * 2;
    float4 xv[TCH][2];
#pragma unroll
    for (int u = 0; u < TCH; ++u) { xv[u][0] = __ldg(xp); xv[u][1] = __ldg(xp + 1); xp += src_plane; }
    const float s = a.split_scale;
    // v * s = fma(fma(ca, x, cb), ga * s, gbv * s): exact for a power-of-two s, so the scale costs nothing per plane
    float ca[8], cb[8], ga[8], gbv[8];
    {
        const float4* c = reinterpret_cast<const float4*>(a.coef + (long long)b * a.C * 2) + c8 * 4;
        const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
        ca[0] = c0.x; cb[0] = c0.y; ca[1] = c0.z; cb[1] = c0.w; ca[2] = c1.x; cb[2] = c1.y; ca[3] = c1.z; cb[3] = c1.w;
        ca[4] = c2.x; cb[4] = c2.y; ca[5] = c2.z; cb[5] = c2.w; ca[6] = c3.x; cb[6] = c3.y; ca[7] = c3.z; cb[7] = c3.w;
        const float4* g = reinterpret_cast<const float4*>(a.gb + (long long)b * a.H * a.W * 2 * a.C) + (long long)hw * (C8 * 4) + c8 * 2;
        const float4 g0 = __ldg(g), g1 = __ldg(g + 1), b0 = __ldg(g + C8 * 2), b1 = __ldg(g + C8 * 2 + 1);
        ga[0] = 1.f + g0.x; ga[1] = 1.f + g0.y; ga[2] = 1.f + g0.z; ga[3] = 1.f + g0.w;
        ga[4] = 1.f + g1.x; ga[5] = 1.f + g1.y; ga[6] = 1.f + g1.z; ga[7] = 1.f + g1.w;
        gbv[0] = b0.x; gbv[1] = b0.y; gbv[2] = b0.z; gbv[3] = b0.w; gbv[4] = b1.x; gbv[5] = b1.y; gbv[6] = b1.z; gbv[7] = b1.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) { ga[j] *= s; gbv[j] *= s; }
    }
    float c2a[8], c2b[8];
    if (SECOND) {
        const float4* c = reinterpret_cast<const float4*>(a.coef_b + (long long)b * a.C * 2) + c8 * 4;
        const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
        c2a[0] = c0.x; c2b[0] = c0.y; c2a[1] = c0.z; c2b[1] = c0.w; c2a[2] = c1.x; c2b[2] = c1.y; c2a[3] = c1.z; c2b[3] = c1.w;
        c2a[4] = c2.x; c2b[4] = c2.y; c2a[5] = c2.z; c2b[5] = c2.w; c2a[6] = c3.x; c2b[6] = c3.y; c2a[7] = c3.z; c2b[7] = c3.w;
    }
    uint4* oh = reinterpret_cast<uint4*>(a.out_hi) + (long long)b * a.T * per_plane + i;
    uint4* ol = reinterpret_cast<uint4*>(a.out_lo) + (long long)b * a.T * per_plane + i;
    uint4* o2h = SECOND ? reinterpret_cast<uint4*>(a.outb_hi) + (long long)b * a.T * per_plane + i : nullptr;
    uint4* o2l = SECOND ? reinterpret_cast<uint4*>(a.outb_lo) + (long long)b * a.T * per_plane + i : nullptr;
    for (int ts0 = 0; ts0 < Ts; ts0 += TCH) {
        const bool more = ts0 + TCH < Ts;
#pragma unroll
        for (int u = 0; u < TCH; ++u) {
            const float v[8] = {xv[u][0].x, xv[u][0].y, xv[u][0].z, xv[u][0].w, xv[u][1].x, xv[u][1].y, xv[u][1].z, xv[u][1].w};
            __half2 hh[4], ll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // the generic kernels' operation order -- fma(A, x, B), fma(v, 1 + gamma, beta), lrelu, scale, split -- with the
                // (power-of-two) scale folded into gamma / beta: bit-identical results
                float f0 = fmaf(ca[2 * j], v[2 * j], cb[2 * j]), f1 = fmaf(ca[2 * j + 1], v[2 * j + 1], cb[2 * j + 1]);
                f0 = fmaf(f0, ga[2 * j], gbv[2 * j]); f1 = fmaf(f1, ga[2 * j + 1], gbv[2 * j + 1]);
                f0 = fmaxf(f0, 0.2f * f0); f1 = fmaxf(f1, 0.2f * f1);
                split_f16x2(f0, f1, hh[j], ll[j]);
            }
#pragma unroll
            for (int r = 0; r < UT; ++r) {
                *oh = *reinterpret_cast<const uint4*>(hh); oh += per_plane;
                *ol = *reinterpret_cast<const uint4*>(ll); ol += per_plane;
            }
            if (SECOND) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    split_f16x2(fmaf(c2a[2 * j], v[2 * j], c2b[2 * j]) * s, fmaf(c2a[2 * j + 1], v[2 * j + 1], c2b[2 * j + 1]) * s, hh[j], ll[j]);
                *o2h = *reinterpret_cast<const uint4*>(hh); o2h += per_plane;
                *o2l = *reinterpret_cast<const uint4*>(ll); o2l += per_plane;
            }
            if (more) { xv[u][0] = __ldg(xp); xv[u][1] = __ldg(xp + 1); xp += src_plane; }   // plane ts + TCH into the freed registers
        }
    }
}

// Map-free variant (AdaIN / GroupNorm-affine / plain activation passes): one (b, t) plane per blockIdx.y -- a pure
// stream, which DRAM serves better than 16 interleaved plane streams (measured: 5.3 vs 4.5 TB/s).  A thread's channel group
// is fixed (the grid stride is a multiple of C/8), so its 8 coefficient pairs are loaded ONCE instead of once per element
// (they cost 4 of the 6 load instructions of an iteration: 3.7 TB/s at C = 128 against 6.0 TB/s without coefficients), two
// elements are in flight per thread.  (Streaming stores, st.global.cs, for the fp16 pair were measured slower: 1.25 -> 1.32 /
// 1.51 -> 1.75 ms on the two big SPADE passes, profiles/r02_bench_ab.txt.)
// ACT >= 0: compile-time activation (none / ReLU / LeakyReLU) and a power-of-two split scale, folded into the coefficients
// (exact), with the SPADE kernel's instruction diet; ACT = -1: the generic form (run-time activation, any scale).  Same bits.
template <bool COEF, int ACT>
__global__ void __launch_bounds__(256, 4) modulate8_split_plane_kernel(const ModArgs a, int c8_shift, int w_shift) {
    pdl_launch_dependents();
    pdl_wait();
    const int C8 = a.C >> 3;
    const int plane = blockIdx.y;                 // b * T + t
    const int b = plane / a.T, t = plane - b * a.T;
    const int Ts = a.T / a.ut, Hs = a.H / a.uh, Ws = a.W / a.uw;
    const int per_plane = a.H * a.W * C8;
    const float4* xp = reinterpret_cast<const float4*>(a.x) + ((long long)(b * Ts + t / a.ut) * Hs * Ws) * (C8 * 2);
    uint4* oh = reinterpret_cast<uint4*>(a.out_hi) + (long long)plane * per_plane;
    uint4* ol = reinterpret_cast<uint4*>(a.out_lo) + (long long)plane * per_plane;
    const float s = a.split_scale;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;    // stride % C8 == 0 (host-checked)
    const int c8 = i0 & (C8 - 1);
    float ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ca[j] = 1.f; cb[j] = 0.f; }
    constexpr bool has_coef = COEF;
    if (has_coef) {
        const float4* c = reinterpret_cast<const float4*>(a.coef + (long long)b * a.C * 2) + c8 * 4;
        const float4 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
        ca[0] = c0.x; cb[0] = c0.y; ca[1] = c0.z; cb[1] = c0.w; ca[2] = c1.x; cb[2] = c1.y; ca[3] = c1.z; cb[3] = c1.w;
        ca[4] = c2.x; cb[4] = c2.y; ca[5] = c2.z; cb[5] = c2.w; ca[6] = c3.x; cb[6] = c3.y; ca[7] = c3.z; cb[7] = c3.w;
        if (ACT >= 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ca[j] *= s; cb[j] *= s; }
        }
    }
    auto src_of = [&](int i) {
        const int hw = i >> c8_shift;
        const int w = hw & (a.W - 1), h = hw >> w_shift;
        return ((h / a.uh) * Ws + (w / a.uw)) * (C8 * 2) + c8 * 2;
    };
    auto finish = [&](int i, const float4& x0, const float4& x1) {
        float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        __half2 hh[4], ll[4];
        if (ACT >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f0 = has_coef ? fmaf(ca[2 * j], v[2 * j], cb[2 * j]) : v[2 * j] * s;
                float f1 = has_coef ? fmaf(ca[2 * j + 1], v[2 * j + 1], cb[2 * j + 1]) : v[2 * j + 1] * s;
                if (ACT == ACT_RELU) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                if (ACT == ACT_LRELU02) { f0 = fmaxf(f0, 0.2f * f0); f1 = fmaxf(f1, 0.2f * f1); }
                split_f16x2(f0, f1, hh[j], ll[j]);
            }
        } else {
            if (has_coef) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(ca[j], v[j], cb[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float f0 = apply_act(v[2 * j], a.act) * s, f1 = apply_act(v[2 * j + 1], a.act) * s;
                __half h0, h1, l0, l1;
                split_f16(f0, h0, l0); split_f16(f1, h1, l1);
                hh[j] = __halves2half2(h0, h1);
                ll[j] = __halves2half2(l0, l1);
            }
        }
        oh[i] = *reinterpret_cast<const uint4*>(hh);
        ol[i] = *reinterpret_cast<const uint4*>(ll);
    };
    int i = i0;
    if (!COEF) {          // plain activation pass: the one-element loop already streams at 6.0 TB/s
        for (; i < per_plane; i += stride) {
            const int sa = src_of(i);
            const float4 xa0 = __ldg(xp + sa), xa1 = __ldg(xp + sa + 1);
            finish(i, xa0, xa1);
        }
        return;
    }
    for (; i + stride < per_plane; i += 2 * stride) {
        const int sa = src_of(i), sb = src_of(i + stride);
        const float4 xa0 = __ldg(xp + sa), xa1 = __ldg(xp + sa + 1), xb0 = __ldg(xp + sb), xb1 = __ldg(xp + sb + 1);
        finish(i, xa0, xa1);
        finish(i + stride, xb0, xb1);
    }
    if (i < per_plane) {
        const int sa = src_of(i);
        const float4 xa0 = __ldg(xp + sa), xa1 = __ldg(xp + sa + 1);
        finish(i, xa0, xa1);
    }
}

__global__ void mean_from_sums_kernel(const double* __restrict__ sums, float* __restrict__ y, int n, double inv) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (float)(sums[2 * (long long)i] * inv);
}

}  // namespace

int launch_channel_stats(const float* x, double* sums, int B, long long V, int C, cudaStream_t stream) {
    I2V_REQUIRE(C % 4 == 0 && C <= 2048, "channel_stats: C=%d must be a multiple of 4 and <= 2048", C);
    I2V_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B * C, stream));
    const int C4 = C / 4;
    const int lanes_c = C4 < STATS_THREADS ? C4 : STATS_THREADS;
    const int rows = STATS_THREADS / lanes_c;
    // voxels per thread: 256 on large tensors; small ones (8x8 / 4x4 planes of the decoder, the embedder's deep layers) are cut
    // finer so that the grid still covers the machine about twice -- a (1, B) grid of 128-voxel serial loops ran at 0.4 TB/s
    int ept = STATS_ELEMS_PER_THREAD;
    while (ept > 8 && (long long)B * ((V + (long long)rows * ept - 1) / ((long long)rows * ept)) < 2 * kNumSMs) ept >>= 1;
    const long long chunk = (long long)rows * ept;
    ProfScope ps(PROF_STATS, 3.0 * (double)B * V * C, 4.0 * (double)B * V * C, stream);
    dim3 grid(ceil_div(V, chunk), B);
    I2V_CHECK_CUDA(launch_k(channel_stats_kernel, grid, dim3(STATS_THREADS), sizeof(double) * 2 * C, stream, x, sums, V, C, lanes_c, rows, ept));
    return 0;
}

int launch_norm_coeffs(const double* sums, float* coef, int B, int C, long long V, int groups, float eps,
                       const float* gamma, const float* beta, const float* mod, cudaStream_t stream) {
    I2V_REQUIRE(groups == 0 || C % groups == 0, "norm_coeffs: C=%d not divisible by groups=%d", C, groups);
    I2V_REQUIRE((gamma == nullptr) == (beta == nullptr), "norm_coeffs: gamma and beta go together");
    const int n = B * C;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    I2V_CHECK_CUDA(launch_k(norm_coeffs_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, stream, sums, coef, B, C, (double)V, groups, eps, gamma, beta, mod));
    return 0;
}

int launch_modulate(const ModArgs& a, cudaStream_t stream) {
    I2V_REQUIRE(a.C % 4 == 0, "modulate: C=%d must be a multiple of 4", a.C);
    I2V_REQUIRE(a.T % a.ut == 0 && a.H % a.uh == 0 && a.W % a.uw == 0, "modulate: upsample factors must divide dims");
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    I2V_REQUIRE(a.outb_hi == nullptr || (a.coef_b && a.outb_lo && a.out_hi && a.r == nullptr && a.ut == 1 && a.uh == 1 && a.uw == 1 &&
                                         a.C % 8 == 0 && pow2(a.C / 8) && pow2(a.W)),
                "modulate: the second result needs the 8-channel split path without upsampling");
    if (a.out_hi != nullptr && a.r == nullptr && a.out_f32 == nullptr && a.C % 8 == 0 && pow2(a.C / 8) && a.C / 8 <= 256 && pow2(a.W) &&
        (long long)a.T * a.H * a.W * (a.C / 8) < (1ll << 31) && (long long)a.B * a.T < 65536) {
        const int per_plane = a.H * a.W * (a.C / 8);
        int bx = (per_plane + 255) / 256;
        const int cap = (kNumSMs * 16 + a.B - 1) / a.B;            // ~16 CTAs of work per SM overall
        if (bx > cap) bx = cap < 1 ? 1 : cap;
        const double tot = (double)a.B * a.T * per_plane * 8;
        ProfScope ps(PROF_MODULATE, 4.0 * tot, 4.0 * (tot + tot / ((double)a.ut * a.uh * a.uw)) + (a.gb ? 8.0 * tot / a.T : 0.0), stream);
        int sexp = 0;
        const bool scale_pow2 = a.split_scale > 0.f && std::frexp(a.split_scale, &sexp) == 0.5f;
        const bool spade_form = a.gb != nullptr && a.coef != nullptr && a.act == ACT_LRELU02 && (a.ut == 1 || a.ut == 2) &&
                                (a.T / a.ut) % 2 == 0 && scale_pow2 && (a.outb_hi == nullptr || a.ut == 1) && tune().mod_spade != 0;
        if (spade_form) {
            const dim3 grid((per_plane + 255) / 256, a.B);
            const int cs = ilog2(a.C / 8), wsft = ilog2(a.W);
            const bool four = (a.T / a.ut) % 4 == 0;
#define I2V_SPADE(SECOND_, UT_, TCH_) I2V_CHECK_CUDA(launch_k(modulate8_spade_kernel<SECOND_, UT_, TCH_>, grid, dim3(256), 0, stream, a, cs, wsft))
            if (a.outb_hi != nullptr) { if (four) I2V_SPADE(true, 1, 4); else I2V_SPADE(true, 1, 2); }
            else if (a.ut == 1) { if (four) I2V_SPADE(false, 1, 4); else I2V_SPADE(false, 1, 2); }
            else { if (four) I2V_SPADE(false, 2, 4); else I2V_SPADE(false, 2, 2); }
#undef I2V_SPADE
        } else if (a.gb != nullptr || a.outb_hi != nullptr) {
            I2V_CHECK_CUDA(launch_k(modulate8_split_kernel, dim3(bx, a.B), dim3(256), 0, stream, a, ilog2(a.C / 8), ilog2(a.W)));
        } else {
            const int planes = a.B * a.T;
            I2V_REQUIRE(planes < 65536, "modulate: too many (b, t) planes for one launch (%d)", planes);
            int bp = (per_plane + 255) / 256;
            const int capp = (kNumSMs * 16 + planes - 1) / planes;
            if (bp > capp) bp = capp < 1 ? 1 : capp;
            const dim3 grid(bp, planes);
            const int cs = ilog2(a.C / 8), wsft = ilog2(a.W);
            const bool ct = scale_pow2 && tune().mod_spade != 0;      // compile-time activation + folded scale
#define I2V_PLANE(COEF_, ACT_) I2V_CHECK_CUDA(launch_k(modulate8_split_plane_kernel<COEF_, ACT_>, grid, dim3(256), 0, stream, a, cs, wsft))
            if (a.coef != nullptr) {
                if (ct && a.act == ACT_LRELU02) I2V_PLANE(true, ACT_LRELU02);
                else if (ct && a.act == ACT_RELU) I2V_PLANE(true, ACT_RELU);
                else if (ct && a.act == ACT_NONE) I2V_PLANE(true, ACT_NONE);
                else I2V_PLANE(true, -1);
            } else {
                if (ct && a.act == ACT_LRELU02) I2V_PLANE(false, ACT_LRELU02);
                else if (ct && a.act == ACT_RELU) I2V_PLANE(false, ACT_RELU);
                else if (ct && a.act == ACT_NONE) I2V_PLANE(false, ACT_NONE);
                else I2V_PLANE(false, -1);
            }
#undef I2V_PLANE
        }
        return 0;
    }
    const long long total4 = (long long)a.B * a.T * a.H * a.W * (a.C / 4);
    long long blocks = (total4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PROF_MODULATE, 4.0 * (double)total4 * 4,
                 4.0 * 4 * ((double)total4 * (1.0 + (a.r ? 1.0 : 0.0)) + (double)total4 / ((double)a.ut * a.uh * a.uw)), stream);
    I2V_CHECK_CUDA(launch_k(modulate_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, a, total4));
    return 0;
}

int launch_mean_from_sums(const double* sums, float* y, int B, int C, long long V, cudaStream_t stream) {
    const int n = B * C;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    I2V_CHECK_CUDA(launch_k(mean_from_sums_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, stream, sums, y, n, 1.0 / (double)V));
    return 0;
}

}  // namespace i2v
