// Generic fp32 implicit-GEMM convolution (channels-last), the exact-arithmetic engine of the path.
//
// Covers every convolution of the sampling path that is not a stride-1 tensor-core shape: the
// embedder's strided 2-D convs (AE.py:131-141 -> torchvision resnet50), the 3-D encoder's strided
// 3x3x3 / (3,7,7) convs (resnet3D.py:166,185-193), SPADE's 3->128 conv (normalization_layer.py:13)
// and, as the fp32 reference engine, the decoder's Conv3d stack (decoder.py:15-25,84).
//
//   GEMM view: M = B*To*Ho*Wo output voxels, N = Cout, K = taps*Cin.
//   x  : [B, Ti, Hi, Wi, Cin]  fp32 channels-last (2-D convs use T = 1)
//   w  : [taps, Cout, Cin]     fp32, taps ordered (kt, kh, kw)  (repacked at load time)
//   y  : [B, To, Ho, Wo, Cout] or, for the final image conv, [B, To, Cout, Ho, Wo]
//        (decoder.py:120 returns x.transpose(1,2) of an NCTHW tensor).
// Epilogue: + bias[n], + residual read through a nearest-upsample index map (the shortcut of a
// GeneratorBlock lives at the pre-upsample resolution, decoder.py:40,102-114), activation.
//
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, register-prefetch double buffering.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <bool VEC>
__global__ void __launch_bounds__(NT, 2) conv_simt_kernel(const ConvArgs a) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    pdl_launch_dependents();
    pdl_wait();

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)a.B * a.To * a.Ho * a.Wo;
    const int K = a.kt * a.kh * a.kw * a.Cin;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- loader roles: thread owns rows (lr, lr+64) of the A tile and column ln of the B tile
    const int lr = tid & 63, lk = tid >> 6;   // lk in 0..3 : which float4 (VEC) / k-quarter (scalar)
    int rb[2], rt[2], rh[2], rw[2];
    bool rvalid[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        long long m = m0 + lr + 64 * i;
        rvalid[i] = m < M;
        long long mm = rvalid[i] ? m : 0;
        int wo = (int)(mm % a.Wo); mm /= a.Wo;
        int ho = (int)(mm % a.Ho); mm /= a.Ho;
        int to = (int)(mm % a.To); mm /= a.To;
        rb[i] = (int)mm;
        rt[i] = to * a.st - a.pt; rh[i] = ho * a.sh - a.ph; rw[i] = wo * a.sw - a.pw;
    }
    const int ln = n0 + lr;           // B-tile column handled by this thread
    const bool nvalid = ln < a.Cout;

    float4 ra[2];   // VEC: one float4 per row ; scalar: 4 consecutive k per row
    float4 rbv;

    auto load_tiles = [&](int k0) {
        if (VEC) {
            // Cin % 16 == 0: the 16-wide k chunk lies inside one tap, contiguous in memory
            const int tap = k0 / a.Cin, c0 = k0 - tap * a.Cin + lk * 4;
            const int dw = tap % a.kw, dh = (tap / a.kw) % a.kh, dt = tap / (a.kw * a.kh);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int ti = rt[i] + dt, hi = rh[i] + dh, wi = rw[i] + dw;
                const bool ok = rvalid[i] && (unsigned)ti < (unsigned)a.Ti && (unsigned)hi < (unsigned)a.Hi &&
                                (unsigned)wi < (unsigned)a.Wi;
                if (ok) {
                    const long long off = ((((long long)rb[i] * a.Ti + ti) * a.Hi + hi) * a.Wi + wi) * a.Cin + c0;
                    ra[i] = __ldg(reinterpret_cast<const float4*>(a.x + off));
                } else {
                    ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (nvalid) {
                const long long off = ((long long)tap * a.Cout + ln) * a.Cin + c0;
                rbv = __ldg(reinterpret_cast<const float4*>(a.w + off));
            } else {
                rbv = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            float va[2][4], vb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lk * 4 + j;
                const bool kok = k < K;
                const int kk = kok ? k : 0;
                const int tap = kk / a.Cin, c = kk - tap * a.Cin;
                const int dw = tap % a.kw, dh = (tap / a.kw) % a.kh, dt = tap / (a.kw * a.kh);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int ti = rt[i] + dt, hi = rh[i] + dh, wi = rw[i] + dw;
                    const bool ok = kok && rvalid[i] && (unsigned)ti < (unsigned)a.Ti &&
                                    (unsigned)hi < (unsigned)a.Hi && (unsigned)wi < (unsigned)a.Wi;
                    va[i][j] = ok ? __ldg(a.x + ((((long long)rb[i] * a.Ti + ti) * a.Hi + hi) * a.Wi + wi) * a.Cin + c)
                                  : 0.f;
                }
                vb[j] = (kok && nvalid) ? __ldg(a.w + ((long long)tap * a.Cout + ln) * a.Cin + c) : 0.f;
            }
            ra[0] = make_float4(va[0][0], va[0][1], va[0][2], va[0][3]);
            ra[1] = make_float4(va[1][0], va[1][1], va[1][2], va[1][3]);
            rbv = make_float4(vb[0], vb[1], vb[2], vb[3]);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][lk * 4 + 0][lr + 64 * i] = ra[i].x;
            As[buf][lk * 4 + 1][lr + 64 * i] = ra[i].y;
            As[buf][lk * 4 + 2][lr + 64 * i] = ra[i].z;
            As[buf][lk * 4 + 3][lr + 64 * i] = ra[i].w;
        }
        Bs[buf][lk * 4 + 0][lr] = rbv.x;
        Bs[buf][lk * 4 + 1][lr] = rbv.y;
        Bs[buf][lk * 4 + 2][lr] = rbv.z;
        Bs[buf][lk * 4 + 3][lr] = rbv.w;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // split-K: blockIdx.z owns a contiguous range of k-steps (deep-K / small-M layers of the encoders would
    // otherwise run on a handful of CTAs with a serial, latency-bound K loop)
    const int nk_all = (K + BK - 1) / BK;
    const int ksplit = gridDim.z;
    const int per = (nk_all + ksplit - 1) / ksplit;
    const int kb0 = blockIdx.z * per;
    const int kb1 = min(nk_all, kb0 + per);
    const int nk = kb1 > kb0 ? kb1 - kb0 : 0;
    if (nk > 0) {
        load_tiles(kb0 * BK);
        store_tiles(0);
    }
    __syncthreads();
    for (int kb = 0; kb < nk; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nk) load_tiles((kb0 + kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- split-K reduction: every split parks its partial tile; the LAST one to arrive sums all of them in
    // split order (a fixed order: the result does not depend on which CTA happens to be last) and runs the epilogue
    if (ksplit > 1) {
        const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
        float* part = a.splitk_ws + ((size_t)tile_id * ksplit + blockIdx.z) * (BM * BN);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(part + (ty * 8 + i) * BN + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __threadfence();
        __syncthreads();
        __shared__ int s_last;
        if (tid == 0) {
            const unsigned prev = atomicAdd(a.splitk_counters + tile_id, 1u);
            s_last = (prev == (unsigned)ksplit - 1);
            if (s_last) a.splitk_counters[tile_id] = 0;      // self-resetting: ready for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const float* base = a.splitk_ws + (size_t)tile_id * ksplit * (BM * BN);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int z = 0; z < ksplit; ++z) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (size_t)z * (BM * BN) + (ty * 8 + i) * BN + tx * 4));
                acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
            }
        }
    }

    // ---- epilogue
    const int nbase = n0 + tx * 4;
    float bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bj[j] = (a.bias != nullptr && nbase + j < a.Cout) ? __ldg(a.bias + nbase + j) : 0.f;
    const bool vec_store = (a.out_mode == 0) && (a.Cout % 4 == 0) && (nbase + 3 < a.Cout);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
        long long mm = m;
        const int wo = (int)(mm % a.Wo); mm /= a.Wo;
        const int ho = (int)(mm % a.Ho); mm /= a.Ho;
        const int to = (int)(mm % a.To); mm /= a.To;
        const int b = (int)mm;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bj[j];
        if (a.res != nullptr) {
            const int Tr = a.To / a.res_ut, Hr = a.Ho / a.res_uh, Wr = a.Wo / a.res_uw;
            const long long roff =
                ((((long long)b * Tr + to / a.res_ut) * Hr + ho / a.res_uh) * Wr + wo / a.res_uw) * a.Cout;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nbase + j < a.Cout) v[j] += __ldg(a.res + roff + nbase + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], a.act);
        if (a.y_hi != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nbase + j < a.Cout) {
                    __half hh, ll;
                    split_f16(v[j] * a.split_scale, hh, ll);
                    a.y_hi[m * a.Cout + nbase + j] = hh;
                    a.y_lo[m * a.Cout + nbase + j] = ll;
                }
        } else if (a.out_mode == 0) {
            float* dst = a.y + m * a.Cout + nbase;
            if (vec_store) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (nbase + j < a.Cout) dst[j] = v[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nbase + j < a.Cout)
                    a.y[((((long long)b * a.To + to) * a.Cout + nbase + j) * a.Ho + ho) * a.Wo + wo] = v[j];
        }
    }
}

}  // namespace

int launch_conv_simt(const ConvArgs& a, cudaStream_t stream) {
    const long long M = (long long)a.B * a.To * a.Ho * a.Wo;
    I2V_REQUIRE(M > 0 && a.Cout > 0 && a.Cin > 0, "conv_simt: empty problem");
    I2V_REQUIRE(a.res == nullptr || (a.To % a.res_ut == 0 && a.Ho % a.res_uh == 0 && a.Wo % a.res_uw == 0),
                "conv_simt: residual upsample factors must divide the output size");
    const double K_ = (double)a.kt * a.kh * a.kw * a.Cin;
    ProfScope ps(PROF_CONV_SIMT, 2.0 * (double)M * a.Cout * K_,
                 4.0 * ((double)a.B * a.Ti * a.Hi * a.Wi * a.Cin + (double)M * a.Cout + K_ * a.Cout), stream);
    dim3 grid(ceil_div(M, BM), ceil_div(a.Cout, BN), 1);
    // split-K when the tile grid cannot fill the machine and K is deep; bounded by the caller's scratch
    if (a.splitk_ws != nullptr && a.splitk_counters != nullptr) {
        const int tiles = grid.x * grid.y;
        const int nk = ((int)K_ + BK - 1) / BK;
        int ks = (2 * kNumSMs + tiles - 1) / tiles;
        if (ks > nk / 4) ks = nk / 4;                       // at least 4 k-steps per split
        const size_t per_tile = (size_t)BM * BN * sizeof(float);
        while (ks > 1 && (size_t)tiles * ks * per_tile > a.splitk_ws_bytes) --ks;
        if (tiles > a.splitk_max_tiles) ks = 1;
        if (ks > 1) grid.z = ks;
    }
    if (a.Cin % 16 == 0)
        I2V_CHECK_CUDA(launch_k(conv_simt_kernel<true>, grid, dim3(NT), 0, stream, a));
    else
        I2V_CHECK_CUDA(launch_k(conv_simt_kernel<false>, grid, dim3(NT), 0, stream, a));
    return 0;
}

}  // namespace i2v
