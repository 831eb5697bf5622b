// Tensor-core convolution engine for sm_100a: stride-1 "same" Conv3d/Conv2d as an implicit GEMM on
// tcgen05.mma with TMEM accumulators, operands staged by TMA, fp32-grade accuracy from an
// error-compensated fp16 split.
//
//   Covers (decoder.py:15-25,84; normalization_layer.py:14-15): conv_0 / conv_1 (3x3x3), conv_s (1x1x1),
//   conv_img (3x3x3, Cout=3), SPADE's gamma|beta conv (3x3, T=1).
//
// GEMM view.  M = output voxels, N = Cout, K = taps x Cin.  A CTA owns a 128 x n_tile output tile whose
// 128 rows are a BOX of the channels-last activation tensor [B,T,H,W,C]: (bw x bh x bt x bb) voxels.
// For tap (dt,dh,dw) the A operand is the same box shifted by (dt-pt, dh-ph, dw-pw): one 5-D TMA load
// whose out-of-bounds elements are zero-filled by the hardware -- that IS the im2col, including the
// zero padding, and the 128B-swizzled tile TMA writes is exactly the K-major UMMA operand layout.
// Weights [taps, Cout, Cin] come in through a 3-D TMA map (box kc x n_tile x 1).
//
// Precision.  The parity bar is fp32 (1e-4 relative through ~14 stacked convs), out of reach of one
// fp16/bf16/tf32 product.  Both operands are pre-split x = (hi + lo)/s with hi = fp16(s x),
// lo = fp16(s x - hi) (power-of-two scales, exact), and the tile accumulates
// hi*hi + hi*lo + lo*hi in fp32 TMEM: 22 significand bits per operand at 3 kind::f16 MMAs, i.e. 1.5x the
// cost of one TF32 pass and half the cost of 3xTF32.  (`terms`=1 keeps only hi*hi: fp16-grade fast mode.)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> scale, bias, residual through the nearest-upsample map,
// activation -> global).  smem full/empty mbarrier ring between producer and MMA, one tmem_full barrier
// between MMA and epilogue.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include "ptx_sm100.cuh"

namespace i2v {

namespace {

constexpr int TC_THREADS = 192;
constexpr int TILE_M = 128;

struct ConvTcKArgs {
    const float* bias; const float* res; const float* scale_ptr; float* y;
    int B, T, H, W, Cin, Cout;
    int kt, kh, kw;
    int bw, bh, bt, bb;
    int tiles_w, tiles_h, tiles_t;
    int n_tile, kc, stages, terms, nacc;
    int res_ut, res_uh, res_uw, act, out_mode;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
               const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const ConvTcKArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t rb = (uint32_t)a.kc * 2;                                   // bytes per operand row
    const uint32_t a_bytes = TILE_M * rb;
    const uint32_t b_bytes = ((uint32_t)a.n_tile * rb + 1023u) & ~1023u;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
    uint64_t* empty = full + a.stages;
    uint64_t* tmem_full = empty + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x;
    const int tw = tile % a.tiles_w; tile /= a.tiles_w;
    const int th = tile % a.tiles_h; tile /= a.tiles_h;
    const int tt = tile % a.tiles_t;
    const int tb = tile / a.tiles_t;
    const int w0 = tw * a.bw, h0 = th * a.bh, t0 = tt * a.bt, b0 = tb * a.bb;
    const int n0 = blockIdx.y * a.n_tile;
    const int taps = a.kt * a.kh * a.kw, cchunks = a.Cin / a.kc, iters = taps * cchunks;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(a.n_tile * a.nacc)) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mAh); ptx::prefetch_tensormap(&mBh);
        if (a.terms > 1) { ptx::prefetch_tensormap(&mAl); ptx::prefetch_tensormap(&mBl); }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < a.stages; ++s) { ptx::mbar_init(full + s, 1); ptx::mbar_init(empty + s, 1); }
            ptx::mbar_init(tmem_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, ncols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer (one lane)
        if (lane == 0) {
            const uint32_t tx = (a.terms > 1 ? 2u : 1u) * (a_bytes + (uint32_t)a.n_tile * rb);
            for (int it = 0; it < iters; ++it) {
                const int s = it % a.stages;
                const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
                ptx::mbar_wait(empty + s, ph ^ 1u);
                const int tap = it / cchunks, c0 = (it - tap * cchunks) * a.kc;
                const int dw = tap % a.kw, dh = (tap / a.kw) % a.kh, dt = tap / (a.kw * a.kh);
                const int cw = w0 + dw - a.kw / 2, ch = h0 + dh - a.kh / 2, ct = t0 + dt - a.kt / 2;
                uint8_t* st = smem + (size_t)s * stage_bytes;
                ptx::mbar_expect_tx(full + s, tx);
                ptx::tma_load_5d(st, &mAh, full + s, c0, cw, ch, ct, b0);
                ptx::tma_load_3d(st + 2 * a_bytes, &mBh, full + s, c0, n0, tap);
                if (a.terms > 1) {
                    ptx::tma_load_5d(st + a_bytes, &mAl, full + s, c0, cw, ch, ct, b0);
                    ptx::tma_load_3d(st + 2 * a_bytes + b_bytes, &mBl, full + s, c0, n0, tap);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t idesc = ptx::make_idesc_f16(TILE_M, a.n_tile);
            const int ksteps = a.kc / 16;
            for (int it = 0; it < iters; ++it) {
                const int s = it % a.stages;
                const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
                ptx::mbar_wait(full + s, ph);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t sb = sa + 2 * a_bytes;
                // round-robin over `nacc` TMEM accumulators: the tensor core's fp32 accumulate truncates, so
                // the chain of dependent adds per accumulator is cut nacc-fold and the partial sums are
                // combined with round-to-nearest fp32 adds in the epilogue
                const uint32_t tacc = tmem_base + (uint32_t)((it % a.nacc) * a.n_tile);
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t off = (uint32_t)k * 32u;       // 16 fp16 along K inside the swizzle span
                    const uint64_t dAh = ptx::make_kmajor_desc(sa + off, rb);
                    const uint64_t dBh = ptx::make_kmajor_desc(sb + off, rb);
                    ptx::mma_f16_ss(tacc, dAh, dBh, idesc, (it >= a.nacc || k > 0) ? 1u : 0u);
                    if (a.terms > 1) {
                        const uint64_t dAl = ptx::make_kmajor_desc(sa + a_bytes + off, rb);
                        const uint64_t dBl = ptx::make_kmajor_desc(sb + b_bytes + off, rb);
                        ptx::mma_f16_ss(tacc, dAh, dBl, idesc, 1u);
                        ptx::mma_f16_ss(tacc, dAl, dBh, idesc, 1u);
                    }
                }
                ptx::mma_commit(empty + s);        // frees the smem stage once these MMAs have read it
            }
            ptx::mma_commit(tmem_full);            // accumulator complete
        }
    } else {
        // ================================ epilogue (4 warps, one TMEM lane quarter each)
        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();
        const int q = warp & 3;
        const int m = q * 32 + lane;
        int r = m;
        const int wi = r % a.bw; r /= a.bw;
        const int hi = r % a.bh; r /= a.bh;
        const int ti = r % a.bt;
        const int bi = r / a.bt;
        const int b = b0 + bi, t = t0 + ti, h = h0 + hi, w = w0 + wi;
        const bool valid = b < a.B;
        const float scale = __ldg(a.scale_ptr);
        const long long vox = (((long long)b * a.T + t) * a.H + h) * a.W + w;
        long long roff = 0;
        if (a.res != nullptr && valid) {
            const int Tr = a.T / a.res_ut, Hr = a.H / a.res_uh, Wr = a.W / a.res_uw;
            roff = ((((long long)b * Tr + t / a.res_ut) * Hr + h / a.res_uh) * Wr + w / a.res_uw) * a.Cout;
        }
        for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
            uint32_t rr[16];
            float accv[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, rr);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) accv[j] = __uint_as_float(rr[j]);
            for (int ai = 1; ai < a.nacc; ++ai) {
                ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ai * a.n_tile + c0), rr);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) accv[j] += __uint_as_float(rr[j]);
            }
            const int nb = n0 + c0;
            if (!valid || nb >= a.Cout) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = nb + j;
                float x = accv[j] * scale;
                if (n < a.Cout) {
                    if (a.bias != nullptr) x += __ldg(a.bias + n);
                    if (a.res != nullptr) x += __ldg(a.res + roff + n);
                }
                v[j] = apply_act(x, a.act);
            }
            if (a.out_mode == 0) {
                float* dst = a.y + vox * a.Cout + nb;
                if ((a.Cout & 3) == 0 && nb + 15 < a.Cout) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (nb + j < a.Cout) dst[j] = v[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (nb + j < a.Cout)
                        a.y[((((long long)b * a.T + t) * a.Cout + nb + j) * a.H + h) * a.W + w] = v[j];
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, ncols);
}

__global__ void split_fp16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, float scale,
                                  long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i] * scale;
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, int row_bytes) {
    EncodeTiledFn fn = encode_fn();
    I2V_REQUIRE(fn != nullptr, "conv_tc: cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    I2V_REQUIRE(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return 0;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

bool conv_tc_supported(int B, int T, int H, int W, int Cin, int Cout, int kt, int kh, int kw) {
    (void)B; (void)Cout;
    if (!(kt == 1 || kt == 3) || !(kh == 1 || kh == 3) || !(kw == 1 || kw == 3)) return false;
    if (Cin % 16 != 0) return false;
    if (!is_pow2(W) || !is_pow2(H) || !is_pow2(T)) return false;
    return true;
}

int launch_split_fp16(const float* x, __half* hi, __half* lo, float scale, long long n, cudaStream_t stream) {
    long long blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    ProfScope ps(PROF_OTHER, 0, 0, stream);
    split_fp16_kernel<<<(int)blocks, 256, 0, stream>>>(x, hi, lo, scale, n);
    I2V_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_conv_tc(const ConvTcArgs& h, cudaStream_t stream) {
    I2V_REQUIRE(conv_tc_supported(h.B, h.T, h.H, h.W, h.Cin, h.Cout, h.kt, h.kh, h.kw),
                "conv_tc: unsupported shape B=%d T=%d H=%d W=%d Cin=%d k=(%d,%d,%d)", h.B, h.T, h.H, h.W, h.Cin, h.kt, h.kh, h.kw);
    I2V_REQUIRE(h.terms == 1 || h.terms == 3, "conv_tc: terms must be 1 or 3");
    I2V_REQUIRE(h.cout_pad % 16 == 0 && h.cout_pad >= h.Cout, "conv_tc: weights must be padded to a multiple of 16 output rows");
    ConvTcKArgs a;
    a.bias = h.bias; a.res = h.res; a.scale_ptr = h.scale_ptr; a.y = h.y;
    a.B = h.B; a.T = h.T; a.H = h.H; a.W = h.W; a.Cin = h.Cin; a.Cout = h.Cout;
    a.kt = h.kt; a.kh = h.kh; a.kw = h.kw;
    a.kc = h.Cin % 64 == 0 ? 64 : (h.Cin % 32 == 0 ? 32 : 16);
    // box: 128 rows = bw x bh x bt x bb voxels
    int rem = TILE_M;
    a.bw = h.W < rem ? h.W : rem; rem /= a.bw;
    a.bh = h.H < rem ? h.H : rem; rem /= a.bh;
    a.bt = h.T < rem ? h.T : rem; rem /= a.bt;
    a.bb = rem;
    a.tiles_w = h.W / a.bw; a.tiles_h = h.H / a.bh; a.tiles_t = h.T / a.bt;
    const int tiles_b = (h.B + a.bb - 1) / a.bb;
    // fp32-grade mode keeps N <= 128 so that four TMEM accumulators fit (see the MMA issuer)
    const int n_cap = h.terms == 3 ? 128 : 256;
    a.n_tile = h.cout_pad < n_cap ? h.cout_pad : n_cap;
    a.terms = h.terms;
    {
        const int iters = h.kt * h.kh * h.kw * (h.Cin / a.kc);
        int nacc = 512 / a.n_tile;
        if (nacc > 4) nacc = 4;
        if (nacc > iters) nacc = iters;
        a.nacc = nacc;
    }
    a.res_ut = h.res_ut; a.res_uh = h.res_uh; a.res_uw = h.res_uw; a.act = h.act; a.out_mode = h.out_mode;
    I2V_REQUIRE(h.res == nullptr || (h.T % h.res_ut == 0 && h.H % h.res_uh == 0 && h.W % h.res_uw == 0),
                "conv_tc: residual upsample factors must divide the output size");
    const int rb = a.kc * 2;
    const size_t a_bytes = (size_t)TILE_M * rb, b_bytes = ((size_t)a.n_tile * rb + 1023) & ~(size_t)1023;
    const size_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const size_t budget = 220 * 1024;
    int stages = (int)((budget - 2048) / stage_bytes);
    if (stages > 8) stages = 8;
    I2V_REQUIRE(stages >= 2, "conv_tc: tile does not fit two pipeline stages");
    a.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;

    CUtensorMap mAh, mAl, mBh, mBl;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)h.T, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.W * h.Cin * 2, (cuuint64_t)h.H * h.W * h.Cin * 2,
                                  (cuuint64_t)h.T * h.H * h.W * h.Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)a.kc, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bt, (cuuint32_t)a.bb};
        if (int rc = encode_map(&mAh, h.x_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mAl, h.terms > 1 ? h.x_lo : h.x_hi, 5, dims, st, box, rb)) return rc;
    }
    {
        const int taps = h.kt * h.kh * h.kw;
        const cuuint64_t dims[3] = {(cuuint64_t)h.Cin, (cuuint64_t)h.cout_pad, (cuuint64_t)taps};
        const cuuint64_t st[2] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.cout_pad * h.Cin * 2};
        const cuuint32_t box[3] = {(cuuint32_t)a.kc, (cuuint32_t)a.n_tile, 1};
        if (int rc = encode_map(&mBh, h.w_hi, 3, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mBl, h.terms > 1 ? h.w_lo : h.w_hi, 3, dims, st, box, rb)) return rc;
    }
    static bool attr_set = false;
    if (!attr_set) {
        I2V_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const long long M = (long long)h.B * h.T * h.H * h.W;
    const double K_ = (double)h.kt * h.kh * h.kw * h.Cin;
    ProfScope ps(PROF_CONV, 2.0 * (double)M * h.Cout * K_, 4.0 * ((double)M * h.Cin + (double)M * h.Cout + K_ * h.Cout), stream);
    dim3 grid((unsigned)(a.tiles_w * a.tiles_h * a.tiles_t * tiles_b), (unsigned)((h.cout_pad + a.n_tile - 1) / a.n_tile));
    conv_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(mAh, mAl, mBh, mBl, a);
    I2V_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace i2v
