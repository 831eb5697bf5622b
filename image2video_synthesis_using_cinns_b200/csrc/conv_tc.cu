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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue, two per TMEM lane quarter (tcgen05.ld -> transpose through shared memory -> scale, bias,
// residual through the nearest-upsample map, activation, per-(sample, channel) statistics -> global).  smem full/empty
// mbarrier ring between producer and MMA, a tmem_full barrier between MMA and epilogue; the halo kernel
// (conv_tc_halo_kernel, the one that carries the decoder) walks its tiles persistently and adds an epi_done barrier
// that hands TMEM and the transpose tiles back to the MMA issuer and the producer.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include "ptx_sm100.cuh"

namespace i2v {

namespace {

constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

// optional phase timestamps (i2v_debug_conv_tc_timestamps): 16 x u64 per CTA, %globaltimer in ns
__device__ unsigned long long* g_dbg = nullptr;
__device__ int g_dbg_ctas = 0;
__device__ __forceinline__ void dbg_stamp(int slot, int tile) {      // one row of stamps per output tile
    if (g_dbg != nullptr && tile < g_dbg_ctas && blockIdx.y == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_dbg[(size_t)tile * 16 + slot] = t;
    }
}
constexpr int TILE_M = 128;
constexpr int kMaxChain = 300;   // longest run of truncating hi*hi MMAs into one TMEM accumulator (see launch_conv_tc_halo)

struct ConvTcKArgs {
    const float* bias; const float* res; const float* scale_ptr; float* y; double* stats;
    int B, T, H, W, Cin, Cout;
    int kt, kh, kw;
    int bw, bh, bt, bb;
    int tiles_w, tiles_h, tiles_t;
    int n_tile, kc, stages, terms, nacc;   // nacc accumulators per (sub-)tile ...
    int nmain;                             // ... of which the first nmain take hi*hi, the rest the small cross terms
    int res_ut, res_uh, res_uw, act, out_mode;
    int cc_lo, cc_hi;             // channel-chunk range [cc_lo, cc_hi) of this launch (K split across launches)
    // 1: temporal phase form (see halo_tile): T is the OUTPUT frame count, the input holds T/2 planes, tiles walk the INPUT
    // planes, blockIdx.z is the output phase p (frame 2j + p) and the weights are the 2 x 2 phase-combined temporal taps
    int t_phase;
};

struct EpiArgs {
    const float* bias; const float* res; float* y;
    int T, H, W, Cout, res_ut, res_uh, res_uw, act, out_mode;
    int contig;       // the 128 rows of a (sub-)tile are consecutive voxels of the output tensor
};

// Finish 16 consecutive output columns [nb, nb+16) of ONE output row held in registers: scale, bias, residual,
// activation, store (channels-last, or the (B,T,C,H,W) frame layout of conv_img when out_mode == 1).
__device__ __forceinline__ void row_finish(const EpiArgs& e, const float (&accv)[16], int nb, float scale, long long vox, long long roff,
                                           int b, int t, int h, int w) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = nb + j;
        float x = accv[j] * scale;
        if (n < e.Cout) {
            if (e.bias != nullptr) x += __ldg(e.bias + n);
            if (e.res != nullptr) x += __ldg(e.res + roff + n);
        }
        v[j] = apply_act(x, e.act);
    }
    if (e.out_mode == 0) {
        float* dst = e.y + vox * e.Cout + nb;
        if ((e.Cout & 3) == 0 && nb + 15 < e.Cout) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (nb + j < e.Cout) dst[j] = v[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (nb + j < e.Cout) e.y[((((long long)b * e.T + t) * e.Cout + nb + j) * e.H + h) * e.W + w] = v[j];
    }
}

// One output row (voxel) per thread: sum `nacc` TMEM accumulators (column stride `acc_stride`), then
// scale, bias, residual through the nearest-upsample map, activation, store.  tcgen05.ld is warp-collective,
// so every lane walks all column chunks and only the stores are predicated.
__device__ __forceinline__ void epilogue_row(const EpiArgs& e, uint32_t tmem_lane_base, int n_tile, int nacc, int acc_stride, int n0,
                                             float scale, int b, int t, int h, int w, bool valid) {
    const long long vox = (((long long)b * e.T + t) * e.H + h) * e.W + w;
    long long roff = 0;
    if (e.res != nullptr && valid) {
        const int Tr = e.T / e.res_ut, Hr = e.H / e.res_uh, Wr = e.W / e.res_uw;
        roff = ((((long long)b * Tr + t / e.res_ut) * Hr + h / e.res_uh) * Wr + w / e.res_uw) * e.Cout;
    }
    for (int c0 = 0; c0 < n_tile; c0 += 16) {
        uint32_t rr[16];
        float accv[16];
        ptx::tmem_ld_32x32b_x16(tmem_lane_base + (uint32_t)c0, rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) accv[j] = __uint_as_float(rr[j]);
        for (int ai = 1; ai < nacc; ++ai) {
            ptx::tmem_ld_32x32b_x16(tmem_lane_base + (uint32_t)(ai * acc_stride + c0), rr);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) accv[j] += __uint_as_float(rr[j]);
        }
        const int nb = n0 + c0;
        if (!valid || nb >= e.Cout) continue;
        row_finish(e, accv, nb, scale, vox, roff, b, t, h, w);
    }
}

// Branch-free write-out for the common case: linear epilogue (ACT_NONE), every lane owns a whole float4 column group
// of N_IT real rows (slice width 32 or 64 columns, all inside Cout).  The generic loop below carries per-row
// predicates and a runtime activation switch, which ptxas turns into branches; with ONE epilogue warp per scheduler
// that code ran at ~14 cycles per instruction (ncu: stall_wait / no_inst / branch_resolving).  Here everything is
// unrolled straight-line code: all residual loads first, then rows in independent groups.
template <int N_IT, bool HAS_RES>
__device__ __forceinline__ void writeout_fast(const EpiArgs& e, const float* stile, int ld, int c, int n, int rsub, long long vox_lane,
                                              long long roff_lane, const float4 b4, float (&ssum)[4], float (&ssq)[4]) {
    constexpr int RPI = 32 / N_IT;               // rows per iteration (lanes_per_row = N_IT)
    // the 32 rows of a warp slice are consecutive voxels (a box row is a full-width run, see the tile shapes), so the
    // output row pointer is linear in r; only the residual goes through the (non-linear) upsample map
    float* const y0 = e.y + __shfl_sync(0xffffffffu, vox_lane, 0) * e.Cout + n;
    constexpr int BATCH = 8;                     // residual float4 in flight per lane (register budget: 168 per thread)
#pragma unroll
    for (int k0 = 0; k0 < N_IT; k0 += BATCH) {
        float4 rv[HAS_RES ? BATCH : 1];
        if (HAS_RES) {
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const long long roff = __shfl_sync(0xffffffffu, roff_lane, (k0 + k) * RPI + rsub);
                rv[k] = __ldg(reinterpret_cast<const float4*>(e.res + roff + n));
            }
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int r = (k0 + k) * RPI + rsub;
            const float4 a4 = *reinterpret_cast<const float4*>(stile + r * ld + c);
            float4 o4 = make_float4(a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w);
            if (HAS_RES) { o4.x += rv[k].x; o4.y += rv[k].y; o4.z += rv[k].z; o4.w += rv[k].w; }
            *reinterpret_cast<float4*>(y0 + r * e.Cout) = o4;
            ssum[0] += o4.x; ssum[1] += o4.y; ssum[2] += o4.z; ssum[3] += o4.w;
            ssq[0] = fmaf(o4.x, o4.x, ssq[0]); ssq[1] = fmaf(o4.y, o4.y, ssq[1]);
            ssq[2] = fmaf(o4.z, o4.z, ssq[2]); ssq[3] = fmaf(o4.w, o4.w, ssq[3]);
        }
    }
}

// Write-out half of the coalesced epilogue: `stile` holds this warp's 32 rows x ncols scaled sums (row stride
// ncols+4); lanes own fixed column groups, rows are walked with shuffled voxel / residual offsets.
__device__ __forceinline__ void epilogue_writeout(const EpiArgs& e, const float* stile, int col0, int ncols, int n0, int lane,
                                                  long long vox_lane, long long roff_lane, bool want_stats, float (&ssum)[4],
                                                  float (&ssq)[4]) {
    const int ld = ncols + 4;
    // lane -> (row sub-index, column group); c4n float4 per row
    const int c4n = ncols >> 2;
    const int lanes_per_row = c4n < 32 ? c4n : 32;
    const int rows_per_iter = 32 / lanes_per_row;
    const int rsub = lane / lanes_per_row, cl = lane - rsub * lanes_per_row;
    const bool vec_ok = (e.Cout & 3) == 0;
    for (int cg = cl; cg < c4n; cg += 32) {          // more than one pass only when the slice is wider than 128
        const int c = cg * 4, n = n0 + col0 + c;
        const bool full4 = vec_ok && n + 3 < e.Cout;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.bias != nullptr && n < e.Cout) {
            if (full4) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n));
            else {
                b4.x = __ldg(e.bias + n);
                if (n + 1 < e.Cout) b4.y = __ldg(e.bias + n + 1);
                if (n + 2 < e.Cout) b4.z = __ldg(e.bias + n + 2);
                if (n + 3 < e.Cout) b4.w = __ldg(e.bias + n + 3);
            }
        }
        // warp-uniform: whole slice inside Cout, 8 or 16 float4 column groups, linear epilogue
        if (e.contig && e.act == ACT_NONE && vec_ok && (c4n == 8 || c4n == 16) && n0 + col0 + ncols <= e.Cout) {
            if (c4n == 16) {
                if (e.res != nullptr) writeout_fast<16, true>(e, stile, ld, c, n, rsub, vox_lane, roff_lane, b4, ssum, ssq);
                else writeout_fast<16, false>(e, stile, ld, c, n, rsub, vox_lane, roff_lane, b4, ssum, ssq);
            } else {
                if (e.res != nullptr) writeout_fast<8, true>(e, stile, ld, c, n, rsub, vox_lane, roff_lane, b4, ssum, ssq);
                else writeout_fast<8, false>(e, stile, ld, c, n, rsub, vox_lane, roff_lane, b4, ssum, ssq);
            }
            continue;
        }
#pragma unroll 4
        for (int i = 0; i < 32; i += rows_per_iter) {
            const int r = (i + rsub) & 31;
            const long long vox = __shfl_sync(0xffffffffu, vox_lane, r);
            const long long roff = __shfl_sync(0xffffffffu, roff_lane, r);
            if (n >= e.Cout || rsub >= rows_per_iter) continue;     // spare lanes when c4n does not divide 32
            const float4 a4 = *reinterpret_cast<const float4*>(stile + r * ld + c);
            float v[4] = {a4.x + b4.x, a4.y + b4.y, a4.z + b4.z, a4.w + b4.w};
            if (full4) {
                if (e.res != nullptr) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(e.res + roff + n));
                    v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
                }
                const float4 o4 = make_float4(apply_act(v[0], e.act), apply_act(v[1], e.act), apply_act(v[2], e.act), apply_act(v[3], e.act));
                *reinterpret_cast<float4*>(e.y + vox * e.Cout + n) = o4;
                if (want_stats) {   // per-(sample, channel) sum / sum of squares of the stored values (feeds the next norm)
                    ssum[0] += o4.x; ssum[1] += o4.y; ssum[2] += o4.z; ssum[3] += o4.w;
                    ssq[0] = fmaf(o4.x, o4.x, ssq[0]); ssq[1] = fmaf(o4.y, o4.y, ssq[1]);
                    ssq[2] = fmaf(o4.z, o4.z, ssq[2]); ssq[3] = fmaf(o4.w, o4.w, ssq[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < e.Cout) {
                        float x = v[j];
                        if (e.res != nullptr) x += __ldg(e.res + roff + n + j);
                        x = apply_act(x, e.act);
                        e.y[vox * e.Cout + n + j] = x;
                        if (want_stats) { ssum[j] += x; ssq[j] = fmaf(x, x, ssq[j]); }
                    }
            }
        }
    }
    __syncwarp();
}

// Coalesced variant for channels-last output: the 32 rows a warp pulls out of TMEM are transposed through a
// padded shared-memory tile (the drained pipeline buffers) so that global stores, the residual read and the
// bias read run along the channel dimension (full 128-byte lines).  The epilogue runs on ONE warp per
// scheduler, i.e. it is bound by dependent-instruction latency, not bandwidth (measured: ~30 us per 256x128
// tile with per-element index arithmetic): each lane therefore owns a fixed column group, row addresses come
// from a warp shuffle of the per-lane voxel index, and the row loop is unrolled for memory-level parallelism.
// `vox_lane` / `roff_lane`: output voxel index and residual element offset of THIS lane's row.
__device__ __forceinline__ void epilogue_warp_coalesced(const EpiArgs& e, float* stile /* [32][ncols+4] */, uint32_t tmem_lane_base,
                                                        int col0, int ncols, int nacc, int acc_stride, int n0, float scale, int lane,
                                                        long long vox_lane, long long roff_lane, bool want_stats, float (&ssum)[4],
                                                        float (&ssq)[4], int dbg_slot = -1, int dbg_tile = 0) {
    // this warp owns tile columns [col0, col0+ncols) of its 32 rows
    const int ld = ncols + 4;
    int c0 = 0;
    for (; c0 + 32 <= ncols; c0 += 32) {
        // up to two accumulators in flight per wait: a TMEM load round trip costs hundreds of cycles
        uint32_t ra[32], rb2[32];
        float accv[32];
        ptx::tmem_ld_32x32b_x32(tmem_lane_base + (uint32_t)(col0 + c0), ra);
        if (nacc > 1) ptx::tmem_ld_32x32b_x32(tmem_lane_base + (uint32_t)(acc_stride + col0 + c0), rb2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) accv[j] = __uint_as_float(ra[j]) + (nacc > 1 ? __uint_as_float(rb2[j]) : 0.f);
        if (nacc > 2) {
            ptx::tmem_ld_32x32b_x32(tmem_lane_base + (uint32_t)(2 * acc_stride + col0 + c0), ra);
            if (nacc > 3) ptx::tmem_ld_32x32b_x32(tmem_lane_base + (uint32_t)(3 * acc_stride + col0 + c0), rb2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) accv[j] += __uint_as_float(ra[j]) + (nacc > 3 ? __uint_as_float(rb2[j]) : 0.f);
        }
        float4* dst = reinterpret_cast<float4*>(stile + lane * ld + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(accv[4 * j] * scale, accv[4 * j + 1] * scale, accv[4 * j + 2] * scale, accv[4 * j + 3] * scale);
    }
    for (; c0 < ncols; c0 += 16) {
        uint32_t rr[16];
        float accv[16];
        ptx::tmem_ld_32x32b_x16(tmem_lane_base + (uint32_t)(col0 + c0), rr);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) accv[j] = __uint_as_float(rr[j]);
        for (int ai = 1; ai < nacc; ++ai) {
            ptx::tmem_ld_32x32b_x16(tmem_lane_base + (uint32_t)(ai * acc_stride + col0 + c0), rr);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) accv[j] += __uint_as_float(rr[j]);
        }
        float4* dst = reinterpret_cast<float4*>(stile + lane * ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(accv[4 * j] * scale, accv[4 * j + 1] * scale, accv[4 * j + 2] * scale, accv[4 * j + 3] * scale);
    }
    __syncwarp();
    if (dbg_slot >= 0 && threadIdx.x == 64) dbg_stamp(dbg_slot, dbg_tile);          // this sub-tile is out of TMEM
    epilogue_writeout(e, stile, col0, ncols, n0, lane, vox_lane, roff_lane, want_stats, ssum, ssq);
    if (dbg_slot >= 0 && threadIdx.x == 64) dbg_stamp(dbg_slot + 1, dbg_tile);      // ... and its stores are issued
}

// flush a lane's column-group partial sums into stats[b, c, {sum, sumsq}] (double, device-wide atomics)
__device__ __forceinline__ void flush_stats(double* stats_b, int Cout, int n0, int col0, int ncols, int lane, const float (&ssum)[4],
                                            const float (&ssq)[4]) {
    const int c4n = ncols >> 2;
    const int lanes_per_row = c4n < 32 ? c4n : 32;
    const int rsub = lane / lanes_per_row, cl = lane - rsub * lanes_per_row;
    float s[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j] = ssum[j]; q[j] = ssq[j]; }
    // lanes that own the same column group (different row phases) combine first: fewer same-address atomics
    if ((lanes_per_row & (lanes_per_row - 1)) == 0) {
        for (int off = lanes_per_row; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
                q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
            }
        }
        if (rsub != 0) return;
    } else if (rsub >= 32 / lanes_per_row) {
        return;
    }
    const int n = n0 + col0 + cl * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (n + j < Cout) {
            atomicAdd(stats_b + 2 * (size_t)(n + j), (double)s[j]);
            atomicAdd(stats_b + 2 * (size_t)(n + j) + 1, (double)q[j]);
        }
}


// ---- kw-stacked ("wstack") epilogue pieces (halo kernel, narrow layers) ------------------------------------
// The three kw taps of a narrow layer (Cout <= 64) are stacked along the MMA's N: one UNSHIFTED activation box
// is multiplied by [W(kw=0) | W(kw=1) | W(kw=2)], so column group kw of accumulator row m' holds the contribution
// of INPUT voxel m' through tap kw.  Output voxel m = m' - (kw-1):
//     out[m] = D1[m] + D0[m-1] * [w > 0] + D2[m+1] * [w < W-1]
// (the predicates are the conv's zero padding; tiles hold whole w-rows, so no halo is needed along w).
// Rows m-1 / m+1 live in the neighbouring lanes: warp shuffles, plus a small shared-memory exchange for the
// first / last lane of each 32-row warp slice (published in `xd0` / `xd2`, consumed after a named barrier).
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// 16 columns [c, c+16) of the three kw groups of this lane's row, summed over the `nacc` accumulators
__device__ __forceinline__ void wstack_chunk16(uint32_t tbase, int n_tile, int nacc, int accw, int c, float (&d0)[16], float (&d1)[16],
                                               float (&d2)[16]) {
    uint32_t r0[16], r1[16], r2[16];
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)c, r0);
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)(n_tile + c), r1);
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)(2 * n_tile + c), r2);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) { d0[j] = __uint_as_float(r0[j]); d1[j] = __uint_as_float(r1[j]); d2[j] = __uint_as_float(r2[j]); }
    for (int ai = 1; ai < nacc; ++ai) {
        const uint32_t tb = tbase + (uint32_t)(ai * accw);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)c, r0);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(n_tile + c), r1);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(2 * n_tile + c), r2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) { d0[j] += __uint_as_float(r0[j]); d1[j] += __uint_as_float(r1[j]); d2[j] += __uint_as_float(r2[j]); }
    }
}

// Same, with the first two accumulators in flight under ONE wait (a TMEM load round trip costs ~0.5 us; conv_img's
// one-row-per-thread write-out is nothing but such round trips).
__device__ __forceinline__ void wstack_chunk16_x2(uint32_t tbase, int n_tile, int nacc, int accw, int c, float (&d0)[16], float (&d1)[16],
                                                  float (&d2)[16]) {
    if (nacc < 2) { wstack_chunk16(tbase, n_tile, nacc, accw, c, d0, d1, d2); return; }
    uint32_t r0[16], r1[16], r2[16], s0[16], s1[16], s2[16];
    const uint32_t tb1 = tbase + (uint32_t)accw;
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)c, r0);
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)(n_tile + c), r1);
    ptx::tmem_ld_32x32b_x16(tbase + (uint32_t)(2 * n_tile + c), r2);
    ptx::tmem_ld_32x32b_x16(tb1 + (uint32_t)c, s0);
    ptx::tmem_ld_32x32b_x16(tb1 + (uint32_t)(n_tile + c), s1);
    ptx::tmem_ld_32x32b_x16(tb1 + (uint32_t)(2 * n_tile + c), s2);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        d0[j] = __uint_as_float(r0[j]) + __uint_as_float(s0[j]);
        d1[j] = __uint_as_float(r1[j]) + __uint_as_float(s1[j]);
        d2[j] = __uint_as_float(r2[j]) + __uint_as_float(s2[j]);
    }
    for (int ai = 2; ai < nacc; ++ai) {
        const uint32_t tb = tbase + (uint32_t)(ai * accw);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)c, r0);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(n_tile + c), r1);
        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(2 * n_tile + c), r2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) { d0[j] += __uint_as_float(r0[j]); d1[j] += __uint_as_float(r1[j]); d2[j] += __uint_as_float(r2[j]); }
    }
}

// in-warp part of the shifted sum; lanes 0 / 31 still miss their cross-warp neighbour (added after the barrier)
__device__ __forceinline__ void wstack_combine(const float (&d0)[16], const float (&d1)[16], const float (&d2)[16], int lane, bool left_ok,
                                               bool right_ok, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float up = __shfl_up_sync(0xffffffffu, d0[j], 1);
        const float dn = __shfl_down_sync(0xffffffffu, d2[j], 1);
        v[j] = d1[j] + ((left_ok && lane > 0) ? up : 0.f) + ((right_ok && lane < 31) ? dn : 0.f);
    }
}

// Gather half of the coalesced epilogue for the stacked layout: fills `stile` like epilogue_warp_coalesced does and
// publishes this warp's edge rows.  xd0 / xd2: [n_tile] floats of this warp's lane quarter (tile-column indexed).
__device__ __forceinline__ void wstack_gather(float* stile, uint32_t tbase, int col0, int ncols, int n_tile, int nacc, int accw, float scale,
                                              int lane, bool left_ok, bool right_ok, float* xd0, float* xd2) {
    const int ld = ncols + 4;
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        float d0[16], d1[16], d2[16], v[16];
        wstack_chunk16(tbase, n_tile, nacc, accw, col0 + c0, d0, d1, d2);
        wstack_combine(d0, d1, d2, lane, left_ok, right_ok, v);
        if (lane == 31) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xd0[col0 + c0 + j] = d0[j];
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xd2[col0 + c0 + j] = d2[j];
        }
        float4* dst = reinterpret_cast<float4*>(stile + lane * ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(v[4 * j] * scale, v[4 * j + 1] * scale, v[4 * j + 2] * scale, v[4 * j + 3] * scale);
    }
}

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
    const int taps = a.kt * a.kh * a.kw, cchunks = a.cc_hi - a.cc_lo, iters = taps * cchunks;
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
    // warp-uniform by construction; the shuffle tells the compiler so (TMEM addresses feed uniform registers)
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // Only now may the successor be scheduled: this CTA already owns its TMEM columns, so a co-resident CTA of the
    // next kernel can never take them first and then sit in its own pdl_wait() while this one starves.
    pdl_launch_dependents();
    pdl_wait();      // prologue above touched no global memory: it overlapped the previous kernel's tail

    // A tap whose shifted box lies entirely in the zero padding contributes nothing (head_0: T = 1, so 18 of
    // the 27 taps): producer, issuer and epilogue all skip it.
    // Valid taps form a contiguous range per dimension; computed once so that the role loops carry no
    // divisions (they run on a single warp and are latency-bound).
    auto tap_range = [](int k, int origin, int box, int extent, int& lo, int& hi) {
        lo = k / 2 - origin - box + 1; if (lo < 0) lo = 0;         // first d with origin + d - k/2 + box > 0
        hi = extent - 1 - origin + k / 2; if (hi > k - 1) hi = k - 1;   // last d with origin + d - k/2 < extent
    };
    int dt_lo, dt_hi, dh_lo, dh_hi, dw_lo, dw_hi;
    // Phase form (g_0.conv_0: 8x8 planes behind a x2 temporal upsample): output frame 2j + p reads the source planes
    // j - 1 + p + q, q = 0, 1, with the pre-summed weight slabs 2p + q; t0 counts SOURCE planes.  With one source plane
    // (T = 2) a single temporal tap per phase survives the range check: a third of the plain form's MMAs.
    const int ph_out = a.t_phase ? (int)blockIdx.z : 0;
    const int t_shift = a.t_phase ? 1 - ph_out : a.kt / 2;        // source plane of temporal tap d: t0 + d - t_shift
    if (a.t_phase) {
        dt_lo = t_shift - t0 - a.bt + 1; if (dt_lo < 0) dt_lo = 0;
        dt_hi = a.T / 2 - 1 - t0 + t_shift; if (dt_hi > 1) dt_hi = 1;
    } else {
        tap_range(a.kt, t0, a.bt, a.T, dt_lo, dt_hi);
    }
    tap_range(a.kh, h0, a.bh, a.H, dh_lo, dh_hi);
    tap_range(a.kw, w0, a.bw, a.W, dw_lo, dw_hi);
    const int n_total = (dt_hi - dt_lo + 1) * (dh_hi - dh_lo + 1) * (dw_hi - dw_lo + 1) * cchunks;   // >= 1: the centre tap
    (void)iters;

    // Role loops run on the WHOLE warp with warp-uniform control flow; only the asynchronous instructions are
    // predicated on one elected lane.  (Running the loop under `if (lane == 0)` makes every operand a
    // divergent value that has to be moved into uniform registers one MMA at a time.)
    if (warp == 0) {
        // ================================ TMA producer
        {
            const uint32_t tx = (a.terms > 1 ? 2u : 1u) * (a_bytes + (uint32_t)a.n_tile * rb);
            int s = 0;
            uint32_t ph = 0;
            for (int dt = dt_lo; dt <= dt_hi; ++dt)
            for (int dh = dh_lo; dh <= dh_hi; ++dh)
            for (int dw = dw_lo; dw <= dw_hi; ++dw)
            for (int cc = a.cc_lo; cc < a.cc_hi; ++cc) {
                const int tap = ((2 * ph_out + dt) * a.kh + dh) * a.kw + dw, c0 = cc * a.kc;     // ph_out = 0 outside the phase form
                const int cw = w0 + dw - a.kw / 2, ch = h0 + dh - a.kh / 2, ct = t0 + dt - t_shift;
                ptx::mbar_wait(empty + s, ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(full + s, tx);
                    ptx::tma_load_5d(st, &mAh, full + s, c0, cw, ch, ct, b0);
                    ptx::tma_load_3d(st + 2 * a_bytes, &mBh, full + s, c0, n0, tap);
                    if (a.terms > 1) {
                        ptx::tma_load_5d(st + a_bytes, &mAl, full + s, c0, cw, ch, ct, b0);
                        ptx::tma_load_3d(st + 2 * a_bytes + b_bytes, &mBl, full + s, c0, n0, tap);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer
        {
            const uint32_t idesc = ptx::make_idesc_f16(TILE_M, a.n_tile);
            const int ksteps = a.kc / 16;
            const uint64_t dproto = ptx::make_kmajor_desc(0, rb);
            const uint32_t dlo = (uint32_t)dproto, dhi = (uint32_t)(dproto >> 32);
            int s = 0, ai = 0, asm_ = 0;
            const int nsmall = a.nacc - a.nmain;
            uint32_t ph = 0;
            for (int n = 0; n < n_total; ++n) {
                ptx::mbar_wait(full + s, ph);
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                // descriptors differ only in their 14-bit start-address field: one 32-bit add per operand
                // (the single issuing thread must stay well ahead of the tensor pipe)
                const uint32_t lah = dlo + (sa >> 4), lal = lah + (a_bytes >> 4);
                const uint32_t lbh = lah + ((2 * a_bytes) >> 4), lbl = lbh + (b_bytes >> 4);
                // round-robin over `nacc` TMEM accumulators: the tensor core's fp32 accumulate truncates, so
                // the chain of dependent adds per accumulator is cut nacc-fold and the partial sums are
                // combined with round-to-nearest fp32 adds in the epilogue
                // hi*hi goes to the "main" accumulators, the 2^-11-sized cross terms hi*lo, lo*hi to their own: a
                // truncating add costs an error relative to the accumulator it lands in, so the small terms no longer
                // spend the main accumulator's precision (and the main chain is 3x shorter)
                const uint32_t tacc = tmem_base + (uint32_t)(ai * a.n_tile);
                const uint32_t tsm = tmem_base + (uint32_t)((a.nmain + asm_) * a.n_tile);
                uint32_t acc_flag = n >= a.nmain ? 1u : 0u;
                uint32_t sm_flag = n >= nsmall ? 1u : 0u;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < ksteps) {
                            const uint32_t o = (uint32_t)k * 2u;          // 16 fp16 = 32 B along K, in 16-byte units
                            ptx::mma_f16_ss(tacc, ptx::desc64(lah + o, dhi), ptx::desc64(lbh + o, dhi), idesc, acc_flag);
                            acc_flag = 1u;
                            if (a.terms > 1) {
                                ptx::mma_f16_ss(tsm, ptx::desc64(lah + o, dhi), ptx::desc64(lbl + o, dhi), idesc, sm_flag);
                                ptx::mma_f16_ss(tsm, ptx::desc64(lal + o, dhi), ptx::desc64(lbh + o, dhi), idesc, 1u);
                                sm_flag = 1u;
                            }
                        }
                    }
                    ptx::mma_commit(empty + s);        // frees the smem stage once these MMAs have read it
                    if (n == n_total - 1) ptx::mma_commit(tmem_full);   // accumulator complete
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
                if (++ai == a.nmain) ai = 0;
                if (nsmall > 0 && ++asm_ == nsmall) asm_ = 0;
            }
        }
    } else {
        // ================================ epilogue (8 warps, two per TMEM lane quarter)
        ptx::mbar_wait_backoff(tmem_full, 0);
        ptx::tc_fence_after();
        const int q = warp & 3, half = (warp - 2) >> 2;       // two warps per TMEM lane quarter: column halves
        const int m = q * 32 + lane;
        int r = m;
        const int wi = r % a.bw; r /= a.bw;
        const int hi = r % a.bh; r /= a.bh;
        const int ti = r % a.bt;
        const int bi = r / a.bt;
        // box dims grow only once the lower ones span the tensor (launch_conv_tc), so the 128 rows are consecutive voxels;
        // tiles of single planes (phase form, T = 2 tiling) keep every 32-row warp slice inside one plane when bw * bh % 32 == 0
        const int t_out = a.t_phase ? 2 * (t0 + ti) + ph_out : t0 + ti;
        const int contig = (!a.t_phase && (a.bt == a.T || a.bb == 1)) || (a.bw * a.bh) % 32 == 0 ? 1 : 0;
        EpiArgs e{a.bias, a.res, a.y, a.T, a.H, a.W, a.Cout, a.res_ut, a.res_uh, a.res_uw, a.act, a.out_mode, contig};
        const int nacc_used = a.nacc;       // host guarantees every accumulator is written by every CTA
        // column split between the two warps of a quarter (multiples of 16)
        const int nh0 = ((a.n_tile / 16 + 1) / 2) * 16;
        const int col0 = half == 0 ? 0 : nh0, ncols = half == 0 ? nh0 : a.n_tile - nh0;
        const size_t stile_bytes = (size_t)32 * (nh0 + 4) * sizeof(float);
        // coalesced path needs every row of the tile to be a real voxel (no batch-dimension overhang)
        if (a.out_mode == 0 && b0 + a.bb <= a.B && 8 * stile_bytes <= (size_t)a.stages * stage_bytes) {
            float* stile = reinterpret_cast<float*>(smem + (size_t)(warp - 2) * stile_bytes);
            const int Tr = a.T / a.res_ut, Hr = a.H / a.res_uh, Wr = a.W / a.res_uw;
            const long long vox_lane = (((long long)(b0 + bi) * a.T + t_out) * a.H + h0 + hi) * a.W + w0 + wi;
            const long long roff_lane =
                ((((long long)(b0 + bi) * Tr + t_out / a.res_ut) * Hr + (h0 + hi) / a.res_uh) * Wr + (w0 + wi) / a.res_uw) * a.Cout;
            if (ncols > 0) {
                float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
                const bool want_stats = a.stats != nullptr;     // host guarantees bb == 1 (one sample per tile) and ncols <= 128
                epilogue_warp_coalesced(e, stile, tmem_base + ((uint32_t)(q * 32) << 16), col0, ncols, nacc_used, a.n_tile, n0,
                                        __ldg(a.scale_ptr), lane, vox_lane, roff_lane, want_stats, ssum, ssq);
                if (want_stats) flush_stats(a.stats + (size_t)b0 * a.Cout * 2, a.Cout, n0, col0, ncols, lane, ssum, ssq);
            }
        } else if (half == 0) {
            epilogue_row(e, tmem_base + ((uint32_t)(q * 32) << 16), a.n_tile, nacc_used, a.n_tile, n0, __ldg(a.scale_ptr), b0 + bi,
                         t_out, h0 + hi, w0 + wi, b0 + bi < a.B);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------------
// v2: 256-row tile + H-halo.  The v1 kernel re-loads the activation box once per tap and is bound by
// L2->SM bandwidth (ncu: 9.4 TB/s of xbar reads, tensor pipe 35 % on g_3.conv_0).  Here a CTA owns a
// (bw x bh2) patch of ONE (b, t) plane = 2 sub-tiles of 128 voxels, and per (kt, kw, channel-chunk) stage
// loads the patch ONCE with a 1-row halo above and below: box (kc, bw, bh2+2).  Because every row of the
// box is a contiguous run of bw voxels, the A operand of sub-tile j under tap kh is the 128-row window
// starting at box row (j*bh2/2 + kh)*bw -- a multiple of 8 rows, i.e. a 1024-byte aligned K-major SW64/
// SW128 descriptor.  One A load therefore serves 3 kh taps x 2 sub-tiles, and one weight load (the 3 kh taps
// of (kt, kw), fetched as a single 5-D box) serves both sub-tiles: half the L2 traffic per MMA of v1.
struct ConvTcHArgs {
    const float* bias; const float* res; const float* scale_ptr; float* y; double* stats;
    int B, T, H, W, Cin, Cout;
    int kt, kw;                   // kh == 3
    int t_phase;                  // 1: input is the temporally x2 nearest-upsampled tensor stored at T/2 (see below)
    int wstack;                   // 1: the 3 kw taps are stacked along N (narrow layers; see wstack_gather)
    int bw, bh2;                  // patch: bw x bh2 voxels (= 256)
    int tiles_w, tiles_h;
    int n_tile, kc, stages, terms, nacc;   // nacc accumulators per (sub-)tile ...
    int nmain;                             // ... of which the first nmain take hi*hi, the rest the small cross terms
    int res_ut, res_uh, res_uw, act, out_mode;
    int cc_lo, cc_hi;             // channel-chunk range [cc_lo, cc_hi) of this launch (K split across launches)
    int flags;                    // bit 0: epilogue warps prefetch their residual rows into L2 while the main loop runs
    // Side input (K extension): `cc2` extra channel chunks of a SECOND activation tensor (same voxel grid) enter through
    // the centre tap only -- a fused 1x1x1 convolution.  This is the GeneratorBlock's learned shortcut
    // conv_s(Norm3D(x)) (decoder.py:44-50) in the blocks that do not upsample: out = conv_1(a1) + conv_s(x_n) is ONE
    // implicit GEMM over K = 27 C_mid + C_in, so the separate shortcut launch, its output tensor and the residual read
    // of this epilogue disappear.  Maps mA2*/mB2* ; weights [kw=3][cout_pad][Cin2] with only the kw = 1 slab non-zero.
    int cc2;
    // Persistent tiles: a CTA walks output tiles blockIdx.x, blockIdx.x + gridDim.x, ... of its N tile.  Barriers, the
    // TMEM allocation and the smem ring live across tiles, and while the epilogue of tile i drains TMEM the producer
    // already loads the first `prefetch_stages` stages of tile i+1.  The 8 transpose tiles of the epilogue sit
    // `stile_per_slot` to a ring buffer in the buffers that tile i filled LAST (0: this launch does not use them).
    int n_tiles, prefetch_stages, stile_per_slot;
};

// Tile coordinates + the temporal tap range of one output tile of the halo kernel.
struct HaloTile { int t, b, w0, h0, ct_base, wt_base, dt_lo, dt_hi, n_main, n_total; };
template <class Args>
__device__ __forceinline__ HaloTile halo_tile(const Args& a, int tile, int cchunks) {
    HaloTile c;
    const int tw = tile % a.tiles_w; tile /= a.tiles_w;
    const int th = tile % a.tiles_h; tile /= a.tiles_h;
    c.t = tile % a.T; c.b = tile / a.T;
    c.w0 = tw * a.bw; c.h0 = th * a.bh2;
    // Temporal taps that fall outside the clip contribute only zero padding: both pipeline ends skip them.
    // t_phase (conv_0 of a block whose input was nearest-upsampled x2 in time, decoder.py:102-111): the input
    // frames 2j and 2j+1 are identical (SPADE's gamma/beta do not depend on t), so
    //   out[2j]   = W0 a[j-1] + (W1+W2) a[j]        out[2j+1] = (W0+W1) a[j] + W2 a[j+1]
    // i.e. TWO temporal taps on the T/2 tensor with per-phase pre-summed weights (loader.py) instead of three
    // on the upsampled one: 2/3 of the MMAs and of the operand traffic, and the upsampled tensor never exists.
    // Source plane of temporal tap dt is ct_base + dt, its weight slab wt_base + dt; the valid dt form a range.
    const int Tin = a.t_phase ? a.T / 2 : a.T;
    c.ct_base = a.t_phase ? (c.t >> 1) - 1 + (c.t & 1) : c.t - a.kt / 2;
    c.wt_base = a.t_phase ? (c.t & 1) * 2 : 0;
    c.dt_lo = c.ct_base < 0 ? -c.ct_base : 0;
    c.dt_hi = (Tin - 1 - c.ct_base) < (a.kt - 1) ? (Tin - 1 - c.ct_base) : (a.kt - 1);
    const int kw_iter = a.wstack ? 1 : a.kw;                      // stacked: one unshifted load covers all kw taps
    c.n_main = (c.dt_hi - c.dt_lo + 1) * kw_iter * cchunks;       // pipeline stages of the 3x3(x3) taps
    c.n_total = c.n_main + a.cc2;                                 // ... plus the side-input chunks (centre tap only)
    return c;
}

// 10 warps = 3 on one scheduler: 16384 / 3 / 32 -> 168 registers per thread is the cap (more needs an 8-warp layout)
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                    const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
                    const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
                    const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, const ConvTcHArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t rb = (uint32_t)a.kc * 2;
    const uint32_t a_rows = (uint32_t)(a.bw * (a.bh2 + 2));
    const uint32_t a_bytes = (a_rows * rb + 1023u) & ~1023u;
    const int nw = a.wstack ? 3 : 1;                                // kw taps per weight row block
    const int accw = nw * a.n_tile;                                 // TMEM columns of one accumulator = the MMA's N
    const uint32_t b_tap = (uint32_t)(nw * a.n_tile) * rb;          // one kh tap; multiple of 1024 (host-checked)
    const uint32_t b_bytes = 3 * b_tap;
    const uint32_t mult = a.terms > 1 ? 2u : 1u;                    // lo words only exist in the 3-term mode
    const uint32_t off_alo = a_bytes, off_bhi = mult * a_bytes, off_blo = mult * a_bytes + b_bytes;
    const uint32_t stage_bytes = mult * (a_bytes + b_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
    uint64_t* empty = full + a.stages;
    uint64_t* tmem_full = empty + a.stages;
    uint64_t* epi_done = tmem_full + 1;       // 8 arrivals per tile: TMEM and the transpose tiles are free again
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) dbg_stamp(0, blockIdx.x);           // CTA start
    const int n0 = blockIdx.y * a.n_tile;
    const int cchunks = a.cc_hi - a.cc_lo;
    const int bh_sub = a.bh2 / 2;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(2 * accw * a.nacc)) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mAh); ptx::prefetch_tensormap(&mBh);
        if (a.terms > 1) { ptx::prefetch_tensormap(&mAl); ptx::prefetch_tensormap(&mBl); }
        if (a.cc2 > 0) {
            ptx::prefetch_tensormap(&mA2h); ptx::prefetch_tensormap(&mB2h);
            if (a.terms > 1) { ptx::prefetch_tensormap(&mA2l); ptx::prefetch_tensormap(&mB2l); }
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < a.stages; ++s) { ptx::mbar_init(full + s, 1); ptx::mbar_init(empty + s, 1); }
            ptx::mbar_init(tmem_full, 1);
            ptx::mbar_init(epi_done, 8);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, ncols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (see conv_tc_kernel)
    // Only now may the successor be scheduled: this CTA already owns its TMEM columns, so a co-resident CTA of the
    // next kernel can never take them first and then sit in its own pdl_wait() while this one starves.
    pdl_launch_dependents();
    pdl_wait();      // prologue above touched no global memory: it overlapped the previous kernel's tail
    if (threadIdx.x == 0) dbg_stamp(1, blockIdx.x);           // prologue done
    const int kw_iter = a.wstack ? 1 : a.kw;
    const int tile0 = blockIdx.x, tstep = gridDim.x;

    if (warp == 0) {
        const uint32_t tx = (a.terms > 1 ? 2u : 1u) * (a_rows * rb + b_bytes);
        // side input: same box (the halo rows ride along unused), one weight slab [nw][n_tile][kc] per chunk
        const uint32_t tx2 = (a.terms > 1 ? 2u : 1u) * (a_rows * rb + b_tap);
        int s = 0;
        uint32_t ph = 0, epi_ph = 0;
        for (int tile = tile0; tile < a.n_tiles; tile += tstep) {
            const HaloTile c = halo_tile(a, tile, cchunks);
            if (lane == 0 && tile != tile0) { dbg_stamp(0, tile); dbg_stamp(1, tile); }    // the producer turns to this tile
            // The first `prefetch_stages` stages of a tile go to ring buffers the previous tile's epilogue does not use:
            // they are loaded while that epilogue still drains TMEM.  The stage after them (or the end of a short tile)
            // waits until the transpose tiles have been read back (epi_done).
            bool epi_waited = tile == tile0;
            int n = 0;
            for (int dt = c.dt_lo; dt <= c.dt_hi; ++dt)
            for (int dw = 0; dw < kw_iter; ++dw)
            for (int cc = a.cc_lo; cc < a.cc_hi; ++cc, ++n) {
                const int ct = c.ct_base + dt, wt = c.wt_base + dt, c0 = cc * a.kc;
                if (!epi_waited && n >= a.prefetch_stages) { ptx::mbar_wait(epi_done, epi_ph); epi_ph ^= 1u; epi_waited = true; }
                ptx::mbar_wait(empty + s, ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const int cw = a.wstack ? c.w0 : c.w0 + dw - a.kw / 2, ch = c.h0 - 1;
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(full + s, tx);
                    ptx::tma_load_5d(st, &mAh, full + s, c0, cw, ch, ct, c.b);
                    ptx::tma_load_5d(st + off_bhi, &mBh, full + s, c0, n0, dw, 0, wt);
                    if (a.terms > 1) {
                        ptx::tma_load_5d(st + off_alo, &mAl, full + s, c0, cw, ch, ct, c.b);
                        ptx::tma_load_5d(st + off_blo, &mBl, full + s, c0, n0, dw, 0, wt);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            for (int cc = 0; cc < a.cc2; ++cc, ++n) {
                const int c0 = cc * a.kc;
                if (!epi_waited && n >= a.prefetch_stages) { ptx::mbar_wait(epi_done, epi_ph); epi_ph ^= 1u; epi_waited = true; }
                ptx::mbar_wait(empty + s, ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(full + s, tx2);
                    ptx::tma_load_5d(st, &mA2h, full + s, c0, c.w0, c.h0 - 1, c.t, c.b);
                    ptx::tma_load_5d(st + off_bhi, &mB2h, full + s, c0, n0, a.wstack ? 0 : 1, 0, 0);
                    if (a.terms > 1) {
                        ptx::tma_load_5d(st + off_alo, &mA2l, full + s, c0, c.w0, c.h0 - 1, c.t, c.b);
                        ptx::tma_load_5d(st + off_blo, &mB2l, full + s, c0, n0, a.wstack ? 0 : 1, 0, 0);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            if (!epi_waited) { ptx::mbar_wait(epi_done, epi_ph); epi_ph ^= 1u; }   // every phase is observed exactly once
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::make_idesc_f16(TILE_M, accw);
        const int ksteps = a.kc / 16;
        const uint64_t dproto = ptx::make_kmajor_desc(0, rb);
        const uint32_t dlo = (uint32_t)dproto, dhi = (uint32_t)(dproto >> 32);
        const uint32_t sub_step = (uint32_t)(bh_sub * a.bw) * rb >> 4, kh_step = (uint32_t)a.bw * rb >> 4;
        const int nsmall = a.nacc - a.nmain;
        int s = 0;
        uint32_t ph = 0, epi_ph = 0;
        for (int tile = tile0; tile < a.n_tiles; tile += tstep) {
            const HaloTile c = halo_tile(a, tile, cchunks);
            if (tile != tile0) {
                // the accumulators are overwritten from the first MMA on: the previous tile must be out of TMEM
                ptx::mbar_wait(epi_done, epi_ph);
                epi_ph ^= 1u;
                ptx::tc_fence_after();
            }
            int ai = 0, asm_ = 0;
            for (int n = 0; n < c.n_total; ++n) {
                ptx::mbar_wait(full + s, ph);
                if (n == 0 && lane == 0) dbg_stamp(2, tile);  // first stage landed
                ptx::tc_fence_after();
                const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t lah = dlo + (sa >> 4), lal = lah + (off_alo >> 4);
                const uint32_t lbh = lah + (off_bhi >> 4), lbl = lah + (off_blo >> 4);
                const uint32_t fresh = n < a.nmain ? 0u : 1u;    // first visit of this main accumulator pair -> overwrite
                uint32_t sm_flag = n < nsmall ? 0u : 1u;         // same for the cross-term accumulators
                // issue order (kh, k, term, sub): back-to-back MMAs target different TMEM accumulators, so a short
                // (N = 64) MMA never waits on the one before it
                const uint32_t tacc0 = tmem_base + (uint32_t)(ai * accw), tacc1 = tacc0 + (uint32_t)(a.nacc * accw);
                // cross terms hi*lo, lo*hi accumulate apart from hi*hi (see conv_tc_kernel); with a single accumulator
                // per sub-tile (stacked N = 192 fills TMEM) they follow hi*hi into the same one (short chains only)
                const uint32_t tsm0 = nsmall > 0 ? tmem_base + (uint32_t)((a.nmain + asm_) * accw) : tacc0;
                const uint32_t tsm1 = tsm0 + (uint32_t)(a.nacc * accw);
                uint32_t acc_flag = fresh;
                const bool ext = n >= c.n_main;                  // side-input stage: centre row only, weight slab 0
                if (ptx::elect_one()) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    if (ext && kh != 1) continue;
                    const uint32_t ao = (uint32_t)kh * kh_step, bo = ext ? 0u : (uint32_t)kh * (b_tap >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < ksteps) {
                            const uint32_t o = (uint32_t)k * 2u;
                            const uint64_t dAh0 = ptx::desc64(lah + ao + o, dhi), dAh1 = ptx::desc64(lah + ao + sub_step + o, dhi);
                            const uint64_t dBh = ptx::desc64(lbh + bo + o, dhi);
                            ptx::mma_f16_ss(tacc0, dAh0, dBh, idesc, acc_flag);
                            ptx::mma_f16_ss(tacc1, dAh1, dBh, idesc, acc_flag);
                            acc_flag = 1u;
                            if (a.terms > 1) {
                                const uint64_t dBl = ptx::desc64(lbl + bo + o, dhi);
                                ptx::mma_f16_ss(tsm0, dAh0, dBl, idesc, sm_flag);
                                ptx::mma_f16_ss(tsm1, dAh1, dBl, idesc, sm_flag);
                                sm_flag = 1u;
                                ptx::mma_f16_ss(tsm0, ptx::desc64(lal + ao + o, dhi), dBh, idesc, 1u);
                                ptx::mma_f16_ss(tsm1, ptx::desc64(lal + ao + sub_step + o, dhi), dBh, idesc, 1u);
                            }
                        }
                    }
                }
                ptx::mma_commit(empty + s);
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
                if (++ai == a.nmain) ai = 0;
                if (nsmall > 0 && ++asm_ == nsmall) asm_ = 0;
            }
            if (ptx::elect_one()) ptx::mma_commit(tmem_full);
            __syncwarp();
            if (lane == 0) dbg_stamp(3, tile);                // last MMA issued
        }
    } else {
        const int nacc_used = a.nacc;       // host guarantees every accumulator is written by every tile
        const int q = warp & 3, half = (warp - 2) >> 2;       // two warps per TMEM lane quarter: column halves
        const int m = q * 32 + lane;
        const int wi = m % a.bw, hi = m / a.bw;
        const int nh0 = ((a.n_tile / 16 + 1) / 2) * 16;
        const int col0 = half == 0 ? 0 : nh0, ncols = half == 0 ? nh0 : a.n_tile - nh0;
        const size_t stile_bytes = (size_t)32 * (nh0 + 4) * sizeof(float);
        const int Tr = a.T / a.res_ut, Hr = a.H / a.res_uh, Wr = a.W / a.res_uw;
        const bool use_stile = a.stile_per_slot > 0;          // coalesced channels-last write-out through transpose tiles
        // edge-row exchange of the stacked form: its own region behind the barriers (the ring belongs to the producer)
        float* const xbuf = reinterpret_cast<float*>(smem + (size_t)a.stages * stage_bytes + 256);   // [slot][quarter][d0|d2][n_tile]
        EpiArgs e{a.bias, a.res, a.y, a.T, a.H, a.W, a.Cout, a.res_ut, a.res_uh, a.res_uw, a.act, a.out_mode, a.bw == a.W ? 1 : 0};
        uint32_t tf_ph = 0;
        int ring = 0;                                         // ring position behind the last stage of the current tile
        float scale = 0.f;
        for (int tile = tile0; tile < a.n_tiles; tile += tstep) {
            const HaloTile c = halo_tile(a, tile, cchunks);
            const int b = c.b, t = c.t, w0 = c.w0, h0 = c.h0;
            ring = (ring + c.n_total) % a.stages;
            if (a.res != nullptr && (a.flags & 1)) {
                // The epilogue warps idle through the main loop: pull the residual rows they will add (the shortcut through
                // its upsample map, or the previous K-split partial sum) from HBM into L2 now, so that the latency-bound
                // residual reads of the epilogue hit L2.
                for (int sub = 0; sub < 2; ++sub) {
                    const int ww = w0 + wi, hh = h0 + sub * bh_sub + hi;
                    const long long roff = ((((long long)b * Tr + t / a.res_ut) * Hr + hh / a.res_uh) * Wr + ww / a.res_uw) * a.Cout;
                    int c_lo = n0 + col0, c_hi = n0 + col0 + ncols;
                    if (c_hi > a.Cout) c_hi = a.Cout;
                    const char* p = reinterpret_cast<const char*>(a.res + roff + c_lo);
                    for (int off = 0; off < (c_hi - c_lo) * 4; off += 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
                }
            }
            if (tile == tile0) scale = __ldg(a.scale_ptr);
            ptx::mbar_wait_backoff(tmem_full, tf_ph);
            tf_ph ^= 1u;
            if (threadIdx.x == 64) dbg_stamp(4, tile);        // accumulators complete
            ptx::tc_fence_after();
            // Transpose tiles live in the ring buffers this tile filled LAST (every TMA landed, every MMA retired): the
            // producer refills those last, so the first stages of the next tile stream in underneath this epilogue.
            float* stile = nullptr;
            if (use_stile) {
                const int idx = warp - 2;
                int slot = ring - 1 - idx / a.stile_per_slot;
                while (slot < 0) slot += a.stages;
                stile = reinterpret_cast<float*>(smem + (size_t)slot * stage_bytes + (size_t)(idx % a.stile_per_slot) * stile_bytes);
            }
            if (a.wstack) {
                // stacked kw taps: shifted sum of the three column groups (see wstack_gather)
                const bool coal = a.out_mode == 0;
                const bool left_ok = wi > 0, right_ok = wi < a.bw - 1;
                const bool fix0 = q > 0 && ((q * 32) % a.bw) > 0, fix31 = q < 3 && ((q * 32 + 31) % a.bw) < a.bw - 1;
                float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
                const bool want_stats = a.stats != nullptr;
                int slot = 0;
                for (int sub = 0; sub < 2; ++sub) {
                    const int ww = w0 + wi, hh = h0 + sub * bh_sub + hi;
                    const long long vox_lane = (((long long)b * a.T + t) * a.H + hh) * a.W + ww;
                    const long long roff_lane = ((((long long)b * Tr + t / a.res_ut) * Hr + hh / a.res_uh) * Wr + ww / a.res_uw) * a.Cout;
                    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * a.nacc * accw);
                    if (coal) {
                        float* xs = xbuf + (size_t)slot * 8 * a.n_tile;
                        if (ncols > 0)
                            wstack_gather(stile, tbase, col0, ncols, a.n_tile, nacc_used, accw, scale, lane, left_ok, right_ok,
                                          xs + (size_t)(q * 2) * a.n_tile, xs + (size_t)(q * 2 + 1) * a.n_tile);
                        epi_bar();
                        if (ncols > 0) {
                            const int ld = ncols + 4;
                            for (int cx = lane; cx < ncols; cx += 32) {
                                if (fix0) stile[cx] += scale * xs[(size_t)((q - 1) * 2) * a.n_tile + col0 + cx];
                                if (fix31) stile[31 * ld + cx] += scale * xs[(size_t)((q + 1) * 2 + 1) * a.n_tile + col0 + cx];
                            }
                            __syncwarp();
                            if (threadIdx.x == 64) dbg_stamp(8 + 2 * sub, tile);
                            epilogue_writeout(e, stile, col0, ncols, n0, lane, vox_lane, roff_lane, want_stats, ssum, ssq);
                            if (threadIdx.x == 64) dbg_stamp(9 + 2 * sub, tile);
                        }
                        slot ^= 1;
                    } else {
                        // one output row per thread (frame layout of conv_img): half-0 warps work, all 8 keep the barrier count
                        for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
                            float* xs = xbuf + (size_t)slot * 8 * a.n_tile;
                            float v[16];
                            if (half == 0) {
                                float d0[16], d1[16], d2[16];
                                wstack_chunk16_x2(tbase, a.n_tile, nacc_used, accw, c0, d0, d1, d2);
                                wstack_combine(d0, d1, d2, lane, left_ok, right_ok, v);
                                if (lane == 31) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) xs[(size_t)(q * 2) * a.n_tile + c0 + j] = d0[j];
                                }
                                if (lane == 0) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) xs[(size_t)(q * 2 + 1) * a.n_tile + c0 + j] = d2[j];
                                }
                            }
                            epi_bar();
                            if (half == 0) {
                                if (lane == 0 && fix0) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] += xs[(size_t)((q - 1) * 2) * a.n_tile + c0 + j];
                                }
                                if (lane == 31 && fix31) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[j] += xs[(size_t)((q + 1) * 2 + 1) * a.n_tile + c0 + j];
                                }
                                if (n0 + c0 < a.Cout) row_finish(e, v, n0 + c0, scale, vox_lane, roff_lane, b, t, hh, ww);
                            }
                            slot ^= 1;
                        }
                    }
                }
                if (coal && want_stats && ncols > 0) flush_stats(a.stats + (size_t)b * a.Cout * 2, a.Cout, n0, col0, ncols, lane, ssum, ssq);
            } else if (use_stile) {
                float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
                const bool want_stats = a.stats != nullptr;
                for (int sub = 0; sub < 2; ++sub) {
                    const int ww = w0 + wi, hh = h0 + sub * bh_sub + hi;
                    const long long vox_lane = (((long long)b * a.T + t) * a.H + hh) * a.W + ww;
                    const long long roff_lane = ((((long long)b * Tr + t / a.res_ut) * Hr + hh / a.res_uh) * Wr + ww / a.res_uw) * a.Cout;
                    if (ncols > 0)
                        epilogue_warp_coalesced(e, stile, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * a.nacc * a.n_tile), col0,
                                                ncols, nacc_used, a.n_tile, n0, scale, lane, vox_lane, roff_lane, want_stats, ssum, ssq,
                                                8 + 2 * sub, tile);
                }
                if (want_stats && ncols > 0) flush_stats(a.stats + (size_t)b * a.Cout * 2, a.Cout, n0, col0, ncols, lane, ssum, ssq);
                if (threadIdx.x == 64) dbg_stamp(12, tile);
            } else if (half == 0) {
                for (int sub = 0; sub < 2; ++sub)
                    epilogue_row(e, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * a.nacc * a.n_tile), a.n_tile, nacc_used,
                                 a.n_tile, n0, scale, b, t, h0 + sub * bh_sub + hi, w0 + wi, true);
            }
            if (threadIdx.x == 64) dbg_stamp(5, tile);        // epilogue stores issued
            // this warp is done with TMEM and with its transpose tile: hand both back (TMA writes the tile's bytes next)
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(epi_done);
            if (threadIdx.x == 64) dbg_stamp(6, tile);        // tile end
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------------
// v3: CTA-pair tiles.  The halo kernel above fills TMEM with ONE tile (2 sub-tiles x (main + cross-term) x 128
// columns), so its epilogue (6.6-12 us of a 14-79 us tile) cannot run under the next tile's MMAs, and its 90 KB
// stages leave room for a 2-deep ring only.  Here the two 128-voxel sub-tiles of the same 256-voxel patch go to the
// two CTAs of a cluster on one TPC, and every MMA is ONE tcgen05.mma.cta_group::2 of M = 256 issued by the leader:
//   * each CTA stages its own sub-tile with its 1-row halo (box bw x (bh_sub + 2)) and HALF of the weight tile (n_tile / 2
//     output rows x 3 kh taps): 44-57 KB per stage instead of 90 -> 4-5 ring buffers in the same shared memory, and the
//     weight tile crosses L2 -> SM once per PAIR;
//   * a CTA's TMEM holds 128 rows x (main + cross) x n_tile columns = half of it: TWO accumulator sets, so the epilogue of
//     tile i drains set i & 1 while the MMAs of tile i + 1 fill the other one (tmem_full / tmem_empty barriers per set);
//   * the epilogue therefore no longer borrows the ring buffers: every thread finishes its own output row straight from
//     registers (128 contiguous bytes per 32-column chunk, 16-byte accesses), and the per-(sample, channel) statistics
//     come out of a 31-shuffle transpose-reduce of each chunk;
//   * stage release (empty) and accumulator-ready (tmem_full) signals are multicast commits onto the same-offset barriers
//     of both CTAs; both CTAs' TMA loads report their bytes to the LEADER's full barrier; the 8 + 8 epilogue warps of the
//     pair arrive on the leader's tmem_empty barrier (remote mbarrier.arrive for the peer).
struct ConvTcPArgs {
    const float* bias; const float* res; const float* scale_ptr; float* y; double* stats;
    int B, T, H, W, Cin, Cout;
    int kt, kw;                   // kh == 3
    int t_phase;
    int wstack;                   // 1: the 3 kw taps are stacked along N (narrow layers; see wstack_gather above)
    int nacc;                     // accumulators per set: 2 = main + cross terms, 1 = shared (terms == 1, or stacked N = 192)
    int out_mode;                 // 1: (B,T,C,H,W) frame layout (conv_img), stacked form only
    int bw, bh2;                  // the PAIR's patch: bw x bh2 voxels (= 256); each CTA owns bh2 / 2 of its rows
    int tiles_w, tiles_h;
    int n_tile, kc, stages, terms;
    int res_ut, res_uh, res_uw, act;
    int cc_lo, cc_hi;
    int cc2;
    int n_tiles;
};

// sum over the warp's 32 rows of each of 32 columns: v[j] of lane r = element (row r, column j); returns to lane c the
// sum of column c.  Each step halves the columns a lane still carries: 16 + 8 + 4 + 2 + 1 shuffles.
__device__ __forceinline__ float column_sums32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = up ? v[j] : v[j + n];
            const float keep = up ? v[j + n] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// 16-column variant: lanes L and L ^ 16 end up with the sum of column L & 15
__device__ __forceinline__ float column_sums16(float (&v)[16], int lane) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
    for (int off = 8, n = 8; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = up ? v[j] : v[j + n];
            const float keep = up ? v[j + n] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                    const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
                    const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
                    const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, const ConvTcPArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t rb = (uint32_t)a.kc * 2;
    const int bh_sub = a.bh2 / 2;
    const uint32_t a_rows = (uint32_t)(a.bw * (bh_sub + 2));
    const uint32_t a_bytes = (a_rows * rb + 1023u) & ~1023u;
    const int nh = a.n_tile / 2;                                    // output channels this CTA stages weights for (its half of N)
    const int nw = a.wstack ? 3 : 1;                                // kw taps per weight row block
    // Stacked form: the MMA's N = 3 n_tile columns are ordered [half][kw][nh], so that each CTA of the pair stages the three
    // kw slabs of ITS nh output channels (one TMA box) and its accumulator columns [kw][nh] sit side by side.
    const uint32_t b_tap = (uint32_t)(nw * nh) * rb;                // one kh tap; multiple of 1024 (host-checked)
    const uint32_t b_bytes = 3 * b_tap;
    const uint32_t mult = a.terms > 1 ? 2u : 1u;
    const uint32_t off_alo = a_bytes, off_bhi = mult * a_bytes, off_blo = mult * a_bytes + b_bytes;
    const uint32_t stage_bytes = (mult * (a_bytes + b_bytes) + 1023u) & ~1023u;   // as pair_config sizes it
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);   // used in the leader only
    uint64_t* empty = full + a.stages;
    uint64_t* tmem_full = empty + a.stages;       // [2] one per accumulator set
    uint64_t* tmem_empty = tmem_full + 2;         // [2] leader only: 16 arrivals (8 epilogue warps of each CTA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* const xbuf = reinterpret_cast<float*>(tmem_slot + 4);    // stacked epilogue: [2 slots][4 quarters][2 halves][d0 | d2][16]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int n0 = blockIdx.y * a.n_tile;
    const int cchunks = a.cc_hi - a.cc_lo;
    const int accw = nw * a.n_tile;                                 // TMEM columns of one accumulator = the MMA's N
    const int set_cols = a.nacc * accw;                             // TMEM columns of one accumulator set
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(2 * set_cols)) ncols <<= 1;
    if (threadIdx.x == 0 && leader) dbg_stamp(0, blockIdx.x >> 1);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mAh); ptx::prefetch_tensormap(&mBh);
        if (a.terms > 1) { ptx::prefetch_tensormap(&mAl); ptx::prefetch_tensormap(&mBl); }
        if (a.cc2 > 0) {
            ptx::prefetch_tensormap(&mA2h); ptx::prefetch_tensormap(&mB2h);
            if (a.terms > 1) { ptx::prefetch_tensormap(&mA2l); ptx::prefetch_tensormap(&mB2l); }
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < a.stages; ++s) { ptx::mbar_init(full + s, 1); ptx::mbar_init(empty + s, 1); }
            ptx::mbar_init(tmem_full, 1); ptx::mbar_init(tmem_full + 1, 1);
            ptx::mbar_init(tmem_empty, 16); ptx::mbar_init(tmem_empty + 1, 16);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc_2cta(tmem_slot, ncols);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();          // both CTAs' barriers exist before any remote arrive / TMA signal
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_launch_dependents();
    pdl_wait();      // nothing above touched global memory
    const int kw_iter = a.wstack ? 1 : a.kw;                        // stacked: one unshifted load covers all kw taps
    const int tile0 = blockIdx.x >> 1, tstep = gridDim.x >> 1;      // the pair walks tiles tile0, tile0 + tstep, ...

    if (warp == 0) {
        // ================================ TMA producer (both CTAs; bytes of both land on the leader's full barrier)
        const uint32_t full0 = ptx::mapa_u32(ptx::smem_u32(full), 0);
        const uint32_t tx = 2u * mult * (a_rows * rb + b_bytes);
        const uint32_t tx2 = 2u * mult * (a_rows * rb + b_tap);
        const int hsub = (int)rank * bh_sub, nsub = n0 + (int)rank * nh;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile0; tile < a.n_tiles; tile += tstep) {
            const HaloTile c = halo_tile(a, tile, cchunks);
            for (int dt = c.dt_lo; dt <= c.dt_hi; ++dt)
            for (int dw = 0; dw < kw_iter; ++dw)
            for (int cc = a.cc_lo; cc < a.cc_hi; ++cc) {
                const int ct = c.ct_base + dt, wt = c.wt_base + dt, c0 = cc * a.kc;
                ptx::mbar_wait(empty + s, ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const int cw = a.wstack ? c.w0 : c.w0 + dw - a.kw / 2, ch = c.h0 + hsub - 1;
                const uint32_t fb = full0 + 8u * (uint32_t)s;
                if (ptx::elect_one()) {
                    if (leader) ptx::mbar_expect_tx(full + s, tx);
                    ptx::tma_load_5d_2cta(st, &mAh, fb, c0, cw, ch, ct, c.b);
                    ptx::tma_load_5d_2cta(st + off_bhi, &mBh, fb, c0, nsub, dw, 0, wt);
                    if (a.terms > 1) {
                        ptx::tma_load_5d_2cta(st + off_alo, &mAl, fb, c0, cw, ch, ct, c.b);
                        ptx::tma_load_5d_2cta(st + off_blo, &mBl, fb, c0, nsub, dw, 0, wt);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            for (int cc = 0; cc < a.cc2; ++cc) {
                const int c0 = cc * a.kc;
                ptx::mbar_wait(empty + s, ph ^ 1u);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const uint32_t fb = full0 + 8u * (uint32_t)s;
                if (ptx::elect_one()) {
                    if (leader) ptx::mbar_expect_tx(full + s, tx2);
                    ptx::tma_load_5d_2cta(st, &mA2h, fb, c0, c.w0, c.h0 + hsub - 1, c.t, c.b);
                    ptx::tma_load_5d_2cta(st + off_bhi, &mB2h, fb, c0, nsub, a.wstack ? 0 : 1, 0, 0);
                    if (a.terms > 1) {
                        ptx::tma_load_5d_2cta(st + off_alo, &mA2l, fb, c0, c.w0, c.h0 + hsub - 1, c.t, c.b);
                        ptx::tma_load_5d_2cta(st + off_blo, &mB2l, fb, c0, nsub, a.wstack ? 0 : 1, 0, 0);
                    }
                }
                __syncwarp();
                if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ================================ MMA issuer (leader only): M = 256 across the pair
            const uint32_t idesc = ptx::make_idesc_f16(2 * TILE_M, accw);
            const int ksteps = a.kc / 16;
            const uint64_t dproto = ptx::make_kmajor_desc(0, rb);
            const uint32_t dlo = (uint32_t)dproto, dhi = (uint32_t)(dproto >> 32);
            const uint32_t kh_step = (uint32_t)a.bw * rb >> 4;
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int tile = tile0; tile < a.n_tiles; tile += tstep, ++it) {
                const HaloTile c = halo_tile(a, tile, cchunks);
                const int set = it & 1;
                if (it >= 2) {
                    // the epilogues of BOTH CTAs must have drained this set (tile it - 2) before it is overwritten
                    ptx::mbar_wait(tmem_empty + set, (uint32_t)(((it - 2) >> 1) & 1));
                    ptx::tc_fence_after();
                }
                // cross terms hi*lo, lo*hi accumulate apart from hi*hi unless the set holds ONE accumulator (stacked N = 192)
                const uint32_t tmain = tmem_base + (uint32_t)(set * set_cols), tcross = a.nacc > 1 ? tmain + (uint32_t)accw : tmain;
                uint32_t acc_flag = 0u, sm_flag = a.nacc > 1 ? 0u : 1u;      // first MMA into each accumulator of this tile overwrites
                for (int n = 0; n < c.n_total; ++n) {
                    ptx::mbar_wait(full + s, ph);
                    if (n == 0 && lane == 0) dbg_stamp(2, tile);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes);
                    const uint32_t lah = dlo + (sa >> 4), lal = lah + (off_alo >> 4);
                    const uint32_t lbh = lah + (off_bhi >> 4), lbl = lah + (off_blo >> 4);
                    const bool ext = n >= c.n_main;                  // side-input stage: centre row only, weight slab 0
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            if (ext && kh != 1) continue;
                            const uint32_t ao = (uint32_t)kh * kh_step, bo = ext ? 0u : (uint32_t)kh * (b_tap >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps) {
                                    const uint32_t o = (uint32_t)k * 2u;
                                    const uint64_t dAh = ptx::desc64(lah + ao + o, dhi), dBh = ptx::desc64(lbh + bo + o, dhi);
                                    ptx::mma_f16_ss_2cta(tmain, dAh, dBh, idesc, acc_flag);
                                    acc_flag = 1u;
                                    if (a.terms > 1) {
                                        ptx::mma_f16_ss_2cta(tcross, dAh, ptx::desc64(lbl + bo + o, dhi), idesc, sm_flag);
                                        sm_flag = 1u;
                                        ptx::mma_f16_ss_2cta(tcross, ptx::desc64(lal + ao + o, dhi), dBh, idesc, 1u);
                                    }
                                }
                            }
                        }
                        ptx::mma_commit_2cta_mc(empty + s, 3);       // frees this ring buffer in BOTH CTAs
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
                if (ptx::elect_one()) ptx::mma_commit_2cta_mc(tmem_full + set, 3);
                __syncwarp();
                if (lane == 0) dbg_stamp(3, tile);
            }
        }
    } else {
        // ================================ epilogue (8 warps per CTA, two per TMEM lane quarter), one output row per thread
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const int wi = m % a.bw, hi = m / a.bw;
        const int ncw = a.n_tile / 2, col0 = half * ncw;                // this warp's columns of the tile
        const int Tr = a.T / a.res_ut, Hr = a.H / a.res_uh, Wr = a.W / a.res_uw;
        const uint32_t te0 = ptx::mapa_u32(ptx::smem_u32(tmem_empty), 0);
        const bool want_stats = a.stats != nullptr;
        float scale = 0.f;
        int it = 0;
        for (int tile = tile0; tile < a.n_tiles; tile += tstep, ++it) {
            const HaloTile c = halo_tile(a, tile, cchunks);
            const int set = it & 1;
            const int hh = c.h0 + (int)rank * bh_sub + hi, ww = c.w0 + wi;
            const long long vox = (((long long)c.b * a.T + c.t) * a.H + hh) * a.W + ww;
            float* yrow = a.y + vox * a.Cout;
            const float* rrow = nullptr;
            if (a.res != nullptr)
                rrow = a.res + ((((long long)c.b * Tr + c.t / a.res_ut) * Hr + hh / a.res_uh) * Wr + ww / a.res_uw) * a.Cout;
            if (it == 0) scale = __ldg(a.scale_ptr);
            ptx::mbar_wait_backoff(tmem_full + set, (uint32_t)((it >> 1) & 1));
            if (threadIdx.x == 64 && leader) dbg_stamp(4, tile);
            ptx::tc_fence_after();
            if (a.wstack) {
                // ---- stacked kw taps: out[m] = D1[m] + D0[m-1] [w > 0] + D2[m+1] [w < W-1]  (see wstack_gather); rows m -+ 1 are
                // the neighbouring lanes, the first / last lane of a 32-row slice gets them from the neighbouring warp through
                // `xbuf` (two alternating slots, one named barrier of the 8 epilogue warps per 16-column chunk)
                const bool left_ok = wi > 0, right_ok = wi < a.bw - 1;
                const bool fix0 = q > 0 && ((q * 32) % a.bw) > 0, fix31 = q < 3 && ((q * 32 + 31) % a.bw) < a.bw - 1;
                const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * set_cols + half * 3 * nh);
                int slot = 0;
                for (int c0 = 0; c0 < nh; c0 += 16) {
                    const int cw16 = nh - c0 < 16 ? nh - c0 : 16;       // 8 when the tile is 16 channels wide (conv_img)
                    uint32_t r0[16], r1[16], r2[16];
                    float d0[16], d1[16], d2[16], v[16];
                    if (cw16 == 16) {
                        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)c0, r0);
                        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(nh + c0), r1);
                        ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(2 * nh + c0), r2);
                    } else {
                        ptx::tmem_ld_32x32b_x8(tb + (uint32_t)c0, r0);
                        ptx::tmem_ld_32x32b_x8(tb + (uint32_t)(nh + c0), r1);
                        ptx::tmem_ld_32x32b_x8(tb + (uint32_t)(2 * nh + c0), r2);
#pragma unroll
                        for (int j = 8; j < 16; ++j) { r0[j] = 0u; r1[j] = 0u; r2[j] = 0u; }
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { d0[j] = __uint_as_float(r0[j]); d1[j] = __uint_as_float(r1[j]); d2[j] = __uint_as_float(r2[j]); }
                    if (a.nacc > 1) {
                        if (cw16 == 16) {
                            ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(accw + c0), r0);
                            ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(accw + nh + c0), r1);
                            ptx::tmem_ld_32x32b_x16(tb + (uint32_t)(accw + 2 * nh + c0), r2);
                        } else {
                            ptx::tmem_ld_32x32b_x8(tb + (uint32_t)(accw + c0), r0);
                            ptx::tmem_ld_32x32b_x8(tb + (uint32_t)(accw + nh + c0), r1);
                            ptx::tmem_ld_32x32b_x8(tb + (uint32_t)(accw + 2 * nh + c0), r2);
                        }
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) { d0[j] += __uint_as_float(r0[j]); d1[j] += __uint_as_float(r1[j]); d2[j] += __uint_as_float(r2[j]); }
                    }
                    wstack_combine(d0, d1, d2, lane, left_ok, right_ok, v);
                    float* xs = xbuf + (size_t)slot * 256;                   // [quarter][half][d0 | d2][16]
                    if (lane == 31) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) xs[((q * 2 + half) * 2) * 16 + j] = d0[j];
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) xs[((q * 2 + half) * 2 + 1) * 16 + j] = d2[j];
                    }
                    epi_bar();
                    if (lane == 0 && fix0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += xs[(((q - 1) * 2 + half) * 2) * 16 + j];
                    }
                    if (lane == 31 && fix31) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += xs[(((q + 1) * 2 + half) * 2 + 1) * 16 + j];
                    }
                    slot ^= 1;
                    const int n = n0 + half * nh + c0;                        // first output channel of this chunk
                    if (a.out_mode == 0) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (4 * g < cw16 && n + 4 * g < a.Cout) {         // Cout % 4 == 0 (host-checked)
                                float o[4] = {v[4 * g] * scale, v[4 * g + 1] * scale, v[4 * g + 2] * scale, v[4 * g + 3] * scale};
                                if (a.bias != nullptr) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n) + g);
                                    o[0] += b4.x; o[1] += b4.y; o[2] += b4.z; o[3] += b4.w;
                                }
                                if (rrow != nullptr) {
                                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + n) + g);
                                    o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
                                }
                                const float4 o4 = make_float4(apply_act(o[0], a.act), apply_act(o[1], a.act), apply_act(o[2], a.act),
                                                              apply_act(o[3], a.act));
                                reinterpret_cast<float4*>(yrow + n)[g] = o4;
                                v[4 * g] = o4.x; v[4 * g + 1] = o4.y; v[4 * g + 2] = o4.z; v[4 * g + 3] = o4.w;
                            } else {
                                v[4 * g] = 0.f; v[4 * g + 1] = 0.f; v[4 * g + 2] = 0.f; v[4 * g + 3] = 0.f;
                            }
                        }
                        if (want_stats) {
                            float sq[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
                            const float s1 = column_sums16(v, lane), s2 = column_sums16(sq, lane);
                            if (lane < 16 && lane < cw16 && n + lane < a.Cout) {
                                double* sp = a.stats + ((size_t)c.b * a.Cout + n + lane) * 2;
                                atomicAdd(sp, (double)s1);
                                atomicAdd(sp + 1, (double)s2);
                            }
                        }
                    } else {
                        // frame layout (B, T, C, H, W) of conv_img: consecutive lanes = consecutive w of one channel plane
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (j < cw16 && n + j < a.Cout) {
                                float o = v[j] * scale;
                                if (a.bias != nullptr) o += __ldg(a.bias + n + j);
                                a.y[((((long long)c.b * a.T + c.t) * a.Cout + n + j) * a.H + hh) * a.W + ww] = apply_act(o, a.act);
                            }
                        }
                    }
                }
            } else {
            const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * set_cols + col0);
            for (int c0 = 0; c0 < ncw; c0 += 32) {                uint32_t ra[32], rc[32];
                float v[32];
                ptx::tmem_ld_32x32b_x32(tb + (uint32_t)c0, ra);
                if (a.nacc > 1) ptx::tmem_ld_32x32b_x32(tb + (uint32_t)(accw + c0), rc);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(ra[j]) + (a.nacc > 1 ? __uint_as_float(rc[j]) : 0.f)) * scale;
                const int n = n0 + col0 + c0;                           // first output channel of this chunk
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    if (n + 4 * g < a.Cout) {                           // Cout % 4 == 0 (host-checked): whole float4 groups
                        if (a.bias != nullptr) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n) + g);
                            v[4 * g] += b4.x; v[4 * g + 1] += b4.y; v[4 * g + 2] += b4.z; v[4 * g + 3] += b4.w;
                        }
                        if (rrow != nullptr) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + n) + g);
                            v[4 * g] += r4.x; v[4 * g + 1] += r4.y; v[4 * g + 2] += r4.z; v[4 * g + 3] += r4.w;
                        }
                        const float4 o4 = make_float4(apply_act(v[4 * g], a.act), apply_act(v[4 * g + 1], a.act),
                                                      apply_act(v[4 * g + 2], a.act), apply_act(v[4 * g + 3], a.act));
                        reinterpret_cast<float4*>(yrow + n)[g] = o4;
                        v[4 * g] = o4.x; v[4 * g + 1] = o4.y; v[4 * g + 2] = o4.z; v[4 * g + 3] = o4.w;
                    } else {
                        v[4 * g] = 0.f; v[4 * g + 1] = 0.f; v[4 * g + 2] = 0.f; v[4 * g + 3] = 0.f;
                    }
                }
                if (want_stats) {
                    // per-(sample, channel) sum / sum of squares of the STORED values (the next normalisation's statistics)
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
                    const float s1 = column_sums32(v, lane), s2 = column_sums32(sq, lane);
                    if (n + lane < a.Cout) {
                        double* sp = a.stats + ((size_t)c.b * a.Cout + n + lane) * 2;
                        atomicAdd(sp, (double)s1);
                        atomicAdd(sp + 1, (double)s2);
                    }
                }
            }
            }
            // this warp is done with its TMEM columns of the set: tell the leader's MMA issuer (remote arrive for the peer)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(te0 + 8u * (uint32_t)set);
            if (threadIdx.x == 64 && leader) dbg_stamp(6, tile);
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();          // no CTA of the pair leaves (or frees TMEM) while its peer may still signal it
    if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, ncols);
}

__global__ void split_fp16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, float scale,
                                  long long n) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        __half h, l;
        split_f16(x[i] * scale, h, l);
        hi[i] = h;
        lo[i] = l;
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

// tuning switches of the halo kernel (i2v_set_option): residual L2 prefetch, persistent tile loop, pipeline depth
int tc_flags() { return tune().tc_flags; }
int tc_persist() { return tune().tc_persist; }
int tc_min_stages() { return tune().tc_min_stages; }

}  // namespace

int conv_tc_set_debug(unsigned long long* buf, int ctas) {
    I2V_CHECK_CUDA(cudaMemcpyToSymbol(g_dbg, &buf, sizeof(buf)));
    I2V_CHECK_CUDA(cudaMemcpyToSymbol(g_dbg_ctas, &ctas, sizeof(ctas)));
    return 0;
}

// True when the epilogue of launch_conv_tc can also produce the per-(sample, channel) sums of its output
// (every 128/256-row tile lies inside one sample and takes the coalesced channels-last path).
bool conv_tc_fuses_stats(int T, int H, int W) { return (long long)T * H * W >= 128; }

// Shapes the 256-row halo kernel takes (whatever the channel counts, as long as conv_tc_supported holds).
bool conv_tc_halo_eligible(int H, int W, int kh) {
    if (kh != 3 || W < 16 || H * W < 256) return false;
    const int bw = W < 128 ? W : 128, bh2 = 256 / bw;
    return bh2 >= 2 && H % bh2 == 0;
}

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
    I2V_CHECK_CUDA(launch_k(split_fp16_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, hi, lo, scale, n));
    return 0;
}

// Tile / pipeline configuration of the halo kernel for a layer; false when the shape is not eligible.
struct HaloCfg { int bw, bh2, n_tile, kc, stages, nw; size_t stage_bytes; bool wstack; };
static bool halo_config(int H, int W, int Cin, int cout_pad, int kh, int kw, int terms, int variant, HaloCfg& c) {
    if (kh != 3 || W < 16 || H * W < 256) return false;
    c.bw = W < 128 ? W : 128;
    c.bh2 = 256 / c.bw;
    if (c.bh2 < 2 || H % c.bh2 != 0) return false;
    const int n_cap = terms == 3 ? 128 : 256;
    c.n_tile = cout_pad < n_cap ? cout_pad : n_cap;
    // kw-stacked form (variant 0 = automatic, 3 = required, 2 = never): narrow layers (Cout <= 64) stack the three kw
    // taps along the MMA's N (N = 3 Cout <= 192).  One unshifted activation load then feeds 9 taps instead of 3, the
    // MMA count drops 3x, and N >= 48 keeps the tensor pipe off the shared-memory A-read floor that N = 64 MMAs
    // sit on.  Needs whole w-rows per tile (W <= 128) so the +-1 voxel shift never leaves the tile.
    bool wstack = variant != 2 && kw == 3 && cout_pad <= 64 && W == c.bw;
    // channel chunk: the largest of 64/32/16 that divides Cin, keeps tap slabs 1024B-aligned and fits >= 2 stages
    const int mult = terms > 1 ? 2 : 1;
    int kc = 0, stages = 0, nw = 1;
    size_t stage_bytes = 0;
    for (int attempt = 0; attempt < 2 && kc == 0; ++attempt) {
        if (attempt == 1) {
            if (!wstack) break;
            wstack = false;                                          // stacked form does not fit: plain halo form
        }
        nw = wstack ? 3 : 1;
        // the widest chunk that still leaves `want` stages in flight; two stages as the last resort (measured: 32-byte
        // rows just to deepen the pipeline cost 7 % on the g_3 layers, profiles/r01_conv_tc_stage_depth.txt)
        for (int want : {tc_min_stages(), 2}) {
            for (int cand : {64, 32, 16}) {
                if (Cin % cand) continue;
                const size_t rb = (size_t)cand * 2;
                if ((nw * c.n_tile * rb) % 1024 != 0) continue;
                const size_t a_bytes = ((size_t)c.bw * (c.bh2 + 2) * rb + 1023) & ~(size_t)1023;
                const size_t sb = mult * (a_bytes + 3 * nw * c.n_tile * rb);
                const int st = (int)((220 * 1024 - 2048) / sb);
                if (st >= want) { kc = cand; stages = st > 6 ? 6 : st; stage_bytes = sb; break; }
            }
            if (kc != 0) break;
        }
    }
    c.kc = kc; c.stages = stages; c.nw = nw; c.stage_bytes = stage_bytes; c.wstack = wstack;
    return kc != 0;
}

// A 3x3x3 layer can take a fused 1x1x1 side input of Cin2 channels (ConvTcHArgs::cc2) when it runs on the halo kernel
// and its channel chunk divides Cin2.
bool conv_tc_side_eligible(int H, int W, int Cin, int Cin2, int cout_pad, int terms) {
    HaloCfg c;
    return halo_config(H, W, Cin, cout_pad, 3, 3, terms, 0, c) && Cin2 > 0 && Cin2 % c.kc == 0;
}

// Tile / pipeline configuration of the CTA-pair kernel; false when the layer is not eligible.
struct PairCfg { int bw, bh2, n_tile, kc, stages, nw, nacc; size_t stage_bytes; bool wstack; };
constexpr size_t kPairSmemBase = 1024 + 256;               // alignment slack + barriers
constexpr size_t kPairSmemStack = 2048;                    // stacked-epilogue edge rows
static bool pair_config(int H, int W, int Cin, int Cout, int cout_pad, int kh, int kw, int terms, int out_mode, bool stack_ok, PairCfg& c) {
    if (kh != 3 || W < 16 || H * W < 256) return false;
    c.bw = W < 128 ? W : 128;
    c.bh2 = 256 / c.bw;
    if (c.bh2 < 2 || H % c.bh2 != 0) return false;
    // narrow layers (Cout <= 64 on whole w-rows): the three kw taps stacked along N (N = 3 Cout), see wstack_gather
    c.wstack = stack_ok && kw == 3 && cout_pad <= 64 && W == c.bw;
    c.nw = c.wstack ? 3 : 1;
    if (c.wstack) {
        c.n_tile = cout_pad;                               // multiple of 16 -> N = 3 n_tile is a multiple of 16 (cta_group::2)
        if (out_mode == 0 && Cout % 4 != 0) return false;
    } else {
        if (out_mode != 0 || Cout % 4 != 0) return false;
        c.n_tile = cout_pad < 128 ? cout_pad : 128;
        if (c.n_tile % 64 != 0) return false;              // two column halves per lane quarter, 32-column epilogue chunks
    }
    const int accw = c.nw * c.n_tile;
    c.nacc = (terms > 1 && 2 * 2 * accw <= 512) ? 2 : 1;   // two accumulator SETS must fit the 512 TMEM columns
    const int mult = terms > 1 ? 2 : 1;
    const int bh_sub = c.bh2 / 2;
    c.kc = 0;
    // the widest channel chunk that still leaves `want` ring buffers (64-byte rows at 4 stages beat 128-byte rows at 2)
    for (int want : {4, 3, 2}) {
        for (int cand : {64, 32, 16}) {
            if (Cin % cand) continue;
            const size_t rb = (size_t)cand * 2;
            // every kh slab of the weight tile starts on a swizzle period (8 rows: 1024 / 512 / 256 bytes for 128 / 64 / 32-byte rows)
            if ((c.nw * c.n_tile / 2) % 8 != 0) continue;
            const size_t a_bytes = ((size_t)c.bw * (bh_sub + 2) * rb + 1023) & ~(size_t)1023;
            const size_t sb = (mult * (a_bytes + 3 * (size_t)(c.nw * c.n_tile / 2) * rb) + 1023) & ~(size_t)1023;
            const int st = (int)((227 * 1024 - kPairSmemBase - (c.wstack ? kPairSmemStack : 0)) / sb);
            if (st >= want) { c.kc = cand; c.stages = st > 8 ? 8 : st; c.stage_bytes = sb; break; }
        }
        if (c.kc != 0) break;
    }
    return c.kc != 0;
}

// v3 eligibility + launch.  Returns 1 if the layer is not eligible (caller falls through to the halo kernel).
// variant 0 / 4: kw-stacked form for narrow layers, plain form otherwise; 5: never stacked.
static int launch_conv_tc_pair(const ConvTcArgs& h, cudaStream_t stream) {
    PairCfg cfg{};
    if (!pair_config(h.H, h.W, h.Cin, h.Cout, h.cout_pad, h.kh, h.kw, h.terms, h.out_mode, h.variant != 5, cfg)) return 1;
    // Narrow layers are bound by L2 -> SM operand traffic, and a pair's 128-voxel sub-tiles carry more halo rows per output
    // row than the single-CTA kernel's 256-voxel tile: measured (profiles/r02_bench_ab.txt) the stacked pair form wins on
    // conv_img (tiny K, epilogue-bound) and loses on g_4.conv_0 -- automatic selection takes it for the frame layout only.
    if (h.variant == 0 && cfg.wstack && !(tune().tc_pair_stack == 2 || (tune().tc_pair_stack == 1 && h.out_mode == 1))) return 1;
    if (tune().tc_pair_stages >= 2 && cfg.stages > tune().tc_pair_stages) cfg.stages = tune().tc_pair_stages;
    if (h.t_phase && !(h.kt == 3 && h.T % 2 == 0)) return 1;
    if (h.stats != nullptr && h.out_mode != 0) return 1;
    const int kt_eff = h.t_phase ? 2 : h.kt;
    const int Tin = h.t_phase ? h.T / 2 : h.T;
    const int kc = cfg.kc, cch = h.Cin / kc;
    if (h.Cin2 != 0 && !(h.x2_hi && h.w2_hi && h.Cin2 % kc == 0 && !h.t_phase && h.kt == 3 && h.kw == 3)) return 1;
    const int cc2 = h.Cin2 / kc;
    // K split across launches: one main accumulator per set, chains of <= kMaxChain truncating MMAs (see launch_conv_tc_halo);
    // a shared accumulator (stacked N = 192) also takes the two cross-term MMAs of every k-step
    int parts = 1;
    if (h.terms > 1) {
        const int kw_iter = cfg.wstack ? 1 : h.kw;
        const int per_k = cfg.nacc > 1 ? 1 : 3;
        const long long chain = (long long)kt_eff * kw_iter * cch * 3 * (kc / 16) * per_k + (long long)cc2 * (kc / 16) * per_k;
        parts = (int)((chain + kMaxChain - 1) / kMaxChain);
        if (parts > cch) parts = cch;
        if (parts < 1) parts = 1;
    }
    if (parts > 1 && (h.act != ACT_NONE || cc2 != 0)) return 1;
    const int cper = (cch + parts - 1) / parts;
    ConvTcPArgs a;
    a.bias = h.bias; a.res = h.res; a.scale_ptr = h.scale_ptr; a.y = h.y; a.stats = h.stats;
    a.B = h.B; a.T = h.T; a.H = h.H; a.W = h.W; a.Cin = h.Cin; a.Cout = h.Cout; a.kt = kt_eff; a.kw = h.kw;
    a.t_phase = h.t_phase ? 1 : 0; a.wstack = cfg.wstack ? 1 : 0; a.nacc = cfg.nacc; a.out_mode = h.out_mode;
    a.bw = cfg.bw; a.bh2 = cfg.bh2; a.tiles_w = h.W / a.bw; a.tiles_h = h.H / a.bh2;
    a.n_tile = cfg.n_tile; a.kc = kc; a.stages = cfg.stages; a.terms = h.terms;
    a.res_ut = h.res_ut; a.res_uh = h.res_uh; a.res_uw = h.res_uw; a.act = h.act;
    a.cc2 = cc2;
    const int rb = kc * 2, bh_sub = a.bh2 / 2;

    CUtensorMap mAh, mAl, mBh, mBl;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)Tin, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.W * h.Cin * 2, (cuuint64_t)h.H * h.W * h.Cin * 2,
                                  (cuuint64_t)Tin * h.H * h.W * h.Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)a.bw, (cuuint32_t)(bh_sub + 2), 1, 1};
        if (int rc = encode_map(&mAh, h.x_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mAl, h.terms > 1 ? h.x_lo : h.x_hi, 5, dims, st, box, rb)) return rc;
    }
    {
        const cuuint64_t row = (cuuint64_t)h.cout_pad * h.Cin * 2;
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.cout_pad, (cuuint64_t)h.kw, 3, (cuuint64_t)(h.t_phase ? 4 : h.kt)};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, row, row * h.kw, row * h.kw * 3};
        // this CTA's half of the N tile; stacked form: its three kw slabs in one box -> smem [kh][kw][n_tile / 2][kc]
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)(a.n_tile / 2), (cuuint32_t)cfg.nw, 3, 1};
        if (int rc = encode_map(&mBh, h.w_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mBl, h.terms > 1 ? h.w_lo : h.w_hi, 5, dims, st, box, rb)) return rc;
    }
    CUtensorMap mA2h = mAh, mA2l = mAl, mB2h = mBh, mB2l = mBl;
    if (cc2 > 0) {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin2, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)h.T, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin2 * 2, (cuuint64_t)h.W * h.Cin2 * 2, (cuuint64_t)h.H * h.W * h.Cin2 * 2,
                                  (cuuint64_t)h.T * h.H * h.W * h.Cin2 * 2};
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)a.bw, (cuuint32_t)(bh_sub + 2), 1, 1};
        if (int rc = encode_map(&mA2h, h.x2_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mA2l, h.terms > 1 ? h.x2_lo : h.x2_hi, 5, dims, st, box, rb)) return rc;
        const cuuint64_t row = (cuuint64_t)h.cout_pad * h.Cin2 * 2;
        const cuuint64_t wd[5] = {(cuuint64_t)h.Cin2, (cuuint64_t)h.cout_pad, 3, 1, 1};
        const cuuint64_t ws_[4] = {(cuuint64_t)h.Cin2 * 2, row, row * 3, row * 3};
        const cuuint32_t wbox[5] = {(cuuint32_t)kc, (cuuint32_t)(a.n_tile / 2), (cuuint32_t)cfg.nw, 1, 1};
        if (int rc = encode_map(&mB2h, h.w2_hi, 5, wd, ws_, wbox, rb)) return rc;
        if (int rc = encode_map(&mB2l, h.terms > 1 ? h.w2_lo : h.w2_hi, 5, wd, ws_, wbox, rb)) return rc;
    }
    static unsigned long long attr_devs = 0;
    I2V_CHECK_CUDA(ensure_max_dyn_smem(conv_tc_pair_kernel, 227 * 1024, attr_devs));
    const size_t smem = (size_t)cfg.stages * cfg.stage_bytes + kPairSmemBase + (cfg.wstack ? kPairSmemStack : 0);
    I2V_REQUIRE(smem <= 227 * 1024, "conv_tc: pair kernel shared memory (%zu bytes) over the 227 KB limit", smem);
    if (h.dry_run) return 0;        // every check above passed
    const long long M = (long long)h.B * h.T * h.H * h.W;
    const double K_ = (double)h.kt * h.kh * h.kw * h.Cin + (double)h.Cin2;
    ProfScope ps(PROF_CONV, 2.0 * (double)M * h.Cout * K_,
                 4.0 * ((double)M * (h.Cin + h.Cin2) + (double)M * h.Cout + K_ * h.Cout), stream);
    // persistent pairs: one per TPC over all N tiles, each walking its share of the 256-voxel output tiles
    const long long n_tiles = (long long)a.tiles_w * a.tiles_h * h.T * h.B;
    const unsigned ny = (unsigned)((h.cout_pad + a.n_tile - 1) / a.n_tile);
    long long pairs = (kNumSMs / 2) / ny > 0 ? (kNumSMs / 2) / ny : 1;
    if (pairs > n_tiles) pairs = n_tiles;
    a.n_tiles = (int)n_tiles;
    dim3 grid((unsigned)(2 * pairs), ny);
    for (int p = 0; p < parts; ++p) {
        a.cc_lo = p * cper;
        a.cc_hi = (p + 1) * cper < cch ? (p + 1) * cper : cch;
        if (a.cc_lo >= a.cc_hi) break;
        const bool last = a.cc_hi == cch;
        if (p > 0) { a.bias = nullptr; a.res = h.y; a.res_ut = a.res_uh = a.res_uw = 1; }
        a.stats = last ? h.stats : nullptr;
        I2V_CHECK_CUDA(launch_k(conv_tc_pair_kernel, grid, dim3(TC_THREADS), smem, stream, mAh, mAl, mBh, mBl, mA2h, mA2l, mB2h, mB2l, a));
    }
    return 0;
}

// v2 eligibility + launch.  Returns 1 if the shape is not eligible (caller falls through to v1).
static int launch_conv_tc_halo(const ConvTcArgs& h, cudaStream_t stream) {
    HaloCfg cfg{};
    if (!halo_config(h.H, h.W, h.Cin, h.cout_pad, h.kh, h.kw, h.terms, h.variant, cfg)) {
        I2V_REQUIRE(h.variant != 3, "conv_tc: shape not eligible for the kw-stacked halo kernel");
        return 1;
    }
    I2V_REQUIRE(h.variant != 3 || cfg.wstack, "conv_tc: shape not eligible for the kw-stacked halo kernel");
    I2V_REQUIRE(!h.t_phase || (h.kt == 3 && h.T % 2 == 0), "conv_tc: t_phase needs a 3-tap temporal kernel and even T");
    const int kt_eff = h.t_phase ? 2 : h.kt;     // temporal taps actually iterated
    const int Tin = h.t_phase ? h.T / 2 : h.T;   // planes of the stored activation tensor
    ConvTcHArgs a;
    a.bw = cfg.bw; a.bh2 = cfg.bh2; a.n_tile = cfg.n_tile;
    const bool wstack = cfg.wstack;
    const int kc = cfg.kc, stages = cfg.stages, nw = cfg.nw;
    const size_t stage_bytes = cfg.stage_bytes;
    a.wstack = wstack ? 1 : 0;
    a.kc = kc; a.stages = stages; a.terms = h.terms;
    int nacc = 512 / (2 * nw * a.n_tile);
    if (nacc > 4) nacc = 4;
    const int cch = h.Cin / kc;
    const int kw_iter = wstack ? 1 : h.kw;
    // K split across launches: the tensor core's fp32 accumulate TRUNCATES, so the error of one accumulator
    // grows with the number of MMAs chained into it (measured ~1e-5 at ~1300 chained MMAs).  Keep every chain
    // at <= kMaxChain MMAs: split the channel chunks over several launches whose partial results are combined by
    // the epilogue's exact fp32 add (launch p > 0 reads y as its residual, in place).
    // accumulators per sub-tile: terms == 3 -> nmain for hi*hi + as many for the cross terms; terms == 1 -> all main
    int nmain = h.terms > 1 ? nacc / 2 : nacc;
    // stacked N = 192 leaves room for ONE accumulator per sub-tile: the cross terms then share it (3x the chain)
    const bool shared_acc = h.terms > 1 && nacc == 1;
    if (shared_acc) nmain = 1;
    if (nmain < 1) return 1;                                                   // needs 2 accumulators per sub-tile
    if (nmain > 2) nmain = 2;
    const int main_per_stage0 = 3 * (kc / 16) * (shared_acc ? 3 : 1);
    // short reductions (conv_img: 36 chained MMAs) do not need a second main accumulator: one fewer TMEM round trip per
    // epilogue chunk
    if (nmain > 1 && (long long)kt_eff * kw_iter * cch * main_per_stage0 + (long long)(h.Cin2 / kc) * (main_per_stage0 / 3) <= kMaxChain / 2)
        nmain = 1;
    const int main_per_stage = 3 * (kc / 16) * (shared_acc ? 3 : 1);           // chained MMAs per stage and sub-tile
    // side input: Cin2 / kc extra stages through the centre tap (one kh row each)
    I2V_REQUIRE(h.Cin2 == 0 || (h.x2_hi && h.w2_hi && h.Cin2 % kc == 0 && !h.t_phase && h.kt == 3 && h.kw == 3),
                "conv_tc: side input needs a 3x3x3 non-phase conv and Cin2 (%d) divisible by the channel chunk %d", h.Cin2, kc);
    const int cc2 = h.Cin2 / kc;
    int parts = 1;
    if (h.terms > 1) {
        const long long chain = ((long long)kt_eff * kw_iter * cch * main_per_stage + (long long)cc2 * (main_per_stage / 3)) / nmain;
        parts = (int)((chain + kMaxChain - 1) / kMaxChain);
        if (parts > cch) parts = cch;
        if (parts < 1) parts = 1;
    }
    I2V_REQUIRE(parts == 1 || h.act == ACT_NONE, "conv_tc: K-split launches need a linear epilogue");
    I2V_REQUIRE(parts == 1 || cc2 == 0, "conv_tc: a side input cannot be combined with K-split launches");
    const int cper = (cch + parts - 1) / parts;
    {
        // every CTA runs at least kw * (chunks of the smallest part) stages: each accumulator must be written
        const int last_chunks = cch - (parts - 1) * cper;
        const int min_stages = kw_iter * (last_chunks < cper ? last_chunks : cper);
        if (nmain > min_stages) nmain = min_stages;
    }
    a.nmain = nmain;
    a.nacc = (h.terms > 1 && !shared_acc) ? 2 * nmain : nmain;
    nacc = a.nacc;
    a.t_phase = h.t_phase ? 1 : 0;
    a.flags = tc_flags();
    a.cc2 = cc2;
    a.bias = h.bias; a.res = h.res; a.scale_ptr = h.scale_ptr; a.y = h.y; a.stats = h.stats;
    I2V_REQUIRE(h.stats == nullptr || (h.out_mode == 0 && a.n_tile <= 256), "conv_tc: fused statistics need channels-last output");
    a.B = h.B; a.T = h.T; a.H = h.H; a.W = h.W; a.Cin = h.Cin; a.Cout = h.Cout; a.kt = kt_eff; a.kw = h.kw;
    a.tiles_w = h.W / a.bw; a.tiles_h = h.H / a.bh2;
    a.res_ut = h.res_ut; a.res_uh = h.res_uh; a.res_uw = h.res_uw; a.act = h.act; a.out_mode = h.out_mode;
    const int rb = kc * 2;

    CUtensorMap mAh, mAl, mBh, mBl;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)Tin, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.W * h.Cin * 2, (cuuint64_t)h.H * h.W * h.Cin * 2,
                                  (cuuint64_t)Tin * h.H * h.W * h.Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)a.bw, (cuuint32_t)(a.bh2 + 2), 1, 1};
        if (int rc = encode_map(&mAh, h.x_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mAl, h.terms > 1 ? h.x_lo : h.x_hi, 5, dims, st, box, rb)) return rc;
    }
    {
        // weights [kt][kh][kw][cout_pad][Cin] viewed as (Cin, cout_pad, kw, kh, kt) (strides ascending): the 3 kh taps of
        // one (kt, kw) arrive as a single box -> smem [kh][n][kc]
        const cuuint64_t row = (cuuint64_t)h.cout_pad * h.Cin * 2;
        // temporal extent of the weight tensor: kt taps, or 2 phases x 2 taps ([phase][tap][kh][kw][cout][cin])
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.cout_pad, (cuuint64_t)h.kw, 3, (cuuint64_t)(h.t_phase ? 4 : h.kt)};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, row, row * h.kw, row * h.kw * 3};
        // stacked form: all 9 (kh, kw) taps of one kt in a single box -> smem [kh][kw][n][kc]
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)a.n_tile, (cuuint32_t)nw, 3, 1};
        if (int rc = encode_map(&mBh, h.w_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mBl, h.terms > 1 ? h.w_lo : h.w_hi, 5, dims, st, box, rb)) return rc;
    }
    CUtensorMap mA2h = mAh, mA2l = mAl, mB2h = mBh, mB2l = mBl;     // placeholders when there is no side input
    if (cc2 > 0) {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin2, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)h.T, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin2 * 2, (cuuint64_t)h.W * h.Cin2 * 2, (cuuint64_t)h.H * h.W * h.Cin2 * 2,
                                  (cuuint64_t)h.T * h.H * h.W * h.Cin2 * 2};
        const cuuint32_t box[5] = {(cuuint32_t)kc, (cuuint32_t)a.bw, (cuuint32_t)(a.bh2 + 2), 1, 1};
        if (int rc = encode_map(&mA2h, h.x2_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mA2l, h.terms > 1 ? h.x2_lo : h.x2_hi, 5, dims, st, box, rb)) return rc;
        const cuuint64_t row = (cuuint64_t)h.cout_pad * h.Cin2 * 2;
        const cuuint64_t wd[5] = {(cuuint64_t)h.Cin2, (cuuint64_t)h.cout_pad, 3, 1, 1};
        const cuuint64_t ws_[4] = {(cuuint64_t)h.Cin2 * 2, row, row * 3, row * 3};
        const cuuint32_t wbox[5] = {(cuuint32_t)kc, (cuuint32_t)a.n_tile, (cuuint32_t)nw, 1, 1};
        if (int rc = encode_map(&mB2h, h.w2_hi, 5, wd, ws_, wbox, rb)) return rc;
        if (int rc = encode_map(&mB2l, h.terms > 1 ? h.w2_lo : h.w2_hi, 5, wd, ws_, wbox, rb)) return rc;
    }
    static unsigned long long attr_devs = 0;
    I2V_CHECK_CUDA(ensure_max_dyn_smem(conv_tc_halo_kernel, 227 * 1024, attr_devs));
    // ring + alignment slack + barriers + (stacked form) 2 slots x 4 quarters x 2 edge rows of n_tile floats
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + (wstack ? 16 * (size_t)a.n_tile * sizeof(float) : 0);
    I2V_REQUIRE(smem <= 227 * 1024, "conv_tc: halo kernel shared memory (%zu bytes) over the 227 KB limit", smem);
    {
        // transpose tiles of the coalesced epilogue: whole tiles per ring buffer, in the buffers a tile fills last
        const size_t nh0 = (size_t)((a.n_tile / 16 + 1) / 2) * 16;
        const size_t stile_bytes = 32 * (nh0 + 4) * sizeof(float);
        int sps = h.out_mode == 0 ? (int)(stage_bytes / stile_bytes) : 0;
        if (sps > 8) sps = 8;
        int epi_slots = sps > 0 ? (8 + sps - 1) / sps : 0;
        if (epi_slots > stages) { sps = 0; epi_slots = 0; }          // does not fit: one-row-per-thread write-out
        I2V_REQUIRE(h.stats == nullptr || sps > 0, "conv_tc: transpose tiles do not fit the pipeline buffers, fused statistics unavailable");
        I2V_REQUIRE(!wstack || h.out_mode != 0 || sps > 0, "conv_tc: stacked epilogue buffers do not fit the pipeline buffers");
        I2V_REQUIRE(!wstack || (long long)a.tiles_w == 1, "conv_tc: stacked form needs whole rows per tile");
        a.stile_per_slot = sps;
        a.prefetch_stages = stages - epi_slots;
    }
    if (h.dry_run) return 0;        // every check above passed
    const long long M = (long long)h.B * h.T * h.H * h.W;
    const double K_ = (double)h.kt * h.kh * h.kw * h.Cin + (double)h.Cin2;
    ProfScope ps(PROF_CONV, 2.0 * (double)M * h.Cout * K_,
                 4.0 * ((double)M * (h.Cin + h.Cin2) + (double)M * h.Cout + K_ * h.Cout), stream);
    // persistent CTAs: one per SM over all N tiles, each walking its share of the output tiles (I2V_TC_PERSIST=0: one
    // CTA per tile, the pre-persistent schedule)
    const long long n_tiles = (long long)a.tiles_w * a.tiles_h * h.T * h.B;
    const unsigned ny = (unsigned)((h.cout_pad + a.n_tile - 1) / a.n_tile);
    long long gx = n_tiles;
    if (tc_persist()) {
        const long long per_y = kNumSMs / ny > 0 ? kNumSMs / ny : 1;
        if (gx > per_y) gx = per_y;
    }
    a.n_tiles = (int)n_tiles;
    dim3 grid((unsigned)gx, ny);
    for (int p = 0; p < parts; ++p) {
        a.cc_lo = p * cper;
        a.cc_hi = (p + 1) * cper < cch ? (p + 1) * cper : cch;
        if (a.cc_lo >= a.cc_hi) break;
        const bool last = a.cc_hi == cch;
        if (p > 0) { a.bias = nullptr; a.res = h.y; a.res_ut = a.res_uh = a.res_uw = 1; }
        a.stats = last ? h.stats : nullptr;
        I2V_CHECK_CUDA(launch_k(conv_tc_halo_kernel, grid, dim3(TC_THREADS), smem, stream, mAh, mAl, mBh, mBl, mA2h, mA2l, mB2h, mB2l, a));
    }
    return 0;
}

int launch_conv_tc(const ConvTcArgs& h, cudaStream_t stream) {
    I2V_REQUIRE(conv_tc_supported(h.B, h.T, h.H, h.W, h.Cin, h.Cout, h.kt, h.kh, h.kw),
                "conv_tc: unsupported shape B=%d T=%d H=%d W=%d Cin=%d k=(%d,%d,%d)", h.B, h.T, h.H, h.W, h.Cin, h.kt, h.kh, h.kw);
    I2V_REQUIRE(h.terms == 1 || h.terms == 3, "conv_tc: terms must be 1 or 3");
    I2V_REQUIRE(h.cout_pad % 16 == 0 && h.cout_pad >= h.Cout, "conv_tc: weights must be padded to a multiple of 16 output rows");
    I2V_REQUIRE(h.res == nullptr || (h.T % h.res_ut == 0 && h.H % h.res_uh == 0 && h.W % h.res_uw == 0),
                "conv_tc: residual upsample factors must divide the output size");
    if ((h.variant == 0 && tune().tc_pair) || h.variant == 4 || h.variant == 5) {
        const int rc = launch_conv_tc_pair(h, stream);
        if (rc <= 0) return rc;     // launched (0) or failed (<0); 1 = not eligible -> the single-CTA kernels below
        I2V_REQUIRE(h.variant != 4 && h.variant != 5, "conv_tc: shape not eligible for the CTA-pair kernel");
    }
    if (h.variant != 1) {
        const int rc = launch_conv_tc_halo(h, stream);
        if (rc <= 0) return rc;     // launched (0) or failed (<0); 1 = shape not eligible -> v1 below
        I2V_REQUIRE(h.variant != 2, "conv_tc: shape not eligible for the halo kernel");
    }
    I2V_REQUIRE(!h.t_phase || (h.kt == 3 && h.T % 2 == 0), "conv_tc: the phase form needs kt = 3 and an even output frame count");
    I2V_REQUIRE(h.Cin2 == 0, "conv_tc: a side input is only implemented by the halo kernel (ask conv_tc_halo_eligible first)");
    ConvTcKArgs a;
    a.bias = h.bias; a.res = h.res; a.scale_ptr = h.scale_ptr; a.y = h.y; a.stats = h.stats;
    a.B = h.B; a.T = h.T; a.H = h.H; a.W = h.W; a.Cin = h.Cin; a.Cout = h.Cout;
    a.kt = h.kt; a.kh = h.kh; a.kw = h.kw;
    a.kc = h.Cin % 64 == 0 ? 64 : (h.Cin % 32 == 0 ? 32 : 16);
    a.t_phase = h.t_phase ? 1 : 0;
    const int Tin = h.t_phase ? h.T / 2 : h.T;       // stored input planes (tiles walk these)
    // box: 128 rows = bw x bh x bt x bb voxels
    int rem = TILE_M;
    a.bw = h.W < rem ? h.W : rem; rem /= a.bw;
    a.bh = h.H < rem ? h.H : rem; rem /= a.bh;
    a.bt = Tin < rem ? Tin : rem;
    // a two-frame clip under a temporal 3-tap kernel (g_0.conv_1: 2 x 8 x 8): a tile that holds BOTH frames needs all three
    // temporal taps, two of them half padding; tiles of one frame (and two samples) skip the tap that falls outside -- 2/3 of
    // the MMAs and operand loads.  The fused statistics (one sample per tile) give way to the separate pass: 8x8 planes.
    if (!h.t_phase && h.kt == 3 && h.T == 2 && a.bt == 2 && (a.bw * a.bh) % 32 == 0 && tune().tc_t2_split) a.bt = 1;
    rem /= a.bt;
    a.bb = rem;
    a.tiles_w = h.W / a.bw; a.tiles_h = h.H / a.bh; a.tiles_t = Tin / a.bt;
    const int tiles_b = (h.B + a.bb - 1) / a.bb;
    I2V_REQUIRE(h.stats == nullptr || (a.bb == 1 && h.out_mode == 0),
                "conv_tc: fused statistics need one sample per tile (ask conv_tc_fuses_stats first)");
    // fp32-grade mode keeps N <= 128 so that four TMEM accumulators fit (see the MMA issuer).  (N = 256 with one main + one
    // cross-term accumulator and twice the K-split parts was measured on g_0, 1024 -> 1024 on 2x8x8 planes: no gain,
    // profiles/r01_bench_ab_v2.txt call 12.)
    const int n_cap = h.terms == 3 ? 128 : 256;
    a.n_tile = h.cout_pad < n_cap ? h.cout_pad : n_cap;
    a.terms = h.terms;
    const int cch1 = h.Cin / a.kc;
    int parts1 = 1;
    {
        int nacc = 512 / a.n_tile;
        if (nacc > 4) nacc = 4;
        int nmain = h.terms > 1 ? nacc / 2 : nacc;        // N <= 128 in the 3-term mode, so nacc >= 4 there
        if (nmain < 1) nmain = 1;
        // K split across launches, same reasoning as the halo kernel (main chains of <= kMaxChain truncating MMAs)
        if (h.terms > 1) {
            // temporal taps that can be live in one tile: 2 in the phase form (1 when there is a single source plane), 2 for
            // the one-frame tiles of a two-frame clip, kt otherwise
            const int kt_live = h.t_phase ? (Tin == 1 ? 1 : 2) : (h.kt == 3 && h.T == 2 && a.bt == 1 ? 2 : h.kt);
            const long long chain = (long long)kt_live * h.kh * h.kw * cch1 * (a.kc / 16) / nmain;
            parts1 = (int)((chain + kMaxChain - 1) / kMaxChain);
            if (parts1 > cch1) parts1 = cch1;
            if (parts1 < 1) parts1 = 1;
        }
        const int cper = (cch1 + parts1 - 1) / parts1;
        const int last_chunks = cch1 - (parts1 - 1) * cper;
        const int min_stages = last_chunks < cper ? last_chunks : cper;     // the centre tap is always valid
        if (nmain > min_stages) nmain = min_stages;
        a.nmain = nmain;
        a.nacc = h.terms > 1 ? 2 * nmain : nmain;
    }
    I2V_REQUIRE(parts1 == 1 || h.act == ACT_NONE, "conv_tc: K-split launches need a linear epilogue");
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
    {
        const size_t nh0 = (size_t)((a.n_tile / 16 + 1) / 2) * 16;
        I2V_REQUIRE(h.stats == nullptr || 8 * 32 * (nh0 + 4) * sizeof(float) <= (size_t)stages * stage_bytes,
                    "conv_tc: transpose tiles do not fit the pipeline buffers, fused statistics unavailable");
    }

    CUtensorMap mAh, mAl, mBh, mBl;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)h.Cin, (cuuint64_t)h.W, (cuuint64_t)h.H, (cuuint64_t)Tin, (cuuint64_t)h.B};
        const cuuint64_t st[4] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.W * h.Cin * 2, (cuuint64_t)h.H * h.W * h.Cin * 2,
                                  (cuuint64_t)Tin * h.H * h.W * h.Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)a.kc, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.bt, (cuuint32_t)a.bb};
        if (int rc = encode_map(&mAh, h.x_hi, 5, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mAl, h.terms > 1 ? h.x_lo : h.x_hi, 5, dims, st, box, rb)) return rc;
    }
    {
        const int taps = (h.t_phase ? 4 : h.kt) * h.kh * h.kw;
        const cuuint64_t dims[3] = {(cuuint64_t)h.Cin, (cuuint64_t)h.cout_pad, (cuuint64_t)taps};
        const cuuint64_t st[2] = {(cuuint64_t)h.Cin * 2, (cuuint64_t)h.cout_pad * h.Cin * 2};
        const cuuint32_t box[3] = {(cuuint32_t)a.kc, (cuuint32_t)a.n_tile, 1};
        if (int rc = encode_map(&mBh, h.w_hi, 3, dims, st, box, rb)) return rc;
        if (int rc = encode_map(&mBl, h.terms > 1 ? h.w_lo : h.w_hi, 3, dims, st, box, rb)) return rc;
    }
    static unsigned long long attr_devs = 0;
    I2V_CHECK_CUDA(ensure_max_dyn_smem(conv_tc_kernel, 227 * 1024, attr_devs));
    if (h.dry_run) return 0;        // every check above passed
    const long long M = (long long)h.B * h.T * h.H * h.W;
    const double K_ = (double)h.kt * h.kh * h.kw * h.Cin;
    ProfScope ps(PROF_CONV_TC1, 2.0 * (double)M * h.Cout * K_, 4.0 * ((double)M * h.Cin + (double)M * h.Cout + K_ * h.Cout), stream);
    dim3 grid((unsigned)(a.tiles_w * a.tiles_h * a.tiles_t * tiles_b), (unsigned)((h.cout_pad + a.n_tile - 1) / a.n_tile),
              h.t_phase ? 2u : 1u);
    const int cper1 = (cch1 + parts1 - 1) / parts1;
    for (int p = 0; p < parts1; ++p) {
        a.cc_lo = p * cper1;
        a.cc_hi = (p + 1) * cper1 < cch1 ? (p + 1) * cper1 : cch1;
        if (a.cc_lo >= a.cc_hi) break;
        const bool last = a.cc_hi == cch1;
        if (p > 0) { a.bias = nullptr; a.res = h.y; a.res_ut = a.res_uh = a.res_uw = 1; }
        a.stats = last ? h.stats : nullptr;
        I2V_CHECK_CUDA(launch_k(conv_tc_kernel, grid, dim3(TC_THREADS), smem, stream, mAh, mAl, mBh, mBl, a));
    }
    return 0;
}

}  // namespace i2v
