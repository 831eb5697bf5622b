// Conditional INN (stage2_cINN) as cluster-resident nets: one thread-block cluster of 16 CTAs carries up to 8 batch rows
// through ALL couplings without ever synchronising with another cluster.
//
// Reference semantics: see flow.cu (flow_blocks.py:31-187, modules.py:9-104).  What changes is the schedule.
//
// The cooperative kernel in flow.cu spreads every Linear layer over all 148 SMs and pays, per dependent layer, a grid-wide
// atomic barrier (~3 us) plus the broadcast of the whole [B, 512] activation matrix into every SM (~9 us): 160 layers ->
// 2.07 ms at B = 64 against a 29 us weight-streaming floor (profiles/r01_flow_phase_timestamps_B64.txt).  Here
//   * the batch is cut into groups of 8 rows, one CLUSTER per group; rows never interact, so clusters never synchronise;
//   * inside a cluster the scale net lives on CTAs 0..7 and the translation net on CTAs 8..15; a CTA owns H/8 output
//     columns of every hidden Linear and 4 of the 32 outputs of the last one;
//   * layer outputs are PUSHED into the peers' shared memory (st.shared::cluster, 16-byte vectors) and handed over with one
//     remote mbarrier arrive per peer (release/acquire at cluster scope) -- no global memory, no grid barrier;
//   * weights do not depend on the data: a producer warp streams the CTA's weight slabs -- repacked at load time into
//     the order they are consumed, k-major so that shared-memory reads are conflict-free (loader.pack_flow "wpack") -- with
//     cp.async.bulk through a 20-deep ring of 8 KB chunks, running as far ahead of the dependency chain as the ring allows;
//   * a hidden layer splits K over the 8 compute warps (each warp = the k-rows of its own chunks, 8 rows x 2 columns of
//     accumulators per lane, float4 activation broadcasts), partial sums meet in shared memory.
// Algorithmic bytes are unchanged (weights once per cluster instead of once per grid: 8 x 189 MB of L2 -> SM traffic at
// B = 64, HBM traffic 189 MB thanks to L2), the dependent-step latency drops from ~12 us to ~2 us.
#include <cuda_runtime.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include "ptx_sm100.cuh"

namespace i2v {

namespace {

constexpr int FC_CLUSTER = 16;                // CTAs per cluster: 8 per net
constexpr int FC_NET = 8;
constexpr int FC_WARPS = 16;                  // compute warps: one per k-chunk of a hidden layer
constexpr int FC_CTHREADS = FC_WARPS * 32;
constexpr int FC_THREADS = FC_CTHREADS + 32;  // + the weight-streaming warp
constexpr int FC_D = 64, FC_HALF = 32;        // latent width of every reference config (z_dim = 64)
constexpr int FC_MAX_RING = 32;
constexpr int FC_MAX_CLUSTERS = 7;            // 16-CTA clusters a B200 co-schedules (cudaOccupancyMaxActiveClusters, measured)

struct FlowClusterArgs {
    const float* wpack;   // [2*n_flows][16][32*H/8 + depth*H*H/8 + 4*H] per-CTA weight stream, k-major (loader.pack_flow_chunks)
    const float* c1;      // [B, n_flows*2*2H]  hoisted conditioning part of the first Linear (+ bias)
    const float* bh;      // [n_flows, 2, depth, 2H]
    const float* bo;      // [n_flows, 2, 64]
    const float* loc; const float* scale;       // [n_flows, 64]
    const int* perm_fwd; const int* perm_bwd;   // [n_flows, 64]
    const float* in; float* out; float* logdet;
    int B, n_flows, H, depth, reverse, ring, rows_per_cluster;
    unsigned char cond_mode[64];
};

// optional phase timestamps (i2v_debug_flow_timestamps): 16 x u64 for coupling #4 of cluster 0, CTA ranks 0 and 9
__device__ unsigned long long* g_fc_dbg = nullptr;
__device__ __forceinline__ void fdbg(int coupling, int slot, uint32_t rank) {
    if (g_fc_dbg != nullptr && coupling == 4 && threadIdx.x == 0 && blockIdx.x < FC_CLUSTER && (rank == 0 || rank == 9)) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_fc_dbg[(rank == 0 ? 0 : 16) + slot] = t;
    }
}

__device__ __forceinline__ float lrelu001(float v) { return v >= 0.f ? v : 0.01f * v; }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(FC_CTHREADS) : "memory"); }   // compute warps only

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Hand-over of a pushed tensor: every compute thread of this CTA has issued its remote stores; one remote arrive per peer of
// `group` (first CTA rank `g0`, `gn` CTAs, this CTA included), then wait until all `gn` peers have arrived here.
__device__ __forceinline__ void exchange(uint64_t* bar, uint32_t parity, int g0, int gn) {
    cbar();
    if ((int)threadIdx.x < gn) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(bar), (uint32_t)(g0 + threadIdx.x)));
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}

// CPL = output columns per lane in the first / hidden layers (H/8 columns per CTA over 32 lanes): 2 for H = 512, 1 for H = 256
// ROWS = batch rows a cluster carries (8, or 10 so that 64 rows fit the 7 clusters a B200 co-schedules)
template <int CPL, int ROWS>
__global__ void __launch_bounds__(FC_THREADS, 1) flow_cluster_kernel(const FlowClusterArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int H = a.H, Cc = H / FC_NET;                 // columns of a hidden layer this CTA owns
    const int KC = H / FC_WARPS;                        // k-rows of one hidden-layer chunk (one chunk per compute warp)
    const int l1_floats = FC_HALF * Cc, hc_floats = KC * Cc, l4_floats = 4 * H;
    const int slot_floats = l1_floats > l4_floats ? l1_floats : l4_floats;      // >= hc_floats
    const int cc_shift = Cc == 64 ? 6 : 5;
    float* ring = reinterpret_cast<float*>(smem_raw);
    float* act0 = ring + (size_t)a.ring * slot_floats;                  // [ROWS][H]
    float* act1 = act0 + ROWS * H;
    float* part = act1 + ROWS * H;                                      // [FC_WARPS][ROWS][Cc]
    float* xs = part + FC_WARPS * ROWS * Cc;                            // [ROWS][64] state
    float* xt = xs + ROWS * FC_D;
    float* stb = xt + ROWS * FC_D;                                      // [ROWS][64]  (s | t) of the current coupling
    float* lds = stb + ROWS * FC_D;                                     // [16] running log-det
    uint64_t* full = reinterpret_cast<uint64_t*>(lds + 16);
    uint64_t* empty = full + FC_MAX_RING;
    uint64_t* act_ready = empty + FC_MAX_RING;                          // [2]
    uint64_t* st_ready = act_ready + 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int net = (int)rank / FC_NET, j = (int)rank % FC_NET;
    const int row0 = (blockIdx.x / FC_CLUSTER) * a.rows_per_cluster;
    const int R = a.B - row0 < a.rows_per_cluster ? a.B - row0 : a.rows_per_cluster;
    const size_t coupling_floats = (size_t)l1_floats + (size_t)a.depth * FC_WARPS * hc_floats + l4_floats;
    const int c1_stride = a.n_flows * 2 * 2 * H;

    if (tid == 0) {
        for (int s = 0; s < a.ring; ++s) { ptx::mbar_init(full + s, 1); ptx::mbar_init(empty + s, 1); }
        ptx::mbar_init(act_ready, FC_NET); ptx::mbar_init(act_ready + 1, FC_NET);
        ptx::mbar_init(st_ready, FC_CLUSTER);
        ptx::fence_barrier_init();
    }
    for (int i = tid; i < 2 * ROWS * H; i += FC_THREADS) act0[i] = 0.f;         // rows >= R stay zero
    for (int i = tid; i < ROWS * FC_D; i += FC_THREADS) { xs[i] = 0.f; stb[i] = 0.f; }
    if (tid < 16) lds[tid] = 0.f;
    ptx::cluster_sync();          // every CTA's barriers and buffers exist before any push / remote arrive
    pdl_launch_dependents();
    pdl_wait();                   // c1 comes from the launch before this one

    if (warp == FC_WARPS) {
        // ================================ weight streamer: data-independent, runs ahead as far as the ring allows
        int slot = 0;
        uint32_t ph = 0;
        auto stream = [&](const float* src, int floats) {
            ptx::mbar_wait(empty + slot, ph ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(full + slot, (uint32_t)floats * 4u);
                bulk_g2s(ring + (size_t)slot * slot_floats, src, (uint32_t)floats * 4u, full + slot);
            }
            __syncwarp();
            if (++slot == a.ring) { slot = 0; ph ^= 1u; }
        };
        for (int step = 0; step < a.n_flows; ++step) {
            const int fl = a.reverse ? a.n_flows - 1 - step : step;
            for (int ci = 0; ci < 2; ++ci) {
                const int i = a.reverse ? 1 - ci : ci;
                const float* src = a.wpack + ((size_t)(fl * 2 + i) * FC_CLUSTER + rank) * coupling_floats;
                stream(src, l1_floats);
                src += l1_floats;
                for (int c = 0; c < a.depth * FC_WARPS; ++c, src += hc_floats) stream(src, hc_floats);
                stream(src, l4_floats);
            }
        }
    } else {
        // ================================ compute warps
        for (int i = tid; i < R * FC_D; i += FC_CTHREADS) xs[i] = __ldg(a.in + (size_t)row0 * FC_D + i);
        cbar();
        int slot = 0;             // ring position of the next chunk (same walk as the streamer: every warp steps over every chunk)
        uint32_t rph = 0;         // ... and its phase
        int lay = 0;              // exchanges so far: buffer = lay & 1, phase = (lay >> 1) & 1
        int cpl_n = 0;            // couplings so far (phase of st_ready)
        const uint32_t my_act0 = ptx::smem_u32(act0), my_act1 = ptx::smem_u32(act1), my_stb = ptx::smem_u32(stb);
        auto chunk_ptr = [&]() { return ring + (size_t)slot * slot_floats; };
        auto chunk_wait = [&]() { ptx::mbar_wait(full + slot, rph); };
        // step to the next chunk; `owner`: this warp hands the ring buffer back to the streamer (ONE arrival per chunk, by the
        // warp that read it last)
        auto chunk_next = [&](bool owner) {
            if (owner) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(empty + slot); }
            if (++slot == a.ring) { slot = 0; rph ^= 1u; }
        };
        // Global-memory operands of the dependency chain (the hoisted conditioning term c1, the biases) are fetched one step
        // early into registers: an L2 round trip (~0.6 us) per layer would otherwise sit on the critical path.
        auto cidx_of = [&](int cq_) {      // coupling sequence number -> coupling index in the weight tensors
            const int st = cq_ >> 1, c_ = cq_ & 1;
            const int fl_ = a.reverse ? a.n_flows - 1 - st : st;
            return fl_ * 2 + (a.reverse ? 1 - c_ : c_);
        };
        float c1_next[CPL];
        auto load_c1 = [&](int cq_) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) c1_next[c] = 0.f;
            if (warp < R && cq_ < 2 * a.n_flows) {
                const float* cp = a.c1 + (size_t)(row0 + warp) * c1_stride + (size_t)cidx_of(cq_) * 2 * H + net * H + j * Cc + lane * CPL;
#pragma unroll
                for (int c = 0; c < CPL; ++c) c1_next[c] = __ldg(cp + c);
            }
        };
        load_c1(0);

        for (int step = 0; step < a.n_flows; ++step) {
            const int fl = a.reverse ? a.n_flows - 1 - step : step;
            const float* loc = a.loc + fl * FC_D;
            const float* scale = a.scale + fl * FC_D;
            const bool cmode = a.cond_mode[fl] != 0;
            if (a.reverse) {
                const int* perm = a.perm_bwd + fl * FC_D;                 // Shuffle^-1
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) xt[i] = xs[(i / FC_D) * FC_D + __ldg(perm + (i % FC_D))];
                cbar();
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) xs[i] = xt[i];
                cbar();
            } else {
                // ActNorm: h = scale * (x + loc); logdet += sum log|scale|; InvLeakyRelu: h *= (h >= 0 ? 1 : 0.9)
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) {
                    const int c = i % FC_D;
                    float v = __ldg(scale + c) * (xs[i] + __ldg(loc + c));
                    xs[i] = v * (v >= 0.f ? 1.f : 0.9f);
                }
                if (warp == 0) {
                    float s = 0.f;
                    for (int c = lane; c < FC_D; c += 32) s += logf(fabsf(__ldg(scale + c)));
                    s = warp_sum(s);
                    if (lane < ROWS) lds[lane] += s;
                }
                cbar();
            }
            for (int ci = 0; ci < 2; ++ci) {
                const int i = a.reverse ? 1 - ci : ci;
                const bool swap_first = a.reverse ? (i % 2 == 0) : (i % 2 != 0);
                if (swap_first) {
                    for (int e = tid; e < ROWS * FC_HALF; e += FC_CTHREADS) {
                        const int b = e / FC_HALF, c = e % FC_HALF;
                        const float lo = xs[b * FC_D + c], hi = xs[b * FC_D + FC_HALF + c];
                        xs[b * FC_D + c] = hi; xs[b * FC_D + FC_HALF + c] = lo;
                    }
                    cbar();
                }
                const int cidx = fl * 2 + i;
                const int cq = step * 2 + ci;      // coupling sequence number (profiling)
                fdbg(cq, 0, rank);
                // ---- layer 1: h1 = lrelu(W1x . x[:, :32] + c1): warp w = row w, lanes = this CTA's columns
                {
                    chunk_wait();
                    fdbg(cq, 1, rank);
                    const float* wc = chunk_ptr();
                    const int col = lane * CPL;
                    float acc[CPL];
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[c] = c1_next[c];
                    load_c1(cq + 1);                 // the next coupling's term, in flight during this coupling
                    if (warp < R) {
                        if (!cmode) {
                            const float* xr = xs + warp * FC_D;
#pragma unroll 8
                            for (int k = 0; k < FC_HALF; ++k) {
                                const float xk = xr[k];
#pragma unroll
                                for (int c = 0; c < CPL; ++c) acc[c] = fmaf(wc[k * Cc + col + c], xk, acc[c]);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[c] = lrelu001(acc[c]);
                        const uint32_t dst = ((lay & 1) ? my_act1 : my_act0) + (uint32_t)(warp * H + j * Cc + col) * 4u;
#pragma unroll
                        for (int p = 0; p < FC_NET; ++p) {
                            const uint32_t ra = ptx::mapa_u32(dst, (uint32_t)(net * FC_NET + p));
                            if (CPL == 2) st_cluster_v2(ra, acc[0], acc[CPL - 1]);
                            else st_cluster_f32(ra, acc[0]);
                        }
                    }
                    fdbg(cq, 2, rank);
                    exchange(act_ready + (lay & 1), (uint32_t)((lay >> 1) & 1), net * FC_NET, FC_NET);      // (its cbar: all warps are
                    chunk_next(warp == 0);                                                                   //  done with the chunk)
                    ++lay;
                    fdbg(cq, 3, rank);
                }
                // ---- hidden layers: K split over the warps, warp w owns chunk w = k-rows [w KC, (w + 1) KC)
                for (int l = 0; l < a.depth; ++l) {
                    const float* hin = ((lay - 1) & 1) ? act1 : act0;
                    float acc[ROWS][CPL];
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[r][c] = 0.f;
                    // this thread's bias for the reduction below (one output per thread, two when ROWS * Cc > 512)
                    const float* bias = a.bh + ((size_t)cidx * a.depth + l) * 2 * H + net * H + j * Cc;
                    const float bias0 = __ldg(bias + (tid & (Cc - 1)));
                    for (int c = 0; c < warp; ++c) chunk_next(false);
                    {
                        chunk_wait();
                        const float* wc = chunk_ptr() + lane * CPL;
                        const float* hk = hin + warp * KC;
#pragma unroll 2
                        for (int kk = 0; kk < KC; kk += 4) {
                            float w[4][CPL];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (CPL == 2) {
                                    const float2 w2 = *reinterpret_cast<const float2*>(wc + (kk + q) * Cc);
                                    w[q][0] = w2.x; w[q][CPL - 1] = w2.y;
                                } else {
                                    w[q][0] = wc[(kk + q) * Cc];
                                }
                            }
#pragma unroll
                            for (int r = 0; r < ROWS; ++r) {
                                const float4 h4 = *reinterpret_cast<const float4*>(hk + r * H + kk);
#pragma unroll
                                for (int cc = 0; cc < CPL; ++cc) {
                                    acc[r][cc] = fmaf(w[0][cc], h4.x, acc[r][cc]); acc[r][cc] = fmaf(w[1][cc], h4.y, acc[r][cc]);
                                    acc[r][cc] = fmaf(w[2][cc], h4.z, acc[r][cc]); acc[r][cc] = fmaf(w[3][cc], h4.w, acc[r][cc]);
                                }
                            }
                        }
                        chunk_next(true);
                    }
                    for (int c = warp + 1; c < FC_WARPS; ++c) chunk_next(false);
                    fdbg(cq, 4 + 4 * l, rank);
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
#pragma unroll
                        for (int cc = 0; cc < CPL; ++cc) part[(warp * ROWS + r) * Cc + lane * CPL + cc] = acc[r][cc];
                    cbar();
                    fdbg(cq, 5 + 4 * l, rank);
                    // reduce the 16 partial sums, bias, LeakyReLU, push this CTA's [rows][Cc] slice to the 8 CTAs of its net
                    const uint32_t dbase = (lay & 1) ? my_act1 : my_act0;
                    for (int o = tid; o < ROWS * Cc; o += FC_CTHREADS) {
                        const int r = o >> cc_shift, col = o & (Cc - 1);       // Cc is a power of two (32 / 64); col == tid & (Cc - 1)
                        float v = 0.f;
#pragma unroll
                        for (int w = 0; w < FC_WARPS; ++w) v += part[(w * ROWS + r) * Cc + col];
                        v = lrelu001(v + bias0);
                        if (r < R) {
                            const uint32_t dst = dbase + (uint32_t)(r * H + j * Cc + col) * 4u;
#pragma unroll
                            for (int p = 0; p < FC_NET; ++p) st_cluster_f32(ptx::mapa_u32(dst, (uint32_t)(net * FC_NET + p)), v);
                        }
                    }
                    fdbg(cq, 6 + 4 * l, rank);
                    exchange(act_ready + (lay & 1), (uint32_t)((lay >> 1) & 1), net * FC_NET, FC_NET);
                    ++lay;
                    fdbg(cq, 7 + 4 * l, rank);
                }
                // ---- last layer: this CTA computes 4 of its net's 32 outputs; chunk layout [k = H][4]; warp w takes k-rows
                // [w KC, (w + 1) KC), one per lane
                {
                    const float* hin = ((lay - 1) & 1) ? act1 : act0;
                    const float bo0 = tid < ROWS * 4 ? __ldg(a.bo + (size_t)cidx * FC_D + net * FC_HALF + 4 * j + (tid & 3)) : 0.f;
                    chunk_wait();
                    const float4* wc = reinterpret_cast<const float4*>(chunk_ptr());
                    constexpr int NACC = ROWS * 4 < 32 ? 32 : ROWS * 4;
                    float acc[NACC];          // [row][4 outputs] (padded to the 32 values of the transpose-reduce)
#pragma unroll
                    for (int q = 0; q < NACC; ++q) acc[q] = 0.f;
                    if (lane < KC) {
                        const int k = warp * KC + lane;
                        const float4 w4 = wc[k];
#pragma unroll
                        for (int r = 0; r < ROWS; ++r) {
                            const float h = hin[r * H + k];
                            acc[r * 4] = w4.x * h; acc[r * 4 + 1] = w4.y * h; acc[r * 4 + 2] = w4.z * h; acc[r * 4 + 3] = w4.w * h;
                        }
                    }
                    // sum over the lanes: the first 32 (row, output) pairs by a transpose-reduce (lane L ends with pair L), the
                    // rest (ROWS > 8) by plain butterflies
#pragma unroll
                    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int q = 0; q < n; ++q) {
                            const float send = up ? acc[q] : acc[q + n];
                            const float keep = up ? acc[q + n] : acc[q];
                            acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    if (lane < ROWS * 4) part[warp * (ROWS * 4) + lane] = acc[0];
#pragma unroll
                    for (int q = 32; q < ROWS * 4; ++q) {
                        const float v = warp_sum(acc[q]);
                        if (lane == 0) part[warp * (ROWS * 4) + q] = v;
                    }
                    cbar();
                    chunk_next(warp == 0);
                    if (tid < ROWS * 4) {
                        float v = 0.f;
#pragma unroll
                        for (int w = 0; w < FC_WARPS; ++w) v += part[w * (ROWS * 4) + tid];
                        const int r = tid >> 2, c = tid & 3;
                        v += bo0;
                        if (r < R) {
                            const uint32_t dst = my_stb + (uint32_t)(r * FC_D + net * FC_HALF + 4 * j + c) * 4u;
#pragma unroll
                            for (int p = 0; p < FC_CLUSTER; ++p) st_cluster_f32(ptx::mapa_u32(dst, (uint32_t)p), v);
                        }
                    }
                    fdbg(cq, 12, rank);
                    exchange(st_ready, (uint32_t)(cpl_n & 1), 0, FC_CLUSTER);
                    ++cpl_n;
                    fdbg(cq, 13, rank);
                }
                // ---- affine update of the kept half (every CTA, on its private copy of the state)
                for (int e = tid; e < ROWS * FC_HALF; e += FC_CTHREADS) {
                    const int b = e / FC_HALF, c = e % FC_HALF;
                    const float s = stb[b * FC_D + c], t = stb[b * FC_D + FC_HALF + c];
                    const float xk = xs[b * FC_D + FC_HALF + c];
                    xs[b * FC_D + FC_HALF + c] = a.reverse ? (xk - t) * expf(-s) : fmaf(xk, expf(s), t);
                }
                if (!a.reverse && warp < ROWS) {
                    const float s = warp_sum(stb[warp * FC_D + lane]);          // sum of the 32 scale outputs of row `warp`
                    if (lane == 0) lds[warp] += s;
                }
                cbar();
                fdbg(cq, 14, rank);
            }
            if (a.reverse) {
                // InvLeakyRelu^-1: h / (h >= 0 ? 1 : 0.9); ActNorm^-1: h / scale - loc   (true divisions, like the reference)
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) {
                    const int c = i % FC_D;
                    float v = xs[i];
                    v = v / (v >= 0.f ? 1.f : 0.9f);
                    xs[i] = v / __ldg(scale + c) - __ldg(loc + c);
                }
                cbar();
            } else {
                const int* perm = a.perm_fwd + fl * FC_D;
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) xt[i] = xs[(i / FC_D) * FC_D + __ldg(perm + (i % FC_D))];
                cbar();
                for (int i = tid; i < ROWS * FC_D; i += FC_CTHREADS) xs[i] = xt[i];
                cbar();
            }
        }
        if (rank == 0) {
            for (int i = tid; i < R * FC_D; i += FC_CTHREADS) a.out[(size_t)row0 * FC_D + i] = xs[i];
            if (a.logdet != nullptr && tid < R) a.logdet[row0 + tid] = lds[tid];
        }
    }
    ptx::cluster_sync();          // nobody leaves while a peer could still push into its shared memory
}

size_t fc_smem_bytes(int H, int rows, int ring) {
    const size_t Cc = H / FC_NET;
    const size_t slot = (FC_HALF * Cc > 4 * (size_t)H ? FC_HALF * Cc : 4 * (size_t)H) * 4;
    return (size_t)ring * slot + 2 * (size_t)rows * H * 4 + (size_t)FC_WARPS * rows * Cc * 4 + 3 * (size_t)rows * FC_D * 4 + 16 * 4 +
           (2 * FC_MAX_RING + 3) * 8 + 64;
}

}  // namespace

int flow_cluster_set_debug(unsigned long long* buf) {
    I2V_CHECK_CUDA(cudaMemcpyToSymbol(g_fc_dbg, &buf, sizeof(buf)));
    return 0;
}

bool flow_cluster_eligible(const FlowWeights& fw) {
    return tune().flow_cluster && fw.wpack != nullptr && fw.d == FC_D && (fw.hidden == 512 || fw.hidden == 256) && fw.depth >= 0 &&
           fw.n_flows <= 64;
}

// Returns 1 when the device cannot co-schedule a 16-CTA cluster of this kernel (caller falls back to the cooperative kernel).
int launch_flow_cluster(const FlowWeights& fw, const float* in, const float* c1, float* out, float* logdet, int B, bool reverse,
                        cudaStream_t stream) {
    const int H = fw.hidden;
    // 8 rows per cluster while the batch fits the clusters the device co-schedules, 10 beyond (64 rows = 7 clusters, one wave);
    // the rows are split evenly over whole waves of clusters, and a batch of <= 7 rows (the scripts' -bs 6) runs the
    // one-row instantiation: one row per cluster
    int rows = B <= 8 * FC_MAX_CLUSTERS ? 8 : 10;
    int rpc = rows;
    {
        const int waves = (B + rows * FC_MAX_CLUSTERS - 1) / (rows * FC_MAX_CLUSTERS);
        const int clusters_wanted = waves * FC_MAX_CLUSTERS;
        rpc = (B + clusters_wanted - 1) / clusters_wanted;          // even split over whole waves
        if (rpc < 1) rpc = 1;
    }
    if (rpc == 1) rows = 1;
    int ring = FC_MAX_RING;
    while (ring > 4 && fc_smem_bytes(H, rows, ring) > 226 * 1024) --ring;
    const size_t smem = fc_smem_bytes(H, rows, ring);
    void (*kernel)(const FlowClusterArgs) =
        H == 512 ? (rows == 1 ? flow_cluster_kernel<2, 1> : rows == 8 ? flow_cluster_kernel<2, 8> : flow_cluster_kernel<2, 10>)
                 : (rows == 1 ? flow_cluster_kernel<1, 1> : rows == 8 ? flow_cluster_kernel<1, 8> : flow_cluster_kernel<1, 10>);
    const int vi = (H == 512 ? 0 : 3) + (rows == 1 ? 0 : rows == 8 ? 1 : 2);
    static unsigned long long attr_devs[6] = {0, 0, 0, 0, 0, 0};
    int dev = 0;
    I2V_CHECK_CUDA(cudaGetDevice(&dev));
    static int cluster_ok[64][6];      // per device / variant: 0 unknown, 1 yes, -1 no
    int& ok = cluster_ok[dev < 64 ? dev : 63][vi];
    if (!(dev < 64 && ((attr_devs[vi] >> dev) & 1ull))) {
        I2V_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        I2V_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        if (dev < 64) attr_devs[vi] |= 1ull << dev;
    }
    const int clusters = (B + rpc - 1) / rpc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * FC_CLUSTER)); cfg.blockDim = dim3(FC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FC_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (ok == 0) {
        int n = 0;
        const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
        ok = (e == cudaSuccess && n >= 1) ? 1 : -1;
        if (e != cudaSuccess) (void)cudaGetLastError();
    }
    if (ok < 0) return 1;
    FlowClusterArgs ka;
    ka.wpack = fw.wpack; ka.c1 = c1; ka.bh = fw.bh; ka.bo = fw.bo; ka.loc = fw.loc; ka.scale = fw.scale;
    ka.perm_fwd = fw.perm_fwd; ka.perm_bwd = fw.perm_bwd; ka.in = in; ka.out = out; ka.logdet = logdet;
    ka.B = B; ka.n_flows = fw.n_flows; ka.H = H; ka.depth = fw.depth; ka.reverse = reverse ? 1 : 0; ka.ring = ring;
    ka.rows_per_cluster = rpc;
    for (int i = 0; i < 64; ++i) ka.cond_mode[i] = (i < fw.n_flows && fw.cond_mode) ? fw.cond_mode[i] : 0;
    const double wbytes = 4.0 * fw.n_flows * 2 * ((double)2 * H * fw.half + (double)fw.depth * 2 * H * H + (double)2 * fw.half * H);
    ProfScope ps(PROF_FLOW, 2.0 * B * wbytes / 4.0, wbytes + 4.0 * B * (2.0 * fw.d + (double)fw.n_flows * 2 * 2 * H), stream);
    I2V_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, ka));
    return 0;
}

}  // namespace i2v
