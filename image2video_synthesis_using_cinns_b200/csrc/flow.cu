// Conditional INN (stage2_cINN) as ONE persistent cooperative kernel per direction.
//
// Reference semantics (file:line in /root/reference):
//   ConditionalFlow.forward                      stage2_cINN/modules/flow_blocks.py:31-57
//   ConditionalFlatDoubleCouplingFlowBlock       flow_blocks.py:117-136   (ActNorm, InvLeakyRelu, coupling, Shuffle)
//   ConditionalDoubleVectorCouplingBlock         flow_blocks.py:77-105    (x1 = x1*exp(s)+t / (x1-t)*exp(-s))
//   BasicFullyConnectedNet                       stage2_cINN/modules/modules.py:9-30 (LeakyReLU slope 0.01)
//   ActNorm                                      modules.py:80,86-88,100
//   InvLeakyRelu (alpha 0.9, logdet reported 0)  flow_blocks.py:172-187
//   Shuffle                                      flow_blocks.py:142-154
//
// Design.  The chain is 2*n_flows couplings x (depth+2) strictly dependent Linear layers whose
// weights (189-200 MB fp32) are each used once per pass: the work is weight streaming plus
// B x 2H x H fp32 FMAs per hidden layer.  One CTA per SM stays resident for the whole pass:
//   * the latent state x[B, d] and the running log-det live in every CTA's shared memory (each CTA
//     applies the cheap element-wise steps -- permutation, swap, affine update, InvLeakyReLU,
//     ActNorm -- redundantly, so the state is never exchanged);
//   * every Linear layer is split by OUTPUT feature across the grid (first half of the CTAs owns the
//     scale net, second half the translation net); a CTA stages its net's input activations
//     [B, H] in shared memory once per layer, one warp per output feature streams that feature's
//     weight row from HBM exactly once (coalesced float4) and reduces with warp shuffles;
//   * layers are separated by a grid-wide barrier (cooperative launch guarantees co-residency);
//   * the conditioning half of every first Linear, W1[:, half:] * cond + b1, does not depend on the
//     state (flow_blocks.py:33-41 feeds the same embedding to every block) and is hoisted out of the
//     sequential chain into one batched GEMM ahead of the kernel (launch_linear).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace i2v {

namespace {

constexpr int FLOW_THREADS = 256;
constexpr int FLOW_WARPS = FLOW_THREADS / 32;
constexpr int FLOW_MAX_ROWS = 64;   // batch rows per launch (shared-memory budget: B*H*4 + 2*B*d*4)

struct FlowKernelArgs {
    FlowWeights w;
    const float* in;      // [B, d]
    const float* c1;      // [B, n_flows*2*2H]  hoisted conditioning part of the first Linear (+ bias)
    float* hbuf0;         // [B, 2H]
    float* hbuf1;         // [B, 2H]
    float* st;            // [B, 2*half]        (s | t) of the current coupling
    float* out;           // [B, d]
    float* logdet;        // [B] or nullptr
    unsigned* bar;        // [2] grid barrier (count, generation), zeroed before launch
    int B;
    int reverse;
    unsigned char cond_mode[64];   // n_flows <= 64
};

// optional phase timestamps (i2v_debug_flow_timestamps): 16 x u64 for coupling #4 of CTAs 0 and 100
__device__ unsigned long long* g_flow_dbg = nullptr;
__device__ __forceinline__ void fdbg(int coupling, int slot) {
    if (g_flow_dbg != nullptr && coupling == 4 && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 100)) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_flow_dbg[(blockIdx.x == 0 ? 0 : 16) + slot] = t;
    }
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Sense-reversing grid barrier.  All CTAs are co-resident (cooperative launch).
__device__ __forceinline__ unsigned atom_add_release(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// Release/acquire instead of full fences: the data exchanged across the barrier is only ever re-read with
// ld.global.cg (L2), so no L1 invalidation (CCTL.IVALL) is needed; bar.sync gives CTA-level cumulativity.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = ld_acquire(bar + 1);
        const unsigned prev = atom_add_release(bar, 1u);
        if (prev == nblocks - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
            atom_add_release(bar + 1, 1u);
        } else {
            while (ld_acquire(bar + 1) == gen) {
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float lrelu001(float v) { return v >= 0.f ? v : 0.01f * v; }

// One hidden Linear (+LeakyReLU) of both nets: hout[b, net*H + o] = lrelu(W[net][o,:] . hin[b, net*H:] + bias).
// The CTA belongs to one net; it stages that net's input rows in shared memory.
// Weight rows of ALL output features this CTA owns in a hidden layer (<= 8: ctas_per_net * 8 >= H), one copy per
// lane-slice in registers.  They are fetched BEFORE the grid barrier that precedes the layer: weights do not
// depend on the data, so their HBM latency hides behind the barrier and the activation staging.
constexpr int FLOW_MAXC = 8;
struct WCols { float4 v[FLOW_MAXC][4]; float bias[FLOW_MAXC]; };

__device__ __forceinline__ void load_wcols(WCols& w, const float* __restrict__ W, const float* __restrict__ bias, int H, int net,
                                           int cta_in_net, int ctas_per_net, int j0) {
    const int H4 = H >> 2, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < FLOW_MAXC; ++c) {
        const int o = cta_in_net + ctas_per_net * (j0 + c);
        const float4* wr = reinterpret_cast<const float4*>(W + ((long long)net * H + (o < H ? o : 0)) * H);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = lane + 32 * q;
            w.v[c][q] = (o < H && k < H4) ? __ldg(wr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        w.bias[c] = o < H ? __ldg(bias + net * H + o) : 0.f;
    }
}

// One hidden Linear (+LeakyReLU) of both nets: hout[b, net*H + o] = lrelu(W[net][o,:] . hin[b, net*H:] + bias).
// The CTA belongs to one net and stages that net's input rows in shared memory.  Work split: every warp takes
// a subset of ROWS and evaluates all of the CTA's output features for them, so an activation row is read from
// shared memory once per 8 features (one-feature-per-warp re-read it per feature and was bound by shared-memory
// bandwidth: measured 14.7 us per layer at B = 64).
__device__ void hidden_layer(const float* __restrict__ W /*[2][H][H]*/, const float* __restrict__ bias /*[2H]*/,
                             const float* hin, float* hout, float* hs /*smem [B][H]*/, int B, int H, int net,
                             int cta_in_net, int ctas_per_net, const WCols& w0) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H4 = H >> 2;
    {
        // stage hin[:, net*H : (net+1)*H] (written by other CTAs before the barrier -> read through L2); eight
        // independent 16-byte loads in flight per thread, each CTA starting at a different row
        const int total = B * H4;
        const int rot = (int)(((long long)cta_in_net * total) / ctas_per_net) / H4 * H4;
        for (int i0 = tid; i0 < total; i0 += FLOW_THREADS * 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int i = i0 + j * FLOW_THREADS;
                if (i < total) {
                    i += rot; if (i >= total) i -= total;
                    const int b = i / H4, k = i - b * H4;
                    v[j] = __ldcg(reinterpret_cast<const float4*>(hin + (long long)b * 2 * H + net * H) + k);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int i = i0 + j * FLOW_THREADS;
                if (i < total) {
                    i += rot; if (i >= total) i -= total;
                    reinterpret_cast<float4*>(hs)[i] = v[j];
                }
            }
        }
    }
    __syncthreads();
    {
        const int j0 = 0;              // launch_flow guarantees ctas_per_net * FLOW_MAXC >= H: one batch of features
        const WCols& w = w0;
        (void)W; (void)bias;
        for (int b = warp; b < B; b += FLOW_WARPS) {
            const float4* hr = reinterpret_cast<const float4*>(hs + (long long)b * H);
            float4 hv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                hv[q] = k < H4 ? hr[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float acc[FLOW_MAXC];
#pragma unroll
            for (int c = 0; c < FLOW_MAXC; ++c) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    s = fmaf(w.v[c][q].x, hv[q].x, s); s = fmaf(w.v[c][q].y, hv[q].y, s);
                    s = fmaf(w.v[c][q].z, hv[q].z, s); s = fmaf(w.v[c][q].w, hv[q].w, s);
                }
                acc[c] = s;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int c = 0; c < FLOW_MAXC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
            }
            if (lane < FLOW_MAXC) {
                float v = acc[0], bo = w.bias[0];
#pragma unroll
                for (int c = 1; c < FLOW_MAXC; ++c)
                    if (lane == c) { v = acc[c]; bo = w.bias[c]; }
                const int o = cta_in_net + ctas_per_net * (j0 + lane);
                if (o < H) hout[(long long)b * 2 * H + net * H + o] = lrelu001(v + bo);
            }
        }
    }
}

__global__ void __launch_bounds__(FLOW_THREADS, 1) flow_kernel(const FlowKernelArgs a) {
    extern __shared__ __align__(16) float smem[];
    const FlowWeights& fw = a.w;
    const int B = a.B, d = fw.d, half = fw.half, H = fw.hidden, depth = fw.depth;
    float* xs = smem;                      // [B][d] state
    float* xt = xs + B * d;                // [B][d] scratch for permutations
    float* lds = xt + B * d;               // [B]    running log-det
    float* hs = lds + ((B + 3) & ~3);      // [B][H] staged activations

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int ctas_per_net = G >> 1;
    const int net = cta < ctas_per_net ? 0 : 1;
    const int cta_in_net = cta - net * ctas_per_net;
    const bool active_net_cta = cta < 2 * ctas_per_net;   // odd grid: last CTA idles in hidden layers
    const int c1_stride = fw.n_flows * 2 * 2 * H;

    for (int i = tid; i < B * d; i += FLOW_THREADS) xs[i] = __ldg(a.in + i);
    for (int i = tid; i < B; i += FLOW_THREADS) lds[i] = 0.f;
    __syncthreads();

    for (int step = 0; step < fw.n_flows; ++step) {
        const int fl = a.reverse ? fw.n_flows - 1 - step : step;
        const float* loc = fw.loc + fl * d;
        const float* scale = fw.scale + fl * d;
        const bool cmode = a.cond_mode[fl] != 0;

        if (a.reverse) {
            // Shuffle^-1: x = x[:, backward_idx]
            const int* perm = fw.perm_bwd + fl * d;
            for (int i = tid; i < B * d; i += FLOW_THREADS) {
                const int b = i / d, c = i - b * d;
                xt[i] = xs[b * d + __ldg(perm + c)];
            }
            __syncthreads();
            for (int i = tid; i < B * d; i += FLOW_THREADS) xs[i] = xt[i];
            __syncthreads();
        } else {
            // ActNorm: h = scale*(x+loc) ; logdet += sum log|scale| ; InvLeakyRelu: h *= (h>=0 ? 1 : 0.9)
            for (int i = tid; i < B * d; i += FLOW_THREADS) {
                const int c = i % d;
                float v = __ldg(scale + c) * (xs[i] + __ldg(loc + c));
                v = v * (v >= 0.f ? 1.f : 0.9f);
                xs[i] = v;
            }
            if (tid < 32) {
                float s = 0.f;
                for (int c = lane; c < d; c += 32) s += logf(fabsf(__ldg(scale + c)));
                s = warp_sum(s);
                for (int b = lane; b < B; b += 32) lds[b] += s;
            }
            __syncthreads();
        }

        for (int ci = 0; ci < 2; ++ci) {
            const int i = a.reverse ? 1 - ci : ci;     // reverse visits coupling 1 then 0
            const bool swap_first = a.reverse ? (i % 2 == 0) : (i % 2 != 0);
            if (swap_first) {   // x = cat(x[:, half:], x[:, :half])
                for (int e = tid; e < B * half; e += FLOW_THREADS) {
                    const int b = e / half, c = e - b * half;
                    const float lo = xs[b * d + c], hi = xs[b * d + half + c];
                    xs[b * d + c] = hi; xs[b * d + half + c] = lo;
                }
                __syncthreads();
            }
            const int cidx = fl * 2 + i;
            const int cq = step * 2 + ci;      // coupling sequence number (profiling)
            fdbg(cq, 0);
            // ---- layer 1: h1 = lrelu(W1x . x[:, :half] + c1)
            {
                const float* w1x = fw.w1x + (long long)cidx * 2 * H * half;
                const float* c1 = a.c1 + (long long)cidx * 2 * H;
                // columns n = cta, cta+G, ... of the 2H outputs
                const int ncols = (2 * H - cta + G - 1) / G;
                for (int e = tid; e < ncols * B; e += FLOW_THREADS) {
                    // column index fastest: lanes of a warp share the row b (shared-memory broadcast of the state;
                    // with b fastest the row stride of 64 floats is a 32-way bank conflict)
                    const int b = e / ncols, n = cta + G * (e - b * ncols);
                    float acc = __ldg(c1 + (long long)b * c1_stride + n);
                    if (!cmode) {
                        const float* wr = w1x + (long long)n * half;
                        const float* xr = xs + b * d;
                        for (int k = 0; k < half; ++k) acc = fmaf(__ldg(wr + k), xr[k], acc);
                    }
                    a.hbuf0[(long long)b * 2 * H + n] = lrelu001(acc);
                }
            }
            fdbg(cq, 1);
            WCols wnext;
            if (active_net_cta && depth > 0)
                load_wcols(wnext, fw.wh + ((long long)cidx * depth) * 2 * H * H, fw.bh + ((long long)cidx * depth) * 2 * H, H, net,
                           cta_in_net, ctas_per_net, 0);
            grid_barrier(a.bar, G);
            fdbg(cq, 2);
            // ---- hidden layers
            const float* hin = a.hbuf0;
            float* hout = a.hbuf1;
            for (int l = 0; l < depth; ++l) {
                if (active_net_cta)
                    hidden_layer(fw.wh + ((long long)cidx * depth + l) * 2 * H * H,
                                 fw.bh + ((long long)cidx * depth + l) * 2 * H, hin, hout, hs, B, H, net, cta_in_net,
                                 ctas_per_net, wnext);
                // prefetch the next hidden layer's weight rows before waiting at the barrier
                if (active_net_cta && l + 1 < depth)
                    load_wcols(wnext, fw.wh + ((long long)cidx * depth + l + 1) * 2 * H * H,
                               fw.bh + ((long long)cidx * depth + l + 1) * 2 * H, H, net, cta_in_net, ctas_per_net, 0);
                fdbg(cq, 3 + 2 * l);
                grid_barrier(a.bar, G);
                fdbg(cq, 4 + 2 * l);
                const float* t = hin; hin = hout; hout = const_cast<float*>(t);
            }
            // ---- last layer: (s | t)[b, r] for r in [0, 2*half).  Only 2*half = 64 outputs exist, so the work is
            // spread over (output, 4-row group) tasks across every warp of the grid (one warp per output serialised
            // all B rows: measured 39 us of an 89 us coupling at B = 64).
            {
                const float* wo = fw.wo + (long long)cidx * 2 * half * H;
                const float* bo = fw.bo + (long long)cidx * 2 * half;
                const int H4 = H >> 2;
                const int R = 2 * half, RG = (B + 3) >> 2;
                for (int task = cta * FLOW_WARPS + warp; task < R * RG; task += G * FLOW_WARPS) {
                    const int r = task % R, b = (task / R) * 4;
                    const int rnet = r < half ? 0 : 1;
                    const float4* wr = reinterpret_cast<const float4*>(wo + (long long)r * H);
                    float4 wv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = lane + 32 * q;
                        wv[q] = k < H4 ? __ldg(wr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float br = __ldg(bo + r);
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        if (b + rr < B) {
                            const float4* hr = reinterpret_cast<const float4*>(hin + (long long)(b + rr) * 2 * H + rnet * H);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int k = lane + 32 * q;
                                if (k < H4) {
                                    const float4 hv = __ldcg(hr + k);
                                    acc[rr] = fmaf(wv[q].x, hv.x, acc[rr]); acc[rr] = fmaf(wv[q].y, hv.y, acc[rr]);
                                    acc[rr] = fmaf(wv[q].z, hv.z, acc[rr]); acc[rr] = fmaf(wv[q].w, hv.w, acc[rr]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) acc[rr] += __shfl_xor_sync(0xffffffffu, acc[rr], off);
                    }
                    if (lane < 4 && b + lane < B) {
                        const float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
                        a.st[(long long)(b + lane) * 2 * half + r] = v + br;
                    }
                }
            }
            fdbg(cq, 7);
            grid_barrier(a.bar, G);
            fdbg(cq, 8);
            // ---- affine update of the kept half (every CTA, on its private copy)
            for (int e = tid; e < B * half; e += FLOW_THREADS) {
                const int b = e / half, c = e - b * half;
                const float s = __ldcg(a.st + (long long)b * 2 * half + c);
                const float t = __ldcg(a.st + (long long)b * 2 * half + half + c);
                const float xk = xs[b * d + half + c];
                xs[b * d + half + c] = a.reverse ? (xk - t) * expf(-s) : fmaf(xk, expf(s), t);
            }
            if (!a.reverse) {
                __syncthreads();
                for (int b = warp; b < B; b += FLOW_WARPS) {
                    float s = 0.f;
                    for (int c = lane; c < half; c += 32) s += __ldcg(a.st + (long long)b * 2 * half + c);
                    s = warp_sum(s);
                    if (lane == 0) lds[b] += s;
                }
            }
            __syncthreads();
            fdbg(cq, 9);
        }

        if (a.reverse) {
            // InvLeakyRelu^-1: h / (h>=0 ? 1 : 0.9) ; ActNorm^-1: h/scale - loc   (true divisions)
            for (int i = tid; i < B * d; i += FLOW_THREADS) {
                const int c = i % d;
                float v = xs[i];
                v = v / (v >= 0.f ? 1.f : 0.9f);
                xs[i] = v / __ldg(scale + c) - __ldg(loc + c);
            }
            __syncthreads();
        } else {
            const int* perm = fw.perm_fwd + fl * d;
            for (int i = tid; i < B * d; i += FLOW_THREADS) {
                const int b = i / d, c = i - b * d;
                xt[i] = xs[b * d + __ldg(perm + c)];
            }
            __syncthreads();
            for (int i = tid; i < B * d; i += FLOW_THREADS) xs[i] = xt[i];
            __syncthreads();
        }
    }

    if (cta == 0) {
        for (int i = tid; i < B * d; i += FLOW_THREADS) a.out[i] = xs[i];
        if (a.logdet != nullptr)
            for (int i = tid; i < B; i += FLOW_THREADS) a.logdet[i] = lds[i];
    }
}

}  // namespace
int flow_set_debug(unsigned long long* buf) {
    I2V_CHECK_CUDA(cudaMemcpyToSymbol(g_flow_dbg, &buf, sizeof(buf)));
    return flow_cluster_set_debug(buf);
}
namespace {
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

size_t flow_workspace_bytes(const FlowWeights& fw, int B) {
    const int rows = B < FLOW_MAX_ROWS ? B : FLOW_MAX_ROWS;
    size_t n = 0;
    n += align256(sizeof(float) * (size_t)B * fw.n_flows * 2 * 2 * fw.hidden);   // c1 (whole batch)
    n += 2 * align256(sizeof(float) * (size_t)rows * 2 * fw.hidden);              // hbuf0/1
    n += align256(sizeof(float) * (size_t)rows * 2 * fw.half);                    // st
    n += 256;                                                                      // barrier
    return n;
}

int launch_flow(const FlowWeights& fw, const float* in, const float* cond, float* out, float* logdet, int B,
                bool reverse, void* ws, size_t ws_bytes, cudaStream_t stream) {
    I2V_REQUIRE(B > 0, "flow: empty batch");
    I2V_REQUIRE(fw.n_flows <= 64, "flow: n_flows=%d > 64", fw.n_flows);
    I2V_REQUIRE(fw.hidden % 4 == 0 && fw.hidden <= 512, "flow: hidden=%d must be a multiple of 4 and <= 512", fw.hidden);
    I2V_REQUIRE(fw.zc % 4 == 0, "flow: (padded) conditioning width %d must be a multiple of 4", fw.zc);
    I2V_REQUIRE(ws_bytes >= flow_workspace_bytes(fw, B), "flow: workspace too small (%zu < %zu)", ws_bytes,
                flow_workspace_bytes(fw, B));
    const int H = fw.hidden, d = fw.d;
    const int rows_max = B < FLOW_MAX_ROWS ? B : FLOW_MAX_ROWS;
    char* p = static_cast<char*>(ws);
    float* c1 = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)B * fw.n_flows * 2 * 2 * H);
    float* hbuf0 = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)rows_max * 2 * H);
    float* hbuf1 = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)rows_max * 2 * H);
    float* st = reinterpret_cast<float*>(p); p += align256(sizeof(float) * (size_t)rows_max * 2 * fw.half);
    unsigned* bar = reinterpret_cast<unsigned*>(p);

    // hoisted conditioning GEMM: c1[b, :] = W1c . cond[b] + b1   (all couplings at once)
    const int n1 = fw.n_flows * 2 * 2 * H;
    if (int rc = launch_linear(cond, fw.w1c, fw.b1, c1, B, fw.zc, n1, ACT_NONE, stream)) return rc;

    if (flow_cluster_eligible(fw)) {
        // cluster-resident nets: no grid barrier, no activation round trip through L2 (flow_cluster.cu)
        const int rc = launch_flow_cluster(fw, in, c1, out, logdet, B, reverse, stream);
        if (rc <= 0) return rc;        // launched or failed; 1 = clusters of 16 cannot be scheduled here -> cooperative kernel
    }
    static unsigned long long attr_devs = 0;
    const size_t smem_max = 200 * 1024;
    I2V_CHECK_CUDA(ensure_max_dyn_smem(flow_kernel, (int)smem_max, attr_devs));
    int dev = 0, sms = 0;
    I2V_CHECK_CUDA(cudaGetDevice(&dev));
    I2V_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    I2V_REQUIRE((sms / 2) * FLOW_MAXC >= fw.hidden, "flow: %d SMs cannot own hidden=%d features with %d per CTA", sms, fw.hidden,
                FLOW_MAXC);

    for (int b0 = 0; b0 < B; b0 += FLOW_MAX_ROWS) {
        const int rows = (B - b0) < FLOW_MAX_ROWS ? (B - b0) : FLOW_MAX_ROWS;
        const size_t smem = sizeof(float) * ((size_t)2 * rows * d + ((rows + 3) & ~3) + (size_t)rows * H);
        I2V_REQUIRE(smem <= smem_max, "flow: shared memory %zu exceeds budget", smem);
        FlowKernelArgs ka;
        ka.w = fw;
        ka.in = in + (size_t)b0 * d;
        ka.c1 = c1 + (size_t)b0 * n1;
        ka.hbuf0 = hbuf0; ka.hbuf1 = hbuf1; ka.st = st;
        ka.out = out + (size_t)b0 * d;
        ka.logdet = logdet ? logdet + b0 : nullptr;
        ka.bar = bar;
        ka.B = rows;
        ka.reverse = reverse ? 1 : 0;
        for (int i = 0; i < 64; ++i) ka.cond_mode[i] = (i < fw.n_flows && fw.cond_mode) ? fw.cond_mode[i] : 0;
        I2V_CHECK_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(unsigned), stream));
        const double wbytes = 4.0 * fw.n_flows * 2 * ((double)2 * H * fw.half + (double)fw.depth * 2 * H * H + (double)2 * fw.half * H);
        ProfScope ps(PROF_FLOW, 2.0 * rows * wbytes / 4.0, wbytes + 4.0 * rows * (2.0 * d + n1), stream);
        void* kargs[] = {&ka};
        I2V_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)flow_kernel, dim3(sms), dim3(FLOW_THREADS), kargs, smem,
                                                   stream));
    }
    return 0;
}

}  // namespace i2v
