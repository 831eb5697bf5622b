// Shared device/host helpers for the sm_100a kernels of the image->video sampling path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace i2v {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Activation applied in epilogues / element-wise passes.
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU02 = 2, ACT_TANH = 3 };

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_LRELU02: return v >= 0.f ? v : 0.2f * v;   // F.leaky_relu(x, 0.2), decoder.py:52
        case ACT_TANH: return tanhf(v);                      // decoder.py:118
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// thread-local error slot surfaced through i2v_last_error()
void set_error(const char* fmt, ...);

#define I2V_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::i2v::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                             cudaGetErrorString(_e));                                     \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

#define I2V_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            ::i2v::set_error(__VA_ARGS__);                                                \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace i2v
