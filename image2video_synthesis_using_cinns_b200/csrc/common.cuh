// Shared device/host helpers for the sm_100a kernels of the image->video sampling path.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <utility>

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

// fp16 saturation of the tensor-core engine's operand split: values beyond +-65504 (|x| > 4094 at the activation
// scale of 16) clip instead of turning into inf and then NaN frames; NaN stays NaN (comparisons are false).
__device__ __forceinline__ float sat_f16(float f) { return f > 65504.f ? 65504.f : (f < -65504.f ? -65504.f : f); }
// hi = fp16(f), lo = fp16(f - hi): the error-compensated operand pair (f already carries the split scale)
__device__ __forceinline__ void split_f16(float f, __half& hi, __half& lo) {
    f = sat_f16(f);
    hi = __float2half_rn(f);
    lo = __float2half_rn(f - __half2float(hi));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch.  Every kernel of the path (except the cooperative flow kernel) is launched with the
// programmatic-stream-serialization attribute, lets its successor be scheduled right away (launch_dependents) and
// only then waits for its predecessor to complete and flush (wait): the successor's launch latency, block scheduling
// and prologue (barrier init, TMEM allocation, tensor-map prefetch) overlap the predecessor's tail.  Correctness
// rests on one rule: no global-memory access before pdl_wait().  Completion is transitive because every kernel
// finishes after its own wait returns.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // api.cu; I2V_PDL=0 turns the launch attribute off (the device-side instructions become no-ops)

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: set it once per device the
// process launches on (`done_mask`: one bit per device ordinal, a static at the call site).
template <class K>
inline cudaError_t ensure_max_dyn_smem(K kernel, int bytes, unsigned long long& done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done_mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) done_mask |= 1ull << dev;
    return e;
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
