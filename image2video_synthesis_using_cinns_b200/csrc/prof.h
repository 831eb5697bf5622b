// Launch accounting + optional CUDA-event timing per kernel family (used by bench.py for the roofline).
#pragma once
#include <cuda_runtime.h>

namespace i2v {

// 0 = the halo tensor-core conv kernel (dominant), 5 = per-tap tensor-core kernel, 6 = fp32 SIMT conv kernel
enum ProfCat : int { PROF_CONV = 0, PROF_STATS = 1, PROF_MODULATE = 2, PROF_FLOW = 3, PROF_OTHER = 4, PROF_CONV_TC1 = 5,
                     PROF_CONV_SIMT = 6, PROF_NCAT = 7 };

// Counts the launch; when profiling is enabled also brackets it with events on `stream`.
struct ProfScope {
    ProfScope(int cat, double flops, double bytes, cudaStream_t stream);
    ~ProfScope();
    int idx_;
    cudaStream_t stream_;
};

}  // namespace i2v
