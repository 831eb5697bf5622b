// Launch accounting + optional CUDA-event timing per kernel family (used by bench.py for the roofline).
#pragma once
#include <cuda_runtime.h>

namespace i2v {

enum ProfCat : int { PROF_CONV = 0, PROF_STATS = 1, PROF_MODULATE = 2, PROF_FLOW = 3, PROF_OTHER = 4, PROF_NCAT = 5 };

// Counts the launch; when profiling is enabled also brackets it with events on `stream`.
struct ProfScope {
    ProfScope(int cat, double flops, double bytes, cudaStream_t stream);
    ~ProfScope();
    int idx_;
    cudaStream_t stream_;
};

}  // namespace i2v
