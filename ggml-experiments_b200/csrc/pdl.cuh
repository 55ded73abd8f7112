// Programmatic dependent launch (PDL) for the kernels of the fused plan.
//
// A forward pass is a chain of ~80 short kernels; between two of them the GPU otherwise drains completely before the next grid
// is even scheduled.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, the next kernel's CTAs are scheduled
// as soon as every CTA of the current one has executed griddepcontrol.launch_dependents (or exited); they run their prologue
// (barrier init, TMEM allocation, constant loads) and then block in griddepcontrol.wait until the previous grid has COMPLETED
// and its memory is visible.  Rules followed here:
//   * every kernel launched through launch_pdl() executes pdl_wait() before its first read of an activation AND before its
//     first global write (the buffer it overwrites may still be read by the previous kernel);
//   * kernels that do not contain pdl_wait() are launched the normal way (full serialisation on both sides);
//   * GGML_B200_NO_PDL=1 turns the attribute off (pdl_wait / pdl_trigger are then no-ops).
// The edges survive CUDA-graph capture (programmatic dependency edges).
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

#include "internal.h"

namespace b200 {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = getenv("GGML_B200_NO_PDL") == nullptr;
    return on;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = block;
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = st;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = 1;
    B200_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

}  // namespace b200
