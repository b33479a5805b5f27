// Shared helpers for libproud_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/proud_slam_b200.h"

namespace pslam {

void set_error(const char *fmt, ...);

#define PSLAM_CHECK_ARG(cond, code, ...)            \
    do {                                            \
        if (!(cond)) {                              \
            ::pslam::set_error(__VA_ARGS__);        \
            return (code);                          \
        }                                           \
    } while (0)

// Launch-error check: never aborts (the reference exit(-1)s, cuda_utils.h:37-48).
#define PSLAM_CHECK_LAUNCH(what)                                                  \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            ::pslam::set_error("%s: %s", what, cudaGetErrorString(e__));          \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// red.global.add.v4.f32 (sm_90+): one L2 reduction for 4 consecutive floats.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

int num_sms();

// "Done once per device": cudaFuncSetAttribute (the opt-in dynamic shared-memory size) and occupancy queries belong to a
// (function, device) pair, so a process that drives several GPUs must repeat them on each.  Races set the same value twice.
struct PerDevice { bool done[64]; int value[64]; };
static inline int current_device()
{
    int dev = 0;
    return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : 0;
}

// Programmatic dependent launch (PSLAM_OPT_PDL).  Every kernel of the fused step starts with pdl_wait() (all earlier grids
// complete, their memory visible; a no-op for a plain launch) followed by pdl_trigger() (the next grid may be scheduled as
// soon as every CTA of this one has started), so consecutive launches overlap their launch latency and prologue but never
// their data.  Because the wait comes first, "my predecessor has started" implies "its predecessors have completed".
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

int pdl_enabled();

template <class... KArgs, class... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args &&>(args))...);
}

}  // namespace pslam
