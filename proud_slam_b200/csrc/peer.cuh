// Peer-memory exchange area and the system-scope loads / stores the cross-GPU kernels use (peer.cu, composite.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/proud_slam_b200.h"

namespace pslam {

constexpr int kArThreads = 512;
constexpr int kArMaxBlocks = 128;

// One per rank, in that rank's memory, written by its peers (zeroed once by the owner before the first exchange).
struct PeerSync {
    unsigned long long epoch_loss;                                   // exchanges completed by the owner (only the owner writes)
    unsigned long long epoch_ar;
    unsigned int ticket, pad_;
    unsigned long long loss_flag[2][PSLAM_MAX_PEERS];                // [parity][source rank] = epoch of the rows below
    double rows[2][PSLAM_MAX_PEERS][16];                             // [parity][source rank] raw loss sums (composite.cu RAW_*)
    unsigned long long ar_flag[2][kArMaxBlocks][PSLAM_MAX_PEERS];    // [entry | exit][block][source rank] = epoch
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float4 *p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// waits until *flag >= epoch; false after ~2 s (a peer that never arrives must not hang the GPU).  Relaxed polling, one acquire
// fence once the flag is seen (an acquire load per poll was measured at ~9 us for a barrier that should cost an NVLink hop).
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long epoch)
{
    const long long t0 = clock64();
    unsigned long long v;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 4000000000ll) return false;
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    return true;
}

int launch_peer_allreduce(const pslam_peer_t *peer, int *fail_flag, cudaStream_t st);

}  // namespace pslam
