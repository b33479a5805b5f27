// Cross-GPU exchanges of a data-parallel mapping iteration, written against NVLink peer memory instead of NCCL calls
// issued from the host (rays sharded, map replicated; the reference itself is single GPU).
//
// Every rank owns two peer-mapped allocations (symmetric memory: the same size on every rank, every rank holds a
// device pointer to every other rank's copy):
//   * `sync` (pslam_peer_sync_bytes()): flags and small payloads that PEERS write into;
//   * `flat` ([flat_count] floats): the flat gradient buffer [E*16 | decoder] the backward kernels add into.
// Two exchanges per iteration, none of them a host-side call:
//   1. the loss closure (src/criterion.py:37-50, 96-112 couples all rays through global means and counts): the single block
//      that reduces this rank's loss partials stores its 16 raw sums into every peer's `rows[parity][rank]`, raises
//      `loss_flag[parity][rank] = epoch` there, waits until its own flags show this epoch from everybody, and closes the loss
//      from the world's rows -- identically on every rank (composite.cu: k_loss_reduce calls peer_loss_exchange);
//   2. the gradient all-reduce: a two-shot kernel -- rank r sums slice r of everybody's buffer (peer loads, fixed rank order,
//      so every rank ends up with bit-identical sums) and stores the result into everybody's slice r (peer stores) -- between
//      an entry and an exit barrier made of per-block flags in `sync`.
// Flags carry a monotonically increasing epoch (kept on the device, so a captured CUDA graph replays correctly); nothing is
// ever reset.  A peer that never shows up is a bounded spin: the kernel gives up after ~2 s, sets bit 4 (value 16) of
// counters[PSLAM_C_OVERFLOW] where a pipeline is attached and carries on with what it has, so a lost rank cannot hang the GPU.
#include "common.cuh"
#include "kernels.h"
#include "peer.cuh"

namespace pslam {

__global__ void __launch_bounds__(kArThreads)
k_peer_allreduce(pslam_peer_t peer, int *__restrict__ fail_flag)
{
    pdl_enter();
    __shared__ unsigned long long s_epoch;
    __shared__ int s_fail;
    const int world = peer.world, rank = peer.rank, b = blockIdx.x, tid = threadIdx.x;
    PeerSync *mine = static_cast<PeerSync *>(peer.sync[rank]);
    if (tid == 0) { s_epoch = ld_acquire_sys(&mine->epoch_ar) + 1ull; s_fail = 0; }
    __syncthreads();
    const unsigned long long e = s_epoch;
    // ---- entry barrier: every rank's kernel has started, i.e. its buffer is final (stream order on that rank) ----
    if (tid < world) {
        st_release_sys(&static_cast<PeerSync *>(peer.sync[tid])->ar_flag[0][b][rank], e);
        if (!spin_until(&mine->ar_flag[0][b][tid], e)) s_fail = 1;
    }
    __syncthreads();
    // ---- slice `rank` of the world's buffers: sum in rank order, store to everybody ----
    const int64_t n4 = peer.flat_count / 4;
    const int64_t lo = n4 * rank / world, hi = n4 * (rank + 1) / world;
    if (!s_fail) {
        // every load of an element is in flight before the first is used: a remote load is a ~2 us NVLink round trip, and the
        // launch gives a thread one or two elements, so the slice costs about one round trip
        const int64_t stride = (int64_t)gridDim.x * kArThreads;
        for (int64_t i = lo + (int64_t)b * kArThreads + tid; i < hi; i += stride) {
            float4 v[PSLAM_MAX_PEERS];
#pragma unroll
            for (int q = 0; q < PSLAM_MAX_PEERS; ++q)
                if (q < world) v[q] = ld_relaxed_sys_v4(reinterpret_cast<const float4 *>(peer.flat[q]) + i);
            float4 acc = v[0];
#pragma unroll
            for (int q = 1; q < PSLAM_MAX_PEERS; ++q)
                if (q < world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
#pragma unroll
            for (int q = 0; q < PSLAM_MAX_PEERS; ++q)
                if (q < world) reinterpret_cast<float4 *>(peer.flat[q])[i] = acc;
        }
    }
    // ---- exit barrier: everybody's slice has landed in this rank's buffer ----
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        st_release_sys(&static_cast<PeerSync *>(peer.sync[tid])->ar_flag[1][b][rank], e);
        if (!spin_until(&mine->ar_flag[1][b][tid], e)) s_fail = 1;
    }
    __syncthreads();
    if (tid == 0) {
        if (s_fail && fail_flag) atomicOr(fail_flag, 16);
        __threadfence();
        if (atomicAdd(&mine->ticket, 1u) == gridDim.x - 1) {   // last block of this launch: the epoch moves on
            mine->ticket = 0u;
            st_release_sys(&mine->epoch_ar, e);
        }
    }
}

int launch_peer_allreduce(const pslam_peer_t *peer, int *fail_flag, cudaStream_t st)
{
    PSLAM_CHECK_ARG(peer && peer->world >= 2 && peer->world <= PSLAM_MAX_PEERS && peer->rank >= 0 && peer->rank < peer->world, PSLAM_E_ARG,
                    "peer_allreduce: world %d / rank %d", peer ? peer->world : 0, peer ? peer->rank : 0);
    PSLAM_CHECK_ARG(peer->flat_count > 0 && peer->flat_count % 4 == 0, PSLAM_E_ARG, "peer_allreduce: flat_count must be a positive multiple of 4");
    for (int q = 0; q < peer->world; ++q)
        PSLAM_CHECK_ARG(peer->sync[q] && peer->flat[q] && ((uintptr_t)peer->flat[q] % 16 == 0), PSLAM_E_ARG, "peer_allreduce: null or misaligned peer pointer");
    // one element (float4) of this rank's slice per thread up to kArMaxBlocks blocks, two or more beyond
    const int64_t slice4 = peer->flat_count / 4 / peer->world;
    int blocks = (int)ceil_div64(slice4, kArThreads);
    blocks = blocks < 8 ? 8 : (blocks > kArMaxBlocks ? kArMaxBlocks : blocks);
    if (blocks > num_sms()) blocks = num_sms();      // every block spins on its peers: all of them must be resident
    launch_chain(k_peer_allreduce, dim3(blocks), dim3(kArThreads), 0, st, *peer, fail_flag);
    PSLAM_CHECK_LAUNCH("peer_allreduce");
    return 0;
}

}  // namespace pslam

using namespace pslam;

extern "C" int64_t pslam_peer_sync_bytes(void) { return (int64_t)sizeof(PeerSync); }

extern "C" int pslam_peer_allreduce(const pslam_peer_t *peer, int *fail_flag, pslam_stream_t stream)
{
    return launch_peer_allreduce(peer, fail_flag, (cudaStream_t)stream);
}
