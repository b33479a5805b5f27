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
    // (no block-wide fence: the barrier orders the block's stores before the flag thread, whose st.release.sys is cumulative;
    //  512 threads x membar.sys was measured at 6.6 us of a 28 us kernel)
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

// ---- low-latency all-reduce for small buffers (<= kLLMaxBytes): data and flag in the same 16-byte store ----------------------
// The two-shot kernel above pays four NVLink hops that are pure latency at 1.5 MB (entry flags, pulled loads = a round trip,
// exit flags; measured 26.7 us at 2 GPUs, of which 13 us are the two flag barriers and 7 us the kernel's own launch + epoch
// bookkeeping; NCCL: 19.4 us).  Here every 16-byte store carries three floats and the epoch, so the RECEIVER sees data and
// "ready" in one transaction and nothing waits for an acknowledgement:
//   1  every rank pushes its contribution to slice j into rank j's staging area  rs[parity][source][.]        (fire and forget)
//   2  rank j adds its own slice and the N-1 staged contributions in rank order (it spins on the epoch word of each element),
//      keeps the sum in its flat buffer and pushes it to every peer's           ag[parity][j][.]              (fire and forget)
//   3  every rank copies the N-1 reduced slices out of its staging area into its flat buffer (spinning on the epoch words).
// Two one-way hops, no barrier, no fence; sums are formed once per slice in a fixed order, hence bit-identical on all ranks.
// Staging slots are rewritten only one step later -- by then the writer has received this rank's phase-2 data of the current
// step, which this rank sends after it has consumed the slots -- and are double-buffered by epoch parity on top of that.
constexpr int64_t kLLMaxBytes = 4 << 20;
__host__ __device__ inline int64_t ll_slice3(int64_t flat_count, int world) { return (flat_count / world + 2) / 3 + 1; }   // packed elements per slice (upper bound)

__device__ __forceinline__ void st_ll(uint4 *p, float a, float b, float c, unsigned flag)
{
    asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4 *p, unsigned flag, float &a, float &b, float &c)
{
    const long long t0 = clock64();
    for (;;) {
        uint4 v;
        asm volatile("ld.relaxed.sys.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
        if (v.w == flag) { a = __uint_as_float(v.x); b = __uint_as_float(v.y); c = __uint_as_float(v.z); return true; }
        if (clock64() - t0 > 4000000000ll) return false;
    }
}

__global__ void __launch_bounds__(kArThreads)
k_peer_allreduce_ll(pslam_peer_t peer, int *__restrict__ fail_flag)
{
    pdl_enter();
    __shared__ unsigned long long s_epoch;
    const int world = peer.world, rank = peer.rank, tid = threadIdx.x;
    PeerSync *mine = static_cast<PeerSync *>(peer.sync[rank]);
    if (tid == 0) s_epoch = ld_acquire_sys(&mine->epoch_ar) + 1ull;
    __syncthreads();
    const unsigned long long e = s_epoch;
    const unsigned flag = (unsigned)(e & 0xffffffffull) | 0x80000000u;      // never the zero the staging area starts with
    const int par = (int)(e & 1ull);
    const int64_t n = peer.flat_count, per = n / world + (n % world ? 1 : 0);   // slice r = [r * per, min(n, (r + 1) * per))
    const int64_t s3 = ll_slice3(n, world);
    // staging area of a rank: rs[2][world][s3] then ag[2][world][s3], 16 bytes each
    auto rs_of = [&](int q, int src) { return static_cast<uint4 *>(peer.stage[q]) + ((int64_t)par * world + src) * s3; };
    auto ag_of = [&](int q, int src) { return static_cast<uint4 *>(peer.stage[q]) + ((int64_t)(2 + par) * world + src) * s3; };
    const int64_t gtid = (int64_t)blockIdx.x * kArThreads + tid, gstride = (int64_t)gridDim.x * kArThreads;
    float *flat = peer.flat[rank];
    auto load3 = [&](int64_t i, int64_t hi, float &a, float &b, float &c) {     // three consecutive floats of the own buffer, zero beyond the slice
        a = i < hi ? flat[i] : 0.f; b = i + 1 < hi ? flat[i + 1] : 0.f; c = i + 2 < hi ? flat[i + 2] : 0.f;
    };
    bool ok = true;
    // ---- 1: push this rank's share of every other slice ----
    for (int dj = 1; dj < world; ++dj) {
        const int j = (rank + dj) % world;
        const int64_t lo = (int64_t)j * per, hi = lo + per < n ? lo + per : n;
        uint4 *dst = rs_of(j, rank);
        for (int64_t t = gtid; lo + 3 * t < hi; t += gstride) {
            float a, b, c;
            load3(lo + 3 * t, hi, a, b, c);
            st_ll(dst + t, a, b, c, flag);
        }
    }
    // ---- 2: reduce the own slice in rank order, keep it, push it to everybody ----
    {
        const int64_t lo = (int64_t)rank * per, hi = lo + per < n ? lo + per : n;
        for (int64_t t = gtid; lo + 3 * t < hi; t += gstride) {
            float acc[3] = {0.f, 0.f, 0.f};
            for (int q = 0; q < world; ++q) {
                float a, b, c;
                if (q == rank) load3(lo + 3 * t, hi, a, b, c);
                else ok = ld_ll(rs_of(rank, q) + t, flag, a, b, c) && ok;
                if (q == 0) { acc[0] = a; acc[1] = b; acc[2] = c; }
                else { acc[0] += a; acc[1] += b; acc[2] += c; }
            }
            const int64_t i = lo + 3 * t;
            flat[i] = acc[0];
            if (i + 1 < hi) flat[i + 1] = acc[1];
            if (i + 2 < hi) flat[i + 2] = acc[2];
            for (int dq = 1; dq < world; ++dq) st_ll(ag_of((rank + dq) % world, rank) + t, acc[0], acc[1], acc[2], flag);
        }
    }
    // ---- 3: the other ranks' reduced slices out of the staging area ----
    for (int dj = 1; dj < world; ++dj) {
        const int j = (rank + dj) % world;
        const int64_t lo = (int64_t)j * per, hi = lo + per < n ? lo + per : n;
        const uint4 *src = ag_of(rank, j);
        for (int64_t t = gtid; lo + 3 * t < hi; t += gstride) {
            float a = 0.f, b = 0.f, c = 0.f;
            ok = ld_ll(src + t, flag, a, b, c) && ok;
            const int64_t i = lo + 3 * t;
            flat[i] = a;
            if (i + 1 < hi) flat[i + 1] = b;
            if (i + 2 < hi) flat[i + 2] = c;
        }
    }
    if (!ok && fail_flag) atomicOr(fail_flag, 16);
    __syncthreads();
    if (tid == 0) {
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
    bool staged = peer->flat_count * 4 <= kLLMaxBytes;
    for (int q = 0; q < peer->world; ++q) staged = staged && peer->stage[q] != nullptr;
    if (staged) {
        const int64_t per3 = ll_slice3(peer->flat_count, peer->world);
        int blocks = (int)ceil_div64(per3, kArThreads);
        blocks = blocks < 8 ? 8 : (blocks > kArMaxBlocks ? kArMaxBlocks : blocks);
        if (blocks > num_sms()) blocks = num_sms();
        launch_chain(k_peer_allreduce_ll, dim3(blocks), dim3(kArThreads), 0, st, *peer, fail_flag);
        PSLAM_CHECK_LAUNCH("peer_allreduce_ll");
        return 0;
    }
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

extern "C" int64_t pslam_peer_stage_bytes(int64_t flat_count, int world)
{
    if (flat_count <= 0 || world < 2 || flat_count * 4 > kLLMaxBytes) return 0;     // larger buffers take the two-shot kernel
    return 4 * (int64_t)world * ll_slice3(flat_count, world) * 16;                  // rs[2][world][s3] + ag[2][world][s3]
}

extern "C" int pslam_peer_allreduce(const pslam_peer_t *peer, int *fail_flag, pslam_stream_t stream)
{
    return launch_peer_allreduce(peer, fail_flag, (cudaStream_t)stream);
}
