// Kernel 2: stratified inverse-CDF sampling inside the hit voxels.
//
// Semantics follow third_party/sparse_voxels/src/sample_gpu.cu:133-239 exactly,
// including the tail loop's indexing quirk: whether a
// ray gets its closing samples depends on its position j inside its group
// ("num_rays > j*P + curr_bin"), the break test reads the hit table of the
// group's FIRST ray, and after the main loop ran out of bins one more sample is
// emitted whose voxel id is read one slot past the ray's hits (-1, or the next
// ray's first voxel when the ray filled all P slots).  Arithmetic is pinned
// with _rn intrinsics: IEEE division, and the one FMA nvcc contracts in the
// reference (z = min + u*(max-min), SURVEY A-Q8).
//
// One `sample_ray` routine serves both layouts through two small policies:
//   * HitView  -- how hit (ray-in-group j', bin) is read (reference [b,n,P]
//                 tensors, or the fused pipeline's slot-major tables addressed
//                 through the rank -> ray map, emulating the reference's
//                 G=200 grouping and 800-ray chunks without materialising them)
//   * Sink     -- where samples go (padded [.., max_steps] rows, a CSR segment,
//                 or nowhere for the counting pass)
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace pslam {

constexpr int kSampleThreads = 64;    // fused path: one ray per thread, 2 warps per block so that 8192 rays cover 128 SMs
constexpr int kGroups = 200;      // voxel_helpers.py:300
constexpr int kChunkRays = 800;   // voxel_helpers.py:331 (4*G)

// optional per-warp timeline of k_sample_warp (pslam_debug_sample_trace), see WARP_TRACE below
__device__ long long *g_sample_trace = nullptr;
__device__ __forceinline__ long long globaltimer_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// The sampling loop for ray j of a group with `num_rays` rays and P hit slots.
template <class HitView, class Noise, class Sink>
__device__ __forceinline__ int sample_ray(const HitView &hv, const Noise &noise, Sink &sink, int j, int num_rays, int P,
                                          float prob0, float steps, float fixed_step_size)
{
    int bin = 0, s = 0;
    float lo_d = hv.tmin(j, 0), hi_d = hv.tmax(j, 0);
    float lo_c = 0.0f, hi_c = prob0;
    float step_size = __fdiv_rn(1.0f, steps);  // (float)(1.0/steps): double rounding is innocuous here
    float z_low = lo_d;
    const int total_steps = (int)ceilf(steps);
    bool done = false;
    if (fixed_step_size > 0.0f) step_size = fixed_step_size;

    // One thread runs one ray, so a step is a chain of dependent latencies (hash, IEEE division) with nothing to hide them
    // behind: the next step's cdf is computed ahead of the division.  (Taking two in-bin steps together was measured and
    // dropped: the extra divergence between lanes on the one- and two-step paths cost more than the second division hid.)
    float cdf_next = __fmul_rn(__fadd_rn(0.0f, noise(0)), step_size);
    for (int step = 0; step < total_steps; ++step) {
        const float cdf = cdf_next;
        while (cdf > hi_c) {
            sink(s, hv.idx(j, bin), __fsub_rn(hi_d, z_low), __fmul_rn(__fadd_rn(hi_d, z_low), 0.5f));
            ++bin; ++s;
            if (bin >= P || hv.idx(j, bin) == -1) { done = true; break; }
            lo_d = hv.tmin(j, bin); hi_d = hv.tmax(j, bin);
            lo_c = hi_c; hi_c = hv.next_cdf(hi_c, j, bin);
            z_low = lo_d;
        }
        if (done) break;
        cdf_next = __fmul_rn(__fadd_rn((float)(step + 1), noise(step + 1)), step_size);   // independent of this step: overlaps the division
        const float u = __fdiv_rn(__fsub_rn(cdf, lo_c), __fsub_rn(hi_c, lo_c));
        const float z = __fmaf_rn(u, __fsub_rn(hi_d, lo_d), lo_d);
        sink(s, hv.idx(j, bin), __fsub_rn(z, z_low), __fmul_rn(__fadd_rn(z, z_low), 0.5f));
        z_low = z; ++s;
    }
    // tail, sample_gpu.cu:224-237 (`~done` is always true)
    while (z_low < hi_d && num_rays > j * P + bin) {
        sink(s, hv.flat_idx(j * P + bin), __fsub_rn(hi_d, z_low), __fmul_rn(__fadd_rn(hi_d, z_low), 0.5f));
        ++bin; ++s;
        if (bin >= P || hv.flat_idx(bin) == -1) break;
        lo_d = hv.tmin(j, bin); hi_d = hv.tmax(j, bin);
        z_low = lo_d;
    }
    return s;
}

// ---- reference layout ---------------------------------------------------------------------
struct RefHits {
    const int *idx_; const float *min_, *max_, *prob_;  // group base pointers
    int P;
    __device__ __forceinline__ int idx(int j, int b) const { return __ldg(idx_ + j * P + b); }
    __device__ __forceinline__ int flat_idx(int f) const { return __ldg(idx_ + f); }
    __device__ __forceinline__ float tmin(int j, int b) const { return __ldg(min_ + j * P + b); }
    __device__ __forceinline__ float tmax(int j, int b) const { return __ldg(max_ + j * P + b); }
    __device__ __forceinline__ float prob(int j, int b) const { return __ldg(prob_ + j * P + b); }
    __device__ __forceinline__ float next_cdf(float hi_c, int j, int b) const { return __fadd_rn(hi_c, prob(j, b)); }
};
struct RefNoise {
    const float *row; int max_steps;
    __device__ __forceinline__ float operator()(int step) const { return step < max_steps ? __ldg(row + step) : 0.5f; }
};
struct PaddedSink {
    int *idx; float *depth, *dist; int max_steps;
    __device__ __forceinline__ void operator()(int s, int vox, float d, float z)
    {
        if (s < max_steps) { idx[s] = vox; dist[s] = d; depth[s] = z; }
    }
};

__global__ void __launch_bounds__(128)
k_inverse_cdf_ref(int b, int num_rays, int P, int max_steps, float fixed_step_size, const int *__restrict__ pts_idx,
                  const float *__restrict__ min_depth, const float *__restrict__ max_depth,
                  const float *__restrict__ noise, const float *__restrict__ probs, const float *__restrict__ steps,
                  int *__restrict__ sampled_idx, float *__restrict__ sampled_depth, float *__restrict__ sampled_dists)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * num_rays) return;
    const int g = (int)(r / num_rays), j = (int)(r % num_rays);
    const int64_t hb = (int64_t)g * num_rays * P;
    RefHits hv{pts_idx + hb, min_depth + hb, max_depth + hb, probs + hb, P};
    RefNoise nz{noise + r * max_steps, max_steps};
    PaddedSink sink{sampled_idx + r * max_steps, sampled_depth + r * max_steps, sampled_dists + r * max_steps, max_steps};
    int s = sample_ray(hv, nz, sink, j, num_rays, P, hv.prob(j, 0), __ldg(steps + r), fixed_step_size);
    s = min(s, max_steps);
    for (int k = s; k < max_steps; ++k) { sink.idx[k] = -1; sink.depth[k] = 0.0f; sink.dist[k] = 0.0f; }  // sample.cpp:80-89
}

// ---- fused layout ---------------------------------------------------------------------------
// Ray of rank q lives in group g = q / n, position jf = q % n (n = ceil(R_h/200)),
// chunk c = jf / 800, j = jf % 800 inside a chunk of nc = min(800, n - 800c) rays
// (voxel_helpers.py:300-343).  Ranks >= R_h are padding = copies of rank 0 (:302-311).
struct FusedHits {
    const int *hit_idx; const float *hit_min, *hit_max; const int *hit_count, *hit_ray;
    int R, Rh, P, chunk_base_rank;  // rank of (g, 800c + 0)
    int own_j, own_r, own_cnt;      // this thread's ray: every (j, bin) access of the sampling loop is to it
    // ... and its hit list is staged in shared memory ([slot][thread], conflict-free): the loop's bin changes happen at
    // different steps in different lanes, so from global memory every one of them was a serialised L2 round trip
    const int *s_idx; const float *s_min, *s_max;
    float total;                    // sum of this ray's segment lengths
    float max_distance;             // value the reference reads in padded slots (voxel_helpers.py:579-580)
    __device__ __forceinline__ int ray_of(int j) const
    {
        int q = chunk_base_rank + j;
        if (q >= Rh) q = 0;
        return __ldg(hit_ray + q);
    }
    __device__ __forceinline__ int idx_r(int r, int b) const
    {
        return (b < __ldg(hit_count + r)) ? __ldg(hit_idx + (int64_t)b * R + r) : -1;
    }
    // own ray: one load instead of three dependent ones (rank -> ray -> count -> slot)
    __device__ __forceinline__ int idx(int, int b) const { return (b < own_cnt) ? s_idx[b * kSampleThreads] : -1; }
    // the tail loop's flat index may land on another ray of the chunk (SURVEY A-Q7)
    __device__ __forceinline__ int flat_idx(int f) const
    {
        const int j = f / P;
        return j == own_j ? idx(j, f % P) : idx_r(ray_of(j), f % P);
    }
    __device__ __forceinline__ float tmin(int, int b) const { return (b < own_cnt) ? s_min[b * kSampleThreads] : max_distance; }
    __device__ __forceinline__ float tmax(int, int b) const { return (b < own_cnt) ? s_max[b * kSampleThreads] : max_distance; }
    __device__ __forceinline__ float prob(int j, int b) const
    {
        return __fdiv_rn(__fsub_rn(tmax(j, b), tmin(j, b)), total);  // voxel_helpers.py:639-643
    }
    // running sum of the bin probabilities: the same left-to-right chain, taken once while the hit list is staged (s_cum)
    // so that a bin change inside the divergent sampling loop costs a shared-memory read instead of an IEEE division
    const float *s_cum;
    __device__ __forceinline__ float next_cdf(float hi_c, int j, int b) const
    {
        return (s_cum && b < own_cnt) ? s_cum[b * kSampleThreads] : __fadd_rn(hi_c, prob(j, b));
    }
};

// Counter-based uniform noise for production runs (no noise tensor in HBM): the (seed, rank) pair is mixed
// once per ray into a 32-bit key, each step costs one 32-bit hash (lowbias32: 2 multiplies, 3 xor-shifts) ->
// (0.001, 0.999) like voxel_helpers.py:328's clamp.  Parity tests pass an explicit tensor instead.
struct HashNoise {
    uint32_t key;
    __device__ __forceinline__ explicit HashNoise(uint64_t k)
    {
        k ^= k >> 30; k *= 0xBF58476D1CE4E5B9ull;
        k ^= k >> 27; k *= 0x94D049BB133111EBull;
        key = (uint32_t)(k ^ (k >> 32));
    }
    __device__ __forceinline__ float operator()(int step) const
    {
        uint32_t x = key + (uint32_t)step * 0x9E3779B9u;
        x ^= x >> 16; x *= 0x7FEB352Du;
        x ^= x >> 15; x *= 0x846CA68Bu;
        x ^= x >> 16;
        const float u = (float)(x >> 8) * (1.0f / 16777216.0f);
        return fminf(fmaxf(u, 0.001f), 0.999f);
    }
};
struct TensorNoise {
    const float *row; int stride;
    __device__ __forceinline__ float operator()(int step) const { return step < stride ? __ldg(row + step) : 0.5f; }
};
struct CountSink {
    int last;
    __device__ __forceinline__ void operator()(int, int v, float, float) { last = v; }
};
struct CsrSink {
    int *vox; float *z, *dist; int *ray; int rank; int room;
    __device__ __forceinline__ void operator()(int s, int v, float d, float zz)
    {
        // a trailing -1 id (A-Q7/A-Q9) is not a sample; it can only be the last emission
        if (v != -1 && s < room) { vox[s] = v; z[s] = zz; dist[s] = fmaxf(d, 0.0f); ray[s] = rank; }
    }
};

// Everything one thread does for the ray of rank q: stage its hit list, rebuild the reference's group / chunk
// indexing, run the sampling loop into `sink`.  Returns the number of emissions.
template <class Sink>
__device__ __forceinline__ int run_ray(const pslam_render_t &p, int q, int Rh, int P, int *s_idx, float *s_min, float *s_max, Sink &sink,
                                       float *s_cum = nullptr)
{
    const int n = (Rh + kGroups - 1) / kGroups;
    const int g = q / n, jf = q % n;
    const int c = jf / kChunkRays, j = jf % kChunkRays;
    const int nc = min(kChunkRays, n - c * kChunkRays);
    const int r = __ldg(p.hit_ray + q);
    const int cnt = min(__ldg(p.hit_count + r), p.n_max);
    for (int b = 0; b < cnt; ++b) {      // independent loads, coalesced over the rays of a warp where ranks are consecutive
        s_idx[b * kSampleThreads] = __ldg(p.hit_idx + (int64_t)b * p.R + r);
        s_min[b * kSampleThreads] = __ldg(p.hit_min + (int64_t)b * p.R + r);
        s_max[b * kSampleThreads] = __ldg(p.hit_max + (int64_t)b * p.R + r);
    }
    // a5: dists, their sum (left-to-right fp32), probs and steps (voxel_helpers.py:639-644)
    float total = 0.0f;
    for (int b = 0; b < cnt; ++b) total = __fadd_rn(total, __fsub_rn(s_max[b * kSampleThreads], s_min[b * kSampleThreads]));
    const float steps = __fdiv_rn(total, p.step_size);
    FusedHits hv;
    hv.hit_idx = p.hit_idx; hv.hit_min = p.hit_min; hv.hit_max = p.hit_max;
    hv.hit_count = p.hit_count; hv.hit_ray = p.hit_ray;
    hv.R = p.R; hv.Rh = Rh; hv.P = P; hv.chunk_base_rank = g * n + c * kChunkRays;
    hv.total = total; hv.max_distance = p.max_distance;
    hv.own_j = j; hv.own_r = r; hv.own_cnt = cnt;
    hv.s_idx = s_idx; hv.s_min = s_min; hv.s_max = s_max;
    hv.s_cum = s_cum;
    if (s_cum) {
        float c = 0.0f;
        for (int b = 0; b < cnt; ++b) {
            const float pr = hv.prob(0, b);
            c = (b == 0) ? pr : __fadd_rn(c, pr);
            s_cum[b * kSampleThreads] = c;
        }
    }
    const float prob0 = hv.prob(j, 0);
    if (p.noise) {
        TensorNoise nz{p.noise + (int64_t)q * p.noise_stride, p.noise_stride};
        return sample_ray(hv, nz, sink, j, nc, P, prob0, steps, -1.0f);
    }
    const HashNoise nz((p.seed + (p.seed_dev ? *p.seed_dev : 0ull)) ^ ((uint64_t)(uint32_t)q * 0xD1B54A32D192ED03ull));
    return sample_ray(hv, nz, sink, j, nc, P, prob0, steps, -1.0f);
}

// Two-pass form (count -> offsets -> write): used when the one-pass kernel's shared-memory buffers do not fit.
template <bool WRITE>
__global__ void __launch_bounds__(kSampleThreads)
k_sample_fused(pslam_render_t p, int *__restrict__ block_counts)
{
    extern __shared__ __align__(16) unsigned char s_hits[];   // [3][n_max][kSampleThreads]: idx, min, max of this block's rays
    int *s_idx = reinterpret_cast<int *>(s_hits) + threadIdx.x;
    float *s_min = reinterpret_cast<float *>(s_hits) + (size_t)p.n_max * kSampleThreads + threadIdx.x;
    float *s_max = s_min + (size_t)p.n_max * kSampleThreads;
    const int Rh = p.counters[PSLAM_C_RH];
    const int P = p.counters[PSLAM_C_P];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    int nsamp = 0;
    if (q < Rh) {
        if (WRITE) {
            const int off = min(p.samp_off[q], p.sample_cap);
            const int room = max(0, min(p.samp_off[q + 1], p.sample_cap) - off);
            CsrSink csr{p.samp_vox + off, p.samp_z + off, p.samp_dist + off, p.samp_ray + off, q, room};
            run_ray(p, q, Rh, P, s_idx, s_min, s_max, csr);
        } else {
            CountSink cs{0};
            const int s = run_ray(p, q, Rh, P, s_idx, s_min, s_max, cs);
            // a trailing -1 id (A-Q7) is not a sample and can only be the last emission
            nsamp = (s > 0 && cs.last == -1) ? s - 1 : s;
            p.samp_off[q] = nsamp;  // counts now; turned into offsets by k_sample_offsets
        }
    }
    if (!WRITE) {
        __shared__ int s_sum[kSampleThreads / 32];
        const int wsum = warp_sum_i(nsamp), wmax = warp_max_i(nsamp);
        if ((threadIdx.x & 31) == 0) {
            s_sum[threadIdx.x >> 5] = wsum;
            if (wmax > 0) atomicMax(p.counters + PSLAM_C_S, wmax);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < kSampleThreads / 32; ++w) t += s_sum[w];
            block_counts[blockIdx.x] = t;
        }
    }
}

// One-pass form, ONE WARP PER RAY.  The reference's loop (sample_gpu.cu:176-222) looks serial -- every step's sample depends on
// the previous one through z_low -- but each step's cdf is a closed form of (step, noise), so the warp takes 32 steps at a time:
//   * lane k computes cdf_k and F(k) = the first bin whose cumulative probability is not below it; the loop's bin pointer never
//     moves back, so bin(k) is the running maximum of F (a warp prefix-max), and z_k follows from the bin's (cum, min, max);
//   * the in-bin sample of step k is emission number k + bin(k) (k earlier in-bin samples + one closing sample per bin left);
//     its z_low is z_{k-1} when step k-1 sat in the same bin (one shuffle), else the bin's entry depth;
//   * the closing samples of the bins crossed between steps k-1 and k are emitted by lane k (numbers k + b);
//   * the first step whose bin runs past the ray's hits is the loop's `done` break: later lanes stay silent.
// The arithmetic of every emission is the reference's, operation for operation (same _rn intrinsics), so samples stay
// bit-identical; only the order in which they are produced changes.  The tail loop (sample_gpu.cu:224-237, the indexing quirk
// SURVEY A-Q7) is short and data-dependent on other rays' hit lists: it runs as written, uniformly in all lanes.
// A block stages the hit lists of its rays cooperatively (coalesced over rays), each warp then samples its rays into a
// shared-memory buffer ([ray][slot]: lane = emission number, conflict-free), the CSR offsets come from a block scan plus a
// decoupled look-back over the per-block totals (state[b]: flag in the high word -- 1 = this block's total, 2 = inclusive
// prefix -- value in the low word; zeroed before the launch), and each warp copies its rays out as contiguous runs.  A ray with
// more emissions than buffer slots reruns straight into global memory.
constexpr int kWarpThreads = 256;               // 8 warps per block
constexpr int kWarpBuf = 96;                    // buffered emissions per ray
struct BufSink {
    int *vox; float *z, *dist;
    __device__ __forceinline__ void operator()(int s, int v, float d, float zz)
    {
        if (s < kWarpBuf) { vox[s] = v; z[s] = zz; dist[s] = fmaxf(d, 0.0f); }
    }
};

// the hit list of one ray in shared memory + what the tail loop needs to reach other rays of its chunk
struct WarpHits {
    const int *idx; const float *tmin_, *tmax_, *cum;   // [cnt]
    int cnt, P, own_j, chunk_base_rank, Rh, R;
    float max_distance;
    const int *hit_idx, *hit_count, *hit_ray;
    __device__ __forceinline__ int own_idx(int b) const { return b < cnt ? idx[b] : -1; }
    __device__ __forceinline__ float tmin(int b) const { return b < cnt ? tmin_[b] : max_distance; }
    __device__ __forceinline__ float tmax(int b) const { return b < cnt ? tmax_[b] : max_distance; }
    __device__ __forceinline__ int flat_idx(int f) const     // FusedHits::flat_idx
    {
        const int jj = f / P, b = f % P;
        if (jj == own_j) return own_idx(b);
        int q = chunk_base_rank + jj;
        if (q >= Rh) q = 0;
        const int r = __ldg(hit_ray + q);
        return (b < __ldg(hit_count + r)) ? __ldg(hit_idx + (int64_t)b * R + r) : -1;
    }
};

// All 32 lanes call this for the same ray.  Returns the number of emissions; `last` = voxel id of the final one.
template <class Noise, class Sink>
__device__ __forceinline__ int warp_sample_ray(const WarpHits &h, const Noise &noise, Sink &sink, int nc, float steps, int &last)
{
    const int lane = threadIdx.x & 31;
    const int cnt = h.cnt;
    const float step_size = __fdiv_rn(1.0f, steps);
    const int T = (int)ceilf(steps);
    // state handed to the tail loop
    int s = 0, bin = 0;
    float z_low = h.tmin(0), hi_d = h.tmax(0);
    int carry_bin = 0;
    float carry_z = 0.0f;
    for (int base = 0; base < T; base += 32) {
        const int k = base + lane;
        const bool act = k < T;
        float cdf = 0.0f;
        int F = carry_bin;
        if (act) {
            cdf = __fmul_rn(__fadd_rn((float)k, noise(k)), step_size);
            while (F < cnt && cdf > h.cum[F]) ++F;
        }
        int b_k = F;                                 // running maximum: the loop's bin pointer never moves back
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, b_k, o);
            if (lane >= o) b_k = max(b_k, t);
        }
        int b_prev = __shfl_up_sync(0xffffffffu, b_k, 1);
        if (lane == 0) b_prev = carry_bin;
        float z = 0.0f;
        if (act && b_k < cnt) {
            const float lo_c = b_k == 0 ? 0.0f : h.cum[b_k - 1], hi_c = h.cum[b_k];
            const float lo = h.tmin_[b_k], hi = h.tmax_[b_k];
            const float u = __fdiv_rn(__fsub_rn(cdf, lo_c), __fsub_rn(hi_c, lo_c));
            z = __fmaf_rn(u, __fsub_rn(hi, lo), lo);
        }
        float z_prev = __shfl_up_sync(0xffffffffu, z, 1);
        if (lane == 0) z_prev = carry_z;
        const bool has_prev = k >= 1;
        const bool live = act && b_prev < cnt;       // no earlier step was the `done` step
        if (live) {
            const int bend = min(b_k, cnt);
            for (int b = b_prev; b < bend; ++b) {    // closing samples of the bins left at this step
                const float zl = (b == b_prev && has_prev) ? z_prev : h.tmin_[b];
                const float hd = h.tmax_[b];
                sink(k + b, h.idx[b], __fsub_rn(hd, zl), __fmul_rn(__fadd_rn(hd, zl), 0.5f));
            }
            if (b_k < cnt) {
                const float zl = (b_k == b_prev && has_prev) ? z_prev : h.tmin_[b_k];
                sink(k + b_k, h.idx[b_k], __fsub_rn(z, zl), __fmul_rn(__fadd_rn(z, zl), 0.5f));
            }
        }
        const unsigned done_mask = __ballot_sync(0xffffffffu, live && b_k >= cnt);
        if (done_mask) {                             // the loop's break: bins ran out at step kd
            const int src = __ffs(done_mask) - 1;
            const int kd = base + src;
            const int bp = __shfl_sync(0xffffffffu, b_prev, src);
            const float zp = __shfl_sync(0xffffffffu, z_prev, src);
            s = kd + cnt; bin = cnt;
            z_low = (bp == cnt - 1 && kd >= 1) ? zp : h.tmin_[cnt - 1];
            hi_d = h.tmax_[cnt - 1];
            break;
        }
        if (base + 32 >= T) {                        // ran out of steps inside bin(T-1)
            const int src = T - 1 - base;
            bin = __shfl_sync(0xffffffffu, b_k, src);
            z_low = __shfl_sync(0xffffffffu, z, src);
            s = T + bin;
            hi_d = h.tmax_[bin];
            break;
        }
        carry_bin = __shfl_sync(0xffffffffu, b_k, 31);
        carry_z = __shfl_sync(0xffffffffu, z, 31);
    }
    // tail, sample_gpu.cu:224-237 (uniform in all lanes, lane 0 stores)
    last = 0;
    const int j = h.own_j, P = h.P;
    while (z_low < hi_d && nc > j * P + bin) {
        const int v = h.flat_idx(j * P + bin);
        if (lane == 0) sink(s, v, __fsub_rn(hi_d, z_low), __fmul_rn(__fadd_rn(hi_d, z_low), 0.5f));
        last = v;
        ++bin; ++s;
        if (bin >= P || h.flat_idx(bin) == -1) break;
        hi_d = h.tmax(bin);
        z_low = h.tmin(bin);
    }
    return s;
}

template <class Sink>
__device__ __forceinline__ int warp_run_ray(const pslam_render_t &p, const WarpHits &h, int q, int nc, float total, Sink &sink, int &last)
{
    const float steps = __fdiv_rn(total, p.step_size);
    if (p.noise) {
        TensorNoise nz{p.noise + (int64_t)q * p.noise_stride, p.noise_stride};
        return warp_sample_ray(h, nz, sink, nc, steps, last);
    }
    const HashNoise nz((p.seed + (p.seed_dev ? *p.seed_dev : 0ull)) ^ ((uint64_t)(uint32_t)q * 0xD1B54A32D192ED03ull));
    return warp_sample_ray(h, nz, sink, nc, steps, last);
}

// optional per-warp timeline (pslam_debug_sample_trace): [block][warp][8] = globaltimer at entry, clock64 at entry / hits staged /
// rays sampled / offsets known / copied out, globaltimer at exit, largest sample count of the warp's rays
#define WARP_TRACE(slot, value)                                                                                                 \
    do {                                                                                                                        \
        if (g_sample_trace && (threadIdx.x & 31) == 0) g_sample_trace[((size_t)blockIdx.x * (kWarpThreads / 32) + (threadIdx.x >> 5)) * 8 + (slot)] = (value); \
    } while (0)

__global__ void __launch_bounds__(kWarpThreads)
k_sample_warp(pslam_render_t p, unsigned long long *__restrict__ state, int rpb)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Rh = p.counters[PSLAM_C_RH];
    const int P = p.counters[PSLAM_C_P];
    const int q0 = blockIdx.x * rpb;
    if (q0 >= Rh) return;                                   // (no later block has rays either: nobody waits for this one)
    WARP_TRACE(0, globaltimer_ns());
    WARP_TRACE(1, clock64());
    const int nm = p.n_max, pitch = nm | 1;                 // odd pitch: staging writes (ray varies fastest) are conflict-free
    int *h_idx = reinterpret_cast<int *>(s_raw);            // [rpb][pitch]
    float *h_min = reinterpret_cast<float *>(h_idx + rpb * pitch);
    float *h_max = h_min + rpb * pitch;
    float *h_cum = h_max + rpb * pitch;
    float *w_pr = h_cum + rpb * pitch;                      // [warps][pitch]: bin probabilities of the ray a warp is preparing
    int *b_vox = reinterpret_cast<int *>(w_pr + (kWarpThreads / 32) * pitch);   // [rpb][kWarpBuf]
    float *b_z = reinterpret_cast<float *>(b_vox + rpb * kWarpBuf);
    float *b_dist = b_z + rpb * kWarpBuf;
    int *s_cnt = reinterpret_cast<int *>(b_dist + rpb * kWarpBuf);              // [rpb] hit count
    int *s_ns = s_cnt + rpb;                                // [rpb] samples of the ray
    int *s_off = s_ns + rpb;                                // [rpb] its CSR offset (unclamped)
    float *s_total = reinterpret_cast<float *>(s_off + rpb);   // [rpb] summed segment length

    // ---- phase 0: stage the hit lists (thread -> ray tid % rpb, slots tid / rpb, + blockDim / rpb, ...) ----
    {
        const int i = tid % rpb, q = q0 + i;
        int cnt = 0, r = 0;
        if (q < Rh) {
            r = __ldg(p.hit_ray + q);
            cnt = min(__ldg(p.hit_count + r), nm);
        }
        if (tid < rpb) s_cnt[i] = cnt;
        for (int b = tid / rpb; b < cnt; b += kWarpThreads / rpb) {
            h_idx[i * pitch + b] = __ldg(p.hit_idx + (int64_t)b * p.R + r);
            h_min[i * pitch + b] = __ldg(p.hit_min + (int64_t)b * p.R + r);
            h_max[i * pitch + b] = __ldg(p.hit_max + (int64_t)b * p.R + r);
        }
    }
    __syncthreads();
    WARP_TRACE(2, clock64());
    // ---- phase 1: every warp samples its rays ----
    const int n = (Rh + kGroups - 1) / kGroups;
    auto hits_of = [&](int i, int q) {
        const int g = q / n, jf = q % n;
        const int c = jf / kChunkRays;
        WarpHits h;
        h.idx = h_idx + i * pitch; h.tmin_ = h_min + i * pitch; h.tmax_ = h_max + i * pitch; h.cum = h_cum + i * pitch;
        h.cnt = s_cnt[i]; h.P = P; h.own_j = jf % kChunkRays; h.chunk_base_rank = g * n + c * kChunkRays; h.Rh = Rh; h.R = p.R;
        h.max_distance = p.max_distance;
        h.hit_idx = p.hit_idx; h.hit_count = p.hit_count; h.hit_ray = p.hit_ray;
        return h;
    };
    auto nc_of = [&](int q) { const int c = (q % n) / kChunkRays; return min(kChunkRays, n - c * kChunkRays); };
    int wmax = 0;
    for (int i = warp; i < rpb; i += kWarpThreads / 32) {
        const int q = q0 + i;
        int nsamp = 0;
        if (q < Rh) {
            WarpHits h = hits_of(i, q);
            const int cnt = h.cnt;
            // a5: dists, their sum (left to right), probs and their running sum (voxel_helpers.py:639-644): every lane runs the same
            // chains on broadcast reads, lane b keeps element b
            float total = 0.0f;
            for (int b = 0; b < cnt; ++b) total = __fadd_rn(total, __fsub_rn(h.tmax_[b], h.tmin_[b]));
            float *pr = w_pr + warp * pitch;
            for (int b = lane; b < cnt; b += 32) pr[b] = __fdiv_rn(__fsub_rn(h.tmax_[b], h.tmin_[b]), total);
            __syncwarp();
            {
                float c = 0.0f;
                const int upto = ((cnt - 1 - lane) >= 0) ? lane + ((cnt - 1 - lane) / 32) * 32 : -1;   // last element this lane keeps
                for (int b = 0; b <= upto; ++b) {
                    c = (b == 0) ? pr[0] : __fadd_rn(c, pr[b]);
                    if ((b & 31) == lane) h_cum[i * pitch + b] = c;
                }
            }
            __syncwarp();
            if (lane == 0) s_total[i] = total;
            BufSink bs{b_vox + i * kWarpBuf, b_z + i * kWarpBuf, b_dist + i * kWarpBuf};
            int last = 0;
            const int s = warp_run_ray(p, h, q, nc_of(q), total, bs, last);
            nsamp = (s > 0 && last == -1) ? s - 1 : s;      // a trailing -1 id (A-Q7) is not a sample and can only be the last emission
            __syncwarp();
        }
        if (lane == 0) s_ns[i] = nsamp;
        wmax = max(wmax, nsamp);
    }
    if (lane == 0 && wmax > 0) atomicMax(p.counters + PSLAM_C_S, wmax);
    __syncthreads();
    WARP_TRACE(3, clock64());
    // ---- phase 2: CSR offsets = block scan + decoupled look-back (warp 0) ----
    if (warp == 0) {
        const int v = lane < rpb ? s_ns[lane] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        const int btotal = __shfl_sync(0xffffffffu, x, 31);
        const int b = blockIdx.x;
        if (lane == 0) {
            const unsigned long long sv = ((unsigned long long)(b == 0 ? 2u : 1u) << 32) | (unsigned)btotal;
            __threadfence();
            atomicExch(state + b, sv);
        }
        int base = 0;
        for (int hi_b = b - 1; hi_b >= 0; hi_b -= 32) {
            const int idx = hi_b - lane;                        // lane 0 looks at the nearest predecessor
            unsigned long long sv = 2ull << 32;                 // out of range: a zero prefix
            if (idx >= 0) {
                do { sv = *reinterpret_cast<volatile unsigned long long *>(state + idx); } while ((sv >> 32) == 0ull);
            }
            const unsigned done = __ballot_sync(0xffffffffu, (sv >> 32) == 2ull);
            const int stop = done ? __ffs(done) - 1 : 32;       // nearest lane that already has its inclusive prefix
            int part = (lane <= stop) ? (int)(unsigned)(sv & 0xffffffffull) : 0;
            part = warp_sum_i(part);
            base += part;
            if (done) break;
        }
        if (lane == 0 && b > 0) {
            __threadfence();
            atomicExch(state + b, (2ull << 32) | (unsigned)(base + btotal));
        }
        if (lane < rpb) s_off[lane] = base + x - v;
    }
    __syncthreads();
    WARP_TRACE(4, clock64());
    // ---- phase 3: copy-out, a ray = one contiguous run per array ----
    for (int i = warp; i < rpb; i += kWarpThreads / 32) {
        const int q = q0 + i;
        if (q >= Rh) break;
        const int off_raw = s_off[i], nsamp = s_ns[i];
        if (lane == 0) {
            p.samp_off[q] = off_raw;
            if (q == Rh - 1) {
                const int total = off_raw + nsamp;
                p.samp_off[Rh] = total;
                p.counters[PSLAM_C_NSAMP] = min(total, p.sample_cap);
                if (total > p.sample_cap) atomicOr(p.counters + PSLAM_C_OVERFLOW, 1);
            }
        }
        const int off = min(off_raw, p.sample_cap);
        const int room = max(0, min(off_raw + nsamp, p.sample_cap) - off);
        if (nsamp <= kWarpBuf) {
            for (int k = lane; k < min(nsamp, room); k += 32) {
                p.samp_vox[off + k] = b_vox[i * kWarpBuf + k];
                p.samp_z[off + k] = b_z[i * kWarpBuf + k];
                p.samp_dist[off + k] = b_dist[i * kWarpBuf + k];
                p.samp_ray[off + k] = q;
            }
        } else {                                            // too long for the buffer: once more, straight into global memory
            WarpHits h = hits_of(i, q);
            CsrSink csr{p.samp_vox + off, p.samp_z + off, p.samp_dist + off, p.samp_ray + off, q, room};
            int last = 0;
            warp_run_ray(p, h, q, nc_of(q), s_total[i], csr, last);
        }
    }
    WARP_TRACE(5, clock64());
    WARP_TRACE(6, globaltimer_ns());
    WARP_TRACE(7, (long long)wmax);
}

// counts -> exclusive offsets (block-local scan + scanned block bases); also writes off[R_h].
__global__ void __launch_bounds__(kSampleThreads)
k_sample_offsets(pslam_render_t p, const int *__restrict__ block_base)
{
    __shared__ int s_warp[kSampleThreads / 32];
    const int Rh = p.counters[PSLAM_C_RH];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = (q < Rh) ? p.samp_off[q] : 0;
    int x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int base = block_base[blockIdx.x];
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    if (q < Rh) p.samp_off[q] = base + x - v;
    if (q == Rh - 1) {
        const int total = base + x;
        p.samp_off[Rh] = total;
        p.counters[PSLAM_C_NSAMP] = min(total, p.sample_cap);
        if (total > p.sample_cap) atomicOr(p.counters + PSLAM_C_OVERFLOW, 1);
    }
}

// rays per block of the warp-per-ray kernel: 32 when that still gives every SM a block, else 16 / 8 (tracking batches)
static int sample_rays_per_block(int R)
{
    static int forced = -1;                       // PSLAM_SAMPLE_RPB=8|16|32: measurement override
    if (forced < 0) {
        const char *e = getenv("PSLAM_SAMPLE_RPB");
        const int v = e ? atoi(e) : 0;
        forced = (v == 8 || v == 16 || v == 32) ? v : 0;
    }
    if (forced) return forced;
    const int sms = num_sms();
    if (ceil_div(R, 16) < sms) return 8;         // tracking batches: one ray per warp, every SM gets a block
    if (ceil_div(R, 16) <= 8 * sms) return 16;   // measured at 8192 rays: 14.7 us (16) vs 19.8 (8) / 19.0 (32)
    return 32;
}

int launch_sample_fused(const pslam_render_t *p, cudaStream_t st)
{
    int *block_counts = p->scratch_i + scratch_i_sample_off(p->R);   // after intersect's block_hits: look-back state, 2 ints per block
    const int rpb = sample_rays_per_block(p->R);
    const int pitch = p->n_max | 1;
    const size_t smem1 = sizeof(int) * ((size_t)4 * rpb * pitch + (size_t)(kWarpThreads / 32) * pitch + (size_t)3 * rpb * kWarpBuf + (size_t)4 * rpb);
    if (smem1 <= 200 * 1024) {
        static PerDevice once = {};
        bool &configured = once.done[current_device()];
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(k_sample_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) { set_error("sample: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            configured = true;
        }
        unsigned long long *state = reinterpret_cast<unsigned long long *>(block_counts + (((uintptr_t)block_counts & 7) ? 1 : 0));
        // (the look-back state was cleared by k_compact_rays, the last kernel of the intersection stage)
        launch_chain(k_sample_warp, dim3(ceil_div(p->R, rpb)), dim3(kWarpThreads), smem1, st, *p, state, rpb);
        PSLAM_CHECK_LAUNCH("sample_warp");
        return 0;
    }
    // hit lists too long for the shared-memory staging of a block: count -> scan -> write, one thread per ray
    const int nb = ceil_div(p->R, kSampleThreads);
    const size_t smem = (size_t)p->n_max * kSampleThreads * 12;
    PSLAM_CHECK_ARG(smem <= 48 * 1024, PSLAM_E_RANGE, "n_max=%d: the per-block hit staging exceeds 48 KB of shared memory", p->n_max);
    k_sample_fused<false><<<nb, kSampleThreads, smem, st>>>(*p, block_counts);
    PSLAM_CHECK_LAUNCH("sample_count");
    if (int rc = scan_partials(block_counts, nb, p->counters + PSLAM_C_TILE2, st)) return rc;
    k_sample_offsets<<<nb, kSampleThreads, 0, st>>>(*p, block_counts);
    PSLAM_CHECK_LAUNCH("sample_offsets");
    k_sample_fused<true><<<nb, kSampleThreads, smem, st>>>(*p, nullptr);
    PSLAM_CHECK_LAUNCH("sample_write");
    return 0;
}

// ---- uniform_ray_sampling (API surface), sample_gpu.cu:13-131 ---------------------------------
__global__ void k_uniform_sampling_ref(int b, int num_rays, int max_hits, int max_steps, float step_size,
                                       const int *__restrict__ pts_idx, const float *__restrict__ min_depth,
                                       const float *__restrict__ max_depth, const float *__restrict__ noise,
                                       int *__restrict__ sampled_idx, float *__restrict__ sampled_depth,
                                       float *__restrict__ sampled_dists)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * num_rays) return;
    const int *pi = pts_idx + r * max_hits;
    const float *mn = min_depth + r * max_hits, *mx = max_depth + r * max_hits, *nz = noise + r * max_steps;
    int *oi = sampled_idx + r * max_steps;
    float *od = sampled_depth + r * max_steps, *os = sampled_dists + r * max_steps;
    for (int k = 0; k < max_steps; ++k) { oi[k] = -1; od[k] = 0.0f; os[k] = 0.0f; }
    int s = 0, ucur = 0, umin = 0, umax = 0;
    float last_min = 0.f, last_max = 0.f, curr = 0.f;
    // merge segment boundaries with the marching samples, :45-95
    while (true) {
        if (umax == max_hits || ucur == max_steps || pi[umax] == -1) break;
        last_min = (umin < max_hits) ? mn[umin] : 10000.0f;
        last_max = (umax < max_hits) ? mx[umax] : 10000.0f;
        if (ucur < max_steps) curr = mn[0] + ((float)ucur + nz[ucur]) * step_size;
        if (s >= max_steps) break;  // the reference would write out of bounds here
        if (last_max <= curr && last_max <= last_min) { od[s] = last_max; oi[s] = pi[umax]; ++umax; ++s; continue; }
        if (curr <= last_min && curr <= last_max) { od[s] = curr; oi[s] = (umin > 0) ? pi[umin - 1] : -1; ++ucur; ++s; continue; }
        if (last_min <= curr && last_min <= last_max) { od[s] = last_min; oi[s] = pi[umin]; ++umin; ++s; continue; }
        break;  // NaN input: none of the three orderings holds (the reference would spin)
    }
    int step = 0;
    umin = 0; umax = 0;
    for (ucur = 0; ucur < max_steps - 1; ++ucur) {  // :97-123
        if (oi[ucur + 1] == -1) break;
        const float l = od[ucur], rr = od[ucur + 1];
        od[ucur] = (l + rr) * 0.5f;
        os[ucur] = rr - l;
        if (umin < max_hits && od[ucur] >= mn[umin] && pi[umin] > -1) ++umin;
        if (umax < max_hits && od[ucur] >= mx[umax] && pi[umax] > -1) ++umax;
        if (umax == max_hits || pi[umax] == -1) break;
        if (umin - 1 == umax && os[ucur] > 0) { od[step] = od[ucur]; os[step] = os[ucur]; oi[step] = oi[ucur]; ++step; }
    }
    for (int k = step; k < max_steps; ++k) oi[k] = -1;
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_debug_sample_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(g_sample_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { set_error("sample_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pslam_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps, float fixed_step_size,
                                          const int *pts_idx, const float *min_depth, const float *max_depth,
                                          const float *uniform_noise, const float *probs, const float *steps,
                                          int *sampled_idx, float *sampled_depth, float *sampled_dists,
                                          pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(b > 0 && num_rays > 0 && max_hits > 0 && max_steps > 0, PSLAM_E_ARG,
                    "sizes must be positive (b=%d num_rays=%d max_hits=%d max_steps=%d)", b, num_rays, max_hits, max_steps);
    PSLAM_CHECK_ARG(pts_idx && min_depth && max_depth && uniform_noise && probs && steps, PSLAM_E_ARG, "null input pointer");
    PSLAM_CHECK_ARG(sampled_idx && sampled_depth && sampled_dists, PSLAM_E_ARG, "null output pointer");
    const int blocks = (int)ceil_div64((int64_t)b * num_rays, 128);
    k_inverse_cdf_ref<<<blocks, 128, 0, (cudaStream_t)stream>>>(
        b, num_rays, max_hits, max_steps, fixed_step_size, pts_idx, min_depth, max_depth, uniform_noise, probs, steps,
        sampled_idx, sampled_depth, sampled_dists);
    PSLAM_CHECK_LAUNCH("inverse_cdf_sampling");
    return 0;
}

extern "C" int pslam_uniform_ray_sampling(int b, int num_rays, int max_hits, int max_steps, float step_size,
                                          const int *pts_idx, const float *min_depth, const float *max_depth,
                                          const float *uniform_noise, int *sampled_idx, float *sampled_depth,
                                          float *sampled_dists, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(b > 0 && num_rays > 0 && max_hits > 0 && max_steps > 0, PSLAM_E_ARG, "sizes must be positive");
    PSLAM_CHECK_ARG(pts_idx && min_depth && max_depth && uniform_noise && sampled_idx && sampled_depth && sampled_dists,
                    PSLAM_E_ARG, "null pointer argument");
    const int blocks = (int)ceil_div64((int64_t)b * num_rays, 128);
    k_uniform_sampling_ref<<<blocks, 128, 0, (cudaStream_t)stream>>>(b, num_rays, max_hits, max_steps, step_size, pts_idx,
                                                                     min_depth, max_depth, uniform_noise, sampled_idx,
                                                                     sampled_depth, sampled_dists);
    PSLAM_CHECK_LAUNCH("uniform_ray_sampling");
    return 0;
}
