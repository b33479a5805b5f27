// Kernel 1: ray vs sparse-voxel-octree intersection (+ the other `grid`
// intersectors kept for API completeness).
//
// Semantics follow the reference kernel
// third_party/sparse_voxels/src/intersect_gpu.cu:75-140 (slab test) and
// :191-270 (DFS): root is row 0, a node is a leaf iff children[k,8]==1, the box
// half-size is 0.5*voxelsize*children[k,8], children are pushed 0..7 so child 7
// is popped first, hits are appended in that DFS order up to n_max.  The slab
// arithmetic keeps the reference's operation order and its __fdividef
// reciprocal so depths are bit-identical.
//
// What is different (B200-first): one thread per ray over a flat grid sized to
// the ray count instead of 256 blocks x <=32 threads over a 256x replicated
// octree; the ray, its reciprocal direction and the traversal stack live in
// registers / shared memory (stack[level][thread], conflict-free) instead of a
// zero-filled 1 KB local array with the ray re-read from global per node; the
// octree is read once through the read-only path and stays L1/L2 resident.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace pslam {

constexpr int kSmemStack = 64;    // 7*L+1 entries cover L <= 9 levels (grid_dim <= 512)
constexpr int kSpillStack = 192;  // beyond that: local memory, total 256 as intersect_gpu.cu:229
constexpr int kIntersectThreads = 128;

struct Ray {
    float o[3], inv[3];
};

__device__ __forceinline__ Ray load_ray(const float *__restrict__ ray_start, const float *__restrict__ ray_dir, int64_t r)
{
    Ray ray;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ray.o[a] = __ldg(ray_start + r * 3 + a);
        ray.inv[a] = __fdividef(1.0f, __ldg(ray_dir + r * 3 + a));  // intersect_gpu.cu:91-101
    }
    return ray;
}

// intersect_gpu.cu:75-140.  Explicit _rn intrinsics pin the reference's
// (c -/+ h - o) * inv evaluation order (no FMA contraction is possible there
// either: SASS of the reference shows FADD,FADD,FMUL).
__device__ __forceinline__ bool slab(const Ray &ray, float cx, float cy, float cz, float half, float &t_lo, float &t_hi)
{
    float lo = 0.0f, hi = 100000.0f;
    const float c[3] = {cx, cy, cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float d_lo = __fmul_rn(__fsub_rn(__fsub_rn(c[a], half), ray.o[a]), ray.inv[a]);
        float d_hi = __fmul_rn(__fsub_rn(__fadd_rn(c[a], half), ray.o[a]), ray.inv[a]);
        if (d_hi < d_lo) { float t = d_lo; d_lo = d_hi; d_hi = t; }
        if (d_hi < lo) return false;
        if (d_lo > hi) return false;
        lo = (d_lo > lo) ? d_lo : lo;
        hi = (d_hi < hi) ? d_hi : hi;
        if (lo > hi) return false;
    }
    t_lo = lo; t_hi = hi;
    return lo > -1.0f;  // "depths.x > -1.0f", intersect_gpu.cu:247
}

// DFS over the flattened octree.  `emit(cnt, k, lo, hi)` stores hit number cnt.
// Returns the hit count, or -1-count when the stack overflowed 256 entries (the
// reference asserts there, intersect_gpu.cu:235).
template <class Emit>
__device__ __forceinline__ int dfs(const Ray &ray, const float *__restrict__ points, const int *__restrict__ children,
                                   float half_voxel, int n_max, int *s_stack, Emit emit)
{
    int spill[kSpillStack];
    const int tid = threadIdx.x, nthr = blockDim.x;
    int top = 0, cnt = 0;
    s_stack[tid] = 0;  // root = row 0, intersect_gpu.cu:232
    while (top > -1 && cnt < n_max) {
        const int k = (top < kSmemStack) ? s_stack[top * nthr + tid] : spill[top - kSmemStack];
        --top;
        const int *ch = children + (int64_t)k * 9;
        const int side = __ldg(ch + 8);
        float lo, hi;
        if (!slab(ray, __ldg(points + (int64_t)k * 3), __ldg(points + (int64_t)k * 3 + 1), __ldg(points + (int64_t)k * 3 + 2),
                  __fmul_rn(half_voxel, (float)side), lo, hi))
            continue;
        if (side == 1) {  // terminal node, intersect_gpu.cu:250
            emit(cnt, k, lo, hi);
            ++cnt;
            continue;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = __ldg(ch + u);
            if (c > -1) {
                ++top;
                if (top < kSmemStack) s_stack[top * nthr + tid] = c;
                else if (top < kSmemStack + kSpillStack) spill[top - kSmemStack] = c;
                else return -1 - cnt;
            }
        }
    }
    return cnt;
}

// ---- reference layout (grid.svo_intersect) --------------------------------------------------
__global__ void __launch_bounds__(kIntersectThreads)
k_svo_intersect_ref(int b, int n, int m, float half_voxel, int n_max, const float *__restrict__ ray_start,
                    const float *__restrict__ ray_dir, const float *__restrict__ points,
                    const int *__restrict__ children, int *__restrict__ idx, float *__restrict__ min_depth,
                    float *__restrict__ max_depth)
{
    extern __shared__ int s_stack[];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * m) return;
    const int bi = (int)(r / m);
    const Ray ray = load_ray(ray_start, ray_dir, r);
    int *oi = idx + r * n_max;
    float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
    int cnt = dfs(ray, points + (int64_t)bi * n * 3, children + (int64_t)bi * n * 9, half_voxel, n_max, s_stack,
                  [&](int c, int k, float lo, float hi) { oi[c] = k; omin[c] = lo; omax[c] = hi; });
    if (cnt < 0) cnt = -1 - cnt;
    for (int l = cnt; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.0f; omax[l] = 0.0f; }  // intersect.cpp:98-106
}

// ---- fused layout: slot-major [n_max, R], sorted by entry depth, trimmed at max_distance ------
// Does the work of voxel_helpers.py:571-588 (fill/sort/gather/trim) per ray in the same kernel.
//
// ONE WARP PER RAY.  The reference's walk is a DFS: one node per step, a chain of dependent loads as long as the number of
// nodes a ray tests.  What the rest of the pipeline consumes, though, is the SET of hit leaves -- sorted by entry depth, ties
// in DFS order, cut to the first n_max in DFS order -- and the DFS order of two leaves is a static property of their paths:
// children are pushed 0..7 and popped 7 first (intersect_gpu.cu:259-265), all leaves sit at the same depth, so leaf a is
// emitted before leaf b iff its path of child indices is lexicographically LARGER.  Every hit carries that path as a key
// (3 bits per level) and the walk is free to run in any order:
//   * the four 8-lane groups of the warp expand up to four pending internal nodes per trip, lane c of a group loading child
//     c's record (id + centre, one 16-byte load: k_build_child_records) and running the slab test -- a ray needs about one
//     trip per octree level instead of one per hit internal node (8-10 instead of ~25 dependent round trips);
//   * hit leaves go straight to the ray's hit list (ballot-compacted), hit internal nodes back on the pending stack;
//   * more than n_max hit leaves: the list keeps the n_max largest keys (= the first n_max of the DFS), replacing its
//     smallest key -- rare, taken one candidate at a time;
//   * at the end every lane ranks its hits by (entry depth, then key descending) = the reference's stable sort of the DFS
//     emission order, and stores them at their rank (hits entering beyond max_distance are the tail of that order and are
//     dropped, voxel_helpers.py:578).
// The slab arithmetic, the child's half size (half_voxel * side, side halving per level exactly like children[c*9+8]) and the
// "depths.x > -1" acceptance are the reference's, so depths stay bit-identical.
constexpr int kWarpRays = 8;                     // rays (= warps) per block
constexpr int kRayThreads = kWarpRays * 32;
constexpr int kCompactRays = 32;                 // rays per block of the compaction that follows (block_hits granularity)
constexpr int kRayStack = 64;                    // pending internal nodes per ray
constexpr int kRayHits = 50;                     // == n_max of the reference (voxel_helpers.py:561)
constexpr size_t kRaySmem = (size_t)kWarpRays * (kRayStack * (4 + 8) + kRayHits * (4 + 4 + 4 + 8));   // x rays per warp

// First warp of a step: folds the previous step's overflow flags (and "no ray hit") into the sticky slot, counts the step and
// clears the per-step counters (include/proud_slam_b200.h: PSLAM_C_STICKY).
__device__ __forceinline__ void step_begin_counters(int *__restrict__ counters, int t)
{
    const int old = t < PSLAM_C_COUNT ? counters[t] : 0;
    const int ov = __shfl_sync(0xffffffffu, old, PSLAM_C_OVERFLOW), rh = __shfl_sync(0xffffffffu, old, PSLAM_C_RH);
    const int steps = __shfl_sync(0xffffffffu, old, PSLAM_C_STEPS);
    if (t < PSLAM_C_COUNT) {
        int v = 0;
        if (t == PSLAM_C_STICKY) v = old | ov | ((steps > 0 && rh == 0) ? 8 : 0);
        if (t == PSLAM_C_STEPS) v = old + 1;
        counters[t] = v;
    }
}
__global__ void k_step_begin(int *__restrict__ counters, int *__restrict__ block_hits, int nbh)
{
    pdl_enter();
    if (threadIdx.x < 32) step_begin_counters(counters, threadIdx.x);
    for (int i = threadIdx.x; i < nbh; i += blockDim.x) block_hits[i] = 0;   // hit-ray counts per kCompactRays rays (k_intersect_* add)
}

// Child records for the walk: rec[node][c] = (row id of child c or -1, centre of that child).  Expanding a node then
// costs ONE dependent 16-byte load per lane instead of two (child id, then its centre); a node's eight records are one
// 128-byte line.  Built once per map generation (PSLAM_F_NODE_CACHE_VALID tells the step that the table is current).
// The row id takes the low 24 bits of .x (the cached walks need N <= 2^24), the top 8 bits say which children the CHILD has.
__global__ void k_build_child_records(int N, const float *__restrict__ points, const int *__restrict__ children, int4 *__restrict__ rec,
                                      int *__restrict__ counters, int *__restrict__ block_hits, int nbh)
{
    pdl_enter();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 32 && counters) step_begin_counters(counters, t);   // first kernel of a step: the step's counters (instead of a memset node in front of the chain)
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < nbh; i += blockDim.x) block_hits[i] = 0;
    if (t >= N * 8) return;
    const int node = t >> 3, c = t & 7;
    const int cid = __ldg(children + (int64_t)node * 9 + c);
    int4 r = make_int4(-1, 0, 0, 0);
    if (cid > -1) {
        // .x = child row (24 bits) | which of ITS children exist (8 bits): the level-synchronous walk queues existing children only
        unsigned mask = 0u;
        if (__ldg(children + (int64_t)cid * 9 + 8) > 1)
            for (int k = 0; k < 8; ++k) mask |= (__ldg(children + (int64_t)cid * 9 + k) > -1 ? 1u : 0u) << k;
        r.x = (int)((unsigned)cid | (mask << 24));
        r.y = __float_as_int(__ldg(points + (int64_t)cid * 3));
        r.z = __float_as_int(__ldg(points + (int64_t)cid * 3 + 1));
        r.w = __float_as_int(__ldg(points + (int64_t)cid * 3 + 2));
    }
    rec[t] = r;
}

// optional per-warp timeline of the walk (pslam_debug_intersect_trace): [block][warp][8] = globaltimer at entry, clock64 at entry /
// after the walk / after the ranking + write-out / at exit, globaltimer at exit, trips of the walk, hit count
__device__ long long *g_intersect_trace = nullptr;
__device__ __forceinline__ void intersect_stamp(long long *tr, int slot, long long v)
{
    if (tr && (threadIdx.x & 31) == 0) tr[((size_t)blockIdx.x * kWarpRays + (threadIdx.x >> 5)) * 8 + slot] = v;   // [block][warp][8]
}
__device__ __forceinline__ long long intersect_globaltimer()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// intersect_gpu.cu:75-140 without branches: the same operations in the same order; once a test has failed the later
// updates of lo / hi are dead, so carrying them on changes nothing that is read.
__device__ __forceinline__ bool slab_nb(const Ray &ray, float cx, float cy, float cz, float half, float &t_lo, float &t_hi)
{
    float lo = 0.0f, hi = 100000.0f;
    bool ok = true;
    const float c[3] = {cx, cy, cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float d_lo = __fmul_rn(__fsub_rn(__fsub_rn(c[a], half), ray.o[a]), ray.inv[a]);
        float d_hi = __fmul_rn(__fsub_rn(__fadd_rn(c[a], half), ray.o[a]), ray.inv[a]);
        const bool sw = d_hi < d_lo;
        const float t0 = sw ? d_hi : d_lo, t1 = sw ? d_lo : d_hi;
        ok = ok && !(t1 < lo) && !(t0 > hi);
        lo = (t0 > lo) ? t0 : lo;
        hi = (t1 < hi) ? t1 : hi;
        ok = ok && !(lo > hi);
    }
    t_lo = lo; t_hi = hi;
    return ok && lo > -1.0f;  // "depths.x > -1.0f", intersect_gpu.cu:247
}

// RPW rays per warp: a ray owns 32 / RPW lanes = 4 / RPW nodes per trip (RPW = 1: shortest dependent chain, for small batches;
// RPW = 2 / 4: better lane utilisation near the root, where a ray has one or two pending nodes, for large ones).
template <bool CACHED, int RPW>
__global__ void __launch_bounds__(kRayThreads, 6)
k_intersect_warp(int R, float half_voxel, int n_max, float max_distance, const float *__restrict__ ray_start,
                 const float *__restrict__ ray_dir, const float *__restrict__ points, const int *__restrict__ children,
                 const int4 *__restrict__ rec, int *__restrict__ hit_idx, float *__restrict__ hit_min, float *__restrict__ hit_max,
                 int *__restrict__ hit_count, int *__restrict__ block_hits, int *__restrict__ counters)
{
    pdl_enter();
    constexpr int SEG = 32 / RPW;                         // lanes of a ray
    constexpr int NPT = SEG / 8;                          // nodes a ray expands per trip
    constexpr int RAYS = kWarpRays * RPW;                 // rays of a block
    long long *const tr = g_intersect_trace;
    if (tr) { intersect_stamp(tr, 0, intersect_globaltimer()); intersect_stamp(tr, 1, clock64()); }
    extern __shared__ __align__(16) unsigned char s_ray[];
    __shared__ int s_blockmax;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = lane / SEG, sl = lane % SEG;          // ray of the warp, lane of the ray
    const int grp = sl >> 3, sub = sl & 7;
    const int slot = warp * RPW + seg;                    // ray of the block
    unsigned long long *st_key = reinterpret_cast<unsigned long long *>(s_ray) + slot * kRayStack;          // [ray][kRayStack]
    unsigned long long *h_key = reinterpret_cast<unsigned long long *>(s_ray) + RAYS * kRayStack + slot * kRayHits;
    int *ibase = reinterpret_cast<int *>(s_ray + (size_t)RAYS * (kRayStack + kRayHits) * 8);
    int *st_id = ibase + slot * kRayStack;                // id | log2(side) << 26
    int *h_id = ibase + RAYS * kRayStack + slot * kRayHits;
    float *h_lo = reinterpret_cast<float *>(ibase + RAYS * kRayStack + RAYS * kRayHits) + slot * kRayHits;
    float *h_hi = h_lo + RAYS * kRayHits;
    if (threadIdx.x == 0) s_blockmax = 0;

    const int r = blockIdx.x * RAYS + slot;
    const bool valid = r < R;
    if (n_max > kRayHits) n_max = kRayHits;
    int top = 0, cnt = 0, trips = 0;
    bool overflow = false;
    Ray ray{};
    if (valid) {
        ray = load_ray(ray_start, ray_dir, r);
        // the root (row 0, intersect_gpu.cu:232)
        const int side = __ldg(children + 8);
        float lo, hi;
        if (slab(ray, __ldg(points), __ldg(points + 1), __ldg(points + 2), __fmul_rn(half_voxel, (float)side), lo, hi)) {
            if (side == 1) {
                if (sl == 0) { h_id[0] = 0; h_lo[0] = lo; h_hi[0] = hi; h_key[0] = 0ull; }
                cnt = 1;
            } else {
                if (sl == 0) { st_id[0] = (31 - __clz(side)) << 26; st_key[0] = 0ull; }
                top = 1;
            }
        }
    }
    __syncwarp();
    const unsigned seg_mask = (SEG == 32) ? 0xffffffffu : (((1u << SEG) - 1u) << (seg * SEG));
    const unsigned lt = ((1u << lane) - 1u) & seg_mask;    // earlier lanes of this ray
    while (__any_sync(0xffffffffu, top > 0)) {
        ++trips;
        const int take = min(top, NPT);
        const bool have = grp < take;
        int e = 0;
        unsigned long long key = 0ull;
        if (have) { e = st_id[top - 1 - grp]; key = st_key[top - 1 - grp]; }
        top -= take;
        __syncwarp();                                        // the slots just read may be overwritten below
        const int level = e >> 26, cur = e & 0x3FFFFFF;
        const int cside = (1 << level) >> 1;
        bool hit = false;
        int cid = -1;
        float lo = 0.f, hi = 0.f;
        if (have) {
            if (CACHED) {
                const int4 r4 = __ldg(rec + (int64_t)cur * 8 + sub);
                cid = r4.x == -1 ? -1 : (r4.x & 0xFFFFFF);
                hit = slab_nb(ray, __int_as_float(r4.y), __int_as_float(r4.z), __int_as_float(r4.w), __fmul_rn(half_voxel, (float)cside), lo, hi) && cid > -1;
            } else {
                cid = __ldg(children + (int64_t)cur * 9 + sub);
                if (cid > -1)
                    hit = slab_nb(ray, __ldg(points + (int64_t)cid * 3), __ldg(points + (int64_t)cid * 3 + 1), __ldg(points + (int64_t)cid * 3 + 2),
                                  __fmul_rn(half_voxel, (float)cside), lo, hi);
            }
        }
        const bool leaf = hit && cside == 1, inner = hit && cside > 1;
        const unsigned long long ckey = (key << 3) | (unsigned long long)sub;
        const unsigned m_in = __ballot_sync(0xffffffffu, inner) & seg_mask, m_leaf = __ballot_sync(0xffffffffu, leaf) & seg_mask;
        if (inner) {
            const int pos = top + __popc(m_in & lt);
            if (pos < kRayStack) {
                st_id[pos] = cid | ((level - 1) << 26);
                st_key[pos] = ckey;
                if (CACHED) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (int64_t)cid * 8));   // its records: one 128-byte line
            } else overflow = true;
        }
        top = min(top + __popc(m_in), kRayStack);
        const int n_new = __popc(m_leaf);
        const bool fits = cnt + n_new <= n_max;
        if (fits) {
            if (leaf) {
                const int pos = cnt + __popc(m_leaf & lt);
                h_id[pos] = cid; h_lo[pos] = lo; h_hi[pos] = hi; h_key[pos] = ckey;
            }
            cnt += n_new;
        }
        // more hit leaves than slots: keep the n_max largest keys = the first n_max leaves of the reference's DFS (rare; one
        // candidate per ray and round, every lane of the warp takes part in the shuffles)
        unsigned rest = fits ? 0u : m_leaf;
        while (__any_sync(0xffffffffu, rest != 0u)) {
            const bool act = rest != 0u;
            const int src = act ? __ffs(rest) - 1 : lane;
            rest &= rest - 1;
            const int c_id = __shfl_sync(0xffffffffu, cid, src);
            const float c_lo = __shfl_sync(0xffffffffu, lo, src), c_hi = __shfl_sync(0xffffffffu, hi, src);
            const unsigned long long c_key = __shfl_sync(0xffffffffu, ckey, src);
            unsigned long long mk = ~0ull;
            int mi = -1;
            if (act && cnt >= n_max) {
                for (int i = sl; i < n_max; i += SEG) {
                    const unsigned long long k2 = h_key[i];
                    if (k2 < mk) { mk = k2; mi = i; }
                }
            }
#pragma unroll
            for (int o = SEG / 2; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                if (ok < mk) { mk = ok; mi = oi; }               // keys of distinct leaves are distinct
            }
            if (act) {
                if (cnt < n_max) {
                    if (sl == 0) { h_id[cnt] = c_id; h_lo[cnt] = c_lo; h_hi[cnt] = c_hi; h_key[cnt] = c_key; }
                    ++cnt;
                } else if (c_key > mk && sl == 0) { h_id[mi] = c_id; h_lo[mi] = c_lo; h_hi[mi] = c_hi; h_key[mi] = c_key; }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    if (tr) intersect_stamp(tr, 2, clock64());
    // rank = position in the stable sort by entry depth of the DFS emission order (SURVEY A-Q3), then the max_distance trim
    int count = 0;
    if (valid) {
        for (int i = sl; i < cnt; i += SEG) {
            const float li = h_lo[i];
            const unsigned long long ki = h_key[i];
            int rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const float lj = h_lo[j];
                rank += (lj < li || (lj == li && h_key[j] > ki)) ? 1 : 0;
            }
            if (!(li > max_distance)) {       // (the dropped hits are the tail of the order: voxel_helpers.py:578)
                const int64_t at = (int64_t)rank * R + r;
                hit_idx[at] = h_id[i]; hit_min[at] = li; hit_max[at] = h_hi[i];
                ++count;
            }
        }
    }
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (valid && sl == 0) hit_count[r] = count;
    if (tr) intersect_stamp(tr, 3, clock64());
    if (sl == 0 && count > 0) atomicMax(&s_blockmax, count);
    if (overflow) atomicOr(counters + PSLAM_C_OVERFLOW, 2);
    const int nhit = __syncthreads_count(sl == 0 && count > 0);
    if (threadIdx.x == 0 && nhit > 0) {
        atomicAdd(block_hits + (blockIdx.x * RAYS) / kCompactRays, nhit);   // (RAYS divides kCompactRays)
        atomicMax(counters + PSLAM_C_P, s_blockmax);
    }
    if (tr) {
        intersect_stamp(tr, 4, clock64()); intersect_stamp(tr, 5, intersect_globaltimer());
        intersect_stamp(tr, 6, trips); intersect_stamp(tr, 7, warp_max_i(count));
    }
}

static int g_walk_mode = 0;
int walk_mode() { return g_walk_mode; }
void set_walk_mode(int m) { g_walk_mode = m ? 1 : 0; }

// ---- block-cooperative level-synchronous walk (PSLAM_OPT_WALK = 1) ---------------------------------------------------------
// The warp-per-ray walk above is bound by instruction issue: near the root a ray has one or two pending nodes, so most lanes of
// its warp idle while the warp still pays the ~180 instructions of a trip.  With the DFS keys the walk can also run breadth
// first for a whole block of rays: the slab tests of ALL its rays at one octree level sit in a shared queue -- one item per
// EXISTING child of a node hit one level up (the child records carry each child's own child mask, so empty octants are never
// tested: ~2.7 tests per expansion instead of 8) -- every lane a test; a hit internal child queues its existing children for
// the next level (warp scan + one atomic per warp), a hit leaf goes to its ray's hit list (one shared-memory atomic for the
// slot).  One trip per octree level for the block.
// A ray with more than n_max hit leaves, or a queue overflow, falls back to the exact warp walk for those rays (walk_ray_warp32:
// the n_max largest keys), so results are the same in every case.
constexpr int kBfsRays = 16, kBfsThreads = 128, kBfsQueue = 1024;
constexpr size_t kBfsSmem = sizeof(unsigned long long) * (2 * kBfsQueue + kBfsRays * kRayHits + (kBfsThreads / 32) * kRayStack) +
                            sizeof(int) * (2 * kBfsQueue + 3 * kBfsRays * kRayHits + (kBfsThreads / 32) * kRayStack + kBfsRays * 8 + 16);

// the exact walk of ONE ray by a whole warp (k_intersect_warp with one ray per warp) into the ray's hit arrays; returns the hit count
template <bool CACHED>
__device__ __forceinline__ int walk_ray_warp32(const Ray &ray, float half_voxel, int n_max, const float *__restrict__ points,
                                               const int *__restrict__ children, const int4 *__restrict__ rec, int *st_id,
                                               unsigned long long *st_key, int *h_id, float *h_lo, float *h_hi, unsigned long long *h_key,
                                               bool &overflow)
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    int top = 0, cnt = 0;
    {
        const int side = __ldg(children + 8);
        float lo, hi;
        if (slab(ray, __ldg(points), __ldg(points + 1), __ldg(points + 2), __fmul_rn(half_voxel, (float)side), lo, hi)) {
            if (side == 1) {
                if (lane == 0) { h_id[0] = 0; h_lo[0] = lo; h_hi[0] = hi; h_key[0] = 0ull; }
                cnt = 1;
            } else {
                if (lane == 0) { st_id[0] = (31 - __clz(side)) << 26; st_key[0] = 0ull; }
                top = 1;
            }
        }
    }
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    while (top > 0) {
        const int take = min(top, 4);
        const bool have = grp < take;
        int e = 0;
        unsigned long long key = 0ull;
        if (have) { e = st_id[top - 1 - grp]; key = st_key[top - 1 - grp]; }
        top -= take;
        __syncwarp();
        const int level = e >> 26, cur = e & 0x3FFFFFF;
        const int cside = (1 << level) >> 1;
        bool hit = false;
        int cid = -1;
        float lo = 0.f, hi = 0.f;
        if (have) {
            if (CACHED) {
                const int4 r4 = __ldg(rec + (int64_t)cur * 8 + sub);
                cid = r4.x == -1 ? -1 : (r4.x & 0xFFFFFF);
                hit = slab_nb(ray, __int_as_float(r4.y), __int_as_float(r4.z), __int_as_float(r4.w), __fmul_rn(half_voxel, (float)cside), lo, hi) && cid > -1;
            } else {
                cid = __ldg(children + (int64_t)cur * 9 + sub);
                if (cid > -1)
                    hit = slab_nb(ray, __ldg(points + (int64_t)cid * 3), __ldg(points + (int64_t)cid * 3 + 1), __ldg(points + (int64_t)cid * 3 + 2),
                                  __fmul_rn(half_voxel, (float)cside), lo, hi);
            }
        }
        const bool leaf = hit && cside == 1, inner = hit && cside > 1;
        const unsigned long long ckey = (key << 3) | (unsigned long long)sub;
        const unsigned m_in = __ballot_sync(0xffffffffu, inner), m_leaf = __ballot_sync(0xffffffffu, leaf);
        if (inner) {
            const int pos = top + __popc(m_in & lt);
            if (pos < kRayStack) { st_id[pos] = cid | ((level - 1) << 26); st_key[pos] = ckey; }
            else overflow = true;
        }
        top = min(top + __popc(m_in), kRayStack);
        unsigned rest = m_leaf;
        while (rest) {                                        // (uniform: every lane holds the same mask)
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
            const int c_id = __shfl_sync(0xffffffffu, cid, src);
            const float c_lo = __shfl_sync(0xffffffffu, lo, src), c_hi = __shfl_sync(0xffffffffu, hi, src);
            const unsigned long long c_key = __shfl_sync(0xffffffffu, ckey, src);
            if (cnt < n_max) {
                if (lane == 0) { h_id[cnt] = c_id; h_lo[cnt] = c_lo; h_hi[cnt] = c_hi; h_key[cnt] = c_key; }
                ++cnt;
            } else {
                unsigned long long mk = ~0ull;
                int mi = -1;
                for (int i = lane; i < n_max; i += 32) {
                    const unsigned long long k2 = h_key[i];
                    if (k2 < mk) { mk = k2; mi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mk, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
                    if (ok < mk) { mk = ok; mi = oi; }
                }
                if (c_key > mk && lane == 0) { h_id[mi] = c_id; h_lo[mi] = c_lo; h_hi[mi] = c_hi; h_key[mi] = c_key; }
            }
            __syncwarp();
        }
        __syncwarp();
    }
    return cnt;
}

template <bool CACHED>
__global__ void __launch_bounds__(kBfsThreads)
k_intersect_bfs(int R, float half_voxel, int n_max, float max_distance, const float *__restrict__ ray_start,
                const float *__restrict__ ray_dir, const float *__restrict__ points, const int *__restrict__ children,
                const int4 *__restrict__ rec, int *__restrict__ hit_idx, float *__restrict__ hit_min, float *__restrict__ hit_max,
                int *__restrict__ hit_count, int *__restrict__ block_hits, int *__restrict__ counters)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char s_bfs[];
    unsigned long long *q_key = reinterpret_cast<unsigned long long *>(s_bfs);                  // [2][kBfsQueue]
    unsigned long long *h_key = q_key + 2 * kBfsQueue;                                           // [rays][kRayHits]
    unsigned long long *f_key = h_key + kBfsRays * kRayHits;                                     // [warps][kRayStack] (fallback walk)
    int *q_node = reinterpret_cast<int *>(f_key + (kBfsThreads / 32) * kRayStack);               // [2][kBfsQueue]: ray << 26 | node row
    int *h_id = q_node + 2 * kBfsQueue;                                                          // [rays][kRayHits]
    float *h_lo = reinterpret_cast<float *>(h_id + kBfsRays * kRayHits);
    float *h_hi = h_lo + kBfsRays * kRayHits;
    int *f_id = reinterpret_cast<int *>(h_hi + kBfsRays * kRayHits);                             // [warps][kRayStack]
    float *s_ray = reinterpret_cast<float *>(f_id + (kBfsThreads / 32) * kRayStack);             // [rays][8]: o, inv
    int *h_cnt = reinterpret_cast<int *>(s_ray + kBfsRays * 8);                                  // [rays] hits found (may exceed n_max)
    __shared__ int s_misc[8];                                                                    // queue lengths [2], queue overflow, block max, hit rays
    int *q_n = s_misc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * kBfsRays;
    if (n_max > kRayHits) n_max = kRayHits;
    if (tid < 8) s_misc[tid] = 0;
    __syncthreads();
    const int root_side = __ldg(children + 8);
    int level = 31 - __clz(root_side);
    // ---- rays and the root (row 0, intersect_gpu.cu:232): a hit internal root queues its existing children ----
    // queue item = one slab test: ray (5 bits) | child index c (3 bits) | PARENT row (24 bits), with the key of the child
    if (tid < kBfsRays) {
        const int r = r0 + tid;
        int cnt = 0;
        if (r < R) {
            const Ray ray = load_ray(ray_start, ray_dir, r);
#pragma unroll
            for (int a = 0; a < 3; ++a) { s_ray[tid * 8 + a] = ray.o[a]; s_ray[tid * 8 + 4 + a] = ray.inv[a]; }
            float lo, hi;
            if (slab(ray, __ldg(points), __ldg(points + 1), __ldg(points + 2), __fmul_rn(half_voxel, (float)root_side), lo, hi)) {
                if (root_side == 1) {
                    h_id[tid * kRayHits] = 0; h_lo[tid * kRayHits] = lo; h_hi[tid * kRayHits] = hi; h_key[tid * kRayHits] = 0ull;
                    cnt = 1;
                } else {
                    for (int c = 0; c < 8; ++c) {
                        if (__ldg(&rec[c].x) == -1) continue;
                        const int at = atomicAdd(q_n, 1);
                        q_node[at] = (tid << 27) | (c << 24);          // parent = row 0
                        q_key[at] = (unsigned long long)c;
                    }
                }
            }
        }
        h_cnt[tid] = cnt;
    }
    __syncthreads();
    // ---- one trip per octree level: the items of a level are the existing children of the nodes hit one level up ----
    int cur = 0;
    for (; level >= 1; --level) {
        const int nq = q_n[cur];
        if (nq == 0) break;
        const int cside = (1 << level) >> 1;
        const float half = __fmul_rn(half_voxel, (float)cside);
        int *qn_next = q_n + (cur ^ 1);
        for (int base = 0; base < nq; base += kBfsThreads) {
            const int t = base + tid;
            bool hit = false;
            int cid = -1, rl = 0;
            unsigned cmask = 0u;
            float lo = 0.f, hi = 0.f;
            unsigned long long ckey = 0ull;
            if (t < nq) {
                const unsigned e = (unsigned)q_node[cur * kBfsQueue + t];
                rl = (int)(e >> 27);
                ckey = q_key[cur * kBfsQueue + t];
                Ray ray;
#pragma unroll
                for (int a = 0; a < 3; ++a) { ray.o[a] = s_ray[rl * 8 + a]; ray.inv[a] = s_ray[rl * 8 + 4 + a]; }
                const int4 r4 = __ldg(rec + (int64_t)(e & 0xFFFFFFu) * 8 + ((e >> 24) & 7u));
                cid = r4.x & 0xFFFFFF;
                cmask = (unsigned)r4.x >> 24;
                hit = slab_nb(ray, __int_as_float(r4.y), __int_as_float(r4.z), __int_as_float(r4.w), half, lo, hi);
            }
            if (cside == 1) {                                 // leaves: straight to the ray's hit list
                if (hit) {
                    const int slot = atomicAdd(h_cnt + rl, 1);
                    if (slot < n_max) {
                        h_id[rl * kRayHits + slot] = cid; h_lo[rl * kRayHits + slot] = lo; h_hi[rl * kRayHits + slot] = hi; h_key[rl * kRayHits + slot] = ckey;
                    }
                }
            } else {                                          // a hit internal node queues its existing children (one atomic per warp)
                const int n_push = hit ? __popc(cmask) : 0;
                int x = n_push;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                const int total = __shfl_sync(0xffffffffu, x, 31);
                int at = 0;
                if (lane == 31 && total) at = atomicAdd(qn_next, total);
                at = __shfl_sync(0xffffffffu, at, 31) + x - n_push;
                unsigned m = hit ? cmask : 0u;
                while (m) {
                    const int c = __ffs(m) - 1;
                    m &= m - 1;
                    if (at < kBfsQueue) {
                        q_node[(cur ^ 1) * kBfsQueue + at] = (rl << 27) | (c << 24) | cid;
                        q_key[(cur ^ 1) * kBfsQueue + at] = (ckey << 3) | (unsigned long long)c;
                    } else s_misc[2] = 1;                     // queue overflow: every ray of the block takes the exact walk
                    ++at;
                }
            }
        }
        __syncthreads();
        if (tid == 0) { q_n[cur] = 0; if (q_n[cur ^ 1] > kBfsQueue) q_n[cur ^ 1] = kBfsQueue; }
        cur ^= 1;
        __syncthreads();
    }
    // ---- per ray: (exact re-walk where needed,) rank by (entry depth, key descending), trim, store ----
    const bool all_exact = s_misc[2] != 0;
    bool overflow = false;
    for (int rl = warp; rl < kBfsRays; rl += kBfsThreads / 32) {
        const int r = r0 + rl;
        if (r >= R) break;
        int *hi_ = h_id + rl * kRayHits;
        float *hl = h_lo + rl * kRayHits, *hh = h_hi + rl * kRayHits;
        unsigned long long *hk = h_key + rl * kRayHits;
        int cnt = h_cnt[rl];
        if (all_exact || cnt > n_max) {
            Ray ray;
#pragma unroll
            for (int a = 0; a < 3; ++a) { ray.o[a] = s_ray[rl * 8 + a]; ray.inv[a] = s_ray[rl * 8 + 4 + a]; }
            __syncwarp();
            cnt = walk_ray_warp32<CACHED>(ray, half_voxel, n_max, points, children, rec, f_id + warp * kRayStack, f_key + warp * kRayStack, hi_, hl, hh, hk, overflow);
            __syncwarp();
        }
        int count = 0;
        for (int i = lane; i < cnt; i += 32) {
            const float li = hl[i];
            const unsigned long long ki = hk[i];
            int rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const float lj = hl[j];
                rank += (lj < li || (lj == li && hk[j] > ki)) ? 1 : 0;
            }
            if (!(li > max_distance)) {       // (the dropped hits are the tail of the order: voxel_helpers.py:578)
                const int64_t at = (int64_t)rank * R + r;
                hit_idx[at] = hi_[i]; hit_min[at] = li; hit_max[at] = hh[i];
                ++count;
            }
        }
        count = warp_sum_i(count);
        if (lane == 0) {
            hit_count[r] = count;
            if (count > 0) { atomicAdd(s_misc + 4, 1); atomicMax(s_misc + 3, count); }
        }
    }
    if (overflow) atomicOr(counters + PSLAM_C_OVERFLOW, 2);
    __syncthreads();
    if (tid == 0 && s_misc[4] > 0) {
        atomicAdd(block_hits + (blockIdx.x * kBfsRays) / kCompactRays, s_misc[4]);   // (kBfsRays divides kCompactRays)
        atomicMax(counters + PSLAM_C_P, s_misc[3]);
    }
}

// Pulls the flattened octree into L2 before the latency-bound traversal: the DFS is a chain of ~70
// dependent loads per ray, so every miss to HBM is paid in full.  One 128-byte line per thread.
__global__ void k_prefetch_l2(const char *__restrict__ a, size_t na, const char *__restrict__ b, size_t nb)
{
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (i < na) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + i));
    if (i < nb) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + i));
}

// Exclusive scan of `nb` block partials by one block; total -> *total_out.
__global__ void __launch_bounds__(1024) k_scan_partials(int *__restrict__ partials, int nb, int *__restrict__ total_out)
{
    pdl_enter();
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nb) ? partials[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < nb) partials[i] = carry + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

// rank of every hit ray (order-preserving compaction, render_helpers.py:390-398).
__global__ void __launch_bounds__(kIntersectThreads)
k_compact_rays(int R, const int *__restrict__ hit_count, const int *__restrict__ block_base, int *__restrict__ hit_ray,
               int *__restrict__ ray_rank, int *__restrict__ zero, int zero_n, int *__restrict__ total_out)
{
    pdl_enter();
    __shared__ int s_base;
    // total_out != NULL: block_base holds the raw per-block hit counts and every block sums its predecessors itself (a few
    // hundred L2-resident ints: cheaper than a scan kernel in the chain); NULL: block_base was scanned by k_scan_partials
    if (threadIdx.x < 32) {
        int acc = 0;
        if (total_out) {
            for (int i = threadIdx.x; i < (int)blockIdx.x; i += 32) acc += block_base[i];
            acc = warp_sum_i(acc);
            if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) *total_out = acc + block_base[blockIdx.x];
        } else acc = block_base[blockIdx.x];
        if (threadIdx.x == 0) s_base = acc;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_n; i += gridDim.x * blockDim.x) zero[i] = 0;   // the sampling kernel's look-back state
    __shared__ int s_warp[kIntersectThreads / 32];   // launched with kCompactRays threads per block (<= kIntersectThreads)
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = (r < R) && hit_count[r] > 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int base = s_base;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    const int rank = base + __popc(ballot & ((1u << lane) - 1u));
    if (r < R) ray_rank[r] = hit ? rank : -1;
    if (hit) hit_ray[rank] = r;
}

// ---- API-surface kernels (not perf targets) ---------------------------------------------------
__global__ void k_aabb_intersect_ref(int b, int n, int m, float half_voxel, int n_max,
                                     const float *__restrict__ ray_start, const float *__restrict__ ray_dir,
                                     const float *__restrict__ points, int *__restrict__ idx,
                                     float *__restrict__ min_depth, float *__restrict__ max_depth)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * m) return;
    const float *pts = points + (int64_t)(r / m) * n * 3;
    const Ray ray = load_ray(ray_start, ray_dir, r);
    int *oi = idx + r * n_max;
    float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
    int cnt = 0;
    for (int k = 0; k < n && cnt < n_max; ++k) {  // intersect_gpu.cu:172-187
        float lo, hi;
        if (slab(ray, __ldg(pts + k * 3), __ldg(pts + k * 3 + 1), __ldg(pts + k * 3 + 2), half_voxel, lo, hi)) {
            oi[cnt] = k; omin[cnt] = lo; omax[cnt] = hi; ++cnt;
        }
    }
    for (int l = cnt; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.0f; omax[l] = 0.0f; }
}

__global__ void k_ball_intersect_ref(int b, int n, int m, float radius, int n_max,
                                     const float *__restrict__ ray_start, const float *__restrict__ ray_dir,
                                     const float *__restrict__ points, int *__restrict__ idx,
                                     float *__restrict__ min_depth, float *__restrict__ max_depth)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * m) return;
    const float *pts = points + (int64_t)(r / m) * n * 3;
    const float x0 = ray_start[r * 3], y0 = ray_start[r * 3 + 1], z0 = ray_start[r * 3 + 2];
    const float xw = ray_dir[r * 3], yw = ray_dir[r * 3 + 1], zw = ray_dir[r * 3 + 2];
    const float radius2 = radius * radius;
    int *oi = idx + r * n_max;
    float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
    int cnt = 0;
    for (int k = 0; k < n && cnt < n_max; ++k) {  // intersect_gpu.cu:51-71
        const float x = pts[k * 3] - x0, y = pts[k * 3 + 1] - y0, z = pts[k * 3 + 2] - z0;
        const float d2 = x * x + y * y + z * z;
        const float proj = x * xw + y * yw + z * zw;
        const float d2_proj = proj * proj;  // pow(., 2)
        const float r2 = d2 - d2_proj;
        if (r2 < radius2) {
            const float depth = sqrtf(d2_proj), blur = sqrtf(radius2 - r2);
            oi[cnt] = k; omin[cnt] = depth - blur; omax[cnt] = depth + blur; ++cnt;
        }
    }
    for (int l = cnt; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.0f; omax[l] = 0.0f; }
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 sub3(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 cross3(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__global__ void k_triangle_intersect_ref(int b, int n, int m, float cagesize, float blur, int n_max,
                                         const float *__restrict__ ray_start, const float *__restrict__ ray_dir,
                                         const float *__restrict__ face_points, int *__restrict__ idx,
                                         float *__restrict__ depth, float *__restrict__ uv)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (int64_t)b * m) return;
    const float *fp = face_points + (int64_t)(r / m) * n * 9;
    const F3 ori = {ray_start[r * 3], ray_start[r * 3 + 1], ray_start[r * 3 + 2]};
    const F3 dir = {ray_dir[r * 3], ray_dir[r * 3 + 1], ray_dir[r * 3 + 2]};
    int *oi = idx + r * n_max;
    float *od = depth + r * n_max * 3, *ouv = uv + r * n_max * 2;
    for (int l = 0; l < n_max; ++l) { oi[l] = -1; od[l * 3] = od[l * 3 + 1] = od[l * 3 + 2] = 0.0f; ouv[l * 2] = ouv[l * 2 + 1] = 0.0f; }
    int cnt = 0;
    for (int k = 0; k < n && cnt < n_max; ++k) {
        // Moeller-Trumbore, intersect_gpu.cu:272-303
        const F3 v0 = {fp[k * 9], fp[k * 9 + 1], fp[k * 9 + 2]};
        const F3 v0v1 = sub3({fp[k * 9 + 3], fp[k * 9 + 4], fp[k * 9 + 5]}, v0);
        const F3 v0v2 = sub3({fp[k * 9 + 6], fp[k * 9 + 7], fp[k * 9 + 8]}, v0);
        const F3 v0O = sub3(ori, v0);
        const F3 dxe = cross3(dir, v0v2);
        const float det = __fdividef(1.0f, dot3(v0v1, dxe));
        float u = dot3(v0O, dxe) * det;
        if (u < 0.0f - blur || u > 1.0f + blur) continue;
        const F3 oxe = cross3(v0O, v0v1);
        float v = dot3(dir, oxe) * det;
        if (v < 0.0f - blur || v > 1.0f + blur) continue;
        if ((u + v) < 0.0f - blur || (u + v) > 1.0f + blur) continue;
        float d = dot3(v0v2, oxe) * det;
        if (!(d > 0)) continue;
        int ki = k;
        for (int l = 0; l < cnt; ++l)  // insertion by depth, intersect_gpu.cu:356-365
            if (d < od[l * 3]) {
                int ti = oi[l]; oi[l] = ki; ki = ti;
                float t = od[l * 3]; od[l * 3] = d; d = t;
                t = ouv[l * 2]; ouv[l * 2] = u; u = t;
                t = ouv[l * 2 + 1]; ouv[l * 2 + 1] = v; v = t;
            }
        oi[cnt] = ki; od[cnt * 3] = d; ouv[cnt * 2] = u; ouv[cnt * 2 + 1] = v; ++cnt;
    }
    for (int l = 0; l < cnt; ++l) {  // :373-386
        od[l * 3 + 1] = (l == 0) ? -cagesize : -fminf(cagesize, 0.5f * (od[l * 3] - od[l * 3 - 3]));
        od[l * 3 + 2] = (l == cnt - 1) ? cagesize : fminf(cagesize, 0.5f * (od[l * 3 + 3] - od[l * 3]));
    }
}

__global__ void k_debug_rcp(const float *__restrict__ in, float *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fdividef(1.0f, in[i]);
}

}  // namespace pslam

using namespace pslam;

static int check_grouped(int b, int n, int m, int n_max, const void *a, const void *c, const void *d, const void *e)
{
    PSLAM_CHECK_ARG(b > 0 && n > 0 && m > 0 && n_max > 0, PSLAM_E_ARG, "sizes must be positive (b=%d n=%d m=%d n_max=%d)", b, n, m, n_max);
    PSLAM_CHECK_ARG(a && c && d && e, PSLAM_E_ARG, "null pointer argument");
    PSLAM_CHECK_ARG((int64_t)b * m * n_max < (int64_t)1 << 40, PSLAM_E_RANGE, "batch too large");
    return 0;
}

extern "C" int pslam_svo_intersect(int b, int n, int m, float voxelsize, int n_max, const float *ray_start,
                                   const float *ray_dir, const float *points, const int *children, int *idx,
                                   float *min_depth, float *max_depth, pslam_stream_t stream)
{
    if (int rc = check_grouped(b, n, m, n_max, ray_start, ray_dir, points, children)) return rc;
    PSLAM_CHECK_ARG(idx && min_depth && max_depth, PSLAM_E_ARG, "null output pointer");
    const int64_t rays = (int64_t)b * m;
    const int blocks = (int)ceil_div64(rays, kIntersectThreads);
    const size_t smem = sizeof(int) * kSmemStack * kIntersectThreads;
    k_svo_intersect_ref<<<blocks, kIntersectThreads, smem, (cudaStream_t)stream>>>(
        b, n, m, (float)(voxelsize * 0.5), n_max, ray_start, ray_dir, points, children, idx, min_depth, max_depth);
    PSLAM_CHECK_LAUNCH("svo_intersect");
    return 0;
}

extern "C" int pslam_aabb_intersect(int b, int n, int m, float voxelsize, int n_max, const float *ray_start,
                                    const float *ray_dir, const float *points, int *idx, float *min_depth,
                                    float *max_depth, pslam_stream_t stream)
{
    if (int rc = check_grouped(b, n, m, n_max, ray_start, ray_dir, points, idx)) return rc;
    PSLAM_CHECK_ARG(min_depth && max_depth, PSLAM_E_ARG, "null output pointer");
    const int blocks = (int)ceil_div64((int64_t)b * m, 128);
    k_aabb_intersect_ref<<<blocks, 128, 0, (cudaStream_t)stream>>>(b, n, m, (float)(voxelsize * 0.5), n_max, ray_start,
                                                                   ray_dir, points, idx, min_depth, max_depth);
    PSLAM_CHECK_LAUNCH("aabb_intersect");
    return 0;
}

extern "C" int pslam_ball_intersect(int b, int n, int m, float radius, int n_max, const float *ray_start,
                                    const float *ray_dir, const float *points, int *idx, float *min_depth,
                                    float *max_depth, pslam_stream_t stream)
{
    if (int rc = check_grouped(b, n, m, n_max, ray_start, ray_dir, points, idx)) return rc;
    PSLAM_CHECK_ARG(min_depth && max_depth, PSLAM_E_ARG, "null output pointer");
    const int blocks = (int)ceil_div64((int64_t)b * m, 128);
    k_ball_intersect_ref<<<blocks, 128, 0, (cudaStream_t)stream>>>(b, n, m, radius, n_max, ray_start, ray_dir, points, idx,
                                                                   min_depth, max_depth);
    PSLAM_CHECK_LAUNCH("ball_intersect");
    return 0;
}

extern "C" int pslam_triangle_intersect(int b, int n, int m, float cagesize, float blur, int n_max,
                                        const float *ray_start, const float *ray_dir, const float *face_points,
                                        int *idx, float *depth, float *uv, pslam_stream_t stream)
{
    if (int rc = check_grouped(b, n, m, n_max, ray_start, ray_dir, face_points, idx)) return rc;
    PSLAM_CHECK_ARG(depth && uv, PSLAM_E_ARG, "null output pointer");
    const int blocks = (int)ceil_div64((int64_t)b * m, 128);
    k_triangle_intersect_ref<<<blocks, 128, 0, (cudaStream_t)stream>>>(b, n, m, cagesize, blur, n_max, ray_start, ray_dir,
                                                                       face_points, idx, depth, uv);
    PSLAM_CHECK_LAUNCH("triangle_intersect");
    return 0;
}

extern "C" int pslam_debug_intersect_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(g_intersect_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { set_error("intersect_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pslam_debug_rcp(const float *in, float *out, int n, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(in && out && n > 0, PSLAM_E_ARG, "bad argument");
    k_debug_rcp<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
    PSLAM_CHECK_LAUNCH("debug_rcp");
    return 0;
}

// Stage-1a of the fused pipeline.
namespace pslam {
int scan_partials(int *partials, int nb, int *total_out, cudaStream_t st)
{
    launch_chain(k_scan_partials, dim3(1), dim3(1024), 0, st, partials, nb, total_out);
    PSLAM_CHECK_LAUNCH("scan_partials");
    return 0;
}

int launch_intersect_fused(const pslam_render_t *p, cudaStream_t st)
{
    const int nb = ceil_div(p->R, kCompactRays);
    int *block_hits = p->scratch_i;  // [nb]
    // the record table costs a pass over the octree per map generation: worth it while the walk is the larger job
    const bool cached = p->node_cache && p->node_cache_bytes >= (int64_t)128 * p->N && ((uintptr_t)p->node_cache % 16 == 0) && p->N <= (1 << 24);
    static int forced = -1;                                // PSLAM_INTERSECT_RPW=1|2|4: measurement override
    if (forced < 0) {
        const char *e = getenv("PSLAM_INTERSECT_RPW");
        const int v = e ? atoi(e) : 0;
        forced = (v == 1 || v == 2 || v == 4) ? v : 0;
    }
    // a ray per warp while every warp of the launch can be resident at once (the dependent chain is all there is: measured at
    // 8192 rays 26.3 us against 28.5 / 31.7 for two / four rays per warp), more rays per warp beyond that (fewer instructions:
    // near the root a ray has one or two pending nodes and most lanes of a whole warp would idle)
    const int rpw = forced ? forced : (p->R <= 64 * num_sms() ? 1 : (p->R <= 256 * num_sms() ? 2 : 4));
    int4 *rec = static_cast<int4 *>(p->node_cache);
    if (cached && !(p->flags & PSLAM_F_NODE_CACHE_VALID)) {
        // (building the record table reads both arrays and leaves the table in L2; it also opens the step: counters)
        launch_chain(k_build_child_records, dim3((int)ceil_div64((int64_t)p->N * 8, 256)), dim3(256), 0, st, p->N, p->centres, p->structure, rec, p->counters, block_hits, nb);
        PSLAM_CHECK_LAUNCH("build_child_records");
    } else {
        // (an L2 prefetch of the child records / corner ids / embedding rows here was measured: no gain -- the walks are bound by
        //  their dependent instruction chains, not by DRAM round trips)
        launch_chain(k_step_begin, dim3(1), dim3(256), 0, st, p->counters, block_hits, nb);
        PSLAM_CHECK_LAUNCH("step_begin");
    }
    const float half_voxel = (float)(p->voxel_size * 0.5);
#define PSLAM_LAUNCH_INTERSECT(C, W)                                                                                                   \
    do {                                                                                                                               \
        static PerDevice once = {};                                                                                                    \
        bool &configured = once.done[current_device()];                                                                                \
        if (!configured) {                                                                                                             \
            cudaError_t e = cudaFuncSetAttribute(k_intersect_warp<C, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRaySmem * W)); \
            if (e != cudaSuccess) { set_error("intersect: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }          \
            configured = true;                                                                                                         \
        }                                                                                                                              \
        launch_chain(k_intersect_warp<C, W>, dim3(ceil_div(p->R, kWarpRays * W)), dim3(kRayThreads), kRaySmem * W, st, p->R, half_voxel, p->n_max, \
                     p->max_distance, p->rays_o, p->rays_d, p->centres, p->structure, (const int4 *)(C ? rec : nullptr), p->hit_idx, p->hit_min,       \
                     p->hit_max, p->hit_count, block_hits, p->counters);                                                              \
    } while (0)
    // PSLAM_OPT_WALK: 0 (default) = a warp per ray, 1 = the block-cooperative level-synchronous walk.  Measured at 8192 rays:
    // 29 us against 31 us (26 us when it tested all eight child slots): with one trip per octree level and two block barriers per
    // trip its few resident warps wait on their own dependent chains, so fewer instructions did not make it faster.
    if (cached && walk_mode() == 1) {
        static PerDevice bfs_once = {};
        bool &bfs_configured = bfs_once.done[current_device()];
        if (!bfs_configured) {
            cudaError_t e = cudaFuncSetAttribute(k_intersect_bfs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBfsSmem);
            if (e != cudaSuccess) { set_error("intersect: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            bfs_configured = true;
        }
        launch_chain(k_intersect_bfs<true>, dim3(ceil_div(p->R, kBfsRays)), dim3(kBfsThreads), kBfsSmem, st, p->R, half_voxel, p->n_max, p->max_distance,
                     p->rays_o, p->rays_d, p->centres, p->structure, (const int4 *)rec, p->hit_idx, p->hit_min, p->hit_max, p->hit_count, block_hits,
                     p->counters);
    } else if (cached) {
        if (rpw == 1) PSLAM_LAUNCH_INTERSECT(true, 1); else if (rpw == 2) PSLAM_LAUNCH_INTERSECT(true, 2); else PSLAM_LAUNCH_INTERSECT(true, 4);
    } else {
        if (rpw == 1) PSLAM_LAUNCH_INTERSECT(false, 1); else if (rpw == 2) PSLAM_LAUNCH_INTERSECT(false, 2); else PSLAM_LAUNCH_INTERSECT(false, 4);
    }
#undef PSLAM_LAUNCH_INTERSECT
    PSLAM_CHECK_LAUNCH("intersect_warp");
    const bool own_scan = nb <= 1024;                   // beyond that the O(nb^2) sums lose to the scan kernel
    if (!own_scan) { if (int rc = scan_partials(block_hits, nb, p->counters + PSLAM_C_RH, st)) return rc; }
    launch_chain(k_compact_rays, dim3(nb), dim3(kCompactRays), 0, st, p->R, p->hit_count, block_hits, p->hit_ray, p->ray_rank,
                 p->scratch_i + scratch_i_sample_off(p->R), scratch_i_sample_len(p->R), own_scan ? p->counters + PSLAM_C_RH : nullptr);
    PSLAM_CHECK_LAUNCH("compact_rays");
    return 0;
}

// The child-record table alone (a caller that manages map generations itself: pslam_build_node_cache).
int launch_build_node_cache(int N, const float *centres, const int *structure, void *node_cache, cudaStream_t st)
{
    launch_chain(k_build_child_records, dim3((int)ceil_div64((int64_t)N * 8, 256)), dim3(256), 0, st, N, centres, structure,
                 static_cast<int4 *>(node_cache), (int *)nullptr, (int *)nullptr, 0);
    PSLAM_CHECK_LAUNCH("build_child_records");
    return 0;
}
}  // namespace pslam
