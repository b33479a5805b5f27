/*
 * oracle/grid_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, OpenMP over rays) of the two CUDA-only kernels on
 * the Proud-SLAM render path.  The reference has no CPU implementation of
 * them (third_party/sparse_voxels/include/utils.h:10-14 rejects non-CUDA
 * tensors), so this file restates the kernels' arithmetic operation by
 * operation, including the quirks listed in SURVEY.md Appendix A.
 *
 *   oracle_svo_intersect         <- third_party/sparse_voxels/src/intersect_gpu.cu:75-140 (slab test)
 *                                   and :191-270 (DFS over the flattened octree)
 *   oracle_inverse_cdf_sampling  <- third_party/sparse_voxels/src/sample_gpu.cu:133-239
 *   oracle_aabb_intersect        <- intersect_gpu.cu:142-189
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (proud_slam_b200/) never does.
 *
 * Parity status: the reference ships no golden vectors for this path
 * (SURVEY.md section 4), so this oracle is pinned (a) against the reference's
 * own CUDA `grid` extension built from /root/reference into oracle/_ref and
 * run on the GPU box (tests/test_gpu_ref_grid.py), and (b) through the
 * reference's own Python (imported from /root/reference in the build
 * container) by tests/golden/make_golden.py.
 *
 * One place cannot be bit-reproduced on a CPU: the slab test uses
 * __fdividef(1.0f, d) (MUFU.RCP, ~1 ulp).  oracle_svo_intersect therefore
 * accepts an optional `inv_dir` array; tests on the GPU box fill it with the
 * device's own reciprocal so that everything else is compared bit for bit.
 * With inv_dir == NULL the IEEE quotient 1.0f/d is used.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_STACK 256 /* intersect_gpu.cu:229 */

/* Slab test, intersect_gpu.cu:75-140.  Returns 1 and the clipped interval, or
 * 0 for the reference's (-1,-1) "miss" value.  Evaluation order and the three
 * early-outs are kept because they decide border cases (NaN/inf compare
 * false). */
static int slab(const float o[3], const float inv[3], const float c[3],
                float half, float *t_lo, float *t_hi)
{
    float lo = 0.0f, hi = 100000.0f;
    for (int a = 0; a < 3; ++a) {
        float d_lo = (c[a] - half - o[a]) * inv[a];
        float d_hi = (c[a] + half - o[a]) * inv[a];
        if (d_hi < d_lo) { float t = d_lo; d_lo = d_hi; d_hi = t; }
        if (d_hi < lo) return 0;
        if (d_lo > hi) return 0;
        lo = (d_lo > lo) ? d_lo : lo;
        hi = (d_hi < hi) ? d_hi : hi;
        if (lo > hi) return 0;
    }
    *t_lo = lo; *t_hi = hi;
    return 1;
}

/* Layout as the reference wrapper (intersect.cpp:83-112): ray_start/ray_dir
 * [b,m,3], points [b,n,3], children [b,n,9], outputs [b,m,n_max].  Outputs
 * are fully initialised here (idx=-1, depths=0) as intersect.cpp:98-106 does.
 * Returns the number of rays whose DFS stack would have tripped the
 * reference's assert (intersect_gpu.cu:235); those rays stop early. */
int oracle_svo_intersect(int b, int n, int m, float voxelsize, int n_max,
                         const float *ray_start, const float *ray_dir,
                         const float *points, const int *children,
                         const float *inv_dir,
                         int *idx, float *min_depth, float *max_depth)
{
    const float half_voxel = (float)(voxelsize * 0.5); /* :222 (double 0.5) */
    int overflow = 0;
#pragma omp parallel for schedule(static) reduction(+ : overflow)
    for (long r = 0; r < (long)b * m; ++r) {
        const int bi = (int)(r / m);
        const float *pts = points + (size_t)bi * n * 3;
        const int *ch = children + (size_t)bi * n * 9;
        const float *o = ray_start + r * 3, *d = ray_dir + r * 3;
        int *oi = idx + r * n_max;
        float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
        for (int l = 0; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.0f; omax[l] = 0.0f; }

        float inv[3];
        for (int a = 0; a < 3; ++a)
            inv[a] = inv_dir ? inv_dir[r * 3 + a] : 1.0f / d[a];

        int stack[ORACLE_STACK];
        int top = 0, cnt = 0;
        stack[0] = 0; /* root is row 0, :232 */
        while (top > -1 && cnt < n_max) {
            if (top >= ORACLE_STACK) { ++overflow; break; }
            const int k = stack[top--];
            const int side = ch[k * 9 + 8];
            float lo, hi;
            /* "depths.x > -1.0f" (:247): a hit interval has lo >= 0. */
            if (!slab(o, inv, pts + k * 3, half_voxel * (float)side, &lo, &hi)) continue;
            if (!(lo > -1.0f)) continue;
            if (side == 1) { /* terminal node, :250 */
                oi[cnt] = k; omin[cnt] = lo; omax[cnt] = hi; ++cnt;
                continue;
            }
            for (int u = 0; u < 8; ++u) /* push 0..7 => pop 7 first, :259-265 */
                if (ch[k * 9 + u] > -1) {
                    ++top;
                    if (top < ORACLE_STACK) stack[top] = ch[k * 9 + u];
                }
        }
    }
    return overflow;
}

/* Brute force over all n boxes, intersect_gpu.cu:142-189 (no children, every
 * box has half = voxelsize/2). */
void oracle_aabb_intersect(int b, int n, int m, float voxelsize, int n_max,
                           const float *ray_start, const float *ray_dir,
                           const float *points, const float *inv_dir,
                           int *idx, float *min_depth, float *max_depth)
{
    const float half_voxel = (float)(voxelsize * 0.5);
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)b * m; ++r) {
        const int bi = (int)(r / m);
        const float *pts = points + (size_t)bi * n * 3;
        const float *o = ray_start + r * 3, *d = ray_dir + r * 3;
        int *oi = idx + r * n_max;
        float *omin = min_depth + r * n_max, *omax = max_depth + r * n_max;
        for (int l = 0; l < n_max; ++l) { oi[l] = -1; omin[l] = 0.0f; omax[l] = 0.0f; }
        float inv[3];
        for (int a = 0; a < 3; ++a)
            inv[a] = inv_dir ? inv_dir[r * 3 + a] : 1.0f / d[a];
        int cnt = 0;
        for (int k = 0; k < n && cnt < n_max; ++k) {
            float lo, hi;
            if (!slab(o, inv, pts + k * 3, half_voxel, &lo, &hi)) continue;
            if (!(lo > -1.0f)) continue;
            oi[cnt] = k; omin[cnt] = lo; omax[cnt] = hi; ++cnt;
        }
    }
}

/* sample_gpu.cu:133-239.  Grouped layout exactly as the reference call:
 * pts_idx/min_depth/max_depth/probs [b,num_rays,max_hits], steps
 * [b,num_rays], noise and outputs [b,num_rays,max_steps].  Outputs are
 * initialised as sample.cpp:80-89 (idx=-1, depth=dists=0).
 *
 * Quirk Q7 is restated literally: the tail loop's guard
 * compares the ray COUNT with a flat hit offset, `~done` is always true, and
 * the break test reads pts_idx[curr_bin] of the group's first ray.  Reads of
 * pts_idx[H+curr_bin] with curr_bin == max_hits therefore land on the next
 * ray's first hit.  Writes are bounds-guarded against max_steps (the
 * reference relies on max_steps = ceil(steps).max()+P being large enough,
 * voxel_helpers.py:320); a guarded-away write is counted in the return value
 * so that tests can assert it never happens. */
int oracle_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps,
                                float fixed_step_size,
                                const int *pts_idx, const float *min_depth,
                                const float *max_depth, const float *noise,
                                const float *probs, const float *steps,
                                int *sampled_idx, float *sampled_depth,
                                float *sampled_dists)
{
    int clipped = 0;
    const long total = (long)b * num_rays;
    for (long i = 0; i < total * max_steps; ++i) {
        sampled_idx[i] = -1; sampled_depth[i] = 0.0f; sampled_dists[i] = 0.0f;
    }
#pragma omp parallel for schedule(static) reduction(+ : clipped)
    for (long r = 0; r < total; ++r) {
        const int bi = (int)(r / num_rays), j = (int)(r % num_rays);
        const int *g_idx = pts_idx + (size_t)bi * num_rays * max_hits;
        const float *g_min = min_depth + (size_t)bi * num_rays * max_hits;
        const float *g_max = max_depth + (size_t)bi * num_rays * max_hits;
        const float *g_prob = probs + (size_t)bi * num_rays * max_hits;
        const float *g_noise = noise + (size_t)bi * num_rays * max_steps;
        int *o_idx = sampled_idx + (size_t)bi * num_rays * max_steps;
        float *o_depth = sampled_depth + (size_t)bi * num_rays * max_steps;
        float *o_dist = sampled_dists + (size_t)bi * num_rays * max_steps;
        /* flat extent of this group's hit table: reads past it would be into
         * the next group (or past the tensor); the guard j*P+bin < num_rays
         * keeps them inside, see DESIGN.md. */
        const int H = j * max_hits, K = j * max_steps;
        int bin = 0, s = 0;
        float lo_d = g_min[H], hi_d = g_max[H];
        float lo_c = 0.0f, hi_c = g_prob[H];
        const float st = steps[(size_t)bi * num_rays + j];
        float step_size = (float)(1.0 / (double)st); /* :174 */
        float z_low = lo_d;
        const int total_steps = (int)ceil((double)st);
        int done = 0;
        if (fixed_step_size > 0.0f) step_size = fixed_step_size;

#define EMIT(IDX, DIST, DEPTH)                                        \
    do {                                                              \
        if (s < max_steps) {                                          \
            o_idx[K + s] = (IDX); o_dist[K + s] = (DIST);             \
            o_depth[K + s] = (DEPTH);                                 \
        } else ++clipped;                                             \
    } while (0)

        for (int step = 0; step < total_steps; ++step) {
            const float nz = (step < max_steps) ? g_noise[K + step] : 0.5f;
            const float cdf = ((float)step + nz) * step_size;
            while (cdf > hi_c) {
                EMIT(g_idx[H + bin], hi_d - z_low, (hi_d + z_low) * 0.5f);
                ++bin; ++s;
                if (bin >= max_hits || g_idx[H + bin] == -1) { done = 1; break; }
                lo_d = g_min[H + bin]; hi_d = g_max[H + bin];
                lo_c = hi_c; hi_c = hi_c + g_prob[H + bin];
                z_low = lo_d;
            }
            if (done) break;
            const float u = (cdf - lo_c) / (hi_c - lo_c);
            const float z = fmaf(u, hi_d - lo_d, lo_d); /* nvcc contracts :214 */
            EMIT(g_idx[H + bin], z - z_low, (z + z_low) * 0.5f);
            z_low = z; ++s;
        }
        /* tail, :224-237 */
        while (z_low < hi_d && num_rays > H + bin) {
            EMIT(g_idx[H + bin], hi_d - z_low, (hi_d + z_low) * 0.5f);
            ++bin; ++s;
            if (bin >= max_hits || g_idx[bin] == -1) break;
            lo_d = g_min[H + bin]; hi_d = g_max[H + bin];
            z_low = lo_d;
        }
#undef EMIT
    }
    return clipped;
}

int oracle_abi_version(void) { return 1; }
