// Kernel 5: SDF -> weight volumetric compositing, the Criterion losses and
// their backward, on CSR ray segments (one warp per hit ray).
//
// Math:
//   sdf2weights + compositing   src/variations/render_helpers.py:510-545
//   Criterion.forward / get_masks / get_sdf_loss   src/criterion.py:16-116
//
// The reference pads every ray to S = max samples per ray with sdf=1, z=10,
// colour=0 (render_helpers.py:510-511, voxel_helpers.py:24,655) and those pads
// take part in the sign-change search and in the loss counts/denominators
// (SURVEY A-Q11/Q12).  Here nothing is padded: the pads' contributions are
// added in closed form per ray ((S - S_q) identical pad elements).
//
// Three launches: k_composite_fwd (per-ray render + per-block loss partials),
// k_loss_finalize (one block: global sums, the tracking median gate, loss
// scalars + the coefficients backward needs), k_composite_bwd (per-ray
// gradient w.r.t. every sample's (r,g,b,sdf)).
#include "common.cuh"
#include "kernels.h"

namespace pslam {

constexpr int kCompThreads = 256;
constexpr int kCompWarps = kCompThreads / 32;
constexpr float kPadZ = 10.0f;    // MAX_DEPTH, voxel_helpers.py:24
constexpr float kPadSdf = 1.0f;   // render_helpers.py:510

// loss[] slots beyond the five public values: coefficients for backward
enum { L_THRESH = 5, L_CDEPTH = 6, L_CFS = 7, L_CSDF = 8, L_CCOLOR = 9, L_NVALID = 10, L_NFS = 11, L_NSDF = 12, L_MEDIAN = 13 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct RaySeg {
    int beg, cnt;   // CSR segment
    int ray;        // ray id
    int ind;        // index of the first sign change (0 if none)
    float z_min;
};

// first k with s_k*s_{k+1} < 0 over the reference's padded row of width S
__device__ __forceinline__ int first_sign_change(const float *__restrict__ out, int beg, int cnt, int S, int lane)
{
    for (int base = 0; base < cnt; base += 32) {
        const int k = base + lane;
        bool flag = false;
        if (k < cnt) {
            const float s = __ldg(out + (size_t)(beg + k) * 4 + 3);
            if (k + 1 < cnt) flag = __fmul_rn(__ldg(out + (size_t)(beg + k + 1) * 4 + 3), s) < 0.0f;
            else if (cnt < S) flag = __fmul_rn(kPadSdf, s) < 0.0f;   // last valid sample against the first pad
        }
        const unsigned b = __ballot_sync(0xffffffffu, flag);
        if (b) return base + __ffs(b) - 1;
    }
    return 0;
}

__global__ void __launch_bounds__(kCompThreads)
k_composite_fwd(pslam_render_t p, float *__restrict__ part_f, int *__restrict__ part_i)
{
    __shared__ float s_f[kCompWarps][3];
    __shared__ int s_i[kCompWarps][2];
    const int Rh = p.counters[PSLAM_C_RH];
    const int S = p.counters[PSLAM_C_S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kCompWarps + warp;
    float sum_color = 0.f, sum_fs = 0.f, sum_sdf = 0.f;
    int n_fs = 0, n_sdf = 0;
    if (q < Rh) {
        const int beg = p.samp_off[q], cnt = min(p.samp_off[q + 1], p.sample_cap) - beg;
        const int ray = p.hit_ray[q];
        const bool has_tgt = p.target_depth != nullptr;   // render only (no Criterion) when absent
        const float gt = has_tgt ? __ldg(p.target_depth + ray) : 0.0f;
        const float tau = p.truncation;
        const int ind = (cnt > 0) ? first_sign_change(p.samp_out, beg, cnt, S, lane) : 0;
        const float z_min = (cnt > 0) ? __ldg(p.samp_z + beg + ind) : kPadZ;
        const float z_cut = __fadd_rn(z_min, tau);
        // pass A: U = sum u_k + 1e-8, loss sums over the valid samples
        float U = 0.0f;
        for (int k = lane; k < cnt; k += 32) {
            const float s = __ldg(p.samp_out + (size_t)(beg + k) * 4 + 3);
            const float z = __ldg(p.samp_z + beg + k);
            const float x = s / tau;
            const float a = sigmoidf_(x) * sigmoidf_(-x);
            if (z < z_cut) U += a;
            if (!has_tgt) continue;
            // get_masks, criterion.py:78-102
            const bool front = z < __fsub_rn(gt, tau), back = z > __fadd_rn(gt, tau);
            const bool sm = !front && !back && gt > 0.0f && gt < p.max_depth;
            if (front) { ++n_fs; const float d = s - 1.0f; sum_fs = fmaf(d, d, sum_fs); }
            if (sm) { ++n_sdf; const float d = (z + s * tau) - gt; sum_sdf = fmaf(d, d, sum_sdf); }
        }
        U = warp_sum(U) + 1e-8f;
        // pass B: w_k = u_k / U, rgb, depth
        float r = 0.f, g = 0.f, b = 0.f, dep = 0.f;
        for (int k = lane; k < cnt; k += 32) {
            const float4 o = __ldg(reinterpret_cast<const float4 *>(p.samp_out + (size_t)(beg + k) * 4));
            const float z = __ldg(p.samp_z + beg + k);
            const float x = o.w / tau;
            const float w = (z < z_cut) ? (sigmoidf_(x) * sigmoidf_(-x)) / U : 0.0f;
            if (p.samp_w) p.samp_w[beg + k] = w;
            r = fmaf(w, o.x, r); g = fmaf(w, o.y, g); b = fmaf(w, o.z, b); dep = fmaf(w, z, dep);
        }
        r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); dep = warp_sum(dep);
        // pass C (tracking): depth variance for the median gate, criterion.py:45-48
        float var = 0.0f;
        if (p.flags & PSLAM_F_TRACKING) {
            for (int k = lane; k < cnt; k += 32) {
                const float s = __ldg(p.samp_out + (size_t)(beg + k) * 4 + 3);
                const float z = __ldg(p.samp_z + beg + k);
                const float x = s / tau;
                const float w = (z < z_cut) ? (sigmoidf_(x) * sigmoidf_(-x)) / U : 0.0f;
                const float d = dep - z;
                var = fmaf(w, d * d, var);
            }
            var = warp_sum(var);
        }
        if (lane == 0) {
            // the (S - cnt) pads of this row: z = 10, sdf = 1
            const int npad = S - cnt;
            if (npad > 0 && has_tgt) {
                const bool front = kPadZ < __fsub_rn(gt, tau), back = kPadZ > __fadd_rn(gt, tau);
                const bool sm = !front && !back && gt > 0.0f && gt < p.max_depth;
                if (front) n_fs += npad;   // (1*1 - 1)^2 = 0 adds nothing to the sum
                if (sm) { n_sdf += npad; const float d = (kPadZ + kPadSdf * tau) - gt; sum_sdf += (float)npad * d * d; }
            }
            float *ro = p.ray_out + (size_t)q * 8;
            ro[0] = r; ro[1] = g; ro[2] = b; ro[3] = dep; ro[4] = z_min; ro[5] = U;
            if (has_tgt) {
                const float *tc = p.target_rgb + (size_t)ray * 3;
                sum_color = fabsf(__ldg(tc) - r) + fabsf(__ldg(tc + 1) - g) + fabsf(__ldg(tc + 2) - b);
                ro[6] = fabsf(gt - dep) / sqrtf(var + 1e-10f);
                ro[7] = (gt > 0.01f && gt < p.max_depth) ? 1.0f : 0.0f;
            } else {
                ro[6] = 0.0f; ro[7] = 0.0f;
            }
        }
        sum_fs = warp_sum(sum_fs); sum_sdf = warp_sum(sum_sdf);
        n_fs = warp_sum_i(n_fs); n_sdf = warp_sum_i(n_sdf);
    }
    if (lane == 0) {
        s_f[warp][0] = sum_color; s_f[warp][1] = sum_fs; s_f[warp][2] = sum_sdf;
        s_i[warp][0] = n_fs; s_i[warp][1] = n_sdf;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.0f;
        for (int w = 0; w < kCompWarps; ++w) t += s_f[w][threadIdx.x];
        part_f[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    } else if (threadIdx.x < 5) {
        int t = 0;
        for (int w = 0; w < kCompWarps; ++w) t += s_i[w][threadIdx.x - 3];
        part_i[(size_t)blockIdx.x * 2 + threadIdx.x - 3] = t;
    }
}

// block-wide sums (1024 threads), deterministic order
__device__ __forceinline__ float block_sum_f(float v, float *s_buf)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < 32; ++w) t += s_buf[w];
    return t;
}
__device__ __forceinline__ long long block_sum_ll(long long v, long long *s_buf)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < 32; ++w) t += s_buf[w];
    return t;
}

__global__ void __launch_bounds__(1024)
k_loss_finalize(pslam_render_t p, const float *__restrict__ part_f, const int *__restrict__ part_i, int nblocks)
{
    __shared__ float s_f[32];
    __shared__ long long s_ll[32];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_rank;
    const int Rh = p.counters[PSLAM_C_RH];
    const int S = p.counters[PSLAM_C_S];
    const int tid = threadIdx.x;
    float c = 0.f, fs = 0.f, sd = 0.f;
    long long nfs = 0, nsdf = 0;
    for (int b = tid; b < nblocks; b += 1024) {
        c += part_f[(size_t)b * 4]; fs += part_f[(size_t)b * 4 + 1]; sd += part_f[(size_t)b * 4 + 2];
        nfs += part_i[(size_t)b * 2]; nsdf += part_i[(size_t)b * 2 + 1];
    }
    c = block_sum_f(c, s_f); fs = block_sum_f(fs, s_f); sd = block_sum_f(sd, s_f);
    nfs = block_sum_ll(nfs, s_ll); nsdf = block_sum_ll(nsdf, s_ll);

    // tracking: lower median of tmp = |dd|/sqrt(var) over the hit rays (torch.median), by
    // 4 x 8-bit radix select on the float bit patterns (all values are >= 0)
    float thresh = __int_as_float(0x7f800000), median = 0.0f;
    if ((p.flags & PSLAM_F_TRACKING) && Rh > 0) {
        if (tid == 0) { s_prefix = 0u; s_rank = (unsigned)((Rh - 1) / 2); }
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (tid < 256) s_hist[tid] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix;
            const unsigned himask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
            for (int q = tid; q < Rh; q += 1024) {
                const unsigned bits = __float_as_uint(p.ray_out[(size_t)q * 8 + 6]);
                if ((bits & himask) == prefix) atomicAdd(&s_hist[(bits >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned rank = s_rank, acc = 0u;
                int bin = 0;
                for (; bin < 255; ++bin) {
                    if (acc + s_hist[bin] > rank) break;
                    acc += s_hist[bin];
                }
                s_rank = rank - acc;
                s_prefix = prefix | ((unsigned)bin << shift);
            }
            __syncthreads();
        }
        median = __uint_as_float(s_prefix);
        thresh = 10.0f * median;
    }
    // depth loss over valid (and gated) rays, criterion.py:41-50
    float dsum = 0.0f;
    long long nvalid = 0;
    for (int q = tid; q < Rh; q += 1024) {
        const float *ro = p.ray_out + (size_t)q * 8;
        const bool ok = ro[7] != 0.0f && (!(p.flags & PSLAM_F_TRACKING) || ro[6] < thresh);
        if (ok) { dsum += fabsf(__ldg(p.target_depth + p.hit_ray[q]) - ro[3]); ++nvalid; }
    }
    dsum = block_sum_f(dsum, s_f);
    nvalid = block_sum_ll(nvalid, s_ll);
    if (tid == 0) {
        const float n = (float)Rh * (float)S;     // elements of the reference's padded [R_h,S] tensors
        const float fnfs = (float)nfs, fnsdf = (float)nsdf;
        const float fs_w = 1.0f - fnfs / (fnfs + fnsdf), sdf_w = 1.0f - fnsdf / (fnfs + fnsdf);
        const float color = c / (3.0f * (float)Rh);
        const float depth = dsum / (float)nvalid;
        const float fs_loss = (fs / n) * fs_w, sdf_loss = (sd / n) * sdf_w;
        p.loss[PSLAM_L_COLOR] = color; p.loss[PSLAM_L_DEPTH] = depth;
        p.loss[PSLAM_L_FS] = fs_loss; p.loss[PSLAM_L_SDF] = sdf_loss;
        p.loss[PSLAM_L_TOTAL] = p.w_rgb * color + p.w_depth * depth + p.w_fs * fs_loss + p.w_sdf * sdf_loss;
        p.loss[L_THRESH] = thresh;
        p.loss[L_CDEPTH] = p.w_depth / (float)nvalid;
        p.loss[L_CFS] = p.w_fs * fs_w * 2.0f / n;
        p.loss[L_CSDF] = p.w_sdf * sdf_w * 2.0f * p.truncation / n;
        p.loss[L_CCOLOR] = p.w_rgb / (3.0f * (float)Rh);
        p.loss[L_NVALID] = (float)nvalid; p.loss[L_NFS] = fnfs; p.loss[L_NSDF] = fnsdf; p.loss[L_MEDIAN] = median;
    }
}

__device__ __forceinline__ float signf_(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(kCompThreads)
k_composite_bwd(pslam_render_t p)
{
    const int Rh = p.counters[PSLAM_C_RH];
    const int S = p.counters[PSLAM_C_S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kCompWarps + warp;
    if (q >= Rh) return;
    const int beg = p.samp_off[q], cnt = min(p.samp_off[q + 1], p.sample_cap) - beg;
    if (cnt <= 0) return;
    const int ray = p.hit_ray[q];
    const float gt = __ldg(p.target_depth + ray);
    const float tau = p.truncation;
    const float *ro = p.ray_out + (size_t)q * 8;
    const float U = ro[5], z_cut = __fadd_rn(ro[4], tau), dep = ro[3];
    (void)S;
    // dL/d(rendered colour, depth), SURVEY Appendix C
    const float *tc = p.target_rgb + (size_t)ray * 3;
    const float cc = p.loss[L_CCOLOR];
    const float g_r = cc * signf_(ro[0] - __ldg(tc)), g_g = cc * signf_(ro[1] - __ldg(tc + 1)), g_b = cc * signf_(ro[2] - __ldg(tc + 2));
    const bool ok = ro[7] != 0.0f && (!(p.flags & PSLAM_F_TRACKING) || ro[6] < p.loss[L_THRESH]);
    const float g_d = ok ? p.loss[L_CDEPTH] * signf_(dep - gt) : 0.0f;
    const float cfs = p.loss[L_CFS], csdf = p.loss[L_CSDF];
    // dot = sum_j w_j g_w[j]
    float dot = 0.0f;
    for (int k = lane; k < cnt; k += 32) {
        const float4 o = __ldg(reinterpret_cast<const float4 *>(p.samp_out + (size_t)(beg + k) * 4));
        const float z = __ldg(p.samp_z + beg + k);
        const float x = o.w / tau;
        const float w = (z < z_cut) ? (sigmoidf_(x) * sigmoidf_(-x)) / U : 0.0f;
        dot = fmaf(w, g_r * o.x + g_g * o.y + g_b * o.z + g_d * z, dot);
    }
    dot = warp_sum(dot);
    for (int k = lane; k < cnt; k += 32) {
        const float4 o = __ldg(reinterpret_cast<const float4 *>(p.samp_out + (size_t)(beg + k) * 4));
        const float z = __ldg(p.samp_z + beg + k);
        const float x = o.w / tau;
        const float sp = sigmoidf_(x), sn = sigmoidf_(-x);
        const bool in = z < z_cut;
        const float w = in ? (sp * sn) / U : 0.0f;
        const float g_w = g_r * o.x + g_g * o.y + g_b * o.z + g_d * z;
        float g_s = in ? ((g_w - dot) / U) * (sp * sn * (sn - sp) / tau) : 0.0f;
        const bool front = z < __fsub_rn(gt, tau), back = z > __fadd_rn(gt, tau);
        const bool sm = !front && !back && gt > 0.0f && gt < p.max_depth;
        if (front) g_s = fmaf(cfs, o.w - 1.0f, g_s);
        if (sm) g_s = fmaf(csdf, (z + o.w * tau) - gt, g_s);
        *reinterpret_cast<float4 *>(p.samp_gout + (size_t)(beg + k) * 4) = make_float4(w * g_r, w * g_g, w * g_b, g_s);
    }
}

__global__ void k_zero_f(float *__restrict__ a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0.0f;
}

int launch_composite_forward(const pslam_render_t *p, cudaStream_t st)
{
    const int nb = ceil_div(p->R, kCompWarps);
    float *part_f = p->scratch_f;                                // [nb,4]
    int *part_i = p->scratch_i + 2 * (ceil_div(p->R, 128) + 8);  // after the two scan-partial arrays: [nb,2]
    k_composite_fwd<<<nb, kCompThreads, 0, st>>>(*p, part_f, part_i);
    PSLAM_CHECK_LAUNCH("composite_fwd");
    if (p->target_depth && p->target_rgb) {
        k_loss_finalize<<<1, 1024, 0, st>>>(*p, part_f, part_i, nb);
        PSLAM_CHECK_LAUNCH("loss_finalize");
    }
    return 0;
}

int launch_composite_backward(const pslam_render_t *p, cudaStream_t st)
{
    if (p->flags & PSLAM_F_GRAD_RAYS) {
        k_zero_f<<<ceil_div(p->R * 3, 256), 256, 0, st>>>(p->g_rays_o, p->R * 3);
        k_zero_f<<<ceil_div(p->R * 3, 256), 256, 0, st>>>(p->g_rays_d, p->R * 3);
        PSLAM_CHECK_LAUNCH("zero_ray_grads");
    }
    k_composite_bwd<<<ceil_div(p->R, kCompWarps), kCompThreads, 0, st>>>(*p);
    PSLAM_CHECK_LAUNCH("composite_bwd");
    return 0;
}

}  // namespace pslam
