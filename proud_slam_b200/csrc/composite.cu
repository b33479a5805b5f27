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
#include "peer.cuh"

namespace pslam {

constexpr int kCompThreads = 256;
constexpr int kCompWarps = kCompThreads / 32;
constexpr float kPadZ = 10.0f;    // MAX_DEPTH, voxel_helpers.py:24
constexpr float kPadSdf = 1.0f;   // render_helpers.py:510

// loss[] slots beyond the five public values: coefficients for backward
enum { L_THRESH = 5, L_CDEPTH = 6, L_CFS = 7, L_CSDF = 8, L_CCOLOR = 9, L_NVALID = 10, L_NFS = 11, L_NSDF = 12, L_MEDIAN = 13 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

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
    pdl_enter();
    __shared__ float s_f[kCompWarps][6];
    __shared__ int s_i[kCompWarps][7];
    const int Rh = p.counters[PSLAM_C_RH];
    const int S = p.counters[PSLAM_C_S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kCompWarps + warp;
    float sum_color = 0.f, sum_fs = 0.f, sum_sdf = 0.f, sum_dep = 0.f;   // sum_dep / n_valid: depth loss without the tracking gate
    int n_fs = 0, n_sdf = 0, n_valid = 0;
    // pad terms, linear in S so that they can be closed after a cross-rank max of S:
    //   pads of a ray = (S - cnt) copies of (z=10, sdf=1)  ->  S*x0 - x1 with x1 = x0*cnt
    float pad_d0 = 0.f, pad_d1 = 0.f;           // sum over rays of [sdf-mask pad] d^2 (and * cnt)
    int pad_f0 = 0, pad_f1 = 0, pad_m0 = 0, pad_m1 = 0;
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
            if (has_tgt) {
                const bool front = kPadZ < __fsub_rn(gt, tau), back = kPadZ > __fadd_rn(gt, tau);
                const bool sm = !front && !back && gt > 0.0f && gt < p.max_depth;
                if (front) { pad_f0 = 1; pad_f1 = cnt; }   // (1*1 - 1)^2 = 0 adds nothing to the fs sum
                if (sm) {
                    const float d = (kPadZ + kPadSdf * tau) - gt;
                    pad_m0 = 1; pad_m1 = cnt; pad_d0 = d * d; pad_d1 = d * d * (float)cnt;
                }
            }
            float *ro = p.ray_out + (size_t)q * 8;
            ro[0] = r; ro[1] = g; ro[2] = b; ro[3] = dep; ro[4] = z_min; ro[5] = U;
            if (has_tgt) {
                const float *tc = p.target_rgb + (size_t)ray * 3;
                sum_color = fabsf(__ldg(tc) - r) + fabsf(__ldg(tc + 1) - g) + fabsf(__ldg(tc + 2) - b);
                ro[6] = fabsf(gt - dep) / sqrtf(var + 1e-10f);
                ro[7] = (gt > 0.01f && gt < p.max_depth) ? 1.0f : 0.0f;
                if (ro[7] != 0.0f) { sum_dep = fabsf(gt - dep); n_valid = 1; }
            } else {
                ro[6] = 0.0f; ro[7] = 0.0f;
            }
        }
        sum_fs = warp_sum(sum_fs); sum_sdf = warp_sum(sum_sdf);
        n_fs = warp_sum_i(n_fs); n_sdf = warp_sum_i(n_sdf);
    }
    if (lane == 0) {
        s_f[warp][0] = sum_color; s_f[warp][1] = sum_fs; s_f[warp][2] = sum_sdf; s_f[warp][3] = pad_d0; s_f[warp][4] = pad_d1;
        s_f[warp][5] = sum_dep;
        s_i[warp][0] = n_fs; s_i[warp][1] = n_sdf; s_i[warp][2] = pad_f0; s_i[warp][3] = pad_f1;
        s_i[warp][4] = pad_m0; s_i[warp][5] = pad_m1; s_i[warp][6] = n_valid;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.0f;
        for (int w = 0; w < kCompWarps; ++w) t += s_f[w][threadIdx.x];
        part_f[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
    } else if (threadIdx.x >= 8 && threadIdx.x < 15) {
        int t = 0;
        for (int w = 0; w < kCompWarps; ++w) t += s_i[w][threadIdx.x - 8];
        part_i[(size_t)blockIdx.x * 8 + threadIdx.x - 8] = t;
    }
}

// block-wide sums (1024 threads) of N values in double, deterministic order; the totals are valid in thread 0 only
template <int N>
__device__ __forceinline__ void block_sums_d(double (&v)[N], double *s_buf /* [32][N] */, double *s_tot /* [N] */)
{
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) s_buf[(threadIdx.x >> 5) * N + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += s_buf[w * N + threadIdx.x];
        s_tot[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = s_tot[k];
}

__device__ void loss_coeffs_body(const pslam_render_t &p, const double *rows, int nrows);

// Stage A of the loss: this rank's raw sums -> raw[16] (double).  Slots: see RAW_* below.
enum { RAW_COLOR = 0, RAW_FS, RAW_SDF, RAW_D0, RAW_D1, RAW_NFS, RAW_F0, RAW_F1, RAW_NSDF, RAW_M0, RAW_M1, RAW_RH,
       RAW_DEPTH, RAW_NVALID, RAW_S, RAW_THRESH };

__global__ void __launch_bounds__(1024)
k_loss_reduce(pslam_render_t p, const float *__restrict__ part_f, const int *__restrict__ part_i, int nblocks, int prologue)
{
    pdl_enter();
    if (prologue) {   // the backward follows in the same step: its prologue (k_bwd_prologue) rides along here
        if (threadIdx.x == 0) p.counters[PSLAM_C_TILE] = 0;
        if (p.flags & PSLAM_F_GRAD_RAYS)
            for (int i = threadIdx.x; i < p.R * 3; i += 1024) { p.g_rays_o[i] = 0.0f; p.g_rays_d[i] = 0.0f; }
    }
    __shared__ double s_d[32 * 13], s_tot[13];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_rank;
    const int Rh = p.counters[PSLAM_C_RH];
    const int tid = threadIdx.x;
    double acc[13];   // f[0..5] (colour, fs, sdf, pad d0, pad d1, ungated depth), n[0..6] (.., ungated valid count)
#pragma unroll
    for (int k = 0; k < 13; ++k) acc[k] = 0.0;
    for (int b = tid; b < nblocks; b += 1024) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += (double)part_f[(size_t)b * 8 + k];
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[6 + k] += (double)part_i[(size_t)b * 8 + k];
    }
    block_sums_d<13>(acc, s_d, s_tot);
    const double *f = acc, *n = acc + 6;

    // tracking: lower median of tmp = |dd|/sqrt(var) over the hit rays (torch.median), by
    // 4 x 8-bit radix select on the float bit patterns (all values are >= 0)
    float thresh = __int_as_float(0x7f800000);
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
            if (tid < 32) {
                // bin of the wanted rank: warp-parallel prefix over the 256 counts (8 per lane) instead of a serial walk
                const unsigned rank = s_rank;
                unsigned c[8], sum = 0u;
#pragma unroll
                for (int k = 0; k < 8; ++k) { c[k] = s_hist[tid * 8 + k]; sum += c[k]; }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += y;
                }
                const unsigned excl = incl - sum;
                const bool mine = rank >= excl && rank < incl;          // exactly one lane (the counts sum to more than rank)
                const unsigned owner = __ballot_sync(0xffffffffu, mine);
                if (mine) {
                    unsigned acc = excl;
                    int k = 0;
                    for (; k < 7; ++k) {
                        if (acc + c[k] > rank) break;
                        acc += c[k];
                    }
                    s_rank = rank - acc;
                    s_prefix = prefix | ((unsigned)(tid * 8 + k) << shift);
                } else if (owner == 0u && tid == 31) {                  // (cannot happen for rank < count; mirrors the old walk's last bin)
                    s_rank = rank - (incl - c[7]);
                    s_prefix = prefix | (255u << shift);
                }
            }
            __syncthreads();
        }
        thresh = 10.0f * __uint_as_float(s_prefix);
    }
    // depth loss over valid (and gated) rays, criterion.py:41-50; without the tracking gate k_composite_fwd already summed it
    double dn[2] = {f[5], n[6]};
    if (p.flags & PSLAM_F_TRACKING) {
        dn[0] = 0.0; dn[1] = 0.0;
        for (int q = tid; q < Rh; q += 1024) {
            const float *ro = p.ray_out + (size_t)q * 8;
            if (ro[7] != 0.0f && ro[6] < thresh) { dn[0] += (double)fabsf(__ldg(p.target_depth + p.hit_ray[q]) - ro[3]); dn[1] += 1.0; }
        }
        block_sums_d<2>(dn, s_d, s_tot);
    }
    const double dsum = dn[0], nvalid = dn[1];
    if (tid == 0) {
        double *raw = p.loss_raw;
        raw[RAW_COLOR] = f[0]; raw[RAW_FS] = f[1]; raw[RAW_SDF] = f[2]; raw[RAW_D0] = f[3]; raw[RAW_D1] = f[4];
        raw[RAW_NFS] = n[0]; raw[RAW_NSDF] = n[1]; raw[RAW_F0] = n[2]; raw[RAW_F1] = n[3]; raw[RAW_M0] = n[4]; raw[RAW_M1] = n[5];
        raw[RAW_RH] = (double)Rh; raw[RAW_DEPTH] = dsum; raw[RAW_NVALID] = nvalid;
        raw[RAW_S] = (double)p.counters[PSLAM_C_S]; raw[RAW_THRESH] = (double)thresh;
        // single rank: close the loss right here (stage B) instead of a launch of its own
        if (!(p.flags & PSLAM_F_DEFER_LOSS) && p.peer.world <= 1) loss_coeffs_body(p, raw, 1);
    }
    if (!(p.flags & PSLAM_F_DEFER_LOSS) && p.peer.world > 1) {
        // all ranks: every rank's raw sums to every rank over NVLink peer memory, then the same closure everywhere (peer.cu)
        __shared__ unsigned long long s_epoch;
        __shared__ int s_fail;
        const int world = p.peer.world, rank = p.peer.rank;
        PeerSync *mine = static_cast<PeerSync *>(p.peer.sync[rank]);
        if (tid == 0) { s_epoch = ld_acquire_sys(&mine->epoch_loss) + 1ull; s_fail = 0; __threadfence(); }
        __syncthreads();
        const unsigned long long e = s_epoch;
        const int par = (int)(e & 1ull);
        if (tid < world * 16) static_cast<PeerSync *>(p.peer.sync[tid >> 4])->rows[par][rank][tid & 15] = p.loss_raw[tid & 15];
        __syncthreads();               // (the flag threads' st.release.sys below is cumulative over what the barrier ordered before them)
        if (tid < world) {
            st_release_sys(&static_cast<PeerSync *>(p.peer.sync[tid])->loss_flag[par][rank], e);
            if (!spin_until(&mine->loss_flag[par][tid], e)) s_fail = 1;
        }
        __syncthreads();
        if (tid == 0) {
            if (s_fail) { atomicOr(p.counters + PSLAM_C_OVERFLOW, 16); loss_coeffs_body(p, p.loss_raw, 1); }
            else loss_coeffs_body(p, &mine->rows[par][0][0], world);
            st_release_sys(&mine->epoch_loss, e);
        }
    }
}

// Stage B: raw sums of all ranks ([nrows,16] doubles; nrows = 1 on one GPU) -> loss values and
// the coefficients backward needs.  Sums add over ranks, S is the maximum, the pad terms close as
// S*x0 - x1.  Final arithmetic in fp32 like criterion.py.
__global__ void k_loss_coeffs(pslam_render_t p, const double *__restrict__ rows, int nrows)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    loss_coeffs_body(p, rows, nrows);
}

__device__ void loss_coeffs_body(const pslam_render_t &p, const double *rows, int nrows)
{
    double t[16];
    for (int k = 0; k < 16; ++k) t[k] = 0.0;
    for (int r = 0; r < nrows; ++r) {
        for (int k = 0; k < RAW_S; ++k) t[k] += rows[(size_t)r * 16 + k];
        t[RAW_S] = fmax(t[RAW_S], rows[(size_t)r * 16 + RAW_S]);
    }
    t[RAW_THRESH] = rows[RAW_THRESH];
    const double S = t[RAW_S];
    const float Rh = (float)t[RAW_RH];
    const float n = Rh * (float)S;     // elements of the reference's padded [R_h,S] tensors
    const float fnfs = (float)(t[RAW_NFS] + S * t[RAW_F0] - t[RAW_F1]);
    const float fnsdf = (float)(t[RAW_NSDF] + S * t[RAW_M0] - t[RAW_M1]);
    const float fs_sum = (float)t[RAW_FS];
    const float sdf_sum = (float)(t[RAW_SDF] + S * t[RAW_D0] - t[RAW_D1]);
    const float nvalid = (float)t[RAW_NVALID];
    const float fs_w = 1.0f - fnfs / (fnfs + fnsdf), sdf_w = 1.0f - fnsdf / (fnfs + fnsdf);
    const float color = (float)t[RAW_COLOR] / (3.0f * Rh);
    const float depth = (float)t[RAW_DEPTH] / nvalid;
    const float fs_loss = (fs_sum / n) * fs_w, sdf_loss = (sdf_sum / n) * sdf_w;
    p.loss[PSLAM_L_COLOR] = color; p.loss[PSLAM_L_DEPTH] = depth;
    p.loss[PSLAM_L_FS] = fs_loss; p.loss[PSLAM_L_SDF] = sdf_loss;
    p.loss[PSLAM_L_TOTAL] = p.w_rgb * color + p.w_depth * depth + p.w_fs * fs_loss + p.w_sdf * sdf_loss;
    p.loss[L_THRESH] = (float)t[RAW_THRESH];
    p.loss[L_CDEPTH] = p.w_depth / nvalid;
    p.loss[L_CFS] = p.w_fs * fs_w * 2.0f / n;
    p.loss[L_CSDF] = p.w_sdf * sdf_w * 2.0f * p.truncation / n;
    p.loss[L_CCOLOR] = p.w_rgb / (3.0f * Rh);
    p.loss[L_NVALID] = nvalid; p.loss[L_NFS] = fnfs; p.loss[L_NSDF] = fnsdf; p.loss[L_MEDIAN] = (float)t[RAW_THRESH] * 0.1f;
}

__device__ __forceinline__ float signf_(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

// max |dL/d(sample outputs)| of this backward (bit pattern, counters[PSLAM_C_TILE]; reset by k_bwd_prologue): the 3xF16
// decoder backward takes its power-of-two gradient scale from it instead of running a reduction of its own
__device__ __forceinline__ void publish_gmax(const pslam_render_t &p, float gmax, int lane)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 0.0f) atomicMax(reinterpret_cast<unsigned *>(p.counters + PSLAM_C_TILE), __float_as_uint(gmax));
}

__global__ void __launch_bounds__(kCompThreads)
k_composite_bwd(pslam_render_t p)
{
    pdl_enter();
    const int Rh = p.counters[PSLAM_C_RH];
    const int S = p.counters[PSLAM_C_S];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kCompWarps + warp;
    if (q >= Rh) return;
    const int beg = p.samp_off[q], cnt = min(p.samp_off[q + 1], p.sample_cap) - beg;
    float gmax = 0.0f;
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
        gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(w * g_r), fabsf(w * g_g)), fmaxf(fabsf(w * g_b), fabsf(g_s))));
    }
    publish_gmax(p, gmax, lane);
}

// Backward of compositing for ARBITRARY upstream gradients (the autograd route of the drop-in
// render_rays: a caller-side Criterion produced dL/d(color, depth, sdf, weights)).  g_color [R_h,3],
// g_depth [R_h] by rank; g_sdf / g_weight per sample in CSR order (either may be NULL).
__global__ void __launch_bounds__(kCompThreads)
k_composite_bwd_ext(pslam_render_t p, const float *__restrict__ g_color, const float *__restrict__ g_depth,
                    const float *__restrict__ g_sdf, const float *__restrict__ g_weight)
{
    const int Rh = p.counters[PSLAM_C_RH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kCompWarps + warp;
    if (q >= Rh) return;
    const int beg = p.samp_off[q], cnt = min(p.samp_off[q + 1], p.sample_cap) - beg;
    float gmax = 0.0f;
    if (cnt <= 0) return;
    const float tau = p.truncation;
    const float *ro = p.ray_out + (size_t)q * 8;
    const float U = ro[5], z_cut = __fadd_rn(ro[4], tau);
    const float g_r = g_color ? g_color[q * 3] : 0.f, g_g = g_color ? g_color[q * 3 + 1] : 0.f, g_b = g_color ? g_color[q * 3 + 2] : 0.f;
    const float g_d = g_depth ? g_depth[q] : 0.f;
    float dot = 0.0f;
    for (int k = lane; k < cnt; k += 32) {
        const float4 o = __ldg(reinterpret_cast<const float4 *>(p.samp_out + (size_t)(beg + k) * 4));
        const float z = __ldg(p.samp_z + beg + k);
        const float x = o.w / tau;
        const float w = (z < z_cut) ? (sigmoidf_(x) * sigmoidf_(-x)) / U : 0.0f;
        dot = fmaf(w, g_r * o.x + g_g * o.y + g_b * o.z + g_d * z + (g_weight ? g_weight[beg + k] : 0.f), dot);
    }
    dot = warp_sum(dot);
    for (int k = lane; k < cnt; k += 32) {
        const float4 o = __ldg(reinterpret_cast<const float4 *>(p.samp_out + (size_t)(beg + k) * 4));
        const float z = __ldg(p.samp_z + beg + k);
        const float x = o.w / tau;
        const float sp = sigmoidf_(x), sn = sigmoidf_(-x);
        const bool in = z < z_cut;
        const float w = in ? (sp * sn) / U : 0.0f;
        const float g_w = g_r * o.x + g_g * o.y + g_b * o.z + g_d * z + (g_weight ? g_weight[beg + k] : 0.f);
        float g_s = in ? ((g_w - dot) / U) * (sp * sn * (sn - sp) / tau) : 0.0f;
        if (g_sdf) g_s += g_sdf[beg + k];
        *reinterpret_cast<float4 *>(p.samp_gout + (size_t)(beg + k) * 4) = make_float4(w * g_r, w * g_g, w * g_b, g_s);
        gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(w * g_r), fabsf(w * g_g)), fmaxf(fabsf(w * g_b), fabsf(g_s))));
    }
    publish_gmax(p, gmax, lane);
}

// before a backward: zero the ray-gradient accumulators and reset the gradient maximum
__global__ void k_bwd_prologue(pslam_render_t p)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) p.counters[PSLAM_C_TILE] = 0;
    if ((p.flags & PSLAM_F_GRAD_RAYS) && i < p.R * 3) { p.g_rays_o[i] = 0.0f; p.g_rays_d[i] = 0.0f; }
}

__global__ void k_zero_f(float *__restrict__ a, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0.0f;
}

int launch_composite_forward(const pslam_render_t *p, cudaStream_t st, int fold_prologue)
{
    const int nb = ceil_div(p->R, kCompWarps);
    float *part_f = p->scratch_f;                                // [nb,8]
    int *part_i = p->scratch_i + scratch_i_composite_off(p->R);  // after the two scan-partial arrays: [nb,8]
    launch_chain(k_composite_fwd, dim3(nb), dim3(kCompThreads), 0, st, *p, part_f, part_i);
    PSLAM_CHECK_LAUNCH("composite_fwd");
    if (p->target_depth && p->target_rgb) {
        launch_chain(k_loss_reduce, dim3(1), dim3(1024), 0, st, *p, part_f, part_i, nb, fold_prologue);
        PSLAM_CHECK_LAUNCH("loss_reduce");
    }
    return 0;
}

int launch_loss_coeffs(const pslam_render_t *p, const double *rows, int nrows, cudaStream_t st)
{
    k_loss_coeffs<<<1, 32, 0, st>>>(*p, rows, nrows);
    PSLAM_CHECK_LAUNCH("loss_coeffs");
    return 0;
}

int launch_composite_backward_ext(const pslam_render_t *p, const float *g_color, const float *g_depth, const float *g_sdf,
                                  const float *g_weight, cudaStream_t st)
{
    launch_chain(k_bwd_prologue, dim3((p->flags & PSLAM_F_GRAD_RAYS) ? ceil_div(p->R * 3, 256) : 1), dim3(256), 0, st, *p);
    PSLAM_CHECK_LAUNCH("bwd_prologue");
    k_composite_bwd_ext<<<ceil_div(p->R, kCompWarps), kCompThreads, 0, st>>>(*p, g_color, g_depth, g_sdf, g_weight);
    PSLAM_CHECK_LAUNCH("composite_bwd_ext");
    return 0;
}

int launch_composite_backward(const pslam_render_t *p, cudaStream_t st, int prologue_done)
{
    if (!prologue_done) {
        launch_chain(k_bwd_prologue, dim3((p->flags & PSLAM_F_GRAD_RAYS) ? ceil_div(p->R * 3, 256) : 1), dim3(256), 0, st, *p);
        PSLAM_CHECK_LAUNCH("bwd_prologue");
    }
    launch_chain(k_composite_bwd, dim3(ceil_div(p->R, kCompWarps)), dim3(kCompThreads), 0, st, *p);
    PSLAM_CHECK_LAUNCH("composite_bwd");
    return 0;
}

}  // namespace pslam
