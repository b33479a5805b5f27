// SE(3) pose at the boundary of the tracking loop (src/se3pose.py:24-34, 62-91; render_helpers.py:679-761):
// ray assembly from the 6-vector pose (t, w) and the pose's Adam step from dL/d(rays_o, rays_d).
//
// The reference keeps this in torch: Rodrigues' formula with sin(x)/x and (1-cos x)/x^2 evaluated by their 11-term
// Maclaurin series, autograd through ~100 tiny ops, torch.optim.Adam -- about 250 elementwise launches per tracking
// iteration, more GPU time than the whole render step at 1024 rays.  Here: one kernel builds the rays, one block
// reduces the ray gradients to dL/d(R, t), applies the analytic derivative of the same series (both series are even in
// theta, so R is a polynomial in w and no division by theta appears anywhere) and performs Adam's update in place on
// the tensors torch.optim.Adam(capturable=True) owns (exp_avg, exp_avg_sq, step), so the optimizer object the caller
// gets back stays consistent.
#include "common.cuh"
#include "kernels.h"

namespace pslam {

constexpr int kSeries = 11;
// sin(x)/x = sum (-1)^i x^(2i)/(2i+1)!   ;   (1-cos x)/x^2 = sum (-1)^i x^(2i)/(2i+2)!
__device__ __constant__ float cSerA[kSeries] = {1.0f, -1.0f / 6, 1.0f / 120, -1.0f / 5040, 1.0f / 362880, -1.0f / 39916800,
                                                1.0f / 6227020800.0f, -1.0f / 1307674368000.0f, 1.0f / 355687428096000.0f,
                                                -1.0f / 121645100408832000.0f, 1.0f / 51090942171709440000.0f};
__device__ __constant__ float cSerB[kSeries] = {0.5f, -1.0f / 24, 1.0f / 720, -1.0f / 40320, 1.0f / 3628800, -1.0f / 479001600,
                                                1.0f / 87178291200.0f, -1.0f / 20922789888000.0f, 1.0f / 6402373705728000.0f,
                                                -1.0f / 2432902008176640000.0f, 1.0f / 1124000727777607680000.0f};

// A(theta), B(theta) and their derivatives with respect to theta^2 (x2 = theta^2)
__device__ __forceinline__ void series(float x2, float &A, float &B, float &dA, float &dB)
{
    A = 0.f; B = 0.f; dA = 0.f; dB = 0.f;
    float p = 1.0f, pm = 0.0f;       // x2^i and i * x2^(i-1)
#pragma unroll
    for (int i = 0; i < kSeries; ++i) {
        A = fmaf(cSerA[i], p, A); B = fmaf(cSerB[i], p, B);
        dA = fmaf(cSerA[i], pm, dA); dB = fmaf(cSerB[i], pm, dB);
        pm = (float)(i + 1) * p;
        p *= x2;
    }
}

// R = I + A [w]x + B [w]x^2   (row-major 3x3)
__device__ __forceinline__ void rotation(const float w[3], float R[9])
{
    const float x2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    float A, B, dA, dB;
    series(x2, A, B, dA, dB);
    const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float k2 = K[a * 3] * K[b] + K[a * 3 + 1] * K[3 + b] + K[a * 3 + 2] * K[6 + b];
            R[a * 3 + b] = (a == b ? 1.0f : 0.0f) + A * K[a * 3 + b] + B * k2;
        }
}

// rays_o[i] = t, rays_d[i] = R d_cam[idx[i]], targets gathered with the same indices
__global__ void k_track_assemble(int n, const float *__restrict__ pose, const long long *__restrict__ idx, const float *__restrict__ dirs,
                                 const float *__restrict__ rgb_all, const float *__restrict__ depth_all, float *__restrict__ rays_o,
                                 float *__restrict__ rays_d, float *__restrict__ rgb, float *__restrict__ depth)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w[3] = {pose[3], pose[4], pose[5]};
    float R[9];
    rotation(w, R);
    const long long s = idx[i];
    const float d0 = __ldg(dirs + s * 3), d1 = __ldg(dirs + s * 3 + 1), d2 = __ldg(dirs + s * 3 + 2);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[i * 3 + a] = R[a * 3] * d0 + R[a * 3 + 1] * d1 + R[a * 3 + 2] * d2;
        rays_o[i * 3 + a] = pose[a];
    }
    if (rgb) { rgb[i * 3] = __ldg(rgb_all + s * 3); rgb[i * 3 + 1] = __ldg(rgb_all + s * 3 + 1); rgb[i * 3 + 2] = __ldg(rgb_all + s * 3 + 2); }
    if (depth) depth[i] = __ldg(depth_all + s);
}

// one block: dL/dt = sum g_o, dL/dR[a][b] = sum g_d[a] d_cam[b]  ->  dL/dw through dR/dw  ->  Adam
__global__ void __launch_bounds__(256) k_track_pose_step(int n, float *__restrict__ pose, const long long *__restrict__ idx,
                                                         const float *__restrict__ dirs, const float *__restrict__ g_o,
                                                         const float *__restrict__ g_d, float *__restrict__ exp_avg,
                                                         float *__restrict__ exp_avg_sq, float *__restrict__ step, float lr, float beta1,
                                                         float beta2, float omb1, float omb2, float eps, float step_value, float *__restrict__ grad_out,
                                                         unsigned long long *__restrict__ iter_counter, const int *__restrict__ hit_count,
                                                         unsigned char *__restrict__ hit_mask)
{
    pdl_enter();
    // end-of-iteration bookkeeping of a captured tracking iteration (instead of three torch launches): which rays hit the map
    // (track_frame's third return value, render_helpers.py:741) and the device-side iteration count the next iteration's pixel
    // selection and sampling noise are seeded with
    if (hit_mask && hit_count)
        for (int i = threadIdx.x; i < n; i += 256) hit_mask[i] = hit_count[i] > 0 ? 1 : 0;
    if (iter_counter && threadIdx.x == 255) *iter_counter += 1ull;
    __shared__ float s_part[8][12];
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const long long s = idx[i];
        const float d[3] = {__ldg(dirs + s * 3), __ldg(dirs + s * 3 + 1), __ldg(dirs + s * 3 + 2)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            acc[a] += g_o[i * 3 + a];
            const float g = g_d[i * 3 + a];
#pragma unroll
            for (int b = 0; b < 3; ++b) acc[3 + a * 3 + b] = fmaf(g, d[b], acc[3 + a * 3 + b]);
        }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 12; ++k) s_part[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x != 0) return;
    float G[12];
    for (int k = 0; k < 12; ++k) {
        float t = 0.0f;
        for (int wv = 0; wv < 8; ++wv) t += s_part[wv][k];
        G[k] = t;
    }
    // dR/dw_k = 2 w_k (dA K + dB K^2) + A E_k + B (E_k K + K E_k),  E_k = [e_k]x,  dA = dA/d(theta^2)
    const float w[3] = {pose[3], pose[4], pose[5]};
    const float x2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    float A, B, dA, dB;
    series(x2, A, B, dA, dB);
    const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    float K2[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) K2[a * 3 + b] = K[a * 3] * K[b] + K[a * 3 + 1] * K[3 + b] + K[a * 3 + 2] * K[6 + b];
    float grad[6] = {G[0], G[1], G[2], 0.f, 0.f, 0.f};
    for (int k = 0; k < 3; ++k) {
        float E[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (k == 0) { E[5] = -1.f; E[7] = 1.f; }
        if (k == 1) { E[2] = 1.f; E[6] = -1.f; }
        if (k == 2) { E[1] = -1.f; E[3] = 1.f; }
        float gk = 0.0f;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                float ek = 0.f, ke = 0.f;
                for (int c = 0; c < 3; ++c) { ek = fmaf(E[a * 3 + c], K[c * 3 + b], ek); ke = fmaf(K[a * 3 + c], E[c * 3 + b], ke); }
                const float dR = 2.0f * w[k] * (dA * K[a * 3 + b] + dB * K2[a * 3 + b]) + A * E[a * 3 + b] + B * (ek + ke);
                gk = fmaf(G[3 + a * 3 + b], dR, gk);
            }
        grad[3 + k] = gk;
    }
    // torch.optim.Adam (no weight decay, no amsgrad), capturable form: the step count is a float tensor on the device
    // step == NULL: the optimizer keeps its count on the host (non-capturable Adam) and passes the new value
    const float t = step ? *step + 1.0f : step_value;
    if (step) *step = t;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    for (int k = 0; k < 6; ++k) {
        const float g = grad[k];
        const float m = exp_avg[k] = beta1 * exp_avg[k] + omb1 * g;           // omb = 1 - beta, rounded once from double like torch
        const float v = exp_avg_sq[k] = beta2 * exp_avg_sq[k] + omb2 * g * g;
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        pose[k] -= (lr / bc1) * (m / denom);
        if (grad_out) grad_out[k] = g;
    }
}

// Pixel selection of a tracking / mapping iteration: n DISTINCT pixels out of hw, uniformly (the reference draws them with
// a Gumbel top-k over uniform weights, src/utils/sample_util.py:4-20, i.e. uniformly without replacement; its cost there
// is a sort-like pass over all H*W pixels per frame and iteration).  Here thread i evaluates a keyed pseudo-random
// PERMUTATION of [0, hw) at i: a 4-round Feistel network on the next even number of bits, cycle-walked back into range.
// Distinct inputs give distinct outputs by construction, so no rejection, no sort and no scratch memory.
__device__ __forceinline__ uint32_t feistel_round(uint32_t x, uint32_t k)
{
    x = (x ^ k) * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u;
    x ^= x >> 13;
    return x;
}
// element i of a keyed permutation of [0, hw): a 4-round Feistel network over the smallest even bit width that covers hw, cycle-
// walked back into range; (seed, seed_dev) pick the permutation
__device__ __forceinline__ long long permuted_pixel(int i, long long hw, unsigned long long seed, const unsigned long long *__restrict__ seed_dev)
{
    unsigned long long key = seed + (seed_dev ? *seed_dev : 0ull);
    key ^= key >> 30; key *= 0xBF58476D1CE4E5B9ull;
    key ^= key >> 27; key *= 0x94D049BB133111EBull;
    key ^= key >> 31;
    int bits = 2;
    while (bits < 62 && (1ll << bits) < hw) bits += 2;       // even width >= log2(hw)
    const int hb = bits >> 1;
    const uint32_t hmask = (hb >= 32) ? 0xffffffffu : ((1u << hb) - 1u);
    unsigned long long x = (unsigned long long)i;
    do {
        uint32_t L = (uint32_t)(x >> hb) & hmask, Rr = (uint32_t)x & hmask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t t = L ^ (feistel_round(Rr, (uint32_t)(key >> (16 * r)) + 0x632BE5ABu * (uint32_t)r) & hmask);
            L = Rr; Rr = t;
        }
        x = ((unsigned long long)L << hb) | Rr;
    } while ((long long)x >= hw);
    return (long long)x;
}

__global__ void k_sample_pixels(int n, long long hw, unsigned long long seed, const unsigned long long *__restrict__ seed_dev,
                                long long *__restrict__ idx)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = permuted_pixel(i, hw, seed, seed_dev);
}

// pixel selection + ray assembly of a tracking iteration in one launch (k_sample_pixels + k_track_assemble)
__global__ void k_track_sample_assemble(int n, long long hw, unsigned long long seed, const unsigned long long *__restrict__ seed_dev,
                                        const float *__restrict__ pose, const float *__restrict__ dirs, const float *__restrict__ rgb_all,
                                        const float *__restrict__ depth_all, long long *__restrict__ idx, float *__restrict__ rays_o,
                                        float *__restrict__ rays_d, float *__restrict__ rgb, float *__restrict__ depth)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = permuted_pixel(i, hw, seed, seed_dev);
    idx[i] = s;
    const float w[3] = {pose[3], pose[4], pose[5]};
    float R[9];
    rotation(w, R);
    const float d0 = __ldg(dirs + s * 3), d1 = __ldg(dirs + s * 3 + 1), d2 = __ldg(dirs + s * 3 + 2);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[i * 3 + a] = R[a * 3] * d0 + R[a * 3 + 1] * d1 + R[a * 3 + 2] * d2;
        rays_o[i * 3 + a] = pose[a];
    }
    if (rgb) { rgb[i * 3] = __ldg(rgb_all + s * 3); rgb[i * 3 + 1] = __ldg(rgb_all + s * 3 + 1); rgb[i * 3 + 2] = __ldg(rgb_all + s * 3 + 2); }
    if (depth) depth[i] = __ldg(depth_all + s);
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_sample_pixels(int n, long long hw, unsigned long long seed, const unsigned long long *seed_dev, long long *idx,
                                   pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(n > 0 && hw > 0 && idx, PSLAM_E_ARG, "sample_pixels: bad argument");
    PSLAM_CHECK_ARG((long long)n <= hw, PSLAM_E_RANGE, "sample_pixels: cannot draw %d distinct pixels out of %lld", n, hw);
    launch_chain(k_sample_pixels, dim3(ceil_div(n, 128)), dim3(128), 0, (cudaStream_t)stream, n, hw, seed, seed_dev, idx);
    PSLAM_CHECK_LAUNCH("sample_pixels");
    return 0;
}

extern "C" int pslam_track_assemble(int n, const float *pose6, const long long *idx, const float *rays_d_cam, const float *rgb_all,
                                    const float *depth_all, float *rays_o, float *rays_d, float *rgb, float *depth, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(n > 0 && pose6 && idx && rays_d_cam && rays_o && rays_d, PSLAM_E_ARG, "track_assemble: bad argument");
    PSLAM_CHECK_ARG((rgb == nullptr || rgb_all) && (depth == nullptr || depth_all), PSLAM_E_ARG, "track_assemble: targets without their source");
    launch_chain(k_track_assemble, dim3(ceil_div(n, 128)), dim3(128), 0, (cudaStream_t)stream, n, pose6, idx, rays_d_cam, rgb_all, depth_all, rays_o, rays_d, rgb, depth);
    PSLAM_CHECK_LAUNCH("track_assemble");
    return 0;
}

extern "C" int pslam_track_sample_assemble(int n, long long hw, unsigned long long seed, const unsigned long long *seed_dev, const float *pose6,
                                           const float *rays_d_cam, const float *rgb_all, const float *depth_all, long long *idx, float *rays_o,
                                           float *rays_d, float *rgb, float *depth, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(n > 0 && hw > 0 && pose6 && idx && rays_d_cam && rays_o && rays_d, PSLAM_E_ARG, "track_sample_assemble: bad argument");
    PSLAM_CHECK_ARG((long long)n <= hw, PSLAM_E_RANGE, "track_sample_assemble: cannot draw %d distinct pixels out of %lld", n, hw);
    PSLAM_CHECK_ARG((rgb == nullptr || rgb_all) && (depth == nullptr || depth_all), PSLAM_E_ARG, "track_sample_assemble: targets without their source");
    launch_chain(k_track_sample_assemble, dim3(ceil_div(n, 128)), dim3(128), 0, (cudaStream_t)stream, n, hw, seed, seed_dev, pose6, rays_d_cam, rgb_all, depth_all,
                 idx, rays_o, rays_d, rgb, depth);
    PSLAM_CHECK_LAUNCH("track_sample_assemble");
    return 0;
}

extern "C" int pslam_track_pose_step(int n, float *pose6, const long long *idx, const float *rays_d_cam, const float *g_rays_o,
                                     const float *g_rays_d, float *exp_avg, float *exp_avg_sq, float *step, double step_value, double lr,
                                     double beta1, double beta2, double eps, float *grad_out, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(n > 0 && pose6 && idx && rays_d_cam && g_rays_o && g_rays_d && exp_avg && exp_avg_sq && (step || step_value >= 1.0),
                    PSLAM_E_ARG, "track_pose_step: bad argument");
    launch_chain(k_track_pose_step, dim3(1), dim3(256), 0, (cudaStream_t)stream, n, pose6, idx, rays_d_cam, g_rays_o, g_rays_d, exp_avg, exp_avg_sq, step, (float)lr,
                                                           (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)step_value, grad_out,
                                                           (unsigned long long *)nullptr, (const int *)nullptr, (unsigned char *)nullptr);
    PSLAM_CHECK_LAUNCH("track_pose_step");
    return 0;
}

extern "C" int pslam_track_pose_step_iter(int n, float *pose6, const long long *idx, const float *rays_d_cam, const float *g_rays_o,
                                          const float *g_rays_d, float *exp_avg, float *exp_avg_sq, float *step, double lr, double beta1,
                                          double beta2, double eps, unsigned long long *iter_counter, const int *hit_count,
                                          unsigned char *hit_mask, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(n > 0 && pose6 && idx && rays_d_cam && g_rays_o && g_rays_d && exp_avg && exp_avg_sq && step, PSLAM_E_ARG,
                    "track_pose_step_iter: bad argument");
    PSLAM_CHECK_ARG((hit_mask == nullptr) == (hit_count == nullptr), PSLAM_E_ARG, "track_pose_step_iter: hit_mask and hit_count go together");
    launch_chain(k_track_pose_step, dim3(1), dim3(256), 0, (cudaStream_t)stream, n, pose6, idx, rays_d_cam, g_rays_o, g_rays_d, exp_avg, exp_avg_sq, step, (float)lr,
                 (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, 0.0f, (float *)nullptr, iter_counter, hit_count, hit_mask);
    PSLAM_CHECK_LAUNCH("track_pose_step_iter");
    return 0;
}
