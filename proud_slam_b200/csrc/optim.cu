// Optimizer step of the mapping loop as one launch.
//
// Reference: src/mapping.py:81-82 (torch.optim.Adam over the dense [E,16] embedding table and over the decoder) stepped in
// src/variations/render_helpers.py:667-676 after every iteration: ~15 foreach kernels and five passes over E*16 floats although
// only the rows under the current rays (about 10 % at the Replica configuration) have ever seen a gradient.
//
// pslam_adam_step updates up to 16 tensors in one kernel, in place on the tensors torch.optim.Adam owns (param, exp_avg,
// exp_avg_sq, step), with torch's arithmetic (no weight decay, no amsgrad):
//     m = b1 m + (1 - b1) g      v = b2 v + (1 - b2) g^2      p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Rows of a row-structured tensor (the embedding table, row = 16 floats) that have NEVER received a gradient are skipped: their
// state is zero and so is their gradient, for which the update above is exactly zero, so the result equals the dense step bit for
// bit in what it changes; a byte per row remembers "has state".  The gradient can be cleared in the same pass (the kernels
// of the next iteration accumulate into it), which replaces the zero_grad / fill launches.
#include <math.h>

#include "common.cuh"

namespace pslam {

constexpr int kAdamMaxTensors = 16;
struct AdamTable {
    pslam_adam_tensor_t t[kAdamMaxTensors];
    long long first_vec[kAdamMaxTensors + 1];   // prefix of float4 counts
    int count;
    float b1, b2, omb1, omb2, eps, step_value;
    float host_step_size_scale, host_bc2s;   // host-side step count: 1 / (1 - b1^t) and sqrt(1 - b2^t) evaluated in double, like torch
    int zero_grad;
};

// the step counts first (one thread per tensor): every thread of the update kernel then reads the NEW count
__global__ void k_adam_bump(AdamTable tab)
{
    pdl_enter();
    const int i = threadIdx.x;
    if (i < tab.count && tab.t[i].step) *tab.t[i].step += 1.0f;
}

__global__ void __launch_bounds__(256) k_adam_step(AdamTable tab)
{
    pdl_enter();
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index over all tensors
    int k = 0;
#pragma unroll 1
    while (k + 1 < tab.count && v >= tab.first_vec[k + 1]) ++k;
    const bool in = v < tab.first_vec[tab.count];
    const pslam_adam_tensor_t T = tab.t[in ? k : 0];
    const long long e0 = in ? (v - tab.first_vec[k]) * 4 : 0;                   // first element of this thread's float4
    const long long left = in ? T.n - e0 : 0;                                    // (<= 0 in the padding between two tensors)
    const int nval = left >= 4 ? 4 : (left > 0 ? (int)left : 0);                // (tail of a tensor whose size is not a multiple of 4)
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (nval == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(T.grad + e0);
        g[0] = q.x; g[1] = q.y; g[2] = q.z; g[3] = q.w;
    } else {
        for (int j = 0; j < nval; ++j) g[j] = T.grad[e0 + j];
    }
    bool active = nval > 0;
    // (every tensor starts at a multiple of four float4s, so the four threads of a 16-float row are one aligned lane group;
    // the ballot is taken by the whole warp, whatever tensors its lanes belong to)
    const bool nz = (g[0] != 0.f) | (g[1] != 0.f) | (g[2] != 0.f) | (g[3] != 0.f);
    const unsigned m = __ballot_sync(0xffffffffu, nz && nval > 0);
    if (T.row == 16 && T.row_active) {
        // four consecutive threads hold one row: it is live if it ever had a gradient or has one now
        const int lane = threadIdx.x & 31, grp = lane & ~3;
        const bool row_nz = ((m >> grp) & 0xFu) != 0u;
        const long long row = e0 / 16;
        const bool had = nval > 0 ? T.row_active[row] != 0 : false;
        active = nval > 0 && (had || row_nz);
        if (row_nz && !had && (lane & 3) == 0 && nval > 0) T.row_active[row] = 1;
    }
    if (!active) return;
    float step_size, bc2s;
    if (T.step) {                                  // capturable form: torch evaluates the corrections in fp32 tensor ops
        const float t = *T.step;
        step_size = T.lr / (1.0f - powf(tab.b1, t));
        bc2s = sqrtf(1.0f - powf(tab.b2, t));
    } else {
        step_size = T.lr * tab.host_step_size_scale;
        bc2s = tab.host_bc2s;
    }
    float p[4], m1[4], m2[4];
    if (nval == 4) {
        const float4 a = *reinterpret_cast<const float4 *>(T.param + e0), b = *reinterpret_cast<const float4 *>(T.exp_avg + e0),
                     c = *reinterpret_cast<const float4 *>(T.exp_avg_sq + e0);
        p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; m1[0] = b.x; m1[1] = b.y; m1[2] = b.z; m1[3] = b.w;
        m2[0] = c.x; m2[1] = c.y; m2[2] = c.z; m2[3] = c.w;
    } else {
        for (int j = 0; j < nval; ++j) { p[j] = T.param[e0 + j]; m1[j] = T.exp_avg[e0 + j]; m2[j] = T.exp_avg_sq[e0 + j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nval) {
            m1[j] = tab.b1 * m1[j] + tab.omb1 * g[j];
            m2[j] = tab.b2 * m2[j] + tab.omb2 * g[j] * g[j];
            p[j] -= step_size * (m1[j] / (sqrtf(m2[j]) / bc2s + tab.eps));
        }
    }
    if (nval == 4) {
        *reinterpret_cast<float4 *>(T.param + e0) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4 *>(T.exp_avg + e0) = make_float4(m1[0], m1[1], m1[2], m1[3]);
        *reinterpret_cast<float4 *>(T.exp_avg_sq + e0) = make_float4(m2[0], m2[1], m2[2], m2[3]);
        if (tab.zero_grad) *reinterpret_cast<float4 *>(T.grad + e0) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int j = 0; j < nval; ++j) {
            T.param[e0 + j] = p[j]; T.exp_avg[e0 + j] = m1[j]; T.exp_avg_sq[e0 + j] = m2[j];
            if (tab.zero_grad) T.grad[e0 + j] = 0.f;
        }
    }
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_adam_step(const pslam_adam_tensor_t *tensors, int count, double step_value, double beta1, double beta2, double eps,
                               int zero_grad, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(tensors && count > 0 && count <= kAdamMaxTensors, PSLAM_E_ARG, "adam_step: 1..%d tensors", kAdamMaxTensors);
    AdamTable tab = {};
    long long vecs = 0;
    for (int i = 0; i < count; ++i) {
        const pslam_adam_tensor_t &t = tensors[i];
        PSLAM_CHECK_ARG(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.n > 0, PSLAM_E_ARG, "adam_step: tensor %d has a null pointer or no elements", i);
        PSLAM_CHECK_ARG(t.step || step_value >= 1.0, PSLAM_E_ARG, "adam_step: tensor %d has no device step count and no host value was given", i);
        PSLAM_CHECK_ARG(((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) % 16 == 0, PSLAM_E_ALIGN,
                        "adam_step: tensor %d is not 16-byte aligned", i);
        PSLAM_CHECK_ARG(t.row == 0 || (t.row == 16 && t.n % 16 == 0), PSLAM_E_RANGE, "adam_step: row must be 0 (dense) or 16 with n a multiple of 16");
        tab.t[i] = t;
        vecs = (vecs + 3) / 4 * 4;
        tab.first_vec[i] = vecs;
        vecs += (t.n + 3) / 4;
    }
    tab.first_vec[count] = vecs;
    tab.count = count;
    tab.b1 = (float)beta1; tab.b2 = (float)beta2; tab.omb1 = (float)(1.0 - beta1); tab.omb2 = (float)(1.0 - beta2);
    tab.eps = (float)eps; tab.step_value = (float)step_value; tab.zero_grad = zero_grad;
    if (step_value >= 1.0) {
        tab.host_step_size_scale = (float)(1.0 / (1.0 - pow(beta1, step_value)));
        tab.host_bc2s = (float)sqrt(1.0 - pow(beta2, step_value));
    }
    cudaStream_t st = (cudaStream_t)stream;
    launch_chain(k_adam_bump, dim3(1), dim3(32), 0, st, tab);
    PSLAM_CHECK_LAUNCH("adam_bump");
    launch_chain(k_adam_step, dim3((unsigned)ceil_div64(vecs, 256)), dim3(256), 0, st, tab);
    PSLAM_CHECK_LAUNCH("adam_step");
    return 0;
}
