// Trilinear stages shared by the stand-alone kernels (field_bf.cu) and the fused backward (field_bw.cu).
#pragma once
#include "field.cuh"

namespace pslam {

// position of sample s inside its voxel, p = (x - centre) / voxel_size + 0.5 (render_helpers.py:105-156), with the reference's
// operation order
__device__ __forceinline__ void sample_position(const FieldParams &p, int s, int &vox, int &ray, float &z, float &px, float &py, float &pz)
{
    vox = __ldg(p.samp_vox + s);
    z = __ldg(p.samp_z + s);
    ray = __ldg(p.hit_ray + __ldg(p.samp_ray + s));
    const float x = __fadd_rn(__ldg(p.rays_o + ray * 3 + 0), __fmul_rn(__ldg(p.rays_d + ray * 3 + 0), z));
    const float y = __fadd_rn(__ldg(p.rays_o + ray * 3 + 1), __fmul_rn(__ldg(p.rays_d + ray * 3 + 1), z));
    const float zz = __fadd_rn(__ldg(p.rays_o + ray * 3 + 2), __fmul_rn(__ldg(p.rays_d + ray * 3 + 2), z));
    px = __fadd_rn(__fdiv_rn(__fsub_rn(x, __ldg(p.centres + (size_t)vox * 3 + 0)), p.voxel_size), 0.5f);
    py = __fadd_rn(__fdiv_rn(__fsub_rn(y, __ldg(p.centres + (size_t)vox * 3 + 1)), p.voxel_size), 0.5f);
    pz = __fadd_rn(__fdiv_rn(__fsub_rn(zz, __ldg(p.centres + (size_t)vox * 3 + 2)), p.voxel_size), 0.5f);
}

constexpr int kScatWarps = 8, kScatWPitch = 9, kScatGPitch = 16;
constexpr int kScatWarpFloats = 32 * (kScatWPitch + kScatGPitch + 8);   // shared memory of one warp: weights [32][9] + gradient rows [32][16] + corner ids [32][8]

// Backward of the lookup for the 32 consecutive samples [s0, s0 + 32) by one warp, with warp-aggregated reductions:
//   phase 1  lane = sample: position, the 8 corner weights and the sample's feature-gradient row go to shared memory; with
//            ray gradients, the lane also takes the 8 dot products <g, corner row> and the per-ray sums are reduced over the
//            lanes of a ray (a segmented shuffle reduction: a ray's samples are consecutive) before they touch memory;
//   phase 2  lane = (corner, quarter of the feature row): for every voxel fragment of the 32 samples (~7 consecutive samples
//            share a voxel) the lane adds up w[k][corner] * g[k][quarter] and issues ONE red.v4 (the corner's row id comes from
//            shared memory, where phase 1 left it: no dependent global load per fragment).
// The stand-alone kernel was bound by L2 reductions (32 red.v4 per sample, 6.1 M per mapping iteration); this issues ~5.6.
// FRESH: g_feat was written earlier in the SAME kernel (fused backward): read it through L2, not the read-only path.
template <bool FRESH>
__device__ __forceinline__ void tri_scatter_warp(const FieldParams &p, const float *__restrict__ g_feat, int s0, int nsamp, float *buf)
{
    float *w = buf, *g = buf + 32 * kScatWPitch;
    int *rows = reinterpret_cast<int *>(buf + 32 * (kScatWPitch + kScatGPitch));
    const int lane = threadIdx.x & 31;
    const int s = s0 + lane;
    const bool live = s < nsamp;
    int vox = -1, ray = -1;
    float z = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    float4 gq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        sample_position(p, s, vox, ray, z, px, py, pz);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 *src = reinterpret_cast<const float4 *>(g_feat + (size_t)s * 16 + j * 4);
            gq[j] = FRESH ? __ldcg(src) : __ldg(src);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4 *>(g + lane * kScatGPitch + j * 4) = gq[j];
    float gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
        w[lane * kScatWPitch + i] = (wx * wy) * wz;
        const int row = live ? __ldg(p.vertex_idx + (size_t)vox * 8 + i) : 0;
        rows[lane * 8 + i] = row;
        if (p.grad_rays && live) {
            float d = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p.emb + (size_t)row * 16 + j * 4));
                d = fmaf(gq[j].w, v.w, fmaf(gq[j].z, v.z, fmaf(gq[j].y, v.y, fmaf(gq[j].x, v.x, d))));
            }
            gp[0] += d * ((i & 4) ? 1.0f : -1.0f) * (wy * wz);
            gp[1] += d * ((i & 2) ? 1.0f : -1.0f) * (wx * wz);
            gp[2] += d * ((i & 1) ? 1.0f : -1.0f) * (wx * wy);
        }
    }
    if (p.grad_rays) {
        // per-ray sums over the lanes of a ray: after the sweep the first lane of every ray fragment holds its total
        float so[3], sd[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { so[a] = gp[a] / p.voxel_size; sd[a] = z * so[a]; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int r2 = __shfl_down_sync(0xffffffffu, ray, o);
            const bool take = (lane + o < 32) && r2 == ray;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float x = __shfl_down_sync(0xffffffffu, so[a], o), y = __shfl_down_sync(0xffffffffu, sd[a], o);
                if (take) { so[a] += x; sd[a] += y; }
            }
        }
        const int rprev = __shfl_up_sync(0xffffffffu, ray, 1);
        if (live && (lane == 0 || rprev != ray)) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                atomicAdd(p.g_rays_o + ray * 3 + a, so[a]);
                atomicAdd(p.g_rays_d + ray * 3 + a, sd[a]);
            }
        }
    }
    if (!p.grad_emb) { __syncwarp(); return; }
    const int vprev = __shfl_up_sync(0xffffffffu, vox, 1), rprev2 = __shfl_up_sync(0xffffffffu, ray, 1);
    unsigned heads = __ballot_sync(0xffffffffu, live && (lane == 0 || vprev != vox || rprev2 != ray));
    const int nlive = __popc(__ballot_sync(0xffffffffu, live));
    __syncwarp();
    const int ci = lane >> 2, cq = lane & 3;
    while (heads) {
        const int start = __ffs(heads) - 1;
        heads &= heads - 1;
        const int end = heads ? __ffs(heads) - 1 : nlive;
        const int row = rows[start * 8 + ci];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = start; k < end; ++k) {
            const float wk = w[k * kScatWPitch + ci];
            const float4 gk = *reinterpret_cast<const float4 *>(g + k * kScatGPitch + cq * 4);
            acc.x = fmaf(wk, gk.x, acc.x); acc.y = fmaf(wk, gk.y, acc.y); acc.z = fmaf(wk, gk.z, acc.z); acc.w = fmaf(wk, gk.w, acc.w);
        }
        red_add_v4(p.g_emb + (size_t)row * 16 + cq * 4, acc.x, acc.y, acc.z, acc.w);
    }
    __syncwarp();      // the next pass of this warp rewrites w / g
}

}  // namespace pslam
