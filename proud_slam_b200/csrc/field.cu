// Kernels 3+4: trilinear lookup of voxel-corner embeddings fused with the
// SDF/colour decoder MLP, forward and backward (fp32 SIMT build).
//
// Math follows the reference's Python:
//   * get_features_vox / trilinear_interp  src/variations/render_helpers.py:105-156, 47-59
//       p = (x - centre)/vs + .5 ; w_i = prod_a (q_ia ? p_a : 1-p_a), corner i = (i>>2&1, i>>1&1, i&1)
//       f = sum_i w_i * emb[vertex_idx[vox, i]]
//   * Decoder.get_values                    src/variations/nrgbd.py:116-135
//       h1 = relu(W1 f + b1); h2 = relu(W2 h1 + b2); o = W3 h2 + b3 (sdf = o[0], t = o[1:129]);
//       hc = relu(W4 [t; f] + b4); rgb = sigmoid(W5 hc + b5); output order (r,g,b,sdf)
//
// What is different from the reference (which runs ~20 torch kernels per 10 000-sample
// chunk and keeps 657 floats/sample of activations in HBM for autograd): one persistent
// CTA per SM walks tiles of M samples; the gathered features and every activation of the
// tile live in shared memory ([feature][sample], sample-contiguous); weights are streamed
// from L2 through a cp.async double buffer; backward recomputes the forward in shared
// memory and overwrites each activation with its gradient in place, so nothing but the
// [P,4] decoder output and its gradient ever touches HBM.  The embedding-gradient scatter
// is aggregated over runs of consecutive samples that share a voxel before it is sent to
// L2 as red.global.add.v4.f32.
#include "common.cuh"
#include "field.cuh"
#include "kernels.h"

namespace pslam {

constexpr int kFieldThreads = 256;
constexpr int kKC = 16;  // reduction rows staged per cp.async chunk

template <int W>
struct FieldCfg {
    static_assert(W == 128 || W == 256, "decoder width must be 128 or 256");
    static constexpr int M = (W == 128) ? 64 : 32;  // samples per tile
    static constexpr int PITCH = M + 4;             // == 4 (mod 32) words: conflict-free row-strided float4 reads
    static constexpr int RM = M / 16;               // rows (samples) per thread in the 16x16 thread grid
    static constexpr int CH = kFieldThreads / M;    // threads per sample in the gather phases
    static constexpr int FPT = 16 / CH;             // embedding floats per gather thread
    static constexpr int PARTS = kFieldThreads / M; // reduction slices for the narrow heads
    // activation rows
    static constexpr int rF = 0, rH1 = 16, rH2 = 16 + W, rT = 16 + 2 * W, rHC = 144 + 2 * W, rGF = 144 + 3 * W;
    static constexpr int ROWS = 160 + 3 * W;
    // float offsets of the other shared arrays
    static constexpr int oStage = ROWS * PITCH;
    static constexpr int oP = oStage + 2 * kKC * 128;  // p[3][M]
    static constexpr int oZ = oP + 3 * M;              // z[M]
    static constexpr int oOut = oZ + M;                // out[4][M]  (r,g,b,sdf)
    static constexpr int oGo = oOut + 4 * M;           // go[4][M]   (d pre-sigmoid rgb, d sdf)
    static constexpr int oRed = oGo + 4 * M;           // red[PARTS][4][M]
    static constexpr int oVox = oRed + PARTS * 4 * M;  // int vox[M]
    static constexpr int oRay = oVox + M;              // int ray[M]  (ray id, -1 = none)
    static constexpr int oGx = oRay + M;               // gx[3][M]  d/dxyz
    static constexpr int FLOATS = oGx + 3 * M;
    static constexpr size_t SMEM = sizeof(float) * FLOATS;
    // packed transposed weights (floats)
    static constexpr int wW1t = 0, wW2t = 16 * W, wW3t = 16 * W + W * W, wW4t = 16 * W + W * W + 128 * W;
    static constexpr int WS = W * W + 288 * W;             // floats of the SIMT pack; the tcgen05 stream follows it
};

// ------------------------------------------------------------------------------------------
// weight packing: transposes so that the forward GEMMs stream [k][n] rows
// ------------------------------------------------------------------------------------------
template <int W>
__global__ void k_pack_decoder(pslam_decoder_t d, float *__restrict__ ws)
{
    using C = FieldCfg<W>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C::WS) return;
    float v;
    if (i < C::wW2t) {            // W1t[k][n] = W1[n][k]
        const int k = i / W, n = i % W;
        v = d.W1[n * 16 + k];
    } else if (i < C::wW3t) {     // W2t[k][n] = W2[n][k]
        const int j = i - C::wW2t, k = j / W, n = j % W;
        v = d.W2[n * W + k];
    } else if (i < C::wW4t) {     // W3t[k][j] = W3[1+j][k]
        const int j = i - C::wW3t, k = j / 128, n = j % 128;
        v = d.W3[(1 + n) * W + k];
    } else {                      // W4t[k][n] = W4[n][k], k over [t(128); f(16)]
        const int j = i - C::wW4t, k = j / W, n = j % W;
        v = d.W4[n * 144 + k];
    }
    ws[i] = v;
}

// ------------------------------------------------------------------------------------------
// tile GEMM building blocks (256 threads as a 16x16 grid: tx = column group, ty = row group)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc[i][j] += sum_k sA[k][ty*RM+i] * B[k][col(j)],  col(j) = tx*4 + (j&3) + 64*(j>>2),
// B = global row-major [K][ldb] (already offset to the 128-column block), K % 16 == 0.
template <int M>
__device__ __forceinline__ void gemm128(const float *sA, int K, const float *__restrict__ gB, int ldb, float *sStage,
                                        float (&acc)[M / 16][8])
{
    constexpr int RM = M / 16, PITCH = M + 4;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    auto prefetch = [&](int k0, int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + h * kFieldThreads;  // 512 float4 per chunk
            const int row = idx >> 5, c4 = idx & 31;
            cp_async16(sStage + buf * (kKC * 128) + row * 128 + c4 * 4, gB + (size_t)(k0 + row) * ldb + c4 * 4);
        }
        cp_async_commit();
    };
    prefetch(0, 0);
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += kKC, buf ^= 1) {
        if (k0 + kKC < K) {
            prefetch(k0 + kKC, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float *sB = sStage + buf * (kKC * 128);
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            float a[RM];
            const float *ap = sA + (k0 + kk) * PITCH + ty * RM;
            if (RM == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(ap);
                a[0] = t.x; a[1] = t.y; a[2 % RM] = t.z; a[3 % RM] = t.w;
            } else {
                const float2 t = *reinterpret_cast<const float2 *>(ap);
                a[0] = t.x; a[1] = t.y;
            }
            const float4 b0 = *reinterpret_cast<const float4 *>(sB + kk * 128 + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(sB + kk * 128 + 64 + tx * 4);
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// narrow block: acc[i] += sum_k sA[k][ty*RM+i] * B[k][tx], B global [K][ldb] read through L1.
template <int M>
__device__ __forceinline__ void gemm16(const float *sA, int K, const float *__restrict__ gB, int ldb, float (&acc)[M / 16])
{
    constexpr int RM = M / 16, PITCH = M + 4;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float b = __ldg(gB + (size_t)k * ldb + tx);
        const float *ap = sA + k * PITCH + ty * RM;
#pragma unroll
        for (int i = 0; i < RM; ++i) acc[i] = fmaf(ap[i], b, acc[i]);
    }
}

// dW[nb+ty+16a][kb+tx+16b] += sum_m sG[ty+16a][m] * sX[tx+16b][m]   (a,b < 8)
template <int M>
__device__ __forceinline__ void wgrad128(const float *sG, const float *sX, float *__restrict__ dW, int ldw)
{
    constexpr int PITCH = M + 4;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;
#pragma unroll 1
    for (int m = 0; m < M; m += 4) {
        float4 g[8], x[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) g[a] = *reinterpret_cast<const float4 *>(sG + (ty + 16 * a) * PITCH + m);
#pragma unroll
        for (int b = 0; b < 8; ++b) x[b] = *reinterpret_cast<const float4 *>(sX + (tx + 16 * b) * PITCH + m);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                acc[a][b] = fmaf(g[a].x, x[b].x, acc[a][b]);
                acc[a][b] = fmaf(g[a].y, x[b].y, acc[a][b]);
                acc[a][b] = fmaf(g[a].z, x[b].z, acc[a][b]);
                acc[a][b] = fmaf(g[a].w, x[b].w, acc[a][b]);
            }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) atomicAdd(dW + (size_t)(ty + 16 * a) * ldw + tx + 16 * b, acc[a][b]);
}

// dW[ty+16a][tx] += sum_m sG[ty+16a][m] * sX[tx][m]   (16 input columns)
template <int M>
__device__ __forceinline__ void wgrad16(const float *sG, const float *sX, float *__restrict__ dW, int ldw)
{
    constexpr int PITCH = M + 4;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
    for (int m = 0; m < M; m += 4) {
        const float4 x = *reinterpret_cast<const float4 *>(sX + tx * PITCH + m);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const float4 g = *reinterpret_cast<const float4 *>(sG + (ty + 16 * a) * PITCH + m);
            acc[a] = fmaf(g.x, x.x, acc[a]); acc[a] = fmaf(g.y, x.y, acc[a]);
            acc[a] = fmaf(g.z, x.z, acc[a]); acc[a] = fmaf(g.w, x.w, acc[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) atomicAdd(dW + (size_t)(ty + 16 * a) * ldw + tx, acc[a]);
}

// db[r] += sum_m sG[r][m] for r < rows (one thread per row)
template <int M>
__device__ __forceinline__ void bias_grad(const float *sG, int rows, float *__restrict__ db)
{
    constexpr int PITCH = M + 4;
    for (int r = threadIdx.x; r < rows; r += kFieldThreads) {
        float s = 0.0f;
#pragma unroll 4
        for (int m = 0; m < M; ++m) s += sG[r * PITCH + m];
        atomicAdd(db + r, s);
    }
}

// trilinear corner weight and its derivative factors
__device__ __forceinline__ float corner_w(int i, float px, float py, float pz)
{
    const float wx = (i & 4) ? px : 1.0f - px;
    const float wy = (i & 2) ? py : 1.0f - py;
    const float wz = (i & 1) ? pz : 1.0f - pz;
    return (wx * wy) * wz;
}

// ------------------------------------------------------------------------------------------
// the fused tile kernel
// ------------------------------------------------------------------------------------------
template <int W, bool BWD>
__global__ void __launch_bounds__(kFieldThreads, 1) k_field(FieldParams p)
{
    using C = FieldCfg<W>;
    constexpr int M = C::M, PITCH = C::PITCH, RM = C::RM;
    extern __shared__ __align__(16) float smem[];
    float *sAct = smem;
    float *sStage = smem + C::oStage;
    float *sP = smem + C::oP, *sZ = smem + C::oZ, *sOut = smem + C::oOut, *sGo = smem + C::oGo, *sRed = smem + C::oRed;
    int *sVox = reinterpret_cast<int *>(smem + C::oVox), *sRay = reinterpret_cast<int *>(smem + C::oRay);
    float *sGx = smem + C::oGx;
    float *sF = sAct + C::rF * PITCH, *sH1 = sAct + C::rH1 * PITCH, *sH2 = sAct + C::rH2 * PITCH;
    float *sT = sAct + C::rT * PITCH, *sHC = sAct + C::rHC * PITCH, *sGF = sAct + C::rGF * PITCH;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + M - 1) / M;
    const float *ws = p.ws;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s0 = tile * M;
        // ---------------- phase 0: features of the tile -> sF[16][M] ----------------
        {
            const int sl = tid / C::CH, c = tid % C::CH;  // local sample, chunk
            const int s = s0 + sl;
            float f[C::FPT];
#pragma unroll
            for (int e = 0; e < C::FPT; ++e) f[e] = 0.0f;
            if (p.feat) {
                if (s < nsamp) {
#pragma unroll
                    for (int e = 0; e < C::FPT; ++e) f[e] = __ldg(p.feat + (size_t)s * 16 + c * C::FPT + e);
                }
            } else {
                int vox = -1, ray = -1;
                float px = 0.f, py = 0.f, pz = 0.f, z = 0.f;
                if (s < nsamp) {
                    vox = __ldg(p.samp_vox + s);
                    z = __ldg(p.samp_z + s);
                    ray = __ldg(p.hit_ray + __ldg(p.samp_ray + s));
                    // xyz = o + d*z (render_helpers.py:436-437, separate mul and add)
                    const float x = __fadd_rn(__ldg(p.rays_o + ray * 3 + 0), __fmul_rn(__ldg(p.rays_d + ray * 3 + 0), z));
                    const float y = __fadd_rn(__ldg(p.rays_o + ray * 3 + 1), __fmul_rn(__ldg(p.rays_d + ray * 3 + 1), z));
                    const float zz = __fadd_rn(__ldg(p.rays_o + ray * 3 + 2), __fmul_rn(__ldg(p.rays_d + ray * 3 + 2), z));
                    // p = (x - c)/vs + .5 (render_helpers.py:91-93)
                    px = __fadd_rn(__fdiv_rn(__fsub_rn(x, __ldg(p.centres + (size_t)vox * 3 + 0)), p.voxel_size), 0.5f);
                    py = __fadd_rn(__fdiv_rn(__fsub_rn(y, __ldg(p.centres + (size_t)vox * 3 + 1)), p.voxel_size), 0.5f);
                    pz = __fadd_rn(__fdiv_rn(__fsub_rn(zz, __ldg(p.centres + (size_t)vox * 3 + 2)), p.voxel_size), 0.5f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = __ldg(p.vertex_idx + (size_t)vox * 8 + i);
                        const float w = corner_w(i, px, py, pz);
                        const float *er = p.emb + (size_t)row * 16 + c * C::FPT;
                        if (C::FPT == 4) {
                            const float4 v = __ldg(reinterpret_cast<const float4 *>(er));
                            f[0] = fmaf(w, v.x, f[0]); f[1] = fmaf(w, v.y, f[1]);
                            f[2 % C::FPT] = fmaf(w, v.z, f[2 % C::FPT]); f[3 % C::FPT] = fmaf(w, v.w, f[3 % C::FPT]);
                        } else {
                            const float2 v = __ldg(reinterpret_cast<const float2 *>(er));
                            f[0] = fmaf(w, v.x, f[0]); f[1] = fmaf(w, v.y, f[1]);
                        }
                    }
                }
                if (c == 0) {
                    sP[sl] = px; sP[M + sl] = py; sP[2 * M + sl] = pz;
                    sZ[sl] = z; sVox[sl] = vox; sRay[sl] = ray;
                }
            }
#pragma unroll
            for (int e = 0; e < C::FPT; ++e) sF[(c * C::FPT + e) * PITCH + sl] = f[e];
        }
        __syncthreads();

        // ---------------- forward ----------------
        // L1: 16 -> W, relu
#pragma unroll 1
        for (int cb = 0; cb < W / 128; ++cb) {
            float acc[RM][8] = {};
            gemm128<M>(sF, 16, ws + C::wW1t + cb * 128, W, sStage, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = cb * 128 + tx * 4 + (j & 3) + 64 * (j >> 2);
                const float b = __ldg(p.dec.b1 + n);
#pragma unroll
                for (int i = 0; i < RM; ++i) sH1[n * PITCH + ty * RM + i] = fmaxf(acc[i][j] + b, 0.0f);
            }
        }
        __syncthreads();
        // L2: W -> W, relu
#pragma unroll 1
        for (int cb = 0; cb < W / 128; ++cb) {
            float acc[RM][8] = {};
            gemm128<M>(sH1, W, ws + C::wW2t + cb * 128, W, sStage, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = cb * 128 + tx * 4 + (j & 3) + 64 * (j >> 2);
                const float b = __ldg(p.dec.b2 + n);
#pragma unroll
                for (int i = 0; i < RM; ++i) sH2[n * PITCH + ty * RM + i] = fmaxf(acc[i][j] + b, 0.0f);
            }
        }
        __syncthreads();
        // L3: W -> 128 features (rows 1..128 of W3), no activation
        {
            float acc[RM][8] = {};
            gemm128<M>(sH2, W, ws + C::wW3t, 128, sStage, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = tx * 4 + (j & 3) + 64 * (j >> 2);
                const float b = __ldg(p.dec.b3 + 1 + n);
#pragma unroll
                for (int i = 0; i < RM; ++i) sT[n * PITCH + ty * RM + i] = acc[i][j] + b;
            }
        }
        // sdf head: row 0 of W3 (partial sums over PARTS slices of k)
        {
            const int sl = tid % M, part = tid / M;
            constexpr int KS = W / C::PARTS;
            float s = 0.0f;
#pragma unroll 8
            for (int k = part * KS; k < (part + 1) * KS; ++k) s = fmaf(sH2[k * PITCH + sl], __ldg(p.dec.W3 + k), s);
            sRed[(part * 4 + 3) * M + sl] = s;
        }
        __syncthreads();
        // L4: [t(128); f(16)] -> W, relu
#pragma unroll 1
        for (int cb = 0; cb < W / 128; ++cb) {
            float acc[RM][8] = {};
            gemm128<M>(sT, 128, ws + C::wW4t + cb * 128, W, sStage, acc);
            gemm128<M>(sF, 16, ws + C::wW4t + 128 * W + cb * 128, W, sStage, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = cb * 128 + tx * 4 + (j & 3) + 64 * (j >> 2);
                const float b = __ldg(p.dec.b4 + n);
#pragma unroll
                for (int i = 0; i < RM; ++i) sHC[n * PITCH + ty * RM + i] = fmaxf(acc[i][j] + b, 0.0f);
            }
        }
        __syncthreads();
        // rgb head: W -> 3
        {
            const int sl = tid % M, part = tid / M;
            constexpr int KS = W / C::PARTS;
            float r = 0.0f, g = 0.0f, b = 0.0f;
#pragma unroll 8
            for (int k = part * KS; k < (part + 1) * KS; ++k) {
                const float h = sHC[k * PITCH + sl];
                r = fmaf(h, __ldg(p.dec.W5 + k), r);
                g = fmaf(h, __ldg(p.dec.W5 + W + k), g);
                b = fmaf(h, __ldg(p.dec.W5 + 2 * W + k), b);
            }
            sRed[(part * 4 + 0) * M + sl] = r;
            sRed[(part * 4 + 1) * M + sl] = g;
            sRed[(part * 4 + 2) * M + sl] = b;
        }
        __syncthreads();
        if (tid < M) {
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float s = 0.0f;
#pragma unroll
                for (int part = 0; part < C::PARTS; ++part) s += sRed[(part * 4 + c) * M + tid];
                o[c] = s;
            }
            const float r = 1.0f / (1.0f + expf(-(o[0] + __ldg(p.dec.b5 + 0))));
            const float g = 1.0f / (1.0f + expf(-(o[1] + __ldg(p.dec.b5 + 1))));
            const float b = 1.0f / (1.0f + expf(-(o[2] + __ldg(p.dec.b5 + 2))));
            const float sdf = o[3] + __ldg(p.dec.b3);
            sOut[tid] = r; sOut[M + tid] = g; sOut[2 * M + tid] = b; sOut[3 * M + tid] = sdf;
            if (!BWD && s0 + tid < nsamp)
                *reinterpret_cast<float4 *>(p.out + (size_t)(s0 + tid) * 4) = make_float4(r, g, b, sdf);
        }
        if (!BWD) {
            __syncthreads();
            continue;
        }

        // ---------------- backward ----------------
        if (tid < M) {
            float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + tid < nsamp) go = __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)(s0 + tid) * 4));
            const float r = sOut[tid], g = sOut[M + tid], b = sOut[2 * M + tid];
            sGo[tid] = go.x * (1.0f - r) * r;           // sigmoid backward: grad*(1-y)*y
            sGo[M + tid] = go.y * (1.0f - g) * g;
            sGo[2 * M + tid] = go.z * (1.0f - b) * b;
            sGo[3 * M + tid] = go.w;
        }
        __syncthreads();
        // (a) dW5, db5 from hc
        if (p.grad_dec) {
            for (int k = tid; k < W; k += kFieldThreads) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
                for (int m = 0; m < M; ++m) {
                    const float h = sHC[k * PITCH + m];
                    a0 = fmaf(sGo[m], h, a0); a1 = fmaf(sGo[M + m], h, a1); a2 = fmaf(sGo[2 * M + m], h, a2);
                }
                atomicAdd(p.g_dec.W5 + k, a0); atomicAdd(p.g_dec.W5 + W + k, a1); atomicAdd(p.g_dec.W5 + 2 * W + k, a2);
            }
            if (tid >= 128 && tid < 132) {  // db5[0..2], db3[0]
                const int c = tid - 128;
                float s = 0.0f;
                for (int m = 0; m < M; ++m) s += sGo[c * M + m];
                atomicAdd(c < 3 ? p.g_dec.b5 + c : p.g_dec.b3, s);
            }
        }
        __syncthreads();
        // (b) g_hc = relu'(hc) * W5^T g_o5, in place
        for (int e = tid; e < W * M; e += kFieldThreads) {
            const int k = e / M, m = e % M;
            const float h = sHC[k * PITCH + m];
            const float g = __ldg(p.dec.W5 + k) * sGo[m] + __ldg(p.dec.W5 + W + k) * sGo[M + m] +
                            __ldg(p.dec.W5 + 2 * W + k) * sGo[2 * M + m];
            sHC[k * PITCH + m] = (h > 0.0f) ? g : 0.0f;
        }
        __syncthreads();
        // (c) dW4 [W][144], db4
        if (p.grad_dec) {
#pragma unroll 1
            for (int nb = 0; nb < W / 128; ++nb) {
                wgrad128<M>(sHC + nb * 128 * PITCH, sT, p.g_dec.W4 + (size_t)nb * 128 * 144, 144);
                wgrad16<M>(sHC + nb * 128 * PITCH, sF, p.g_dec.W4 + (size_t)nb * 128 * 144 + 128, 144);
            }
            bias_grad<M>(sHC, W, p.g_dec.b4);
        }
        __syncthreads();
        // (d) g_t = W4[:, :128]^T g_hc -> sT ; g_f(part 2) = W4[:, 128:]^T g_hc -> sGF
        {
            float acc[RM][8] = {};
            gemm128<M>(sHC, W, p.dec.W4, 144, sStage, acc);
            float acc16[RM] = {};
            gemm16<M>(sHC, W, p.dec.W4 + 128, 144, acc16);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = tx * 4 + (j & 3) + 64 * (j >> 2);
#pragma unroll
                for (int i = 0; i < RM; ++i) sT[n * PITCH + ty * RM + i] = acc[i][j];
            }
#pragma unroll
            for (int i = 0; i < RM; ++i) sGF[tx * PITCH + ty * RM + i] = acc16[i];
        }
        __syncthreads();
        // (e) dW3 [129][W], db3[1:]
        if (p.grad_dec) {
#pragma unroll 1
            for (int kb = 0; kb < W / 128; ++kb)
                wgrad128<M>(sT, sH2 + kb * 128 * PITCH, p.g_dec.W3 + W + kb * 128, W);
            for (int k = tid; k < W; k += kFieldThreads) {  // row 0 (sdf)
                float a = 0.0f;
#pragma unroll 4
                for (int m = 0; m < M; ++m) a = fmaf(sGo[3 * M + m], sH2[k * PITCH + m], a);
                atomicAdd(p.g_dec.W3 + k, a);
            }
            bias_grad<M>(sT, 128, p.g_dec.b3 + 1);
        }
        __syncthreads();
        // (f) g_h2 = relu'(h2) * (W3[1:]^T g_t + W3[0] g_sdf), in place
#pragma unroll 1
        for (int cb = 0; cb < W / 128; ++cb) {
            float acc[RM][8] = {};
            gemm128<M>(sT, 128, p.dec.W3 + W + cb * 128, W, sStage, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = cb * 128 + tx * 4 + (j & 3) + 64 * (j >> 2);
                const float w0 = __ldg(p.dec.W3 + n);
#pragma unroll
                for (int i = 0; i < RM; ++i) {
                    const int m = ty * RM + i;
                    const float h = sH2[n * PITCH + m];
                    sH2[n * PITCH + m] = (h > 0.0f) ? fmaf(w0, sGo[3 * M + m], acc[i][j]) : 0.0f;
                }
            }
        }
        __syncthreads();
        // (g) dW2, db2
        if (p.grad_dec) {
#pragma unroll 1
            for (int nb = 0; nb < W / 128; ++nb)
#pragma unroll 1
                for (int kb = 0; kb < W / 128; ++kb)
                    wgrad128<M>(sH2 + nb * 128 * PITCH, sH1 + kb * 128 * PITCH, p.g_dec.W2 + (size_t)nb * 128 * W + kb * 128, W);
            bias_grad<M>(sH2, W, p.g_dec.b2);
        }
        __syncthreads();
        // (h) g_h1 = relu'(h1) * W2^T g_h2, in place
        {
            float accs[W / 128][RM][8] = {};
#pragma unroll
            for (int cb = 0; cb < W / 128; ++cb) gemm128<M>(sH2, W, p.dec.W2 + cb * 128, W, sStage, accs[cb]);
#pragma unroll
            for (int cb = 0; cb < W / 128; ++cb)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = cb * 128 + tx * 4 + (j & 3) + 64 * (j >> 2);
#pragma unroll
                    for (int i = 0; i < RM; ++i) {
                        const int m = ty * RM + i;
                        const float h = sH1[n * PITCH + m];
                        sH1[n * PITCH + m] = (h > 0.0f) ? accs[cb][i][j] : 0.0f;
                    }
                }
        }
        __syncthreads();
        // (i) dW1 [W][16], db1
        if (p.grad_dec) {
#pragma unroll 1
            for (int nb = 0; nb < W / 128; ++nb) wgrad16<M>(sH1 + nb * 128 * PITCH, sF, p.g_dec.W1 + (size_t)nb * 128 * 16, 16);
            bias_grad<M>(sH1, W, p.g_dec.b1);
        }
        // (j) g_f = W1^T g_h1 + g_f(part 2)
        {
            float acc16[RM] = {};
            gemm16<M>(sH1, W, p.dec.W1, 16, acc16);
#pragma unroll
            for (int i = 0; i < RM; ++i) sGF[tx * PITCH + ty * RM + i] += acc16[i];
        }
        __syncthreads();
        // (k) outputs of the tile
        if (p.g_feat) {
            for (int e = tid; e < M * 16; e += kFieldThreads) {
                const int sl = e / 16, j = e % 16;
                if (s0 + sl < nsamp) p.g_feat[(size_t)(s0 + sl) * 16 + j] = sGF[j * PITCH + sl];
            }
        }
        if (!p.feat && (p.grad_emb || p.grad_rays)) {
            // trilinear backward, aggregated over runs of samples that share (voxel, ray)
            const int sl = tid / C::CH, c = tid % C::CH;
            const int vox = sVox[sl], ray = sRay[sl];
            const bool head = vox >= 0 && (sl == 0 || sVox[sl - 1] != vox || sRay[sl - 1] != ray);
            // the CH lanes of one sample take the same path; shuffles name only those lanes
            const unsigned gmask = ((1u << C::CH) - 1u) << ((tid & 31) & ~(C::CH - 1));
            if (head) {
                int rows[8];
                float e[8][C::FPT];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    rows[i] = __ldg(p.vertex_idx + (size_t)vox * 8 + i);
                    if (p.grad_rays) {
#pragma unroll
                        for (int q = 0; q < C::FPT; ++q) e[i][q] = __ldg(p.emb + (size_t)rows[i] * 16 + c * C::FPT + q);
                    }
                }
                float ge[8][C::FPT] = {};
                float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
                for (int t = sl; t < M && sVox[t] == vox && sRay[t] == ray; ++t) {
                    const float px = sP[t], py = sP[M + t], pz = sP[2 * M + t];
                    float gf[C::FPT];
#pragma unroll
                    for (int q = 0; q < C::FPT; ++q) gf[q] = sGF[(c * C::FPT + q) * PITCH + t];
                    float gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
                        const float w = (wx * wy) * wz;
#pragma unroll
                        for (int q = 0; q < C::FPT; ++q) ge[i][q] = fmaf(w, gf[q], ge[i][q]);
                        if (p.grad_rays) {
                            float d = 0.0f;
#pragma unroll
                            for (int q = 0; q < C::FPT; ++q) d = fmaf(gf[q], e[i][q], d);
                            // full dot over the 16 features: sum the CH chunk lanes of this sample
#pragma unroll
                            for (int o = 1; o < C::CH; o <<= 1) d += __shfl_xor_sync(gmask, d, o, C::CH);
                            gp[0] += d * ((i & 4) ? 1.0f : -1.0f) * (wy * wz);
                            gp[1] += d * ((i & 2) ? 1.0f : -1.0f) * (wx * wz);
                            gp[2] += d * ((i & 1) ? 1.0f : -1.0f) * (wx * wy);
                        }
                    }
                    if (p.grad_rays) {
                        const float z = sZ[t];
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const float gx = gp[a] / p.voxel_size;
                            go[a] += gx; gd[a] = fmaf(z, gx, gd[a]);
                        }
                    }
                }
                if (p.grad_emb) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float *dst = p.g_emb + (size_t)rows[i] * 16 + c * C::FPT;
                        if (C::FPT == 4) red_add_v4(dst, ge[i][0], ge[i][1], ge[i][2 % C::FPT], ge[i][3 % C::FPT]);
                        else { atomicAdd(dst, ge[i][0]); atomicAdd(dst + 1, ge[i][1]); }
                    }
                }
                if (p.grad_rays && c == 0) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) { atomicAdd(p.g_rays_o + ray * 3 + a, go[a]); atomicAdd(p.g_rays_d + ray * 3 + a, gd[a]); }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// standalone trilinear kernels (API surface: get_features_vox and its backward)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_trilinear_fwd(int np, const float *__restrict__ xyz, const int *__restrict__ vox_idx, const float *__restrict__ centres,
                const int *__restrict__ vertex_idx, const float *__restrict__ emb, float vs, float *__restrict__ feat)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = t >> 2, c = t & 3;
    if (s >= np) return;
    const int vox = __ldg(vox_idx + s);
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 0), __ldg(centres + (size_t)vox * 3 + 0)), vs), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 1), __ldg(centres + (size_t)vox * 3 + 1)), vs), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 2), __ldg(centres + (size_t)vox * 3 + 2)), vs), 0.5f);
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = __ldg(vertex_idx + (size_t)vox * 8 + i);
        const float w = corner_w(i, px, py, pz);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)row * 16 + c * 4));
        f.x = fmaf(w, v.x, f.x); f.y = fmaf(w, v.y, f.y); f.z = fmaf(w, v.z, f.z); f.w = fmaf(w, v.w, f.w);
    }
    *reinterpret_cast<float4 *>(feat + (size_t)s * 16 + c * 4) = f;
}

__global__ void __launch_bounds__(256)
k_trilinear_bwd(int np, const float *__restrict__ xyz, const int *__restrict__ vox_idx, const float *__restrict__ centres,
                const int *__restrict__ vertex_idx, const float *__restrict__ emb, float vs,
                const float *__restrict__ g_feat, float *__restrict__ g_emb, float *__restrict__ g_xyz)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = min(t >> 2, np - 1), c = t & 3;
    const bool live = (t >> 2) < np;
    const int vox = __ldg(vox_idx + s);
    const float px = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 0), __ldg(centres + (size_t)vox * 3 + 0)), vs), 0.5f);
    const float py = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 1), __ldg(centres + (size_t)vox * 3 + 1)), vs), 0.5f);
    const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(__ldg(xyz + (size_t)s * 3 + 2), __ldg(centres + (size_t)vox * 3 + 2)), vs), 0.5f);
    const float4 g = __ldg(reinterpret_cast<const float4 *>(g_feat + (size_t)s * 16 + c * 4));
    float gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = __ldg(vertex_idx + (size_t)vox * 8 + i);
        const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
        const float w = (wx * wy) * wz;
        if (g_emb && live) red_add_v4(g_emb + (size_t)row * 16 + c * 4, w * g.x, w * g.y, w * g.z, w * g.w);
        if (g_xyz) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(emb + (size_t)row * 16 + c * 4));
            float d = g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
            d += __shfl_xor_sync(0xffffffffu, d, 1, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2, 4);
            gp[0] += d * ((i & 4) ? 1.0f : -1.0f) * (wy * wz);
            gp[1] += d * ((i & 2) ? 1.0f : -1.0f) * (wx * wz);
            gp[2] += d * ((i & 1) ? 1.0f : -1.0f) * (wx * wy);
        }
    }
    if (g_xyz && live && c < 3) g_xyz[(size_t)s * 3 + c] = gp[c] / vs;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int W, bool BWD>
static int launch_field_t(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    using C = FieldCfg<W>;
    static PerDevice once = {};
    bool &configured = once.done[current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_field<W, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { set_error("field: cudaFuncSetAttribute(%zu B): %s", C::SMEM, cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int tiles = ceil_div(max_samples, C::M);
    const int grid = tiles < num_sms() ? (tiles > 0 ? tiles : 1) : num_sms();
    k_field<W, BWD><<<grid, kFieldThreads, C::SMEM, st>>>(fp);
    PSLAM_CHECK_LAUNCH(BWD ? "field_backward" : "field_forward");
    return 0;
}

// simt_needed = false: a width-128 tensor-core build whose backward is known to have its wgrad workspace (or no decoder
// gradients), so nothing will run the SIMT kernels on these weights
static int pack_decoder(const pslam_decoder_t &d, float *ws, cudaStream_t st, bool simt_needed = true, int *range_flag = nullptr)
{
    if (d.width != 128 || decoder_mode() == 1) simt_needed = true;
    if (simt_needed) {
        if (d.width == 128) k_pack_decoder<128><<<ceil_div(FieldCfg<128>::WS, 256), 256, 0, st>>>(d, ws);
        else k_pack_decoder<256><<<ceil_div(FieldCfg<256>::WS, 256), 256, 0, st>>>(d, ws);
        PSLAM_CHECK_LAUNCH("pack_decoder");
    }
    if (d.width == 128 && decoder_mode() == 0) return tc_pack_decoder(d, ws + FieldCfg<128>::WS, st);
    if (d.width == 128 && decoder_mode() == 2) return bf_pack_decoder(d, ws + FieldCfg<128>::WS, st, range_flag);
    if (d.width == 256 && decoder_mode() == 2) return w2_pack_decoder(d, ws + FieldCfg<256>::WS, st, range_flag);
    return 0;
}
static_assert(FieldCfg<256>::WS == kSimt256Floats, "field.cuh: kSimt256Floats");
static const float *tc_region(const pslam_decoder_t &d, const float *ws) { return ws + (d.width == 128 ? FieldCfg<128>::WS : FieldCfg<256>::WS); }

static int launch_field(const FieldParams &fp, bool bwd, int max_samples, cudaStream_t st, int part = 0)
{
    if (fp.dec.width == 128 && decoder_mode() == 0) {
        if (!bwd) return tc_launch_field_forward(fp, max_samples, st);
        // tensor-core backward needs the wgrad scratch when decoder gradients are wanted
        if (!fp.grad_dec || (fp.wg_scratch && fp.wg_scratch_bytes >= tc_wgrad_scratch_bytes(max_samples)))
            return tc_launch_field_backward(fp, max_samples, st, part);
    }
    if (fp.dec.width == 128 && decoder_mode() == 2) {
        if (!bwd) return bf_launch_field_forward(fp, max_samples, st, part);
        if (!fp.grad_dec || (fp.wg_scratch && fp.wg_scratch_bytes >= bf_wgrad_scratch_bytes(max_samples)))
            return bf_launch_field_backward(fp, max_samples, st, part);
    }
    if (fp.dec.width == 128) return bwd ? launch_field_t<128, true>(fp, max_samples, st) : launch_field_t<128, false>(fp, max_samples, st);
    if (decoder_mode() == 2 && w2_usable(fp, max_samples, bwd))
        return bwd ? w2_launch_field_backward(fp, max_samples, st, part) : w2_launch_field_forward(fp, max_samples, st, part);
    return bwd ? launch_field_t<256, true>(fp, max_samples, st) : launch_field_t<256, false>(fp, max_samples, st);
}

static FieldParams params_from_render(const pslam_render_t *p)
{
    FieldParams fp{};
    fp.nsamp = 0; fp.nsamp_dev = p->counters + PSLAM_C_NSAMP;
    fp.rays_o = p->rays_o; fp.rays_d = p->rays_d; fp.hit_ray = p->hit_ray;
    fp.samp_ray = p->samp_ray; fp.samp_vox = p->samp_vox; fp.samp_z = p->samp_z;
    fp.centres = p->centres; fp.vertex_idx = p->vertex_idx; fp.emb = p->emb; fp.voxel_size = p->voxel_size;
    fp.feat = nullptr; fp.dec = p->dec; fp.ws = p->dec_ws; fp.ws_tc = tc_region(p->dec, p->dec_ws); fp.out = p->samp_out;
    fp.g_out = p->samp_gout; fp.g_feat = nullptr; fp.g_dec = p->g_dec; fp.g_emb = p->g_emb;
    fp.g_rays_o = p->g_rays_o; fp.g_rays_d = p->g_rays_d;
    fp.wg_scratch = static_cast<unsigned char *>(p->wgrad_ws); fp.wg_scratch_bytes = (size_t)p->wgrad_ws_bytes;
    fp.grad_dec = (p->flags & PSLAM_F_GRAD_DEC) ? 1 : 0;
    fp.grad_emb = (p->flags & PSLAM_F_GRAD_EMB) ? 1 : 0;
    fp.grad_rays = (p->flags & PSLAM_F_GRAD_RAYS) ? 1 : 0;
    fp.paired = 1;
    fp.gmax_ready = reinterpret_cast<uint32_t *>(p->counters + PSLAM_C_TILE);
    fp.range_flag = p->counters + PSLAM_C_OVERFLOW;
    return fp;
}

// The weight re-pack of a step depends on nothing the step computes: pslam_render_step forks it onto the side stream behind the
// intersection stage (whose first kernel clears the counters the pack's range flag lives in), underneath the sampling kernel.
// *packed = the event the forward must wait for, or NULL when the pack was not forked (other decoder builds, no side stream).
int fork_decoder_pack(const pslam_render_t *p, cudaStream_t st, cudaEvent_t *packed)
{
    *packed = nullptr;
    if (p->dec.width != 128 || decoder_mode() != 2) return 0;
    const bool simt = (p->flags & PSLAM_F_GRAD_DEC) && !(p->wgrad_ws && (size_t)p->wgrad_ws_bytes >= (size_t)pslam_wgrad_ws_bytes(p->sample_cap));
    SideStream *side = simt ? nullptr : side_stream();
    if (!side) return 0;
    if (cudaEventRecord(side->fork, st) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork, 0) != cudaSuccess) {
        set_error("decoder pack: stream fork: %s", cudaGetErrorString(cudaGetLastError()));
        return PSLAM_E_ARG;
    }
    if (int rc = pack_decoder(p->dec, p->dec_ws, side->stream, false, p->counters + PSLAM_C_OVERFLOW)) return rc;
    if (cudaEventRecord(side->join, side->stream) != cudaSuccess) { set_error("decoder pack: stream join: %s", cudaGetErrorString(cudaGetLastError())); return PSLAM_E_ARG; }
    *packed = side->join;
    return 0;
}

int launch_field_forward(const pslam_render_t *p, cudaStream_t st, int part)
{
    if (part == 4) part = 0;   // part 4: the weights were packed by fork_decoder_pack
    else if (part != 3 && part != 5)   // part 3 / 5 (profiling): the decoder / gather kernel alone, after a full forward of the same arguments
    {
        // the SIMT weights are only read by a SIMT backward: decoder gradients wanted without a large enough workspace
        const bool simt = (p->flags & PSLAM_F_GRAD_DEC) && !(p->wgrad_ws && (size_t)p->wgrad_ws_bytes >= (size_t)pslam_wgrad_ws_bytes(p->sample_cap));
        if (int rc = pack_decoder(p->dec, p->dec_ws, st, simt, p->counters + PSLAM_C_OVERFLOW)) return rc;
    }
    return launch_field(params_from_render(p), false, p->sample_cap, st, part);
}

int launch_field_backward(const pslam_render_t *p, cudaStream_t st, int part)
{
    // dec_ws was packed by the forward of the same step
    return launch_field(params_from_render(p), true, p->sample_cap, st, part);
}

}  // namespace pslam

using namespace pslam;

static int check_decoder(const pslam_decoder_t *dec)
{
    PSLAM_CHECK_ARG(dec, PSLAM_E_ARG, "null decoder");
    PSLAM_CHECK_ARG(dec->width == 128 || dec->width == 256, PSLAM_E_RANGE, "decoder width %d not supported (128 or 256)", dec->width);
    PSLAM_CHECK_ARG(dec->W1 && dec->b1 && dec->W2 && dec->b2 && dec->W3 && dec->b3 && dec->W4 && dec->b4 && dec->W5 && dec->b5,
                    PSLAM_E_ARG, "null decoder parameter");
    PSLAM_CHECK_ARG(((uintptr_t)dec->W1 | (uintptr_t)dec->W2 | (uintptr_t)dec->W3 | (uintptr_t)dec->W4) % 16 == 0, PSLAM_E_ALIGN,
                    "decoder weights must be 16-byte aligned");
    return 0;
}

extern "C" int64_t pslam_decoder_ws_count(int width)
{
    return width == 128 ? FieldCfg<128>::WS + kTcPackFloats : width == 256 ? FieldCfg<256>::WS + kW2PackFloats : -1;
}

extern "C" int pslam_trilinear_fwd(int np, const float *xyz, const int *vox_idx, const float *centres, const int *vertex_idx,
                                   const float *emb, float voxel_size, float *feat, pslam_stream_t stream)
{
    if (np == 0) return 0;
    PSLAM_CHECK_ARG(np > 0 && xyz && vox_idx && centres && vertex_idx && emb && feat, PSLAM_E_ARG, "trilinear_fwd: bad argument");
    PSLAM_CHECK_ARG(((uintptr_t)emb | (uintptr_t)feat) % 16 == 0, PSLAM_E_ALIGN, "emb/feat must be 16-byte aligned");
    k_trilinear_fwd<<<(int)ceil_div64((int64_t)np * 4, 256), 256, 0, (cudaStream_t)stream>>>(np, xyz, vox_idx, centres, vertex_idx, emb,
                                                                                       voxel_size, feat);
    PSLAM_CHECK_LAUNCH("trilinear_fwd");
    return 0;
}

extern "C" int pslam_trilinear_bwd(int np, const float *xyz, const int *vox_idx, const float *centres, const int *vertex_idx,
                                   const float *emb, float voxel_size, const float *g_feat, float *g_emb, float *g_xyz,
                                   pslam_stream_t stream)
{
    if (np == 0) return 0;
    PSLAM_CHECK_ARG(np > 0 && xyz && vox_idx && centres && vertex_idx && emb && g_feat, PSLAM_E_ARG, "trilinear_bwd: bad argument");
    PSLAM_CHECK_ARG(((uintptr_t)emb | (uintptr_t)g_feat | (uintptr_t)g_emb) % 16 == 0, PSLAM_E_ALIGN, "emb/g_feat/g_emb must be 16-byte aligned");
    k_trilinear_bwd<<<(int)ceil_div64((int64_t)np * 4, 256), 256, 0, (cudaStream_t)stream>>>(np, xyz, vox_idx, centres, vertex_idx, emb,
                                                                                       voxel_size, g_feat, g_emb, g_xyz);
    PSLAM_CHECK_LAUNCH("trilinear_bwd");
    return 0;
}

extern "C" int pslam_decoder_fwd(int np, const pslam_decoder_t *dec, const float *feat, float *ws, float *out, pslam_stream_t stream)
{
    if (int rc = check_decoder(dec)) return rc;
    if (np == 0) return 0;
    PSLAM_CHECK_ARG(np > 0 && feat && ws && out, PSLAM_E_ARG, "decoder_fwd: bad argument");
    PSLAM_CHECK_ARG(((uintptr_t)ws | (uintptr_t)out) % 16 == 0, PSLAM_E_ALIGN, "ws/out must be 16-byte aligned");
    if (int rc = pack_decoder(*dec, ws, (cudaStream_t)stream)) return rc;
    FieldParams fp{};
    fp.nsamp = np; fp.feat = feat; fp.dec = *dec; fp.ws = ws; fp.ws_tc = tc_region(*dec, ws); fp.out = out;
    return launch_field(fp, false, np, (cudaStream_t)stream);
}

extern "C" int64_t pslam_wgrad_ws_bytes_w(int max_samples, int width)
{
    if (width == 256) return (int64_t)w2_scratch_bytes(max_samples);
    return pslam_wgrad_ws_bytes(max_samples);
}

extern "C" int64_t pslam_wgrad_ws_bytes(int max_samples)
{
    const size_t a = tc_wgrad_scratch_bytes(max_samples), b = bf_wgrad_scratch_bytes(max_samples);   // either build may be selected later
    return (int64_t)(a > b ? a : b);
}

extern "C" int pslam_decoder_bwd(int np, const pslam_decoder_t *dec, const float *feat, float *ws, const float *g_out, float *g_feat,
                                 const pslam_decoder_grad_t *grad, void *wgrad_ws, int64_t wgrad_ws_bytes, pslam_stream_t stream)
{
    if (int rc = check_decoder(dec)) return rc;
    if (np == 0) return 0;
    PSLAM_CHECK_ARG(np > 0 && feat && ws && g_out, PSLAM_E_ARG, "decoder_bwd: bad argument");
    PSLAM_CHECK_ARG(((uintptr_t)ws | (uintptr_t)g_out) % 16 == 0, PSLAM_E_ALIGN, "ws/g_out must be 16-byte aligned");
    if (grad)
        PSLAM_CHECK_ARG(grad->W1 && grad->b1 && grad->W2 && grad->b2 && grad->W3 && grad->b3 && grad->W4 && grad->b4 && grad->W5 && grad->b5,
                        PSLAM_E_ARG, "decoder_bwd: null gradient pointer");
    if (grad)
        PSLAM_CHECK_ARG(((uintptr_t)grad->W1 | (uintptr_t)grad->W2 | (uintptr_t)grad->W3 | (uintptr_t)grad->W4) % 16 == 0, PSLAM_E_ALIGN,
                        "decoder weight gradients must be 16-byte aligned");
    if (int rc = pack_decoder(*dec, ws, (cudaStream_t)stream)) return rc;
    FieldParams fp{};
    fp.nsamp = np; fp.feat = feat; fp.dec = *dec; fp.ws = ws; fp.ws_tc = tc_region(*dec, ws); fp.g_out = g_out; fp.g_feat = g_feat;
    if (grad) { fp.g_dec = *grad; fp.grad_dec = 1; }
    PSLAM_CHECK_ARG((uintptr_t)wgrad_ws % 16 == 0, PSLAM_E_ALIGN, "wgrad_ws must be 16-byte aligned");
    fp.wg_scratch = static_cast<unsigned char *>(wgrad_ws); fp.wg_scratch_bytes = wgrad_ws ? (size_t)wgrad_ws_bytes : 0;
    return launch_field(fp, true, np, (cudaStream_t)stream);
}
