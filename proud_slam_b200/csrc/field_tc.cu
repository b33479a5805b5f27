// Kernel 4 on the 5th-generation tensor cores: the width-128 SDF/colour decoder
// (src/variations/nrgbd.py:116-135) fused with the trilinear corner-embedding lookup
// (src/variations/render_helpers.py:105-156, 47-59), forward pass.
//
// Design (sm_100a only):
//   * tile = 128 samples = the 128 lanes of tensor memory; one persistent CTA per SM.
//   * every layer is D[128 x N] = A[128 x K] * W^T with A (activations) read from TENSOR MEMORY
//     (tcgen05.mma, A-from-TMEM form) and W streamed from L2 into shared memory by 1-D bulk TMA
//     copies in the exact byte order the MMA wants (weights are re-packed once per iteration by
//     k_tc_pack), so no activation ever touches shared or global memory:
//         accumulators --tcgen05.ld--> registers (bias, ReLU, hi/lo split) --tcgen05.st--> next A.
//   * fp32-equivalent accuracy on TF32 tensor cores by operand splitting (3xTF32):
//         a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   hi = rn_tf32(x), lo = rn_tf32(x - hi)
//     accumulated in fp32 in tensor memory (the reference runs cuBLAS SGEMM with TF32 off;
//     the 1e-4 parity bound of the tests holds with ~100x margin).
//   * warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warps 2-5 =
//     128 worker threads, one per sample row: gather + interpolate the 8 corner embeddings,
//     run the layer epilogues, write (r,g,b,sdf).
//   TMEM columns: A_hi [0,144)  A_lo [144,288)  D [288,432).  Layer inputs: L1 reads the 16
//   features kept at A columns [128,144); L4 reads [t(128); f(16)] = columns [0,144) in place.
#include "field.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace pslam {

using namespace umma;

namespace tc {
constexpr int kThreads = 192;
constexpr int kStages = 8;
constexpr int kStageBytes = 18432;   // 144 rows x 16 k x 4 B x (hi, lo)
constexpr int kTmemCols = 512;
constexpr int cAHI = 0, cALO = 144, cD = 288;
constexpr int kLayers = 5;
// forward layers: output columns N, reduction K, A column offset
__device__ __constant__ int cN[kLayers] = {128, 128, 144, 128, 16};
__device__ __constant__ int cK[kLayers] = {16, 128, 128, 144, 128};
__device__ __constant__ int cAoff[kLayers] = {128, 0, 0, 0, 0};
constexpr int hN[kLayers] = {128, 128, 144, 128, 16};
constexpr int hK[kLayers] = {16, 128, 128, 144, 128};
// shared memory map
constexpr int oBars = kStages * kStageBytes;             // full[8], empty[8], a_ready, mma_done
constexpr int oTmemPtr = oBars + 8 * (2 * kStages + 2);
constexpr int oBias = oTmemPtr + 16;                     // b1[128] b2[128] b3f[128] b4[128] b3_0 b5[3]
constexpr int kSmemBytes = oBias + 4 * (4 * 128 + 4);
}  // namespace tc

// source element of forward layer l at (output row n, reduction index k)
__device__ __forceinline__ float tc_weight(const pslam_decoder_t &d, int l, int n, int k)
{
    switch (l) {
        case 0: return d.W1[n * 16 + k];
        case 1: return d.W2[n * 128 + k];
        case 2: return n < 128 ? d.W3[(1 + n) * 128 + k] : (n == 128 ? d.W3[k] : 0.0f);   // features first, sdf row at 128
        case 3: return d.W4[n * 144 + k];                                                  // k over [t(128); f(16)]
        default: return n < 3 ? d.W5[n * 128 + k] : 0.0f;
    }
}

// Re-packs the decoder into the forward weight stream: layers in order, each as K/16 chunks of
// [hi block | lo block], each block = 4 k-chunks x N rows x 16 B (see umma.cuh).
__global__ void k_tc_pack(pslam_decoder_t d, float *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int base = 0;   // float offset of the layer in `out`
#pragma unroll
    for (int l = 0; l < tc::kLayers; ++l) {
        const int N = tc::cN[l], K = tc::cK[l];
        if (i < N * K) {
            const int n = i / K, k = i % K;
            const int c = k >> 4, kc = (k >> 2) & 3, e = k & 3;
            uint32_t hi, lo;
            tf32_split(tc_weight(d, l, n, k), hi, lo);
            float *chunk = out + base + c * (N * 32);            // chunk = 2 * N * 16 floats
            const int off = kc * (N * 4) + n * 4 + e;
            chunk[off] = __uint_as_float(hi);
            chunk[N * 16 + off] = __uint_as_float(lo);
            return;
        }
        i -= N * K;
        base += 2 * N * K;
    }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(tc::kThreads, 1) k_field_tc_fwd(FieldParams p, const float *__restrict__ wstream)
{
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *empty = full + kStages;
    uint64_t *a_ready = empty + kStages;
    uint64_t *mma_done = a_ready + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oTmemPtr);
    float *sBias = reinterpret_cast<float *>(smem + oBias);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(a_ready, 128);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    for (int i = threadIdx.x; i < 4 * 128 + 4; i += kThreads) {
        float v;
        if (i < 128) v = p.dec.b1[i];
        else if (i < 256) v = p.dec.b2[i - 128];
        else if (i < 384) v = p.dec.b3[1 + i - 256];
        else if (i < 512) v = p.dec.b4[i - 384];
        else if (i == 512) v = p.dec.b3[0];
        else v = p.dec.b5[i - 513];
        sBias[i] = v;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(wstream);
                for (int l = 0; l < kLayers; ++l) {
                    const uint32_t bytes = (uint32_t)cN[l] * 128u;
                    for (int c = 0; c < cK[l] / 16; ++c) {
                        mbar_wait(empty + stage, phase ^ 1);
                        mbar_arrive_expect_tx(full + stage, bytes);
                        bulk_g2s(smem + stage * kStageBytes, src, bytes, full + stage);
                        src += bytes;
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, phase = 0;
            uint32_t uses = 0;   // a_ready phase counter
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < kLayers; ++l) {
                    const int N = cN[l];
                    const uint32_t idesc = idesc_tf32(128, N);
                    mbar_wait(a_ready, uses & 1);
                    ++uses;
                    fence_after_sync();
                    const uint32_t a_hi = tmem + cAHI + cAoff[l], a_lo = tmem + cALO + cAoff[l], d = tmem + cD;
                    for (int c = 0; c < cK[l] / 16; ++c) {
                        mbar_wait(full + stage, phase);
                        fence_after_sync();
                        const uint32_t sb = smem_u32(smem + stage * kStageBytes);
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint32_t k = c * 16 + s * 8;
                            const uint64_t b_hi = bdesc_kmajor(sb + s * (2 * N * 16), N);
                            const uint64_t b_lo = bdesc_kmajor(sb + N * 64 + s * (2 * N * 16), N);
                            mma_tf32_ts(d, a_lo + k, b_hi, idesc, (c | s) ? 1u : 0u);
                            mma_tf32_ts(d, a_hi + k, b_lo, idesc, 1u);
                            mma_tf32_ts(d, a_hi + k, b_hi, idesc, 1u);
                        }
                        mma_commit(empty + stage);   // stage is free once these MMAs have read it
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    mma_commit(mma_done);
                }
            }
        }
    } else {
        // ===================== workers: one thread per sample row =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                  // row of the tile
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t done_uses = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int s = tile * 128 + m;
            // ---- features -> A[:, 128:144) ----
            {
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = 0.0f;
                if (s < nsamp) {
                    if (p.feat) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.feat + (size_t)s * 16 + e));
                            f[e] = v.x; f[e + 1] = v.y; f[e + 2] = v.z; f[e + 3] = v.w;
                        }
                    } else {
                        const int vox = __ldg(p.samp_vox + s);
                        const float z = __ldg(p.samp_z + s);
                        const int ray = __ldg(p.hit_ray + __ldg(p.samp_ray + s));
                        const float x = __fadd_rn(__ldg(p.rays_o + ray * 3 + 0), __fmul_rn(__ldg(p.rays_d + ray * 3 + 0), z));
                        const float y = __fadd_rn(__ldg(p.rays_o + ray * 3 + 1), __fmul_rn(__ldg(p.rays_d + ray * 3 + 1), z));
                        const float zz = __fadd_rn(__ldg(p.rays_o + ray * 3 + 2), __fmul_rn(__ldg(p.rays_d + ray * 3 + 2), z));
                        const float px = __fadd_rn(__fdiv_rn(__fsub_rn(x, __ldg(p.centres + (size_t)vox * 3 + 0)), p.voxel_size), 0.5f);
                        const float py = __fadd_rn(__fdiv_rn(__fsub_rn(y, __ldg(p.centres + (size_t)vox * 3 + 1)), p.voxel_size), 0.5f);
                        const float pz = __fadd_rn(__fdiv_rn(__fsub_rn(zz, __ldg(p.centres + (size_t)vox * 3 + 2)), p.voxel_size), 0.5f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = __ldg(p.vertex_idx + (size_t)vox * 8 + i);
                            const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
                            const float w = (wx * wy) * wz;
                            const float4 *er = reinterpret_cast<const float4 *>(p.emb + (size_t)row * 16);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 v = __ldg(er + e);
                                f[4 * e] = fmaf(w, v.x, f[4 * e]); f[4 * e + 1] = fmaf(w, v.y, f[4 * e + 1]);
                                f[4 * e + 2] = fmaf(w, v.z, f[4 * e + 2]); f[4 * e + 3] = fmaf(w, v.w, f[4 * e + 3]);
                            }
                        }
                    }
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) tf32_split(f[e], hi[e], lo[e]);
                tmem_st16(trow + cAHI + 128, hi);
                tmem_st16(trow + cALO + 128, lo);
                tmem_wait_st();
                fence_before_sync();
                mbar_arrive(a_ready);
            }
            float sdf = 0.0f;
            for (int l = 0; l < kLayers; ++l) {
                mbar_wait(mma_done, done_uses & 1);
                ++done_uses;
                fence_after_sync();
                if (l < 4) {
                    const float *bias = sBias + l * 128;
                    for (int c0 = 0; c0 < 128; c0 += 16) {
                        uint32_t v[16], hi[16], lo[16];
                        tmem_ld16(trow + cD + c0, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float y = __uint_as_float(v[e]) + bias[c0 + e];
                            if (l != 2) y = fmaxf(y, 0.0f);      // L3 (features t) has no activation
                            tf32_split(y, hi[e], lo[e]);
                        }
                        tmem_st16(trow + cAHI + c0, hi);
                        tmem_st16(trow + cALO + c0, lo);
                    }
                    if (l == 2) {   // sdf = row 0 of W3, packed as output column 128
                        uint32_t v[16];
                        tmem_ld16(trow + cD + 128, v);
                        tmem_wait_ld();
                        sdf = __uint_as_float(v[0]) + sBias[512];
                    }
                    tmem_wait_st();
                    fence_before_sync();
                    mbar_arrive(a_ready);
                } else {
                    uint32_t v[16];
                    tmem_ld16(trow + cD, v);
                    tmem_wait_ld();
                    const float r = sigmoid_f(__uint_as_float(v[0]) + sBias[513]);
                    const float g = sigmoid_f(__uint_as_float(v[1]) + sBias[514]);
                    const float b = sigmoid_f(__uint_as_float(v[2]) + sBias[515]);
                    if (s < nsamp) *reinterpret_cast<float4 *>(p.out + (size_t)s * 4) = make_float4(r, g, b, sdf);
                    // D has been read: order it before the next tile's first MMA through a_ready
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tc::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// stand-alone GEMM through the same primitives (unit test of descriptors / TMEM addressing):
// D[128,N] = A[128,K] * B[N,K]^T, K % 8 == 0, N % 16 == 0, both <= 144.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_debug_umma_gemm(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                                                            int N, int K, int split3)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    float *sB = reinterpret_cast<float *>(smem);           // per 8-k step: [hi: 2 chunks x N rows x 4][lo: same]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    // B -> smem in the operand layout, one block of (hi, lo) per MMA k-step
    for (int i = threadIdx.x; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const int st = k >> 3, kc = (k >> 2) & 1, e = k & 3;
        uint32_t hi, lo;
        tf32_split(B[i], hi, lo);
        float *blk = sB + st * (2 * N * 8);
        blk[kc * N * 4 + n * 4 + e] = __uint_as_float(hi);
        blk[N * 8 + kc * N * 4 + n * 4 + e] = __uint_as_float(lo);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    // A row m -> TMEM columns [0,K) hi, [144,144+K) lo
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const float a = (k0 + e < K) ? A[(size_t)m * K + k0 + e] : 0.0f;
            tf32_split(a, hi[e], lo[e]);
        }
        tmem_st16(trow + k0, hi);
        tmem_st16(trow + 144 + k0, lo);
    }
    tmem_wait_st();
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t sb = smem_u32(sB);
        for (int st = 0; st < K / 8; ++st) {
            const uint64_t b_hi = bdesc_kmajor(sb + st * (2 * N * 8 * 4), N);
            const uint64_t b_lo = bdesc_kmajor(sb + st * (2 * N * 8 * 4) + N * 8 * 4, N);
            if (split3) {
                mma_tf32_ts(tmem + 288, tmem + 144 + st * 8, b_hi, idesc, st ? 1u : 0u);
                mma_tf32_ts(tmem + 288, tmem + st * 8, b_lo, idesc, 1u);
                mma_tf32_ts(tmem + 288, tmem + st * 8, b_hi, idesc, 1u);
            } else {
                mma_tf32_ts(tmem + 288, tmem + st * 8, b_hi, idesc, st ? 1u : 0u);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + 288 + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) D[(size_t)m * N + c0 + e] = __uint_as_float(v[e]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
int tc_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st)
{
    int total = 0;
    for (int l = 0; l < tc::kLayers; ++l) total += tc::hN[l] * tc::hK[l];
    k_tc_pack<<<ceil_div(total, 256), 256, 0, st>>>(d, ws_tc);
    PSLAM_CHECK_LAUNCH("tc_pack");
    return 0;
}

int tc_launch_field_forward(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_field_tc_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
        if (e != cudaSuccess) { set_error("field_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int tiles = ceil_div(max_samples, 128);
    const int grid = tiles < num_sms() ? (tiles > 0 ? tiles : 1) : num_sms();
    k_field_tc_fwd<<<grid, tc::kThreads, tc::kSmemBytes, st>>>(fp, fp.ws_tc);
    PSLAM_CHECK_LAUNCH("field_tc_forward");
    return 0;
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_debug_umma_gemm(const float *A, const float *B, float *D, int N, int K, int split3, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(A && B && D, PSLAM_E_ARG, "null pointer");
    PSLAM_CHECK_ARG(N >= 16 && N <= 144 && N % 16 == 0 && K >= 8 && K <= 144 && K % 8 == 0, PSLAM_E_RANGE, "N in 16..144 step 16, K in 8..144 step 8");
    const int smem = 2 * N * K * 4;
    cudaError_t e = cudaFuncSetAttribute(k_debug_umma_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("debug_umma: %s", cudaGetErrorString(e)); return (int)e; }
    k_debug_umma_gemm<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, split3);
    PSLAM_CHECK_LAUNCH("debug_umma_gemm");
    return 0;
}
