// Kernel 4 on the 5th-generation tensor cores: the width-128 SDF/colour decoder
// (src/variations/nrgbd.py:116-135) fused with the trilinear corner-embedding lookup
// (src/variations/render_helpers.py:105-156, 47-59), forward and backward.
//
// Design (sm_100a only):
//   * tile = 128 samples = the 128 lanes of tensor memory; one persistent CTA per SM.
//   * every layer is D[128 x N] = A[128 x K] * W^T with A (activations, or their gradients in the
//     backward chain) read from TENSOR MEMORY (tcgen05.mma, A-from-TMEM form) and W streamed from
//     L2 into shared memory by 1-D bulk TMA copies in the exact byte order the MMA wants (weights
//     are re-packed once per iteration by k_tc_pack), so no activation touches shared memory:
//         accumulators --tcgen05.ld--> registers (bias, ReLU, hi/lo split) --tcgen05.st--> next A.
//   * fp32-equivalent accuracy on TF32 tensor cores by operand splitting (3xTF32):
//         a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,   hi = rn_tf32(x), lo = rn_tf32(x - hi)
//     accumulated in fp32 in tensor memory (the reference runs cuBLAS SGEMM with TF32 off).
//   * warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warps 2-5 =
//     128 worker threads, one per sample row: gather + interpolate the 8 corner embeddings, run the
//     layer epilogues, write (r,g,b,sdf) / scatter the embedding and ray gradients.
//   * backward = forward recompute (ReLU masks kept as bits in registers) + the dgrad chain
//     g5 -> g_hc -> (g_t, g_f) -> g_h2 -> g_h1 -> g_f, same machinery with transposed weight packs.
//     Weight gradients are a different GEMM shape (reduction over SAMPLES): the workers spill each
//     tile's activations and pre-activation gradients to an L2-resident scratch in MMA operand
//     order and k_wgrad_tc contracts them with accumulators resident in tensor memory across all
//     of a CTA's tiles (one flush of red.global.add.v4 per CTA instead of 54k atomics per tile).
//   TMEM columns (k_field_tc): A_hi [0,144)  A_lo [144,288)  D [288,432).
#include "field.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "decoder_layers.cuh"

namespace pslam {

using namespace umma;

namespace tc {
constexpr int kThreads = 352;     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 workers, warp 10 scratch store
constexpr int kWorkers = 256;
constexpr int kStages = 8;        // weight ring depth of the forward kernel
constexpr int kStagesBwd = 4;     // ... of the backward kernel (the rest of shared memory stages the wgrad scratch)
constexpr int kStageBytes = 18432;   // 144 rows x 16 k x 4 B x (hi, lo)
constexpr int kStagingBytes = 65536; // one layer of one tile in scratch order: [4 slices][32 groups][32 samples][16 B]
constexpr int kCluster = 2;          // CTAs that share one multicast weight stream (each loads its share of every chunk)
constexpr uint16_t kClusterMask = (1u << kCluster) - 1u;
constexpr int kTmemCols = 512;
constexpr int cAHI = 0, cALO = 144, cD = 288;
constexpr int kLayersFwd = 5, kLayersAll = 10;
// layers: output columns N, reduction K, A column offset.  0-4 forward (L1..L5), 5-9 dgrad (D5..D1)
__device__ __constant__ int cN[kLayersAll] = {128, 128, 144, 128, 16, 128, 144, 128, 128, 16};
__device__ __constant__ int cK[kLayersAll] = {16, 128, 128, 144, 128, 16, 128, 144, 128, 128};
__device__ __constant__ int cAoff[kLayersAll] = {128, 0, 0, 0, 0, 0, 0, 0, 0, 0};
constexpr int hN[kLayersAll] = {128, 128, 144, 128, 16, 128, 144, 128, 128, 16};
constexpr int hK[kLayersAll] = {16, 128, 128, 144, 128, 16, 128, 144, 128, 128};
// shared memory map: [weight ring][scratch staging x2 (backward only)][barriers][tmem ptr][biases]
template <bool BWD>
struct Smem {
    static constexpr int nStages = BWD ? kStagesBwd : kStages;
    static constexpr int oStaging = nStages * kStageBytes;
    static constexpr int oBars = oStaging + (BWD ? 2 * kStagingBytes : 0);   // full[8], empty[8], a_ready, mma_done, st_full[2], st_free[2]
    static constexpr int oTmemPtr = oBars + 8 * (2 * kStages + 2 + 4);
    static constexpr int oBias = oTmemPtr + 16;                     // b1[128] b2[128] b3f[128] b4[128] b3_0 b5[3]
    static constexpr int bytes = oBias + 4 * (4 * 128 + 4);
};

// wgrad scratch: per tile, per 32-sample slice, 264 groups of (32 lanes x 4 floats); see k_wgrad_tc
constexpr int gF = 0, gH1 = 4, gH2 = 36, gT = 68, gHC = 100, gG1 = 132, gG2 = 164, gG3 = 196, gG4 = 228, gG5 = 260, kGroups = 264;
constexpr size_t kSliceBytes = (size_t)kGroups * 512;
constexpr size_t kTileBytes = 4 * kSliceBytes;
}  // namespace tc

// Re-packs the decoder into the weight stream: layers in order, each as K/16 chunks of
// [hi block | lo block], each block = 4 k-chunks x N rows x 16 B (see umma.cuh).
__global__ void k_tc_pack(pslam_decoder_t d, float *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int base = 0;   // float offset of the layer in `out`
#pragma unroll
    for (int l = 0; l < tc::kLayersAll; ++l) {
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

// One epilogue over this thread's 64 accumulator columns [col0, col0+64): accumulators -> registers ->
// (bias / activation / mask) -> next A operand (hi/lo split, tensor memory) and optionally the wgrad
// scratch.  Two worker warps share each TMEM lane quarter and split the columns in halves.
//   MODE 0: y = relu(D + bias), records y > 0 in mask[]        (forward hidden layers)
//   MODE 1: y = D + bias                                       (forward, no activation)
//   MODE 2: y = mask ? D : 0                                   (dgrad through a ReLU)
//   MODE 3: y = D                                              (dgrad, no activation)
template <int MODE>
__device__ __forceinline__ void epilogue64(uint32_t trow, int col0, const float *bias, uint32_t (&mask)[2], unsigned char *scratch)
{
    using namespace tc;
#pragma unroll
    for (int b = 0; b < 2; ++b) {   // fully unrolled: mask[] must stay in registers
        const int c0 = col0 + 32 * b;
        uint32_t v[32];
        tmem_ld32(trow + cD + c0, v);
        tmem_wait_ld();
        uint32_t bits = 0u;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            float y = __uint_as_float(v[e]);
            if (MODE == 0) { y = fmaxf(y + bias[c0 + e], 0.0f); bits |= (y > 0.0f ? 1u : 0u) << e; }
            if (MODE == 1) y = y + bias[c0 + e];
            if (MODE == 2) y = ((mask[b] >> e) & 1u) ? y : 0.0f;
            v[e] = __float_as_uint(y);
        }
        if (MODE == 0) mask[b] = bits;
        if (scratch) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(scratch + (size_t)(c0 / 4 + j) * 512) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        uint32_t lo[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) { uint32_t h; tf32_split(__uint_as_float(v[e]), h, lo[e]); v[e] = h; }
        tmem_st32(trow + cAHI + c0, v);
        tmem_st32(trow + cALO + c0, lo);
    }
}

// optional timeline trace of CTA 0 (pslam_debug_tc_trace): [tile<4][layer<10][8] clock64 stamps (+ 40 x 8 for k_wgrad_tc)
__device__ long long *g_tc_trace = nullptr;
#define TC_TRACE(tile_i, layer, slot)                                                                   \
    do {                                                                                                \
        if (g_tc_trace && blockIdx.x == 0 && (tile_i) < 4) g_tc_trace[((tile_i) * 10 + (layer)) * 8 + (slot)] = clock64(); \
    } while (0)

template <bool BWD>
__global__ void __cluster_dims__(tc::kCluster, 1, 1) __launch_bounds__(tc::kThreads, 1)
k_field_tc(FieldParams p, const float *__restrict__ wstream)
{
    using namespace tc;
    constexpr int NL = BWD ? kLayersAll : kLayersFwd;
    using SM = Smem<BWD>;
    constexpr int NS = SM::nStages;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::oBars);
    uint64_t *empty = full + kStages;
    uint64_t *a_ready = empty + kStages;
    uint64_t *mma_done = a_ready + 1;
    uint64_t *st_full = mma_done + 1, *st_free = st_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + SM::oTmemPtr);
    float *sBias = reinterpret_cast<float *>(smem + SM::oBias);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    // The CTAs of a cluster consume one shared weight stream in lockstep, so they all run the same number of
    // tile iterations; iterations whose tile index is past the end are dummies (no valid rows, nothing stored).
    const int iters = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t crank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kCluster); }
        mbar_init(a_ready, kWorkers);
        mbar_init(mma_done, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(st_full + i, kWorkers); mbar_init(st_free + i, 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    const bool spill = BWD && p.wg_scratch != nullptr;   // activations / gradients go to the wgrad scratch
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
    cluster_sync();   // every CTA's barriers are initialised before any peer multicasts into them
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
        // each CTA fetches its share of every chunk and multicasts it to the whole cluster: the L2 -> SM weight
        // traffic per SM drops by the cluster size (it was the busiest stream on the SM's L2 port)
        int stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(wstream);
            for (int l = 0; l < NL; ++l) {
                const uint32_t bytes = (uint32_t)cN[l] * 128u, part = bytes / kCluster;
                for (int c = 0; c < cK[l] / 16; ++c) {
                    mbar_wait(empty + stage, phase ^ 1);     // all kCluster CTAs are done reading this stage
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + stage, bytes);
                        bulk_g2s_mcast(smem + stage * kStageBytes + crank * part, src + crank * part, part, full + stage, kClusterMask);
                    }
                    __syncwarp();
                    src += bytes;
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops so that operands stay warp-uniform;
        // one elected lane issues tcgen05.mma / tcgen05.commit) =====================
        int stage = 0, phase = 0;
        uint32_t uses = 0;   // a_ready phase counter
        int tile_i = 0;
        for (int it = 0; it < iters; ++it, ++tile_i) {
            for (int l = 0; l < NL; ++l) {
                const int N = cN[l];
                const uint32_t idesc = idesc_tf32(128, N);
                if (lane == 0) TC_TRACE(tile_i, l, 0);          // MMA warp starts waiting for A
                mbar_wait(a_ready, uses & 1);
                ++uses;
                fence_after_sync();
                if (lane == 0) TC_TRACE(tile_i, l, 1);          // A ready seen
                const uint32_t a_hi = tmem + cAHI + cAoff[l], a_lo = tmem + cALO + cAoff[l], d = tmem + cD;
                const int nchunks = cK[l] / 16;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(full + stage, phase);
                    fence_after_sync();
                    const uint32_t sb = smem_u32(smem + stage * kStageBytes);
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint32_t k = c * 16 + s * 8;
                            const uint64_t b_hi = bdesc_kmajor(sb + s * (2 * N * 16), N);
                            const uint64_t b_lo = bdesc_kmajor(sb + N * 64 + s * (2 * N * 16), N);
                            mma_tf32_ts(d, a_lo + k, b_hi, idesc, (c | s) ? 1u : 0u);
                            mma_tf32_ts(d, a_hi + k, b_lo, idesc, 1u);
                            mma_tf32_ts(d, a_hi + k, b_hi, idesc, 1u);
                        }
                        mma_commit_mcast(empty + stage, kClusterMask);  // this CTA is done with the stage: tell every producer
                        if (c == nchunks - 1) mma_commit(mma_done);
                    }
                    __syncwarp();
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                if (lane == 0) TC_TRACE(tile_i, l, 2);          // all MMAs of the layer issued + committed
            }
        }
    } else if (warp == 10) {
        // ===================== scratch store warp: staged layer (shared memory) -> wgrad scratch by bulk TMA =====================
        if (spill) {
            const int g0s[8] = {gH1, gH2, gT, gHC, gG4, gG3, gG2, gG1};   // order in which the workers produce the layers
            uint32_t sc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {   // dummy iterations store nothing
#pragma unroll
                for (int i = 0; i < 8; ++i, ++sc) {
                    const int b = sc & 1;
                    mbar_wait(st_full + b, (sc >> 1) & 1);
                    if (elect_one()) {
                        const unsigned char *src = smem + SM::oStaging + b * kStagingBytes;
                        for (int q4 = 0; q4 < 4; ++q4)
                            bulk_s2g(p.wg_scratch + ((size_t)tile * 4 + q4) * kSliceBytes + (size_t)g0s[i] * 512, src + q4 * 16384, 16384u);
                        bulk_commit();
                        bulk_wait_read0();                    // the staging buffer may be rewritten
                        mbar_arrive(st_free + b);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) bulk_wait_all0();
            __syncwarp();
        }
    } else {
        // ===================== workers: two threads per sample row (64 accumulator columns each) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which 64 columns
        const int col0 = half * 64;
        const bool lead = half == 0;                  // the row's thread that also gathers / scatters
        const int m = q * 32 + lane;                  // row of the tile
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t done_uses = 0;
        uint32_t nomask[2] = {0u, 0u};
        int tile_i = 0, lcount = -1;                  // trace bookkeeping (warp 2 lane 0 stamps)
        uint32_t sc = 0;                              // staged layers so far (two staging buffers alternate)
        // staging: this thread's 16-byte column of the current buffer ([slice q][group][lane][16 B]); waits until the
        // bulk store that last read the buffer is done
        bool real_tile = true;
        auto stage_begin = [&]() -> unsigned char * {
            if (!spill || !real_tile) return nullptr;
            const int b = sc & 1;
            if (sc >= 2) mbar_wait(st_free + b, ((sc >> 1) - 1) & 1);
            return smem + SM::oStaging + b * kStagingBytes + q * 16384 + lane * 16;
        };
        auto stage_end = [&]() {
            if (!spill || !real_tile) return;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> bulk-copy engine
            mbar_arrive(st_full + (sc & 1));
            ++sc;
        };
        auto layer_done = [&]() {
            mbar_wait(mma_done, done_uses & 1);
            ++done_uses;
            fence_after_sync();
            if (threadIdx.x == 64) TC_TRACE(tile_i, lcount, 3);     // worker sees the accumulators
        };
        auto a_is_ready = [&]() {
            tmem_wait_st();
            fence_before_sync();
            if (threadIdx.x == 64) TC_TRACE(tile_i, lcount + 1, 5);  // worker has produced the A of layer lcount+1
            mbar_arrive(a_ready);
            ++lcount;
        };
        for (int it = 0; it < iters; ++it, ++tile_i, lcount = -1) {
            const int tile = blockIdx.x + it * gridDim.x;
            real_tile = tile < ntiles;
            const int s = real_tile ? tile * 128 + m : nsamp;      // rows of a dummy iteration are all out of range
            unsigned char *scr = nullptr;   // this thread's 16-byte column in the tile's wgrad scratch (small groups go direct)
            if (spill && real_tile) scr = p.wg_scratch + ((size_t)tile * 4 + q) * kSliceBytes + (size_t)lane * 16;
            int vox = -1, ray = -1;
            float z = 0.0f, px = 0.f, py = 0.f, pz = 0.f;
            if (threadIdx.x == 64) TC_TRACE(tile_i, 0, 6);            // gather starts
            // ---- features -> A[:, 128:144) (the lead thread of each row) ----
            if (lead) {
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
                        vox = __ldg(p.samp_vox + s);
                        z = __ldg(p.samp_z + s);
                        ray = __ldg(p.hit_ray + __ldg(p.samp_ray + s));
                        const float x = __fadd_rn(__ldg(p.rays_o + ray * 3 + 0), __fmul_rn(__ldg(p.rays_d + ray * 3 + 0), z));
                        const float y = __fadd_rn(__ldg(p.rays_o + ray * 3 + 1), __fmul_rn(__ldg(p.rays_d + ray * 3 + 1), z));
                        const float zz = __fadd_rn(__ldg(p.rays_o + ray * 3 + 2), __fmul_rn(__ldg(p.rays_d + ray * 3 + 2), z));
                        px = __fadd_rn(__fdiv_rn(__fsub_rn(x, __ldg(p.centres + (size_t)vox * 3 + 0)), p.voxel_size), 0.5f);
                        py = __fadd_rn(__fdiv_rn(__fsub_rn(y, __ldg(p.centres + (size_t)vox * 3 + 1)), p.voxel_size), 0.5f);
                        pz = __fadd_rn(__fdiv_rn(__fsub_rn(zz, __ldg(p.centres + (size_t)vox * 3 + 2)), p.voxel_size), 0.5f);
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
                if (scr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4 *>(scr + (size_t)(gF + j) * 512) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                }
            }
            a_is_ready();
            uint32_t m1[2], m2[2], mc[2];
            // ---- forward ----
            layer_done();
            epilogue64<0>(trow, col0, sBias, m1, stage_begin());                                     // h1
            stage_end();
            a_is_ready();
            layer_done();
            epilogue64<0>(trow, col0, sBias + 128, m2, stage_begin());                               // h2
            stage_end();
            a_is_ready();
            layer_done();
            epilogue64<1>(trow, col0, sBias + 256, nomask, stage_begin());                           // t (no activation)
            stage_end();
            float sdf = 0.0f;
            if (lead) {
                uint32_t v[16];
                tmem_ld16(trow + cD + 128, v);   // sdf = row 0 of W3, packed as output column 128
                tmem_wait_ld();
                sdf = __uint_as_float(v[0]) + sBias[512];
            }
            a_is_ready();
            layer_done();
            epilogue64<0>(trow, col0, sBias + 384, mc, stage_begin());                               // hc
            stage_end();
            a_is_ready();
            layer_done();
            float r = 0.f, g = 0.f, b = 0.f;
            if (lead) {
                uint32_t v[16];
                tmem_ld16(trow + cD, v);
                tmem_wait_ld();
                r = sigmoid_f(__uint_as_float(v[0]) + sBias[513]);
                g = sigmoid_f(__uint_as_float(v[1]) + sBias[514]);
                b = sigmoid_f(__uint_as_float(v[2]) + sBias[515]);
            }
            if (!BWD) {
                if (lead && s < nsamp) *reinterpret_cast<float4 *>(p.out + (size_t)s * 4) = make_float4(r, g, b, sdf);
                continue;   // D has been read; the next tile's first MMA is ordered behind it through a_ready
            }
            // ---- backward ----
            float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lead) {
                if (s < nsamp) go = __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)s * 4));
                // dL/d(pre-sigmoid rgb): grad * (1 - y) * y ; A[:, 0:16) = [g5 r,g,b, 0...]
                const float g5[4] = {go.x * (1.0f - r) * r, go.y * (1.0f - g) * g, go.z * (1.0f - b) * b, go.w};
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) { hi[e] = 0u; lo[e] = 0u; }
#pragma unroll
                for (int e = 0; e < 3; ++e) tf32_split(g5[e], hi[e], lo[e]);
                tmem_st16(trow + cAHI, hi);
                tmem_st16(trow + cALO, lo);
                if (scr) {
                    *reinterpret_cast<float4 *>(scr + (size_t)gG5 * 512) = make_float4(g5[0], g5[1], g5[2], g5[3]);
#pragma unroll
                    for (int j = 1; j < 4; ++j) *reinterpret_cast<float4 *>(scr + (size_t)(gG5 + j) * 512) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            a_is_ready();
            layer_done();
            epilogue64<2>(trow, col0, nullptr, mc, stage_begin());                                   // g_hc
            stage_end();
            a_is_ready();
            layer_done();
            epilogue64<3>(trow, col0, nullptr, nomask, stage_begin());                               // g_t
            stage_end();
            float gf[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) gf[e] = 0.0f;
            if (lead) {
                uint32_t v[16], hi[16], lo[16];
                tmem_ld16(trow + cD + 128, v);   // g_f, part through W4's last 16 input columns
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) { gf[e] = __uint_as_float(v[e]); hi[e] = 0u; lo[e] = 0u; }
                tf32_split(go.w, hi[0], lo[0]);  // A[:, 128] = g_sdf pairs with W3 row 0 (packed at k = 128)
                tmem_st16(trow + cAHI + 128, hi);
                tmem_st16(trow + cALO + 128, lo);
            }
            a_is_ready();
            layer_done();
            epilogue64<2>(trow, col0, nullptr, m2, stage_begin());                                   // g_h2
            stage_end();
            a_is_ready();
            layer_done();
            epilogue64<2>(trow, col0, nullptr, m1, stage_begin());                                   // g_h1
            stage_end();
            a_is_ready();
            layer_done();
            if (!lead) continue;
            {
                uint32_t v[16];
                tmem_ld16(trow + cD, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) gf[e] += __uint_as_float(v[e]);
            }
            if (s >= nsamp) continue;
            if (p.g_feat) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4 *>(p.g_feat + (size_t)s * 16 + 4 * j) = make_float4(gf[4 * j], gf[4 * j + 1], gf[4 * j + 2], gf[4 * j + 3]);
            }
            if (!p.feat && (p.grad_emb || p.grad_rays)) {
                // trilinear backward of this sample
                float gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = __ldg(p.vertex_idx + (size_t)vox * 8 + i);
                    const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
                    const float w = (wx * wy) * wz;
                    if (p.grad_emb) {
                        float *dst = p.g_emb + (size_t)row * 16;
#pragma unroll
                        for (int e = 0; e < 4; ++e) red_add_v4(dst + 4 * e, w * gf[4 * e], w * gf[4 * e + 1], w * gf[4 * e + 2], w * gf[4 * e + 3]);
                    }
                    if (p.grad_rays) {
                        const float4 *er = reinterpret_cast<const float4 *>(p.emb + (size_t)row * 16);
                        float d = 0.0f;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 v = __ldg(er + e);
                            d = fmaf(gf[4 * e], v.x, d); d = fmaf(gf[4 * e + 1], v.y, d);
                            d = fmaf(gf[4 * e + 2], v.z, d); d = fmaf(gf[4 * e + 3], v.w, d);
                        }
                        gp[0] += d * ((i & 4) ? 1.0f : -1.0f) * (wy * wz);
                        gp[1] += d * ((i & 2) ? 1.0f : -1.0f) * (wx * wz);
                        gp[2] += d * ((i & 1) ? 1.0f : -1.0f) * (wx * wy);
                    }
                }
                if (p.grad_rays) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const float gx = gp[a] / p.voxel_size;
                        atomicAdd(p.g_rays_o + ray * 3 + a, gx);
                        atomicAdd(p.g_rays_d + ray * 3 + a, z * gx);
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync();   // no CTA leaves while a peer may still multicast into its shared memory or signal its barriers
    if (warp == 0) tmem_dealloc(tmem, tc::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Weight gradients: dW[n][k] = sum over samples of G[p][n] * A[p][k], per layer, as MMAs whose
// reduction dimension is the SAMPLE index.  Operands come from the scratch k_field_tc<true> wrote:
// per tile and per 32-sample slice, "groups" of 4 features x 32 samples (16 B per sample, written
// coalesced by the sample-per-thread workers).  The loader threads transpose 4x4 blocks in
// registers into the K-major operand layout of umma.cuh (16-byte chunks = 4 consecutive samples
// of one feature), split hi/lo (3xTF32) into shared memory, one thread issues the MMAs (both
// operands from shared memory), and the accumulators stay in tensor memory for ALL tiles of the CTA:
//   cols [0,128) dW2   [128,256) dW3 feature rows   [256,400) dW4   [400,416) dW1
//        [416,432) HC^T*G5 (dW5 rows = columns 0..2)   [432,448) H2^T*G5 (dW3 row 0 = column 3)
// Bias gradients are column sums of the G groups, accumulated by the loader threads.
// ------------------------------------------------------------------------------------------
namespace wg {
constexpr int kXformWarps = 12;
constexpr int kXform = kXformWarps * 32;
constexpr int kThreads = kXform + 64;           // transform warps, then the TMA producer warp, then the MMA issuer warp
constexpr int kStepSamples = 32;                // samples contracted per step (one scratch slice; 4 MMA k-steps)
constexpr int kQuads = kStepSamples / 4;        // 16-byte k-chunks per step
constexpr int kPieces = 32 / kStepSamples;      // steps per 32-sample scratch slice
constexpr int kStepsPerTile = (128 / kStepSamples) * 6;
constexpr int kRawPitch = kStepSamples * 16;    // raw group pitch (unpadded: whole operands arrive as single bulk copies)
constexpr int kRawBytes = 68 * kRawPitch;       // one step's raw fp32 groups: A' (32) + B' (<= 36)
constexpr int kOpBytes = 2 * 272 * kStepSamples * 4;   // hi + lo of A' (128 rows) + B' (<= 144 rows)
constexpr int kRawStages = kStepSamples == 32 ? 2 : 6, kOpStages = 2;
constexpr int oOps = kRawStages * kRawBytes;
constexpr int oBars = oOps + kOpStages * kOpBytes;   // raw_full[6] raw_free[6] op_full[2] op_free[2] all_done, tmem ptr
constexpr int oBiasAcc = oBars + 256;           // float[4][128] column sums of G1..G4 + [16] of G5
constexpr int kSmemBytes = oBiasAcc + 4 * (4 * 128 + 16);
struct Step { int a_group, a_cnt, b_group, b_cnt, b2_group, b2_cnt, dcol; };
// A' (M' = 128 output rows n) , B' (N' columns), accumulator column
__device__ __constant__ Step cSteps[6] = {
    {tc::gG2, 32, tc::gH1, 32, 0, 0, 0},          // dW2[n][k]   = G2^T H1
    {tc::gG3, 32, tc::gH2, 32, 0, 0, 128},        // dW3[1+j][k] = G3^T H2
    {tc::gG4, 32, tc::gT, 32, tc::gF, 4, 256},    // dW4[n][j]   = G4^T [T; F]
    {tc::gG1, 32, tc::gF, 4, 0, 0, 400},          // dW1[n][k]   = G1^T F
    {tc::gHC, 32, tc::gG5, 4, 0, 0, 416},         // D[k][c]     = HC^T G5  (c<3: dW5[c][k])
    {tc::gH2, 32, tc::gG5, 4, 0, 0, 432},         // D[k][c]     = H2^T G5  (c=3: dW3[0][k])
};
}  // namespace wg

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// Three-role pipeline per CTA (persistent over its tiles, 24 steps per tile = 4 slices x 6 GEMMs):
//   TMA producer   : bulk-copies the step's raw fp32 groups from the scratch into a small ring
//   transform warps: raw -> registers (transpose, hi/lo split) -> operand buffers (2 stages)
//   MMA issuer     : 12 MMAs per step (4 k-steps x 3xTF32), commit frees the operand buffer
#define WG_TRACE(g, slot)                                                                           \
    do {                                                                                            \
        if (g_tc_trace && blockIdx.x == 0 && (g) < 40) g_tc_trace[320 + (g) * 8 + (slot)] = clock64();     \
    } while (0)

__global__ void __launch_bounds__(wg::kThreads, 1) k_wgrad_tc(FieldParams p)
{
    using namespace wg;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *raw_full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *raw_free = raw_full + kRawStages, *op_full = raw_free + kRawStages, *op_free = op_full + 2, *all_done = op_free + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oBars + 192);
    float *sBiasAcc = reinterpret_cast<float *>(smem + oBiasAcc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    if (tid == 0) {
        for (int i = 0; i < kRawStages; ++i) { mbar_init(raw_full + i, 1); mbar_init(raw_free + i, kXform); }
        for (int i = 0; i < 2; ++i) { mbar_init(op_full + i, kXform); mbar_init(op_free + i, 1); }
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == kXformWarps + 1) tmem_alloc(tmem_ptr, 512);
    for (int i = tid; i < 4 * 128 + 16; i += kThreads) sBiasAcc[i] = 0.0f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_tiles * kStepsPerTile;

    if (warp == kXformWarps) {
        // ===================== TMA producer =====================
        for (int g = 0; g < nsteps; ++g) {
            const int tl = g / kStepsPerTile, hs = (g / 6) % (4 * kPieces), st = g % 6, rs = g % kRawStages, use = g / kRawStages;
            const Step S = cSteps[st];
            // piece hs of the tile: lanes [kStepSamples (hs % kPieces), +kStepSamples) of slice hs / kPieces
            const unsigned char *slice = p.wg_scratch + ((size_t)(blockIdx.x + (size_t)tl * gridDim.x) * 4 + hs / kPieces) * tc::kSliceBytes +
                                         (hs % kPieces) * (kStepSamples * 16);
            if (use >= 1) mbar_wait(raw_free + rs, (use - 1) & 1);
            if (lane == 0) WG_TRACE(g, 0);
            if (elect_one()) {
                unsigned char *dst = smem + rs * kRawBytes;
                mbar_arrive_expect_tx(raw_full + rs, (uint32_t)(S.a_cnt + S.b_cnt + S.b2_cnt) * 512u);
                bulk_g2s(dst, slice + (size_t)S.a_group * 512, S.a_cnt * 512u, raw_full + rs);
                bulk_g2s(dst + S.a_cnt * 512, slice + (size_t)S.b_group * 512, S.b_cnt * 512u, raw_full + rs);
                if (S.b2_cnt) bulk_g2s(dst + (S.a_cnt + S.b_cnt) * 512, slice + (size_t)S.b2_group * 512, S.b2_cnt * 512u, raw_full + rs);
            }
            __syncwarp();
        }
    } else if (warp == kXformWarps + 1) {
        // ===================== MMA issuer =====================
        uint32_t started = 0u;            // bit st: accumulator of GEMM st has been written
        for (int g = 0; g < nsteps; ++g) {
            const int st = g % 6, ob = g & 1;
            const Step S = cSteps[st];
            const int rowsB = 4 * (S.b_cnt + S.b2_cnt);
            mbar_wait(op_full + ob, (g >> 1) & 1);
            fence_after_sync();
            if (lane == 0) WG_TRACE(g, 4);
            const uint32_t base = smem_u32(smem + oOps + ob * kOpBytes);
            const uint32_t a_hi = base, b_hi = base + 128 * kQuads * 16, a_lo = base + 272 * kQuads * 16, b_lo = a_lo + 128 * kQuads * 16;
            const uint32_t idesc = idesc_tf32(128, rowsB);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kQuads / 2; ++ks) {           // 8 samples = 2 k-chunks per MMA
                    const uint64_t dah = bdesc_kmajor(a_hi + ks * 2 * 128 * 16, 128), dal = bdesc_kmajor(a_lo + ks * 2 * 128 * 16, 128);
                    const uint64_t dbh = bdesc_kmajor(b_hi + ks * 2 * rowsB * 16, rowsB), dbl = bdesc_kmajor(b_lo + ks * 2 * rowsB * 16, rowsB);
                    mma_tf32_ss(tmem + S.dcol, dal, dbh, idesc, (!((started >> st) & 1u) && ks == 0) ? 0u : 1u);
                    mma_tf32_ss(tmem + S.dcol, dah, dbl, idesc, 1u);
                    mma_tf32_ss(tmem + S.dcol, dah, dbh, idesc, 1u);
                }
                mma_commit(op_free + ob);
                if (g == nsteps - 1) mma_commit(all_done);
            }
            __syncwarp();
            if (lane == 0) WG_TRACE(g, 5);
            started |= 1u << st;
        }
    } else {
        // ===================== transform warps =====================
        for (int g = 0; g < nsteps; ++g) {
            const int st = g % 6, rs = g % kRawStages, ob = g & 1;
            const Step S = cSteps[st];
            const int rowsB = 4 * (S.b_cnt + S.b2_cnt);          // B' rows (N'); A' always has 128
            mbar_wait(raw_full + rs, (g / kRawStages) & 1);
            if (tid == 0) WG_TRACE(g, 1);
            if (g >= 2) mbar_wait(op_free + ob, ((g >> 1) - 1) & 1);
            if (tid == 0) WG_TRACE(g, 2);
            const unsigned char *raw = smem + rs * kRawBytes;
            unsigned char *sbuf = smem + oOps + ob * kOpBytes;
            unsigned char *sA_hi = sbuf, *sB_hi = sbuf + 128 * kQuads * 16;        // [k-chunk][rows][16 B]
            unsigned char *sA_lo = sbuf + 272 * kQuads * 16, *sB_lo = sA_lo + 128 * kQuads * 16;
            // Scratch unit = 4 features of ONE sample; the K-major operand wants 16-byte chunks of 4 consecutive
            // SAMPLES of one feature.  A warp takes one feature group per iteration: lane = (4-sample block q,
            // feature c of the group).  The four scalar reads of a lane are issued in a q-dependent rotation so
            // that the 32 lanes always hit 32 distinct banks of the unpadded [sample][4 features] group.
            const int ngroups = S.a_cnt + S.b_cnt + S.b2_cnt;
            const int q = lane >> 2, c = lane & 3, rot = (q >> 1) & 3;
#pragma unroll 2
            for (int gi = warp; gi < ngroups; gi += kXformWarps) {
                const bool isA = gi < S.a_cnt;
                const int rl = (isA ? gi : gi - S.a_cnt) * 4 + c, rows = isA ? 128 : rowsB;
                const float *src = reinterpret_cast<const float *>(raw + (size_t)gi * 512 + q * 64) + c;
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = src[((j + rot) & 3) * 4];      // r[j] = sample 4q + (j + rot) % 4
                // undo the rotation branch-free (two conditional rotations): t[k] = sample 4q + k = r[(k - rot) & 3]
                const bool rot1 = (rot & 1) != 0, rot2 = (rot & 2) != 0;
                const float a0 = rot1 ? r[3] : r[0], a1 = rot1 ? r[0] : r[1], a2 = rot1 ? r[1] : r[2], a3 = rot1 ? r[2] : r[3];
                const float t[4] = {rot2 ? a2 : a0, rot2 ? a3 : a1, rot2 ? a0 : a2, rot2 ? a1 : a3};
                uint4 hi, lo;
                tf32_split(t[0], hi.x, lo.x); tf32_split(t[1], hi.y, lo.y);
                tf32_split(t[2], hi.z, lo.z); tf32_split(t[3], hi.w, lo.w);
                const size_t off = (size_t)q * rows * 16 + (size_t)rl * 16;
                *reinterpret_cast<uint4 *>((isA ? sA_hi : sB_hi) + off) = hi;
                *reinterpret_cast<uint4 *>((isA ? sA_lo : sB_lo) + off) = lo;
                // bias gradients: column sums of the G operands (steps 0..3 carry G2, G3, G4, G1 as A';
                // step 4 carries G5 = (g5 r,g,b, g_sdf) as the first B' group)
                const bool is_g = (st < 4 && isA) || (st == 4 && gi == S.a_cnt);     // warp-uniform
                if (is_g) {
                    float sum = (t[0] + t[1]) + (t[2] + t[3]);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                    if (q == 0) {   // the same thread owns this (step, row) in every slice: no race, fixed order
                        float *acc = (st < 4) ? sBiasAcc + st * 128 + rl : sBiasAcc + 512 + c;
                        *acc += sum;
                    }
                }
            }
            mbar_arrive(raw_free + rs);                                   // raw stage consumed
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> async proxy (MMA)
            mbar_arrive(op_full + ob);
            if (tid == 0) WG_TRACE(g, 3);
        }
        // drain: accumulators -> global gradients (warps 0-3 own the four TMEM lane quarters)
        if (nsteps > 0 && warp < 4) {
            mbar_wait(all_done, 0);
            fence_after_sync();
            const int n = warp * 32 + lane;                          // accumulator row = TMEM lane
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            auto flush = [&](int col0, int ncols, float *dst_row) {   // dst_row: &dW[n][0], ncols % 16 == 0
                for (int c0 = 0; c0 < ncols; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(trow + col0 + c0, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        red_add_v4(dst_row + c0 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
                }
            };
            flush(0, 128, p.g_dec.W2 + (size_t)n * 128);
            flush(128, 128, p.g_dec.W3 + (size_t)(1 + n) * 128);
            flush(256, 144, p.g_dec.W4 + (size_t)n * 144);
            flush(400, 16, p.g_dec.W1 + (size_t)n * 16);
            {
                uint32_t v[16], w[16];
                tmem_ld16(trow + 416, v);
                tmem_ld16(trow + 432, w);
                tmem_wait_ld();
                atomicAdd(p.g_dec.W5 + n, __uint_as_float(v[0]));
                atomicAdd(p.g_dec.W5 + 128 + n, __uint_as_float(v[1]));
                atomicAdd(p.g_dec.W5 + 256 + n, __uint_as_float(v[2]));
                atomicAdd(p.g_dec.W3 + n, __uint_as_float(w[3]));
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (nsteps > 0 && tid < 128) {
        // bias gradients (sBiasAcc rows: step 0 = G2, 1 = G3, 2 = G4, 3 = G1); complete after the barrier above
        const int n = tid;
        atomicAdd(p.g_dec.b2 + n, sBiasAcc[n]);
        atomicAdd(p.g_dec.b3 + 1 + n, sBiasAcc[128 + n]);
        atomicAdd(p.g_dec.b4 + n, sBiasAcc[256 + n]);
        atomicAdd(p.g_dec.b1 + n, sBiasAcc[384 + n]);
        if (n < 3) atomicAdd(p.g_dec.b5 + n, sBiasAcc[512 + n]);
        if (n == 3) atomicAdd(p.g_dec.b3, sBiasAcc[515]);
    }
    if (warp == kXformWarps + 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// stand-alone GEMMs through the same primitives (unit tests of descriptors / TMEM addressing).
// mode 0/1: D[128,N] = A[128,K] * B[N,K]^T with A in tensor memory, B K-major in shared memory
//           (1xTF32 / 3xTF32).  K % 8 == 0, N % 16 == 0, both <= 144.
// mode 4:   same product with BOTH operands K-major in shared memory (the wgrad form), 3xTF32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_debug_umma_gemm(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                                                            int N, int K, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    float *sB = reinterpret_cast<float *>(smem);
    const int warp = threadIdx.x >> 5, m = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    if (mode == 4) {
        // A and B -> smem, K-major, per k-step: [A hi (2 kc x 128 rows x 4)][B hi (2 kc x N x 4)][A lo][B lo]
        const int blk = 2 * (128 + N) * 8;
        for (int i = threadIdx.x; i < (128 + N) * K; i += 128) {
            const int r = i / K, k = i % K;
            const int st = k >> 3, kc = (k >> 2) & 1, e = k & 3;
            uint32_t hi, lo;
            const bool isA = r < 128;
            tf32_split(isA ? A[(size_t)r * K + k] : B[(size_t)(r - 128) * K + k], hi, lo);
            const int rows = isA ? 128 : N, rr = isA ? r : r - 128;
            float *base = sB + st * blk + (isA ? 0 : 128 * 8);
            base[kc * rows * 4 + rr * 4 + e] = __uint_as_float(hi);
            base[(128 + N) * 8 + kc * rows * 4 + rr * 4 + e] = __uint_as_float(lo);
        }
    } else if (mode < 2) {
        // B -> smem in the K-major operand layout, one block of (hi, lo) per MMA k-step
        for (int i = threadIdx.x; i < N * K; i += 128) {
            const int n = i / K, k = i % K;
            const int st = k >> 3, kc = (k >> 2) & 1, e = k & 3;
            uint32_t hi, lo;
            tf32_split(B[i], hi, lo);
            float *blk = sB + st * (2 * N * 8);
            blk[kc * N * 4 + n * 4 + e] = __uint_as_float(hi);
            blk[N * 8 + kc * N * 4 + n * 4 + e] = __uint_as_float(lo);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    if (mode < 2) {
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
    } else {
        // sentinel in D: tells "MMA wrote zeros" from "MMA did not write"
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t sv[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) sv[e] = __float_as_uint(123.0f);
            tmem_st16(trow + 288 + c0, sv);
        }
        tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        const uint32_t sb = smem_u32(sB);
        if (mode == 4) {
            const uint32_t idesc = idesc_tf32(128, N);
            const uint32_t blkb = 2 * (128 + N) * 8 * 4;
            for (int st = 0; st < K / 8; ++st) {
                const uint32_t a_hi = sb + st * blkb, b_hi = a_hi + 128 * 8 * 4;
                const uint32_t a_lo = a_hi + (128 + N) * 8 * 4, b_lo = b_hi + (128 + N) * 8 * 4;
                mma_tf32_ss(tmem + 288, bdesc_kmajor(a_lo, 128), bdesc_kmajor(b_hi, N), idesc, st ? 1u : 0u);
                mma_tf32_ss(tmem + 288, bdesc_kmajor(a_hi, 128), bdesc_kmajor(b_lo, N), idesc, 1u);
                mma_tf32_ss(tmem + 288, bdesc_kmajor(a_hi, 128), bdesc_kmajor(b_hi, N), idesc, 1u);
            }
        } else if (mode < 2) {
            const uint32_t idesc = idesc_tf32(128, N);
            for (int st = 0; st < K / 8; ++st) {
                const uint64_t b_hi = bdesc_kmajor(sb + st * (2 * N * 8 * 4), N);
                const uint64_t b_lo = bdesc_kmajor(sb + st * (2 * N * 8 * 4) + N * 8 * 4, N);
                if (mode == 1) {
                    mma_tf32_ts(tmem + 288, tmem + 144 + st * 8, b_hi, idesc, st ? 1u : 0u);
                    mma_tf32_ts(tmem + 288, tmem + st * 8, b_lo, idesc, 1u);
                    mma_tf32_ts(tmem + 288, tmem + st * 8, b_hi, idesc, 1u);
                } else {
                    mma_tf32_ts(tmem + 288, tmem + st * 8, b_hi, idesc, st ? 1u : 0u);
                }
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
    for (int l = 0; l < tc::kLayersAll; ++l) total += tc::hN[l] * tc::hK[l];
    k_tc_pack<<<ceil_div(total, 256), 256, 0, st>>>(d, ws_tc);
    PSLAM_CHECK_LAUNCH("tc_pack");
    return 0;
}

size_t tc_wgrad_scratch_bytes(int max_samples) { return (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128) * tc::kTileBytes; }

template <bool BWD>
static int launch_tc(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    static PerDevice once = {};
    bool &configured = once.done[current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_field_tc<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Smem<BWD>::bytes);
        if (e != cudaSuccess) { set_error("field_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    // persistent grid = as many whole clusters as can be resident at once (GPC boundaries may strand a few SMs)
    static PerDevice clusters = {};
    int &max_clusters = clusters.value[current_device()];
    if (max_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / tc::kCluster * tc::kCluster);
        cfg.blockDim = dim3(tc::kThreads);
        cfg.dynamicSmemBytes = tc::Smem<BWD>::bytes;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = tc::kCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_field_tc<BWD>, &cfg);
        if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / tc::kCluster; }
        max_clusters = n < num_sms() / tc::kCluster ? n : num_sms() / tc::kCluster;
    }
    const int tiles = ceil_div(max_samples, 128);
    int grid = ceil_div(tiles > 0 ? tiles : 1, tc::kCluster) * tc::kCluster;
    if (grid > max_clusters * tc::kCluster) grid = max_clusters * tc::kCluster;
    k_field_tc<BWD><<<grid, tc::kThreads, tc::Smem<BWD>::bytes, st>>>(fp, fp.ws_tc);
    PSLAM_CHECK_LAUNCH(BWD ? "field_tc_backward" : "field_tc_forward");
    return 0;
}

int tc_launch_field_forward(const FieldParams &fp, int max_samples, cudaStream_t st) { return launch_tc<false>(fp, max_samples, st); }

// backward: dgrad chain (+ trilinear backward); when decoder gradients are wanted the scratch must be
// provided and the wgrad kernel follows on the same stream
int tc_launch_field_backward(const FieldParams &fp_in, int max_samples, cudaStream_t st, int part)
{
    FieldParams fp = fp_in;
    if (!fp.grad_dec) fp.wg_scratch = nullptr;
    if (part != 2)
        if (int rc = launch_tc<true>(fp, max_samples, st)) return rc;
    if (!fp.grad_dec || part == 1) return 0;
    static PerDevice once = {};
    bool &configured = once.done[current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::kSmemBytes);
        if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int tiles = ceil_div(max_samples, 128);
    const int grid = tiles < num_sms() ? (tiles > 0 ? tiles : 1) : num_sms();
    k_wgrad_tc<<<grid, wg::kThreads, wg::kSmemBytes, st>>>(fp);
    PSLAM_CHECK_LAUNCH("wgrad_tc");
    return 0;
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_debug_tc_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(g_tc_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { set_error("tc_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pslam_debug_umma_gemm(const float *A, const float *B, float *D, int N, int K, int mode, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(A && B && D, PSLAM_E_ARG, "null pointer");
    PSLAM_CHECK_ARG(N >= 16 && N <= 144 && N % 16 == 0 && K >= 8 && K <= 144 && K % 8 == 0, PSLAM_E_RANGE, "N in 16..144 step 16, K in 8..144 step 8");
    PSLAM_CHECK_ARG(mode == 0 || mode == 1 || mode == 4, PSLAM_E_RANGE, "mode must be 0, 1 or 4");
    const int smem = mode == 4 ? 2 * (128 + N) * K * 4 : 2 * N * K * 4;
    PSLAM_CHECK_ARG(smem <= 200 * 1024, PSLAM_E_RANGE, "operands do not fit in shared memory");
    cudaError_t e = cudaFuncSetAttribute(k_debug_umma_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("debug_umma: %s", cudaGetErrorString(e)); return (int)e; }
    k_debug_umma_gemm<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, mode);
    PSLAM_CHECK_LAUNCH("debug_umma_gemm");
    return 0;
}
