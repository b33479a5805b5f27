// Kernel 4 on the 5th-generation tensor cores, 3xF16 build, one tile in flight per CTA: the width-128 SDF/colour decoder
// (src/variations/nrgbd.py:116-135) fused with the trilinear corner-embedding lookup
// (src/variations/render_helpers.py:105-156, 47-59), forward and backward.
//
// Same tile / role structure as field_tc.cu (tile = 128 samples = the 128 lanes of tensor memory, A operand
// from tensor memory, weights streamed by multicast bulk TMA, persistent 2-CTA clusters), but every value is
// split into TWO f16 halves kept in their precise window by power-of-two scales (field_bf.cuh, umma.cuh: h16_split2):
//         x = hi + lo,  hi = rn_f16(x), lo = rn_f16(x - hi)          (22 significant bits)
//         a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi                   (kind::f16 MMAs, fp32 accumulate in TMEM)
// which is fp32-equivalent (~2^-21 relative per product) at half the MMA count of 3xTF32.  (bf16 halves were built first
// and rejected: 16 bits, 2e-4 on the rendered sdf.)  What K = 16 per MMA buys over 3xTF32:
//   * half the tcgen05.mma instructions, half the weight-stream bytes, half the tensor-memory columns and tcgen05.st
//     traffic for the A operand (two values per 32-bit column);
//   * hi and lo of one value are 4 bytes together, so the wgrad scratch can hold the operands ALREADY SPLIT
//     and in MMA order at the same 4 B/element as raw fp32.  The scratch is written in the MN-major
//     (sample-major) core-matrix layout, so k_wgrad_bf is nothing but bulk-TMA loads feeding tcgen05.mma
//     with both operands MN-major from shared memory: no transform warps, no transposition, no re-split.
//   TMEM columns (k_field_bf): A_hi [0,72)  A_lo [72,144)  D0 [144,288)  D1 [288,432).
// This file now serves the calls field_pp.cu / field_bw.cu do not take: the recomputing backward (pslam_decoder_bwd, and any
// backward whose forward did not save), launches without feature rows, PSLAM_OPT_TILES = 1; and it owns k_wgrad_bf /
// k_wgrad_finish, the trilinear kernels and the launch logic of the 3xF16 build.
#include "field_bf.cuh"
#include "kernels.h"
#include "trilinear.cuh"
#include <mutex>
#include <type_traits>

namespace pslam {

using namespace umma;


// Re-packs the decoder into the bf16 weight stream: layers in order, each as ceil(K/32) chunks of
// [hi block | lo block], each block = kk/8 k-chunks x N rows x 16 B (8 bf16 along K) -- K-major, no swizzle.
__global__ void k_bf_pack(pslam_decoder_t d, uint16_t *__restrict__ out, int *__restrict__ range_flag)
{
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int base = 0;   // uint16 offset of the layer in `out`
#pragma unroll
    for (int l = 0; l < bf::kLayersPack; ++l) {
        const int N = bf::cN[l], K = bf::cK[l];
        if (i < N * K) {
            const int n = i / K, k = i % K;
            const int ks = k >> 4, c = bf::kstep_chunk(K, ks), kk = bf::chunk_kk(K, c);
            const int kr = bf::kstep_slot(K, ks) * 16 + (k & 15);     // position of k inside its chunk
            uint32_t hi, lo;
            const float ws = bf::kScale * tc_weight(d, l, n, k);
            if (range_flag && fabsf(ws) >= 32752.0f) atomicOr(range_flag, 4);
            h16_split2(ws, 0.0f, hi, lo);
            uint16_t *chunk = out + base + c * (N * bf::kChunkK * 2);
            const int off = (kr >> 3) * (N * 8) + n * 8 + (kr & 7);
            chunk[off] = (uint16_t)(hi & 0xffffu);
            chunk[N * kk + off] = (uint16_t)(lo & 0xffffu);
            return;
        }
        i -= N * K;
        base += 2 * N * K;
    }
}

__global__ void k_grad_scale(const float4 *__restrict__ g_out, int n, const int *__restrict__ n_dev, uint32_t *__restrict__ gscale)
{
    if (n_dev) n = *n_dev;
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v = __ldg(g_out + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(gscale, __float_as_uint(m));   // non-negative floats order like their bits
}

// One epilogue over this thread's 64 accumulator columns [col0, col0+64): accumulators -> registers ->
// (bias / activation / mask) -> hi/lo bf16 pairs -> next A operand (tensor memory, two values per column)
// and optionally the staging buffer of the wgrad scratch (the same packed words, 16 B = one core-matrix row).
//   MODE 0: y = relu(D/16 + bias), records y > 0 in mask[]     (forward hidden layers; bias pre-scaled x16, y = 16 x activation)
//   MODE 1: y = D/16 + bias                                    (forward, no activation)
//   MODE 2: y = mask ? D/16 : 0                                (dgrad through a ReLU; y = Sg x gradient)
//   MODE 3: y = D/16                                           (dgrad, no activation)
// HEAD: this is the hc layer -- its colour head (3 x 128, sigmoid outside) is taken right here on the CUDA cores from the fp32
// values (head[c] += W5[c][col] * y), so the N = 16 tensor-core layer that cost a full issue-bound layer slot for 3 useful
// columns is gone and the next A operand is not written at all.
template <int MODE, bool HEAD = false>
__device__ __forceinline__ void bf_epilogue16(uint32_t trow, uint32_t dcol, int c0, const float *bias, uint32_t &mask, int shift,
                                              unsigned char *stg, float &ymax, const float *w5 = nullptr, float *head = nullptr)
{
    using namespace bf;
    uint32_t v[16];
    tmem_ld16(trow + dcol + c0, v);
    tmem_wait_ld();
    uint32_t bits = 0u;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        float y = __uint_as_float(v[e]);
        if (MODE == 0) { y = fmaxf(fmaf(y, kInvScale, bias[c0 + e]), 0.0f); bits |= (y > 0.0f ? 1u : 0u) << e; }
        if (MODE == 1) y = fmaf(y, kInvScale, bias[c0 + e]);
        if (MODE == 2) y = ((mask >> (shift + e)) & 1u) ? y * kInvScale : 0.0f;
        if (MODE == 3) y = y * kInvScale;
        ymax = fmaxf(ymax, fabsf(y));                // range guard of the f16 operand window (checked once per tile)
        if (HEAD) {
            head[0] = fmaf(w5[c0 + e], y, head[0]);
            head[1] = fmaf(w5[128 + c0 + e], y, head[1]);
            head[2] = fmaf(w5[256 + c0 + e], y, head[2]);
        }
        v[e] = __float_as_uint(y);
    }
    if (MODE == 0) mask = shift ? (mask | (bits << 16)) : bits;     // (shift is 0 or 16; the low half is written first)
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h16_split2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]), hi[e], lo[e]);
    if (stg) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned char *dst = stg + (size_t)(c0 / 8 + j) * 128;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4 *>(dst + 16384) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
    }
    if (!HEAD) {
        tmem_st8(trow + cAHI + c0 / 2, hi);
        tmem_st8(trow + cALO + c0 / 2, lo);
    }
}

// optional timeline trace of CTA 0 (pslam_debug_bf_trace): [tile<4][layer<10][8] clock64 stamps (+ 40 x 8 for k_wgrad_bf)
__device__ long long *g_bf_trace = nullptr;
#define BF_TRACE(tile_i, layer, slot)                                                                   \
    do {                                                                                                \
        if (g_bf_trace && blockIdx.x == 0 && (tile_i) < 4) g_bf_trace[((tile_i) * 10 + (layer)) * 8 + (slot)] = clock64(); \
    } while (0)

// The MMA-issuing thread's state and one layer of the chain: D[128 x N] = A[128 x K] * W^T as K/16 k-steps of three
// MMAs (a_lo*b_hi, a_hi*b_lo, a_hi*b_hi), the weights arriving through the ring in chunks of two k-steps (j, j + 4).
// The workers publish the A operand in four quarters (a_ready[0..3]) and chunk j is issued as soon as quarter j is
// there, i.e. while the rest of the previous epilogue is still running; D alternates between two accumulator buffers
// so that this is safe.
template <int NS>
struct MmaIssuer {
    unsigned char *smem;
    uint64_t *full, *empty, *a_ready, *mma_done;
    uint32_t tmem;
    int stage, phase;
    uint32_t uses;   // layers issued so far: a_ready phase and accumulator buffer
    // one chunk: k-steps at k = kA and (kk == 32) k = kB
    // The weights of the layer's first chunks are waited for up front (weights_ahead), while the issuer would otherwise
    // idle until the epilogue publishes the first quarter of A: the per-chunk critical path is then one barrier wait,
    // six MMAs and a commit.
    __device__ __forceinline__ void weights_ahead(int nch)
    {
        int st = stage, ph = phase;
        for (int j = 0; j < nch && j < NS; ++j) {
            mbar_wait(full + st, ph);
            if (++st == NS) { st = 0; ph ^= 1; }
        }
    }
    __device__ __forceinline__ void chunk(int N, int kA, int kB, int kk, uint32_t a_hi, uint32_t a_lo, uint32_t d, uint32_t idesc, bool first, bool last,
                                          bool waited)
    {
        using namespace bf;
        if (!waited) { mbar_wait(full + stage, phase); fence_after_sync(); }
        // descriptor of the chunk's first 16-byte k-chunk; the others are constant 16-byte-unit offsets from it
        const uint64_t b0 = sdesc(smem_u32(smem + stage * kStageBytes), N * 16, 128);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s * 16 < kk) {
                const uint32_t acol = (uint32_t)(s ? kB : kA) >> 1;
                const uint64_t b_hi = b0 + (uint64_t)((s * (2 * N * 16)) >> 4);
                const uint64_t b_lo = b0 + (uint64_t)((N * kk * 2 + s * (2 * N * 16)) >> 4);
                mma_h16_ts(d, a_lo + acol, b_hi, idesc, (first && s == 0) ? 0u : 1u);
                mma_h16_ts(d, a_hi + acol, b_lo, idesc, 1u);
                mma_h16_ts(d, a_hi + acol, b_hi, idesc, 1u);
            }
        }
        mma_commit_mcast(empty + stage, kClusterMask);  // this CTA is done with the stage: tell every producer
        if (last) mma_commit(mma_done);
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
    __device__ __forceinline__ void layer_rt(int N, int K, int acol0)
    {
        using namespace bf;
        const uint32_t idesc = idesc_h16(128, N);
        uint32_t t = tmem;
        asm volatile("" : "+r"(t));   // opaque: keeps the compiler from hoisting every layer's operand addresses out of the tile loop (64 registers here)
        const uint32_t a_hi = t + cAHI + acol0, a_lo = t + cALO + acol0, d = t + cD + (uses & 1) * kDCols;
        const uint32_t par = uses & 1;
        weights_ahead(K < 64 ? 1 : (K > 128 ? 5 : 4));
        if (K < 64) {
            // a single k-step written by the lead threads alone; every worker still arrives on all four barriers
#pragma unroll
            for (int j = 0; j < 4; ++j) mbar_wait(a_ready + j, par);
            fence_after_sync();
            chunk(N, 0, 0, 16, a_hi, a_lo, d, idesc, true, true, true);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                mbar_wait(a_ready + j, par);
                fence_after_sync();
                chunk(N, 16 * j, 16 * (j + 4), 32, a_hi, a_lo, d, idesc, j == 0, j == 3 && K == 128, j < NS);
            }
            if (K > 128) chunk(N, 128, 128, 16, a_hi, a_lo, d, idesc, false, true, 4 < NS);   // k >= 128: there since before a_ready[3]
        }
        ++uses;
    }
    template <int N, int K, int ACOL>
    __device__ __forceinline__ void layer() { layer_rt(N, K, ACOL); }   // everything folds: N, K are constants here
};

template <int KIND>
__global__ void __cluster_dims__(bf::kCluster, 1, 1) __launch_bounds__(bf::kThreads, 1)
k_field_bf(FieldParams p, const unsigned char *__restrict__ wstream)
{
    pdl_enter();
    using namespace bf;
    constexpr bool kHasFwd = KIND != kBwdSaved, kHasBwd = KIND == kBwdRecompute || KIND == kBwdSaved;
    constexpr int L0 = kHasFwd ? 0 : kLayersFwd, L1 = kHasBwd ? kLayersAll : kLayersFwd;   // layers [L0, L1) of the chain
    using SM = Smem<KIND>;
    constexpr int NS = SM::nStages;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::oBars);
    uint64_t *empty = full + kStages;
    uint64_t *a_ready = empty + kStages;   // [4]: quarter j of the next A operand (k-steps j and j + 4) is in tensor memory
    uint64_t *mma_done = a_ready + 4;
    uint64_t *st_full = mma_done + 1, *st_free = st_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + SM::oTmemPtr);
    float *sBias = reinterpret_cast<float *>(smem + SM::oBias);
    float *sW5 = reinterpret_cast<float *>(smem + SM::oW5), *sHead = reinterpret_cast<float *>(smem + SM::oHead);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    // The CTAs of a cluster consume one shared weight stream in lockstep, so they all run the same number of
    // tile iterations; iterations whose tile index is past the end are dummies (no valid rows, nothing stored).
    const int iters = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t crank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kCluster); }
        for (int i = 0; i < 4; ++i) mbar_init(a_ready + i, kWorkers);
        mbar_init(mma_done, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(st_full + i, kWorkers); mbar_init(st_free + i, 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    // activations / gradients go to the wgrad scratch (the stand-alone backward also runs without decoder gradients)
    const bool spill = p.wg_scratch != nullptr && p.spill_ops != 0;   // (kFwdSave / kBwdSaved without it: ReLU masks only, e.g. tracking)
    for (int i = threadIdx.x; i < 4 * 128 + 4; i += kThreads) {
        float v;
        if (i < 128) v = p.dec.b1[i];
        else if (i < 256) v = p.dec.b2[i - 128];
        else if (i < 384) v = p.dec.b3[1 + i - 256];
        else if (i < 512) v = p.dec.b4[i - 384];
        else if (i == 512) v = p.dec.b3[0];
        else v = p.dec.b5[i - 513];
        sBias[i] = i < 512 ? kScale * v : v;
    }
    for (int i = threadIdx.x; i < 3 * 128; i += kThreads) sW5[i] = p.dec.W5[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    cluster_sync();   // every CTA's barriers are initialised before any peer multicasts into them
    const uint32_t tmem = *tmem_ptr;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
    if (warp == 0) {
        // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
        // each CTA fetches its share of every chunk and multicasts it to the whole cluster
        int stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
            const unsigned char *src = wstream + (kHasFwd ? 0 : kFwdStreamBytes);
            for (int l = L0; l < L1; ++l) {
                const int N = cN[l], K = cK[l];
                if (l == 4 || l == 5) { src += N * K * 4; continue; }   // the colour head and its dgrad do not run on the tensor cores
                for (int k0 = 0; k0 < K; k0 += kChunkK) {
                    const int kk = (K - k0) < kChunkK ? (K - k0) : kChunkK;
                    const uint32_t bytes = (uint32_t)(N * kk * 4), part = bytes / kCluster;
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
        // ===================== MMA issuer: ONE thread, every layer's chunk loop unrolled with its N / K as constants.
        // The issue side is the critical path of a layer (6 MMAs of 64 clk per chunk against what it takes to wait for
        // the chunk, build 4 descriptors and commit), so nothing here is warp-wide (no elect / re-convergence per chunk)
        // and all descriptor arithmetic beyond "stage base + constant" folds at compile time. =====================
        if (lane == 0) {
            MmaIssuer<NS> mi{smem, full, empty, a_ready, mma_done, tmem, 0, 0, 0u};
            for (int it = 0; it < iters; ++it) {
                if constexpr (KIND == kBwdRecompute) {
                    for (int l = 0; l < kLayersAll; ++l) {
                        if (l == 4 || l == 5) continue;   // the colour head and its dgrad (W5, 3 x 128) run on the CUDA cores
                        BF_TRACE(it, l, 0); mi.layer_rt(cN[l], cK[l], cAcol[l]); BF_TRACE(it, l, 2);
                    }
                    continue;
                }
                if constexpr (kHasFwd) {
                    BF_TRACE(it, 0, 0); mi.template layer<128, 16, 64>(); BF_TRACE(it, 0, 2);
                    BF_TRACE(it, 1, 0); mi.template layer<128, 128, 0>(); BF_TRACE(it, 1, 2);
                    BF_TRACE(it, 2, 0); mi.template layer<144, 128, 0>(); BF_TRACE(it, 2, 2);
                    BF_TRACE(it, 3, 0); mi.template layer<128, 144, 0>(); BF_TRACE(it, 3, 2);
                }
                if constexpr (kHasBwd) {
                    BF_TRACE(it, 6, 0); mi.template layer<144, 128, 0>(); BF_TRACE(it, 6, 2);
                    BF_TRACE(it, 7, 0); mi.template layer<128, 144, 0>(); BF_TRACE(it, 7, 2);
                    BF_TRACE(it, 8, 0); mi.template layer<128, 128, 0>(); BF_TRACE(it, 8, 2);
                    BF_TRACE(it, 9, 0); mi.template layer<16, 128, 0>(); BF_TRACE(it, 9, 2);
                }
            }
        }
        __syncwarp();   // the warp must be converged again for the aligned cluster barrier at the end
    } else if (warp == 2) {
        // ===================== scratch store warp: staged operand (shared memory) -> wgrad scratch by bulk TMA =====================
        if (spill && !kDirectSpill) {
            const int ops[kBigOps] = {oH1, oH2, oHC, oG4, oG2, oG1};   // order in which the workers produce the operands
            constexpr int i0 = kHasFwd ? 0 : 3, i1 = kHasBwd ? kBigOps : 3;
            uint32_t sc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {   // dummy iterations store nothing
#pragma unroll
                for (int i = i0; i < i1; ++i, ++sc) {
                    const int b = sc & 1;
                    mbar_wait(st_full + b, (sc >> 1) & 1);
                    if (elect_one()) {
                        bulk_s2g(p.wg_scratch + (size_t)tile * kTileBytes + (size_t)ops[i] * kOpBytes, smem + SM::oStaging + b * kStagingBytes,
                                 (uint32_t)kStagingBytes);
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
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsWorker));
        // ===================== workers: two threads per sample row (64 accumulator columns each) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;             // which 64 columns
        const int col0 = half * 64;
        const bool lead = half == 0;                  // the row's thread that also gathers / scatters
        const int m = q * 32 + lane;                  // row of the tile
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        // this row's 16 bytes inside an operand: [half][plane][kb][fb][m % 8] (big operands: 16 fb, small: 2 fb)
        const int rowoff_big = (m >> 6) * 32768 + ((m >> 3) & 7) * 2048 + (m & 7) * 16;
        const int rowoff_small = (m >> 6) * 4096 + ((m >> 3) & 7) * 256 + (m & 7) * 16;
        uint32_t done_uses = 0;
        uint32_t nomask[2] = {0u, 0u};
        float ymax = 0.0f;                            // largest |operand| this thread produced (x16 activations, xSg gradients)
        uint32_t nlayers = 0;                         // layers consumed so far: which accumulator buffer the next layer_done() refers to
        uint32_t dcol = cD;                           // accumulator buffer of the layer just completed
        const float Sg = kHasBwd ? grad_scale(p.gscale) : 1.0f, invSg = 1.0f / Sg;
        if (kHasBwd && p.finish_zero) {                // clear the wgrad kernel's reduction block (one float4 per thread)
            for (int i = (int)blockIdx.x * (kThreads - 128) + (int)threadIdx.x - 128; i < (int)(kFinishFloats / 4); i += (int)gridDim.x * (kThreads - 128))
                reinterpret_cast<float4 *>(p.finish_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        int tile_i = 0, lcount = L0 - 1;              // trace bookkeeping
        uint32_t sc = 0;                              // staged operands so far (two staging buffers alternate)
        bool real_tile = true;
        int tile_cur = 0;
        const int spill_ops[kBigOps] = {oH1, oH2, oHC, oG4, oG2, oG1};   // order in which the operands are produced
        auto stage_begin = [&]() -> unsigned char * {
            if (!spill || !real_tile) return nullptr;
            if (kDirectSpill) {
                const int op = spill_ops[(kHasFwd ? 0 : 3) + (int)(sc % (uint32_t)((kHasFwd ? 3 : 0) + (kHasBwd ? 3 : 0)))];
                return p.wg_scratch + (size_t)tile_cur * kTileBytes + (size_t)op * kOpBytes + rowoff_big;
            }
            const int b = sc & 1;
            if (sc >= 2) mbar_wait(st_free + b, ((sc >> 1) - 1) & 1);
            return smem + SM::oStaging + b * kStagingBytes + rowoff_big;
        };
        auto stage_end = [&]() {
            if (!spill || !real_tile) return;
            if (kDirectSpill) { ++sc; return; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> bulk-copy engine
            mbar_arrive(st_full + (sc & 1));
            ++sc;
        };
        auto layer_done = [&]() {
            mbar_wait(mma_done, done_uses & 1);
            ++done_uses;
            dcol = cD + (nlayers & 1) * kDCols;
            ++nlayers;
            fence_after_sync();
            if (threadIdx.x == 128) BF_TRACE(tile_i, lcount, 3);     // worker sees the accumulators
        };
        auto a_quarter_ready = [&](int j) {           // quarter j of the next A operand (k-steps j and j + 4) is written
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(a_ready + j);
        };
        auto a_is_ready = [&]() {                     // ... the last quarter, and anything the lead thread adds (k >= 128)
            tmem_wait_st();
            fence_before_sync();
            if (threadIdx.x == 128) BF_TRACE(tile_i, lcount + 1, 5);  // worker has produced the A of layer lcount+1
            mbar_arrive(a_ready + 3);
            ++lcount;
        };
        auto a_small_ready = [&]() {                  // a K = 16 layer reads only what the lead thread wrote: all quarters at once
            a_quarter_ready(0); mbar_arrive(a_ready + 1); mbar_arrive(a_ready + 2);
            a_is_ready();
        };
        // epilogue of one layer in four 16-column batches, each followed by its quarter's signal (the last one by the caller)
        auto epilogue = [&](auto mode, const float *bias, uint32_t (&mask)[2], bool staged) {
            constexpr int MODE = decltype(mode)::value;
            unsigned char *stg = staged ? stage_begin() : nullptr;
            if (threadIdx.x == 128) BF_TRACE(tile_i, lcount, 7);       // staging buffer is free
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bf_epilogue16<MODE>(trow, dcol, col0 + 16 * j, bias, mask[j >> 1], (j & 1) * 16, stg, ymax);
                if (j < 3) a_quarter_ready(j);
            }
            if (staged) stage_end();
            if (threadIdx.x == 128) BF_TRACE(tile_i, lcount, 4);       // epilogue done (operand staged)
        };
        using M0 = std::integral_constant<int, 0>; using M1 = std::integral_constant<int, 1>;
        using M2 = std::integral_constant<int, 2>; using M3 = std::integral_constant<int, 3>;
        // Software prefetch of the NEXT tile's per-row inputs (feature rows for the forward; ReLU masks, rgb and upstream gradient
        // for the saved backward): issued a few layers before the tile ends, so their L2 / HBM latency hides behind the MMAs
        // instead of sitting between two tiles.
        float4 pf[4];
        uint32_t pm[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        float4 po = make_float4(0.f, 0.f, 0.f, 0.f), pgo = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) pf[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto prefetch_tile = [&](int tn) {
            const int sn = tn * 128 + m;
            const bool in = tn < ntiles && sn < nsamp;
            if (kHasFwd && p.feat && lead) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    pf[e] = in ? __ldg(reinterpret_cast<const float4 *>(p.feat + (size_t)sn * 16) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (KIND == kBwdSaved) {
                const uint32_t *mk = p.act_masks + (size_t)(tn < ntiles ? tn : 0) * (kMaskBytes / 4) + half * 256 + m;
                pm[0] = mk[0]; pm[1] = mk[128]; pm[2] = mk[512]; pm[3] = mk[512 + 128]; pm[4] = mk[1024]; pm[5] = mk[1024 + 128];
                po = in ? __ldg(reinterpret_cast<const float4 *>(p.out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                pgo = in ? __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch_tile((int)blockIdx.x);
        for (int it = 0; it < iters; ++it, ++tile_i, lcount = L0 - 1) {
            const int tile = blockIdx.x + it * gridDim.x;
            tile_cur = tile;
            real_tile = tile < ntiles;
            const int s = real_tile ? tile * 128 + m : nsamp;      // rows of a dummy iteration are all out of range
            unsigned char *scr = nullptr;   // this tile's wgrad scratch (the two small operands go direct)
            if (spill && real_tile) scr = p.wg_scratch + (size_t)tile * kTileBytes;
            int vox = -1, ray = -1;
            float z = 0.0f, px = 0.f, py = 0.f, pz = 0.f;
            if (threadIdx.x == 128) BF_TRACE(tile_i, 0, 6);            // gather starts
            uint32_t m1[2], m2[2], mc[2];
            float sdf = 0.0f, r = 0.f, g = 0.f, b = 0.f;
            // ---- features -> A[:, 128:144) (the lead thread of each row); kBwdSaved only needs the sample's position ----
            if (lead) {
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = 0.0f;
                if (s < nsamp) {
                    if (p.feat) {
                        if (kHasFwd) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { f[4 * e] = pf[e].x; f[4 * e + 1] = pf[e].y; f[4 * e + 2] = pf[e].z; f[4 * e + 3] = pf[e].w; }
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
                        if (kHasFwd) {
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
                }
                uint32_t hi[8], lo[8];
                if (kHasFwd) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) ymax = fmaxf(ymax, fabsf(kScale * f[e]));
#pragma unroll
                    for (int e = 0; e < 8; ++e) h16_split2(kScale * f[2 * e], kScale * f[2 * e + 1], hi[e], lo[e]);
                    tmem_st8(trow + cAHI + 64, hi);
                    tmem_st8(trow + cALO + 64, lo);
                }
                if (kHasFwd && scr) {
                    unsigned char *dst = scr + oF + rowoff_small;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        *reinterpret_cast<uint4 *>(dst + j * 128) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<uint4 *>(dst + 2048 + j * 128) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                }
            }
            if constexpr (kHasFwd) {
            a_small_ready();  // layer 1 (K = 16) reads only the features the lead thread just wrote
            // ---- forward ----
            layer_done();
            epilogue(M0{}, sBias, m1, true);                                                            // h1
            a_is_ready();
            layer_done();
            epilogue(M0{}, sBias + 128, m2, true);                                                      // h2
            prefetch_tile(tile + (int)gridDim.x);
            a_is_ready();
            layer_done();
            epilogue(M1{}, sBias + 256, nomask, false);                                                 // t (no activation; not spilled)
            if (lead) {
                uint32_t v[8];
                tmem_ld8(trow + dcol + 128, v);   // sdf = row 0 of W3, packed as output column 128
                tmem_wait_ld();
                sdf = fmaf(__uint_as_float(v[0]), kInvScale * kInvScale, sBias[512]);
            }
            a_is_ready();
            layer_done();
            {
                // hc + colour head: no tensor-core layer follows in the forward direction, so nothing is handed to the MMA thread
                float head[3] = {0.f, 0.f, 0.f};
                unsigned char *stg = stage_begin();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    bf_epilogue16<0, true>(trow, dcol, col0 + 16 * j, sBias + 384, mc[j >> 1], (j & 1) * 16, stg, ymax, sW5, head);
                stage_end();
                sHead[m * 8 + half * 4] = head[0]; sHead[m * 8 + half * 4 + 1] = head[1]; sHead[m * 8 + half * 4 + 2] = head[2];
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");    // the two warps of this lane quarter
                {   // both threads of the row (the backward takes g_hc on the CUDA cores too); fixed order of the two partial sums; y was 16 x hc
                    const float *h0 = sHead + m * 8, *h1 = h0 + 4;
                    r = sigmoid_f(fmaf(h0[0] + h1[0], kInvScale, sBias[513]));
                    g = sigmoid_f(fmaf(h0[1] + h1[1], kInvScale, sBias[514]));
                    b = sigmoid_f(fmaf(h0[2] + h1[2], kInvScale, sBias[515]));
                }
                if (threadIdx.x == 128) BF_TRACE(tile_i, lcount, 4);
            }
            }   // kHasFwd
            if constexpr (!kHasBwd) {
                if (lead && s < nsamp) *reinterpret_cast<float4 *>(p.out + (size_t)s * 4) = make_float4(r, g, b, sdf);
                if (KIND == kFwdSave && real_tile) {
                    // ReLU masks of this thread's 64 columns for the dgrad kernel that follows
                    uint32_t *mk = p.act_masks + (size_t)tile * (kMaskBytes / 4) + half * 256 + m;
                    mk[0] = m1[0]; mk[128] = m1[1]; mk[512] = m2[0]; mk[512 + 128] = m2[1]; mk[1024] = mc[0]; mk[1024 + 128] = mc[1];
                }
                continue;   // D has been read; the next tile's first MMA is ordered behind it through a_ready
            }
            if constexpr (KIND == kBwdSaved) {
                m1[0] = pm[0]; m1[1] = pm[1]; m2[0] = pm[2]; m2[1] = pm[3]; mc[0] = pm[4]; mc[1] = pm[5];
                if (s < nsamp) { r = po.x; g = po.y; b = po.z; }
            }
            // ---- backward ----
            float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < nsamp) go = (KIND == kBwdSaved) ? pgo : __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)s * 4));
            go.x *= Sg; go.y *= Sg; go.z *= Sg; go.w *= Sg;     // the whole chain is linear in g_out: it carries Sg to the end
            // dL/d(pre-sigmoid rgb): grad * (1 - y) * y   (both threads of the row: each needs it for its half of g_hc)
            const float g5[4] = {go.x * (1.0f - r) * r, go.y * (1.0f - g) * g, go.z * (1.0f - b) * b, go.w};
            if (lead) {
                uint32_t hi[8], lo[8];
                h16_split2(g5[0], g5[1], hi[0], lo[0]);
                if (scr) {
                    // wgrad operand G5 = (g5 r, g, b, g_sdf, 0 ...)
                    uint32_t h1w, l1w;
                    h16_split2(g5[2], g5[3], h1w, l1w);
                    unsigned char *dst = scr + oG5 + rowoff_small;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], h1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048) = make_uint4(lo[0], l1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048 + 128) = make_uint4(0u, 0u, 0u, 0u);
                    // bias gradients of the two heads: column sums of G5 (k_wgrad_bf takes the others on the tensor cores)
                    const float s0 = warp_sum(g5[0]), s1 = warp_sum(g5[1]), s2 = warp_sum(g5[2]), s3 = warp_sum(g5[3]);
                    if (lane == 0) {
                        atomicAdd(p.g_dec.b5 + 0, s0 * invSg); atomicAdd(p.g_dec.b5 + 1, s1 * invSg); atomicAdd(p.g_dec.b5 + 2, s2 * invSg);
                        atomicAdd(p.g_dec.b3, s3 * invSg);
                    }
                }
            }
            {
                // g_hc = mask_hc . (W5^T g5): three FMAs per output, taken here instead of a K = 16 tensor-core layer and its round trip
                unsigned char *stg = stage_begin();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c0 = col0 + 16 * j;
                    const uint32_t bits = mc[j >> 1] >> ((j & 1) * 16);
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float y0 = fmaf(g5[2], sW5[256 + c0 + 2 * e], fmaf(g5[1], sW5[128 + c0 + 2 * e], g5[0] * sW5[c0 + 2 * e]));
                        float y1 = fmaf(g5[2], sW5[256 + c0 + 2 * e + 1], fmaf(g5[1], sW5[128 + c0 + 2 * e + 1], g5[0] * sW5[c0 + 2 * e + 1]));
                        y0 = ((bits >> (2 * e)) & 1u) ? y0 : 0.0f;
                        y1 = ((bits >> (2 * e + 1)) & 1u) ? y1 : 0.0f;
                        ymax = fmaxf(ymax, fmaxf(fabsf(y0), fabsf(y1)));
                        h16_split2(y0, y1, hi[e], lo[e]);
                    }
                    if (stg) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            unsigned char *dst = stg + (size_t)(c0 / 8 + k) * 128;
                            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                            *reinterpret_cast<uint4 *>(dst + 16384) = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                        }
                    }
                    tmem_st8(trow + cAHI + c0 / 2, hi);
                    tmem_st8(trow + cALO + c0 / 2, lo);
                    if (j < 3) a_quarter_ready(j);
                }
                stage_end();
            }
            a_is_ready();
            layer_done();
            epilogue(M3{}, nullptr, nomask, false);                                                     // g_t (not spilled)
            float gf[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) gf[e] = 0.0f;
            if (lead) {
                uint32_t v[16], hi[8], lo[8];
                tmem_ld16(trow + dcol + 128, v);   // g_f, part through W4's last 16 input columns
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) gf[e] = __uint_as_float(v[e]);      // 16 x Sg x g_f (first part)
#pragma unroll
                for (int e = 0; e < 8; ++e) { hi[e] = 0u; lo[e] = 0u; }
                h16_split2(go.w, 0.0f, hi[0], lo[0]);  // A[:, 128] = g_sdf pairs with W3 row 0 (packed at k = 128)
                tmem_st8(trow + cAHI + 64, hi);
                tmem_st8(trow + cALO + 64, lo);
            }
            a_is_ready();
            layer_done();
            epilogue(M2{}, nullptr, m2, true);                                                          // g_h2
            if (KIND == kBwdSaved) prefetch_tile(tile + (int)gridDim.x);
            a_is_ready();
            layer_done();
            epilogue(M2{}, nullptr, m1, true);                                                          // g_h1
            a_is_ready();
            layer_done();
            if (!lead) continue;
            {
                uint32_t v[16];
                tmem_ld16(trow + dcol, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) gf[e] = (gf[e] + __uint_as_float(v[e])) * (kInvScale * invSg);   // unscaled g_f
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
        // f16 saturates at 65504: an operand that reached the top of the window (|activation| or |weight| >= 2047, a
        // gradient chain that grew 64x beyond max |g_out|) is reported instead of silently clipped
        if (p.range_flag && ymax >= 32752.0f) atomicOr(p.range_flag, 4);
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync();   // no CTA leaves while a peer may still multicast into its shared memory or signal its barriers
    if (warp == 0) tmem_dealloc(tmem, bf::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Kernel 3 split off the tensor-core kernels for the fused pipeline.  The trilinear lookup is three dependent
// gathers per sample (sample -> voxel -> 8 corner rows -> 8 x 64 B) and its backward 32 vector reductions per
// sample; inside k_field_bf they ran on the 128 lead threads with nothing to overlap them, between two tiles'
// MMAs (traces: 6k of a 21k-clock forward tile, 19k of a 33k-clock backward tile).  As kernels of their own they
// have the whole GPU's memory-level parallelism (4 threads per sample, one 16-byte quarter of the feature row each)
// and the tensor-core kernels read / write plain [P,16] rows (64 B per sample, L2-resident at the mapping sizes).
// Arithmetic and corner order are those of the fused path, so the features are bit-identical.
// ------------------------------------------------------------------------------------------
// A thread quad (one 16-byte quarter of the feature row per thread) walks kTriChunk CONSECUTIVE samples.  Samples are in
// CSR order by ray and sorted by depth, so ~7 neighbours sit in the same voxel: the quad fetches the voxel's corner ids and
// rows once per voxel change (the scatter aggregates the same way, across the lanes of a warp: trilinear.cuh).
constexpr int kTriChunk = 8;

__global__ void __launch_bounds__(256) k_tri_gather(FieldParams p, float *__restrict__ feat)
{
    pdl_enter();
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s0 = (t >> 2) * kTriChunk, c = t & 3;
    if (s0 >= nsamp) return;
    int cur_vox = -1;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int k = 0; k < kTriChunk; ++k) {
        const int s = s0 + k;
        if (s >= nsamp) break;
        int vox, ray;
        float z, px, py, pz;
        sample_position(p, s, vox, ray, z, px, py, pz);
        if (vox != cur_vox) {
            cur_vox = vox;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = __ldg(p.vertex_idx + (size_t)vox * 8 + i);
                v[i] = __ldg(reinterpret_cast<const float4 *>(p.emb + (size_t)row * 16 + c * 4));
            }
        }
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float wx = (i & 4) ? px : 1.0f - px, wy = (i & 2) ? py : 1.0f - py, wz = (i & 1) ? pz : 1.0f - pz;
            const float w = (wx * wy) * wz;
            f.x = fmaf(w, v[i].x, f.x); f.y = fmaf(w, v[i].y, f.y); f.z = fmaf(w, v[i].z, f.z); f.w = fmaf(w, v[i].w, f.w);
        }
        *reinterpret_cast<float4 *>(feat + (size_t)s * 16 + c * 4) = f;
    }
}

// Backward of the lookup with warp-aggregated reductions: tri_scatter_warp (trilinear.cuh), 8 warps of 32 samples per block.
__global__ void __launch_bounds__(kScatWarps * 32) k_tri_scatter(FieldParams p, const float *__restrict__ g_feat)
{
    pdl_enter();
    __shared__ __align__(16) float s_buf[kScatWarps][kScatWarpFloats];
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int warp = threadIdx.x >> 5;
    const int s0 = (blockIdx.x * kScatWarps + warp) * 32;
    if (s0 >= nsamp) return;                                  // whole warps leave
    tri_scatter_warp<false>(p, g_feat, s0, nsamp, s_buf[warp]);
}

// ------------------------------------------------------------------------------------------
// Weight gradients: dW[n][k] = sum over samples of G[p][n] * A[p][k] -- MMAs whose reduction dimension is
// the SAMPLE index.  The scratch holds every operand already split and in the MN-major core-matrix layout,
// so a stage is 2 (+2 small) bulk-TMA copies and the MMAs read them in place (A and B both MN-major from
// shared memory).  Accumulators stay in tensor memory for ALL tiles of the CTA:
//   [0,128) dW2 = G2^T H1      [128,256) M = G4^T H2 (-> dW3, dW4[:, :128] in k_wgrad_finish)
//   [256,272) dW4[:, 128:] = G4^T F    [272,288) dW1 = G1^T F    [288,304) HC^T G5 (cols 0..2 = dW5 rows)
//   [304,320) H2^T G5 (col 3 = dW3 row 0)    [320,368) column sums of G2, G4, G1 (x ones: bias gradients)
// Steps per (tile, 64-sample half): 0: (G2, H1)   1: (G4, H2) + F + G5   2: (G1, HC) + F + G5
// ------------------------------------------------------------------------------------------
namespace wgb {
constexpr int kThreads = 192;                   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 drain
constexpr int kStages = 3;
constexpr int kStepsPerHalf = 3;
constexpr int kHalfBytes = 32768;               // (128-feature operand, 64-sample half): [plane 2][kb 8][fb 16][128 B]
constexpr int kSmallHalf = 4096;                // (16-feature operand, half):           [plane 2][kb 8][fb 2][128 B]
constexpr int oB = kHalfBytes, oFs = 2 * kHalfBytes, oG5s = oFs + kSmallHalf;
constexpr int kStageBytes = oG5s + kSmallHalf;  // 73 728
constexpr int oOnes = kStages * kStageBytes;    // 16 samples x 16 features of bf16 1.0: [kb 2][fb 2][128 B]
constexpr int oBars = oOnes + 512;              // full[3] free[3] all_done, tmem ptr
constexpr int kSmemBytes = oBars + 128;
constexpr int cW2 = 0, cM = 128, cW4f = 256, cW1 = 272, cHCG5 = 288, cH2G5 = 304, cS2 = 320, cS4 = 336, cS1 = 352;
struct Step { int a_op, b_op, use_small; };
__device__ __constant__ Step cSteps[kStepsPerHalf] = {{bf::oG2, bf::oH1, 0}, {bf::oG4, bf::oH2, 1}, {bf::oG1, bf::oHC, 1}};
}  // namespace wgb

#define WGB_TRACE(g, slot)                                                                          \
    do {                                                                                            \
        if (g_bf_trace && blockIdx.x == 0 && (g) < 40) g_bf_trace[320 + (g) * 8 + (slot)] = clock64();     \
    } while (0)

__global__ void __launch_bounds__(wgb::kThreads, 1) k_wgrad_bf(FieldParams p, float *__restrict__ finish)
{
    pdl_enter();
    using namespace wgb;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *freeb = full + kStages, *all_done = freeb + kStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oBars + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(freeb + i, 1); }
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    for (int i = tid; i < 128; i += kThreads) reinterpret_cast<uint32_t *>(smem + oOnes)[i] = 0x3C003C00u;   // f16 1.0 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_tiles * 2 * kStepsPerHalf;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int tl = 0, h = 0, st = 0;
        for (int g = 0; g < nsteps; ++g) {
            const int rs = g % kStages, use = g / kStages;
            const Step S = cSteps[st];
            const unsigned char *tile = p.wg_scratch + (size_t)(blockIdx.x + (size_t)tl * gridDim.x) * bf::kTileBytes;
            if (use >= 1) mbar_wait(freeb + rs, (use - 1) & 1);
            if (lane == 0) WGB_TRACE(g, 0);
            if (elect_one()) {
                unsigned char *dst = smem + rs * kStageBytes;
                mbar_arrive_expect_tx(full + rs, (uint32_t)(2 * kHalfBytes + S.use_small * 2 * kSmallHalf));
                bulk_g2s(dst, tile + (size_t)S.a_op * bf::kOpBytes + (size_t)h * kHalfBytes, kHalfBytes, full + rs);
                bulk_g2s(dst + oB, tile + (size_t)S.b_op * bf::kOpBytes + (size_t)h * kHalfBytes, kHalfBytes, full + rs);
                if (S.use_small) {
                    bulk_g2s(dst + oFs, tile + bf::oF + (size_t)h * kSmallHalf, kSmallHalf, full + rs);
                    bulk_g2s(dst + oG5s, tile + bf::oG5 + (size_t)h * kSmallHalf, kSmallHalf, full + rs);
                }
            }
            __syncwarp();
            if (++st == kStepsPerHalf) { st = 0; if (++h == 2) { h = 0; ++tl; } }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t id128 = idesc_h16(128, 128, 1, 1), id16 = idesc_h16(128, 16, 1, 1);
        const uint64_t ones = sdesc(smem_u32(smem + oOnes), 256, 128);
        int st = 0;
        for (int g = 0; g < nsteps; ++g) {
            const int rs = g % kStages;
            mbar_wait(full + rs, (g / kStages) & 1);
            fence_after_sync();
            if (lane == 0) WGB_TRACE(g, 4);
            const uint32_t base = smem_u32(smem + rs * kStageBytes);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {            // 16 samples = 2 kb blocks per MMA
                    const uint32_t fresh = (g < kStepsPerHalf && ks == 0) ? 0u : 1u;   // first touch of this step's accumulators
                    const uint64_t a_hi = sdesc(base + ks * 4096, 2048, 128), a_lo = sdesc(base + 16384 + ks * 4096, 2048, 128);
                    const uint64_t b_hi = sdesc(base + oB + ks * 4096, 2048, 128), b_lo = sdesc(base + oB + 16384 + ks * 4096, 2048, 128);
                    const uint64_t f_hi = sdesc(base + oFs + ks * 512, 256, 128), f_lo = sdesc(base + oFs + 2048 + ks * 512, 256, 128);
                    const uint64_t g_hi = sdesc(base + oG5s + ks * 512, 256, 128), g_lo = sdesc(base + oG5s + 2048 + ks * 512, 256, 128);
                    auto prod3 = [&](uint32_t dcol, uint64_t xh, uint64_t xl, uint64_t yh, uint64_t yl, uint32_t idesc) {
                        mma_h16_ss(tmem + dcol, xl, yh, idesc, fresh);
                        mma_h16_ss(tmem + dcol, xh, yl, idesc, 1u);
                        mma_h16_ss(tmem + dcol, xh, yh, idesc, 1u);
                    };
                    auto colsum = [&](uint32_t dcol, uint64_t xh, uint64_t xl) {
                        mma_h16_ss(tmem + dcol, xh, ones, id16, fresh);
                        mma_h16_ss(tmem + dcol, xl, ones, id16, 1u);
                    };
                    if (st == 0) {
                        prod3(cW2, a_hi, a_lo, b_hi, b_lo, id128);          // dW2 = G2^T H1
                        colsum(cS2, a_hi, a_lo);                            // db2
                    } else if (st == 1) {
                        prod3(cM, a_hi, a_lo, b_hi, b_lo, id128);           // M = G4^T H2
                        prod3(cW4f, a_hi, a_lo, f_hi, f_lo, id16);          // dW4[:, 128:] = G4^T F
                        prod3(cH2G5, b_hi, b_lo, g_hi, g_lo, id16);         // H2^T G5 (col 3 -> dW3[0])
                        colsum(cS4, a_hi, a_lo);                            // db4
                    } else {
                        prod3(cW1, a_hi, a_lo, f_hi, f_lo, id16);           // dW1 = G1^T F
                        prod3(cHCG5, b_hi, b_lo, g_hi, g_lo, id16);         // HC^T G5 (cols 0..2 -> dW5)
                        colsum(cS1, a_hi, a_lo);                            // db1
                    }
                }
                mma_commit(freeb + rs);
                if (g == nsteps - 1) mma_commit(all_done);
            }
            __syncwarp();
            if (lane == 0) WGB_TRACE(g, 5);
            if (++st == kStepsPerHalf) st = 0;
        }
    } else if (nsteps > 0) {
        // ===================== drain: accumulators -> global gradients (warp & 3 = TMEM lane quarter) =====================
        mbar_wait(all_done, 0);
        fence_after_sync();
        const int qd = warp & 3;
        const int n = qd * 32 + lane;                            // accumulator row = TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
        // accumulators hold 16 x Sg x (sum of products), the column sums Sg x (sum): both factors are powers of two
        const float invSg = 1.0f / grad_scale(p.gscale), cW = bf::kInvScale * invSg;
        auto flush = [&](int col0, int ncols, float *dst_row) {   // dst_row: &dW[n][0], ncols % 16 == 0
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + col0 + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    red_add_v4(dst_row + c0 + 4 * j, cW * __uint_as_float(v[4 * j]), cW * __uint_as_float(v[4 * j + 1]),
                               cW * __uint_as_float(v[4 * j + 2]), cW * __uint_as_float(v[4 * j + 3]));
            }
        };
        flush(cW2, 128, p.g_dec.W2 + (size_t)n * 128);
        flush(cM, 128, finish + (size_t)n * 128);
        flush(cW4f, 16, p.g_dec.W4 + (size_t)n * 144 + 128);
        flush(cW1, 16, p.g_dec.W1 + (size_t)n * 16);
        {
            uint32_t v[8], w[8], bs[8];
            tmem_ld8(trow + cHCG5, v);
            tmem_ld8(trow + cH2G5, w);
            tmem_wait_ld();
            atomicAdd(p.g_dec.W5 + n, cW * __uint_as_float(v[0]));
            atomicAdd(p.g_dec.W5 + 128 + n, cW * __uint_as_float(v[1]));
            atomicAdd(p.g_dec.W5 + 256 + n, cW * __uint_as_float(v[2]));
            atomicAdd(p.g_dec.W3 + n, cW * __uint_as_float(w[3]));
            tmem_ld8(trow + cS2, bs); tmem_wait_ld(); atomicAdd(p.g_dec.b2 + n, invSg * __uint_as_float(bs[0]));
            tmem_ld8(trow + cS4, bs); tmem_wait_ld(); atomicAdd(finish + 128 * 128 + n, invSg * __uint_as_float(bs[0]));
            tmem_ld8(trow + cS1, bs); tmem_wait_ld(); atomicAdd(p.g_dec.b1 + n, invSg * __uint_as_float(bs[0]));
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// The gradients that go through M = G4^T H2 [128,128] and s4 = column sums of G4 (see the scratch layout):
//   blocks 0..127   (j):  dW3[1+j][k] += sum_n W4[n][j] M[n][k]          db3[1+j] += sum_n W4[n][j] s4[n]
//   blocks 128..255 (n):  dW4[n][j]   += sum_k M[n][k] W3[1+j][k] + s4[n] b3[1+j]          db4[n] += s4[n]
// Each block first pulls the whole 128 x 128 matrix it contracts with (M, resp. W3's feature rows) into shared
// memory with all its loads in flight at once; row pitch 129 keeps the second form's column walk conflict-free.
constexpr int kFinishPitch = 129;
constexpr int kFinishSmem = 128 * kFinishPitch * 4;
__global__ void __launch_bounds__(128) k_wgrad_finish(FieldParams p, const float *__restrict__ finish)
{
    pdl_enter();
    extern __shared__ float sMat[];   // [128][129]
    __shared__ float sVec[128], sRed[4];
    const int b = blockIdx.x & 127, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const float *M = finish, *s4 = finish + 128 * 128;
    const bool first = blockIdx.x < 128;
    const float *src = first ? M : p.dec.W3 + 128;      // W3 rows 1..128
#pragma unroll 8
    for (int i = t; i < 128 * 32; i += 128) {           // 128 rows x 32 float4
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + i);
        float *dst = sMat + (i >> 5) * kFinishPitch + (i & 31) * 4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    sVec[t] = first ? p.dec.W4[(size_t)t * 144 + b] : M[(size_t)b * 128 + t];   // W4[n = t][j = b]  /  M[n = b][k = t]
    __syncthreads();
    if (first) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // thread t = column k of M
#pragma unroll 8
        for (int n = 0; n < 128; n += 4) {
            a0 = fmaf(sVec[n], sMat[n * kFinishPitch + t], a0);
            a1 = fmaf(sVec[n + 1], sMat[(n + 1) * kFinishPitch + t], a1);
            a2 = fmaf(sVec[n + 2], sMat[(n + 2) * kFinishPitch + t], a2);
            a3 = fmaf(sVec[n + 3], sMat[(n + 3) * kFinishPitch + t], a3);
        }
        atomicAdd(p.g_dec.W3 + (size_t)(1 + b) * 128 + t, (a0 + a1) + (a2 + a3));
        const float v = warp_sum(sVec[t] * s4[t]);        // db3[1 + b] = sum_n W4[n][b] s4[n]
        if (lane == 0) sRed[warp] = v;
        __syncthreads();
        if (t == 0) atomicAdd(p.g_dec.b3 + 1 + b, (sRed[0] + sRed[1]) + (sRed[2] + sRed[3]));
    } else {
        const float sb = s4[b];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // thread t = output j: walks row t of W3's feature rows
#pragma unroll 8
        for (int k = 0; k < 128; k += 4) {
            a0 = fmaf(sVec[k], sMat[t * kFinishPitch + k], a0);
            a1 = fmaf(sVec[k + 1], sMat[t * kFinishPitch + k + 1], a1);
            a2 = fmaf(sVec[k + 2], sMat[t * kFinishPitch + k + 2], a2);
            a3 = fmaf(sVec[k + 3], sMat[t * kFinishPitch + k + 3], a3);
        }
        atomicAdd(p.g_dec.W4 + (size_t)b * 144 + t, ((a0 + a1) + (a2 + a3)) + sb * p.dec.b3[1 + t]);
        if (t == 0) atomicAdd(p.g_dec.b4 + b, sb);
    }
}

// ------------------------------------------------------------------------------------------
// stand-alone GEMMs through the same primitives (unit tests of descriptors / TMEM packing), 3xBF16.
// mode 0: D[128,N] = A[128,K] * B[N,K]^T, A packed in tensor memory, B K-major in shared memory (the chain form)
// mode 1: D[128,N] = At[K,128]^T * Bt[K,N], both operands MN-major in shared memory (the wgrad form)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_debug_umma_bf(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D,
                                                          int N, int K, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint16_t *s16 = reinterpret_cast<uint16_t *>(smem);
    const int warp = threadIdx.x >> 5, m = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    const int nkb = K / 8;
    const int planeA = nkb * 16 * 64, planeB = nkb * (N / 8) * 64;   // uint16 elements of one MN-major plane
    if (mode == 0) {
        // B -> per 16-k step: [hi: 2 k-chunks x N rows x 8][lo: same]
        for (int i = threadIdx.x; i < N * K; i += 128) {
            const int n = i / K, k = i % K, st = k >> 4, kc = (k >> 3) & 1, e = k & 7;
            uint32_t hi, lo;
            h16_split2(B[i], 0.0f, hi, lo);
            uint16_t *blk = s16 + st * (2 * N * 16);
            blk[kc * N * 8 + n * 8 + e] = (uint16_t)hi;
            blk[N * 16 + kc * N * 8 + n * 8 + e] = (uint16_t)lo;
        }
    } else {
        // At [K][128] -> A planes [kb][fb 16][8 samples][8 features]; Bt [K][N] -> B planes [kb][fb N/8][8][8]
        for (int i = threadIdx.x; i < K * (128 + N); i += 128) {
            const bool isA = i < K * 128;
            const int j = isA ? i : i - K * 128, F = isA ? 128 : N;
            const int k = j / F, f = j % F;
            uint32_t hi, lo;
            h16_split2(isA ? A[j] : B[j], 0.0f, hi, lo);
            uint16_t *pl = s16 + (isA ? 0 : 2 * planeA);
            const int off = (k >> 3) * (F / 8) * 64 + (f >> 3) * 64 + (k & 7) * 8 + (f & 7);
            pl[off] = (uint16_t)hi;
            pl[(isA ? planeA : planeB) + off] = (uint16_t)lo;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_ptr;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    if (mode == 0) {
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h16_split2(A[(size_t)m * K + k0 + 2 * e], A[(size_t)m * K + k0 + 2 * e + 1], hi[e], lo[e]);
            tmem_st8(trow + k0 / 2, hi);
            tmem_st8(trow + 72 + k0 / 2, lo);
        }
        tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after_sync();
        const uint32_t sb = smem_u32(smem);
        if (mode == 0) {
            const uint32_t id_lh = idesc_h16(128, N), id_hl = id_lh, id_hh = id_lh;
            for (int st = 0; st < K / 16; ++st) {
                const uint64_t b_hi = sdesc(sb + st * (2 * N * 16 * 2), N * 16, 128);
                const uint64_t b_lo = sdesc(sb + st * (2 * N * 16 * 2) + N * 16 * 2, N * 16, 128);
                mma_h16_ts(tmem + 144, tmem + 72 + st * 8, b_hi, id_lh, st ? 1u : 0u);
                mma_h16_ts(tmem + 144, tmem + st * 8, b_lo, id_hl, 1u);
                mma_h16_ts(tmem + 144, tmem + st * 8, b_hi, id_hh, 1u);
            }
        } else {
            const uint32_t id_lh = idesc_h16(128, N, 1, 1), id_hl = id_lh, id_hh = id_lh;
            const uint32_t lboA = 16 * 128, lboB = (N / 8) * 128;
            const uint32_t a_hi = sb, a_lo = sb + planeA * 2, b_hi = sb + planeA * 4, b_lo = b_hi + planeB * 2;
            for (int st = 0; st < K / 16; ++st) {
                mma_h16_ss(tmem + 144, sdesc(a_lo + st * 2 * lboA, lboA, 128), sdesc(b_hi + st * 2 * lboB, lboB, 128), id_lh, st ? 1u : 0u);
                mma_h16_ss(tmem + 144, sdesc(a_hi + st * 2 * lboA, lboA, 128), sdesc(b_lo + st * 2 * lboB, lboB, 128), id_hl, 1u);
                mma_h16_ss(tmem + 144, sdesc(a_hi + st * 2 * lboA, lboA, 128), sdesc(b_hi + st * 2 * lboB, lboB, 128), id_hh, 1u);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + 144 + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) D[(size_t)m * N + c0 + e] = __uint_as_float(v[e]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
int bf_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st, int *range_flag)
{
    int total = 0;
    for (int l = 0; l < bf::kLayersPack; ++l) total += bf::hN(l) * bf::hK(l);
    launch_chain(k_bf_pack, dim3(ceil_div(total, 256)), dim3(256), 0, st, d, reinterpret_cast<uint16_t *>(ws_tc), range_flag);
    PSLAM_CHECK_LAUNCH("bf_pack");
    return 0;
}

// scratch = [tiles x kTileBytes operands][kFinishFloats][tiles x kMaskBytes ReLU masks][tiles x 8 kB features][tiles x 8 kB feature gradients]
constexpr size_t kFeatTileBytes = 128 * 16 * sizeof(float);
size_t bf_wgrad_scratch_bytes(int max_samples)
{
    return (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128) * (bf::kTileBytes + bf::kMaskBytes + 2 * kFeatTileBytes) +
           bf::kFinishFloats * sizeof(float);
}
static float *scratch_feat(const FieldParams &fp, int max_samples, int which)   // 0: features, 1: their gradients
{
    const size_t tiles = (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128);
    return reinterpret_cast<float *>(fp.wg_scratch + tiles * (bf::kTileBytes + bf::kMaskBytes) + bf::kFinishFloats * sizeof(float) +
                                     (size_t)which * tiles * kFeatTileBytes);
}
// the fused pipeline with a workspace runs the trilinear stages as kernels of their own (k_tri_gather / k_tri_scatter)
static bool split_trilinear(const FieldParams &fp, int max_samples)
{
    return fp.paired && !fp.feat && fp.wg_scratch && fp.wg_scratch_bytes >= bf_wgrad_scratch_bytes(max_samples);
}
static int launch_tri_gather(FieldParams &fp, int max_samples, cudaStream_t st)
{
    float *feat = scratch_feat(fp, max_samples, 0);
    launch_chain(k_tri_gather, dim3((int)ceil_div64(ceil_div64(max_samples > 0 ? max_samples : 1, kTriChunk) * 4, 256)), dim3(256), 0, st, fp, feat);
    PSLAM_CHECK_LAUNCH("tri_gather");
    fp.feat = feat;
    return 0;
}
static unsigned char *scratch_finish(const FieldParams &fp, int max_samples)
{
    return fp.wg_scratch + (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128) * bf::kTileBytes;
}

// Which scratch holds the activations + masks of the most recent saving forward of a fused pipeline (stream-ordered
// with the backward that consumes them), together with the sample outputs it wrote.  Only `paired` launches (both
// from one pslam_render_t) save / consume; any other launch that writes the same scratch clears the record.
// The record is per device and guarded by a mutex (two host threads may drive two pipelines); what it guards is only the
// host-side decision "the backward may start from the saved masks", the data itself is stream-ordered.
struct SavedRecord { const void *scratch, *out; };
static SavedRecord g_saved[64] = {};
static std::mutex g_saved_mutex;
static int g_save_activations = 1;
void bf_set_save_activations(int on)
{
    std::lock_guard<std::mutex> lock(g_saved_mutex);
    g_save_activations = on ? 1 : 0;
    for (auto &r : g_saved) r = SavedRecord{nullptr, nullptr};
}
static void saved_set(const void *scratch, const void *out)
{
    std::lock_guard<std::mutex> lock(g_saved_mutex);
    g_saved[current_device()] = SavedRecord{scratch, out};
}
// a launch is about to rewrite `scratch` and / or the sample outputs `out`: whatever was saved there is gone
static void saved_invalidate(const void *scratch, const void *out)
{
    std::lock_guard<std::mutex> lock(g_saved_mutex);
    SavedRecord &r = g_saved[current_device()];
    if ((scratch && scratch == r.scratch) || (out && out == r.out)) r = SavedRecord{nullptr, nullptr};
}
static bool saved_matches(const void *scratch, const void *out)
{
    std::lock_guard<std::mutex> lock(g_saved_mutex);
    const SavedRecord &r = g_saved[current_device()];
    return scratch && scratch == r.scratch && out == r.out;
}

template <int KIND>
static int launch_bf(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    static PerDevice once = {};
    bool &configured = once.done[current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_field_bf<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, bf::Smem<KIND>::bytes);
        if (e != cudaSuccess) { set_error("field_bf: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    // persistent grid = as many whole clusters as can be resident at once (GPC boundaries may strand a few SMs)
    static PerDevice clusters = {};
    int &max_clusters = clusters.value[current_device()];
    if (max_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / bf::kCluster * bf::kCluster);
        cfg.blockDim = dim3(bf::kThreads);
        cfg.dynamicSmemBytes = bf::Smem<KIND>::bytes;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = bf::kCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k_field_bf<KIND>, &cfg);
        if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / bf::kCluster; }
        max_clusters = n < num_sms() / bf::kCluster ? n : num_sms() / bf::kCluster;
    }
    const int tiles = ceil_div(max_samples, 128);
    int grid = ceil_div(tiles > 0 ? tiles : 1, bf::kCluster) * bf::kCluster;
    if (grid > max_clusters * bf::kCluster) grid = max_clusters * bf::kCluster;
    launch_chain(k_field_bf<KIND>, dim3(grid), dim3(bf::kThreads), bf::Smem<KIND>::bytes, st, fp, reinterpret_cast<const unsigned char *>(fp.ws_tc));
    static const char *names[4] = {"field_bf_forward", "field_bf_backward", "field_bf_forward_save", "field_bf_backward_saved"};
    PSLAM_CHECK_LAUNCH(names[KIND]);
    return 0;
}

// forward; with decoder gradients to follow (fp.grad_dec) and a scratch, it also spills its activations and ReLU masks
int bf_launch_field_forward(const FieldParams &fp_in, int max_samples, cudaStream_t st, int part)
{
    FieldParams fp = fp_in;
    // the forward saves its ReLU masks for any backward that will follow on this pslam_render_t, and spills the wgrad
    // operands when that backward wants decoder gradients
    const bool save = g_save_activations && fp.paired && (fp.grad_dec || fp.grad_emb || fp.grad_rays) && fp.wg_scratch &&
                      fp.wg_scratch_bytes >= bf_wgrad_scratch_bytes(max_samples);
    fp.spill_ops = fp.grad_dec;
    saved_invalidate(fp.wg_scratch, fp.out);     // every forward rewrites its outputs (and, with a workspace, the feature rows)
    if (split_trilinear(fp, max_samples)) {
        if (part == 3) fp.feat = scratch_feat(fp, max_samples, 0);      // profiling: the rows of the previous full forward
        else if (int rc = launch_tri_gather(fp, max_samples, st)) return rc;
        if (part == 5) return 0;                                        // profiling: the trilinear gather alone
    }
    const bool pp = pp_enabled() && fp.feat != nullptr;   // two tiles in flight per CTA (field_pp.cu): reads feature rows only
    if (!save) return pp ? pp_launch(bf::kFwd, fp, max_samples, st) : launch_bf<bf::kFwd>(fp, max_samples, st);
    fp.act_masks = reinterpret_cast<uint32_t *>(scratch_finish(fp, max_samples) + bf::kFinishFloats * sizeof(float));
    // the fused backward (field_bw.cu) accumulates M = G4^T H2 straight into the reduction block: the forward clears it
    if (pp && fp.grad_dec) fp.finish_zero = reinterpret_cast<float *>(scratch_finish(fp, max_samples));
    if (int rc = pp ? pp_launch(bf::kFwdSave, fp, max_samples, st) : launch_bf<bf::kFwdSave>(fp, max_samples, st)) return rc;
    saved_set(fp.wg_scratch, fp.out);
    return 0;
}

int launch_tri_gather_rows(const FieldParams &fp, float *feat, int max_samples, cudaStream_t st)
{
    launch_chain(k_tri_gather, dim3((int)ceil_div64(ceil_div64(max_samples > 0 ? max_samples : 1, kTriChunk) * 4, 256)), dim3(256), 0, st, fp, feat);
    PSLAM_CHECK_LAUNCH("tri_gather");
    return 0;
}
int launch_tri_scatter_rows(const FieldParams &fp, const float *g_feat, int max_samples, cudaStream_t st)
{
    k_tri_scatter<<<(int)ceil_div64(max_samples > 0 ? max_samples : 1, kScatWarps * 32), kScatWarps * 32, 0, st>>>(fp, g_feat);
    PSLAM_CHECK_LAUNCH("tri_scatter");
    return 0;
}
int launch_grad_scale(const FieldParams &fp, uint32_t *gscale, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(gscale, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) { set_error("grad_scale: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    k_grad_scale<<<num_sms(), 256, 0, st>>>(reinterpret_cast<const float4 *>(fp.g_out), fp.nsamp, fp.nsamp_dev, gscale);
    PSLAM_CHECK_LAUNCH("grad_scale");
    return 0;
}

// one non-blocking side stream + fork / join events per device (created on first use, never destroyed)
SideStream *side_stream()
{
    static SideStream table[64] = {};
    static bool ready[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ready[dev]) {
        SideStream s{};
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        table[dev] = s;
        ready[dev] = true;
    }
    return &table[dev];
}

// backward: dgrad chain (+ trilinear backward); when decoder gradients are wanted the scratch must be
// provided and the wgrad kernel follows on the same stream
int bf_launch_field_backward(const FieldParams &fp_in, int max_samples, cudaStream_t st, int part)
{
    FieldParams fp = fp_in;
    cudaEvent_t joined = nullptr;
    bool fused = false, scatter_fused = false;
    if (!fp.grad_dec && !fp.paired) fp.wg_scratch = nullptr;
    fp.spill_ops = fp.grad_dec;
    // per-launch gradient scale: 4 bytes at the end of the weight-stream region (the f16 stream fills only its first half)
    fp.gscale = reinterpret_cast<uint32_t *>(const_cast<float *>(fp.ws_tc)) + kTcPackFloats - 4;
    const bool gmax_known = fp.paired && fp.gmax_ready;
    if (gmax_known) fp.gscale = fp.gmax_ready;          // k_composite_bwd published max |g_out| while it wrote g_out
    if (part == 4) {   // profiling: the trilinear scatter alone, on the feature-gradient rows of the previous full backward
        if (!split_trilinear(fp, max_samples) || !(fp.grad_emb || fp.grad_rays)) return 0;
        k_tri_scatter<<<(int)ceil_div64(max_samples > 0 ? max_samples : 1, kScatWarps * 32), kScatWarps * 32, 0, st>>>(fp, scratch_feat(fp, max_samples, 1));
        PSLAM_CHECK_LAUNCH("tri_scatter");
        return 0;
    }
    if (part != 2) {
        if (!gmax_known && part != 3) {   // part 3 (profiling): the chain kernel alone; the scale of the previous full backward is still there
            cudaError_t e = cudaMemsetAsync(fp.gscale, 0, sizeof(uint32_t), st);
            if (e != cudaSuccess) { set_error("field_bf: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
            k_grad_scale<<<num_sms(), 256, 0, st>>>(reinterpret_cast<const float4 *>(fp.g_out), fp.nsamp, fp.nsamp_dev, fp.gscale);
            PSLAM_CHECK_LAUNCH("grad_scale");
        }
        const bool saved = g_save_activations && fp.paired && saved_matches(fp.wg_scratch, fp.out);
        if (!saved) saved_invalidate(fp.wg_scratch, nullptr);   // about to be overwritten
        fp.act_masks = fp.wg_scratch ? reinterpret_cast<uint32_t *>(scratch_finish(fp, max_samples) + bf::kFinishFloats * sizeof(float)) : nullptr;
        if (fp.grad_dec && fp.wg_scratch) fp.finish_zero = reinterpret_cast<float *>(scratch_finish(fp, max_samples));
        const bool split = split_trilinear(fp, max_samples);
        FieldParams fps = fp;                           // the scatter kernel wants the sample tables, not the feature rows
        if (split) {
            if (saved) fp.feat = scratch_feat(fp, max_samples, 0);          // (not read: marks the trilinear stages as external)
            else if (int rc = launch_tri_gather(fp, max_samples, st)) return rc;   // the recompute needs the features again
            fp.g_feat = scratch_feat(fp, max_samples, 1);
        }
        if (saved) {
            const bool pp = pp_enabled() && fp.feat != nullptr;
            // decoder gradients wanted: chain + weight gradients in one kernel (nothing is spilled; the forward cleared the reduction block)
            fused = pp && bw_enabled() && fp.grad_dec && fp.spill_ops && part != 1;
            if (fused) {
                // the trilinear backward rides along in the kernel's idle warps (part 3 = the chain kernel alone for profiling: then
                // the stand-alone scatter is simply not launched, as before)
                scatter_fused = split && (fp.grad_emb || fp.grad_rays) && bw_fused_scatter();
                if (int rc = bw_launch(fp, max_samples, reinterpret_cast<float *>(scratch_finish(fp, max_samples)), st, scatter_fused ? &fps : nullptr)) return rc;
            } else if (int rc = pp ? pp_launch(bf::kBwdSaved, fp, max_samples, st) : launch_bf<bf::kBwdSaved>(fp, max_samples, st)) return rc;
        } else {
            if (int rc = launch_bf<bf::kBwdRecompute>(fp, max_samples, st)) return rc;
        }
        if (split && (fp.grad_emb || fp.grad_rays) && part != 3 && !scatter_fused) {
            // the embedding / ray scatter and the weight-gradient kernels both depend on the kernel above only: the scatter
            // (L2 atomics, few threads per SM) runs on a side stream underneath the HBM-bound wgrad kernel
            cudaStream_t ss = st;
            SideStream *side = (fp.grad_dec && part == 0) ? side_stream() : nullptr;
            if (side) {
                if (cudaEventRecord(side->fork, st) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork, 0) != cudaSuccess) {
                    set_error("field_bf: stream fork: %s", cudaGetErrorString(cudaGetLastError()));
                    return PSLAM_E_ARG;
                }
                ss = side->stream;
            }
            k_tri_scatter<<<(int)ceil_div64(max_samples > 0 ? max_samples : 1, kScatWarps * 32), kScatWarps * 32, 0, ss>>>(fps, fp.g_feat);
            PSLAM_CHECK_LAUNCH("tri_scatter");
            if (side) {
                if (cudaEventRecord(side->join, side->stream) != cudaSuccess) { set_error("field_bf: stream join: %s", cudaGetErrorString(cudaGetLastError())); return PSLAM_E_ARG; }
                joined = side->join;
            }
        }
    }
    if (!fp.grad_dec || part == 1 || part == 3) return 0;
    static PerDevice once = {};
    bool &configured = once.done[current_device()];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_bf, cudaFuncAttributeMaxDynamicSharedMemorySize, wgb::kSmemBytes);
        if (e != cudaSuccess) { set_error("wgrad_bf: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int tiles = ceil_div(max_samples > 0 ? max_samples : 1, 128);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    float *finish = reinterpret_cast<float *>(scratch_finish(fp, max_samples));
    if (!fused) {
        if (!fp.finish_zero) {                          // the chain kernel did not clear the reduction block (part 2 on its own grid)
            cudaError_t e = cudaMemsetAsync(finish, 0, bf::kFinishFloats * sizeof(float), st);
            if (e != cudaSuccess) { set_error("wgrad_bf: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        }
        launch_chain(k_wgrad_bf, dim3(grid), dim3(wgb::kThreads), wgb::kSmemBytes, st, fp, finish);
        PSLAM_CHECK_LAUNCH("wgrad_bf");
    }
    static PerDevice finish_once = {};
    bool &finish_configured = finish_once.done[current_device()];
    if (!finish_configured) {
        cudaError_t e2 = cudaFuncSetAttribute(k_wgrad_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinishSmem);
        if (e2 != cudaSuccess) { set_error("wgrad_finish: cudaFuncSetAttribute: %s", cudaGetErrorString(e2)); return (int)e2; }
        finish_configured = true;
    }
    launch_chain(k_wgrad_finish, dim3(256), dim3(128), kFinishSmem, st, fp, finish);
    PSLAM_CHECK_LAUNCH("wgrad_finish");
    if (joined) {                                       // the scatter kernel forked onto the side stream joins here
        cudaError_t e3 = cudaStreamWaitEvent(st, joined, 0);
        if (e3 != cudaSuccess) { set_error("field_bf: cudaStreamWaitEvent: %s", cudaGetErrorString(e3)); return (int)e3; }
    }
    return 0;
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_debug_bf_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(g_bf_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { set_error("bf_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int pslam_debug_umma_gemm_bf(const float *A, const float *B, float *D, int N, int K, int mode, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(A && B && D, PSLAM_E_ARG, "null pointer");
    PSLAM_CHECK_ARG(N >= 16 && N <= 144 && N % 16 == 0 && K >= 16 && K <= 144 && K % 16 == 0, PSLAM_E_RANGE, "N in 16..144 step 16, K in 16..144 step 16");
    PSLAM_CHECK_ARG(mode == 0 || mode == 1, PSLAM_E_RANGE, "mode must be 0 or 1");
    const int smem = mode == 1 ? (128 + N) * K * 4 : N * K * 4;
    cudaError_t e = cudaFuncSetAttribute(k_debug_umma_bf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("debug_umma_bf: %s", cudaGetErrorString(e)); return (int)e; }
    k_debug_umma_bf<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, mode);
    PSLAM_CHECK_LAUNCH("debug_umma_gemm_bf");
    return 0;
}
