// Kernel 4 on the 5th-generation tensor cores, 3xF16 build with TWO 128-sample tiles in flight per CTA
// (src/variations/nrgbd.py:116-135 forward, and its dgrad chain).
//
// field_bf.cu keeps one tile per CTA: a layer's MMAs, its epilogue and the hand-offs between the two form one
// dependent chain, and the tensor pipe idles two thirds of the time (ncu: sm__pipe_tensor_cycles_active 32 %).
// Here the 256 worker threads are two groups of 128 (one thread per sample row, all 128 columns), each group owns a
// tile, and the single MMA-issuing thread alternates between them: while group 0 runs the epilogue of layer l, the
// tensor pipe executes layer l of group 1's tile, and vice versa.  What makes two tiles fit in the 512 columns of
// tensor memory (2 x (A_hi 64 + A_lo 64 + D 128)):
//   * the 16 input features never go to tensor memory: they sit in shared memory as a K-major A operand and the
//     two K = 16 products that read them (layer 1, and the feature columns of layer 4) are SS-mode MMAs;
//   * the sdf head (row 0 of W3) is a dot product of the h2 epilogue on the CUDA cores (one thread holds the whole
//     row), and its dgrad a rank-1 term of the g_h2 epilogue, so layer 3 / its dgrad are plain N = K = 128;
//   * the two N = 16 products of dL/dfeatures (through W4's feature columns and through W1) run as small layers of
//     their own into D[0,16) at the two ends of the backward chain, when D is free.
// Weight stream, operand scales, ReLU-mask layout and the wgrad scratch are those of field_bf.cu (field_bf.cuh), so
// k_wgrad_bf / k_wgrad_finish consume what these kernels spill.  The trilinear stages are the stand-alone kernels
// (k_tri_gather / k_tri_scatter): this build always reads feature rows and writes feature-gradient rows.
#include "field_bf.cuh"
#include "kernels.h"
#include <type_traits>

namespace pslam {

using namespace umma;

namespace pp {
using namespace bf;
constexpr int kPPThreads = 384;           // warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle, warps 4-7 group 0, warps 8-11 group 1
constexpr int kGroups = 2;
constexpr int kRows = 128;                // worker threads per group = rows of a tile
constexpr int kPPStages = 8;
// tensor memory of group g starts at column g * 256
constexpr int cGroup = 256, cHi = 0, cLo = 64, cAcc = 128;
// shared memory
constexpr int kFPlane = 4096;             // K-major [2 k-chunks][128 rows][16 B]
constexpr int oFeat = kPPStages * kStageBytes;                 // [group][hi | lo]
constexpr int oBars = oFeat + kGroups * 2 * kFPlane;           // full[8] empty[8] a_ready[2] mma_done[2]
constexpr int oTmemPtr = oBars + 8 * (2 * kPPStages + 2 * kGroups);
constexpr int oBias = oTmemPtr + 16;                           // b1 b2 b3[1:] b4 (x16) | b3[0] b5[3]
constexpr int oW5 = oBias + 4 * (4 * 128 + 4);                 // W5 [3][128] fp32
constexpr int oW30 = oW5 + 4 * 3 * 128;                        // W3 row 0 (sdf head) [128] fp32
constexpr int kSmemBytes = oW30 + 4 * 128;

// MMA phases of a tile, as indices into the packed weight stream (field_bf.cuh: cN / cK / layer_offset)
//   forward : 0 (features -> h1, SS)   1 (h1 -> h2)   2 (h2 -> t, 128 of the 144 packed rows)   3 (t -> hc, + SS feature chunk)
//   backward: 10 (g_hc -> g_f part, N = 16)   6 (g_hc -> g_t)   7 (g_t -> g_h2, without the sdf chunk)   8 (g_h2 -> g_h1)   9 (g_h1 -> g_f, N = 16)
__host__ __device__ constexpr int n_chunks(int l) { return hK(l) < 64 ? 1 : (l == 3 ? 5 : 4); }
__host__ __device__ constexpr int chunk_bytes(int l, int c) { return hN(l) * ((hK(l) < 64 || c == 4) ? 16 : 32) * 4; }
__host__ __device__ constexpr int chunk_offset(int l, int c) { return layer_offset(l) + c * hN(l) * 32 * 4; }
}  // namespace pp

// 16 accumulator columns [c0, c0 + 16) of this thread's row: accumulators -> registers -> (bias / activation / mask) ->
// f16 hi / lo pairs -> the group's next A operand in tensor memory and, optionally, the wgrad scratch.
//   MODE 0: y = relu(D/16 + bias), records y > 0      MODE 1: y = D/16 + bias
//   MODE 2: y = mask ? D/16 (+ r1 * wx[c]) : 0        MODE 3: y = D/16
//   EXTRA 1: acc[0] += wx[c] * y (sdf head)   EXTRA 2: acc[0..2] += W5[.][c] * y and no A operand is written (colour head)
//   EXTRA 3: the rank-1 term r1 * wx[c] of the sdf head's dgrad is added before the mask
// (32-column batches with the next batch's tcgen05.ld in flight were measured: slower, 116 vs 97 us for the saving forward.)
template <int MODE, int EXTRA>
__device__ __forceinline__ void pp_epi16(uint32_t tg, int c0, const float *bias, uint32_t &mask, int shift, unsigned char *stg, float &ymax,
                                         const float *wx, float *acc, float r1)
{
    using namespace pp;
    uint32_t v[16];
    tmem_ld16(tg + cAcc + c0, v);
    tmem_wait_ld();
    uint32_t bits = 0u;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        float y = __uint_as_float(v[e]);
        if (MODE == 0) { y = fmaxf(fmaf(y, kInvScale, bias[c0 + e]), 0.0f); bits |= (y > 0.0f ? 1u : 0u) << e; }
        if (MODE == 1) y = fmaf(y, kInvScale, bias[c0 + e]);
        if (MODE == 2) {
            y = (EXTRA == 3) ? fmaf(r1, wx[c0 + e], y * kInvScale) : y * kInvScale;
            y = ((mask >> (shift + e)) & 1u) ? y : 0.0f;
        }
        if (MODE == 3) y = y * kInvScale;
        ymax = fmaxf(ymax, fabsf(y));
        if (EXTRA == 1) acc[0] = fmaf(wx[c0 + e], y, acc[0]);
        if (EXTRA == 2) {
            acc[0] = fmaf(wx[c0 + e], y, acc[0]);
            acc[1] = fmaf(wx[128 + c0 + e], y, acc[1]);
            acc[2] = fmaf(wx[256 + c0 + e], y, acc[2]);
        }
        v[e] = __float_as_uint(y);
    }
    if (MODE == 0) mask = shift ? (mask | (bits << 16)) : bits;
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
    if (EXTRA != 2) {
        tmem_st8(tg + cHi + c0 / 2, hi);
        tmem_st8(tg + cLo + c0 / 2, lo);
    }
}

// optional timeline of CTA 0 (pslam_debug_pp_trace): [iteration < 4][who: worker g0, worker g1, issuer g0, issuer g1][16] clock64 stamps
//   worker: 0 tile start, 1 first operand published, 2 + 2i accumulators of phase i seen, 3 + 2i its epilogue done and published
//   issuer: 2i the phase's operand seen, 2i + 1 its MMAs issued and committed
__device__ long long *g_pp_trace = nullptr;
#define PP_TRACE(it_, who_, slot_)                                                                          \
    do {                                                                                                    \
        if (g_pp_trace && blockIdx.x == 0 && (it_) < 4) g_pp_trace[((it_) * 4 + (who_)) * 16 + (slot_)] = clock64(); \
    } while (0)

// The MMA-issuing thread: the weight ring position and, per group, how many phases it has issued.
struct PPIssuer {
    unsigned char *smem;
    uint64_t *full, *empty, *a_ready, *mma_done;
    uint32_t tmem;
    int stage, phase;
    uint32_t uses[pp::kGroups];

    // one phase of group g: all chunks of packed layer L (N output columns taken from row ROW0 of the packed rows)
    template <int L, int N, int ROW0>
    __device__ __forceinline__ void layer(int g, int it, int ph)
    {
        using namespace pp;
        constexpr int NP = hN(L), K = hK(L);
        constexpr int NCH = (L == 7) ? 4 : n_chunks(L);   // the sdf chunk of layer 7 is neither streamed nor issued
        const uint32_t idesc = idesc_h16(128, N);
        uint32_t t = tmem + (uint32_t)g * cGroup;
        asm volatile("" : "+r"(t));
        const uint32_t a_hi = t + cHi, a_lo = t + cLo, d = t + cAcc;
        // features as a K-major shared-memory A operand: [hi plane | lo plane], k-chunk stride 2048 B, 8-row group stride 128 B
        const uint32_t fbase = smem_u32(smem + oFeat + g * 2 * kFPlane);
        const uint64_t f_hi = sdesc(fbase, 2048, 128), f_lo = sdesc(fbase + kFPlane, 2048, 128);
        // The whole warp walks the phase; one elected lane issues.  (Issuing from an `if (lane == 0)` region makes the compiler
        // wrap every UTCHMMA in an ELECT / BRA.U.ANY loop -- the traces showed ~95 clocks per MMA against 64 of execution.)
        mbar_wait(a_ready + g, uses[g] & 1u);
        fence_after_sync();
        if ((threadIdx.x & 31) == 0) PP_TRACE(it, 2 + g, 2 * ph);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            constexpr bool small = K < 64;
            const int kk = (small || c == 4) ? 16 : 32;
            mbar_wait(full + stage, phase);
            fence_after_sync();
            const uint64_t b0 = sdesc(smem_u32(smem + stage * kStageBytes) + ROW0 * 16, NP * 16, 128);
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s * 16 < kk) {
                        const uint64_t b_hi = b0 + (uint64_t)((s * (2 * NP * 16)) >> 4);
                        const uint64_t b_lo = b0 + (uint64_t)((NP * kk * 2 + s * (2 * NP * 16)) >> 4);
                        const uint32_t acc0 = (c == 0 && s == 0) ? 0u : 1u;
                        if (small || c == 4) {                 // K = 16 product on the features (shared-memory A)
                            mma_h16_ss(d, f_lo, b_hi, idesc, acc0);
                            mma_h16_ss(d, f_hi, b_lo, idesc, 1u);
                            mma_h16_ss(d, f_hi, b_hi, idesc, 1u);
                        } else {
                            const uint32_t acol = (uint32_t)(16 * (c + 4 * s)) >> 1;   // chunk c holds the k-steps c and c + 4
                            mma_h16_ts(d, a_lo + acol, b_hi, idesc, acc0);
                            mma_h16_ts(d, a_hi + acol, b_lo, idesc, 1u);
                            mma_h16_ts(d, a_hi + acol, b_hi, idesc, 1u);
                        }
                    }
                }
                mma_commit_mcast(empty + stage, kClusterMask);
                if (c == NCH - 1) mma_commit(mma_done + g);
            }
            __syncwarp();
            if (++stage == kPPStages) { stage = 0; phase ^= 1; }
        }
        if ((threadIdx.x & 31) == 0) PP_TRACE(it, 2 + g, 2 * ph + 1);
        ++uses[g];
    }
};

template <int KIND>
__global__ void __cluster_dims__(bf::kCluster, 1, 1) __launch_bounds__(pp::kPPThreads, 1)
k_field_pp(FieldParams p, const unsigned char *__restrict__ wstream)
{
    pdl_enter();
    using namespace pp;
    static_assert(KIND == kFwd || KIND == kFwdSave || KIND == kBwdSaved, "the recomputing backward stays in field_bf.cu");
    constexpr bool kIsFwd = KIND != kBwdSaved;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *empty = full + kPPStages;
    uint64_t *a_ready = empty + kPPStages;   // [group]: the group's next A operand is complete (and its accumulator columns may be overwritten)
    uint64_t *mma_done = a_ready + kGroups;  // [group]: the phase's accumulators are complete
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oTmemPtr);
    float *sBias = reinterpret_cast<float *>(smem + oBias);
    float *sW5 = reinterpret_cast<float *>(smem + oW5), *sW30 = reinterpret_cast<float *>(smem + oW30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    const int G = (int)gridDim.x;
    // Tile of (iteration it, CTA b, group g) = it * 2G + g * G + b.  The CTAs of a cluster consume one weight stream in
    // lockstep, so every CTA runs the same iterations; an iteration whose second half [it * 2G + G, ..) is empty runs group 0
    // only (the same decision in every CTA), other out-of-range tiles are dummies (no valid rows, nothing stored).
    const int iters = (ntiles + 2 * G - 1) / (2 * G);
    const uint32_t crank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kPPStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kCluster); }
        for (int i = 0; i < kGroups; ++i) { mbar_init(a_ready + i, kRows); mbar_init(mma_done + i, 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    const bool spill = p.wg_scratch != nullptr && p.spill_ops != 0;
    for (int i = threadIdx.x; i < 4 * 128 + 4; i += kPPThreads) {
        float v;
        if (i < 128) v = p.dec.b1[i];
        else if (i < 256) v = p.dec.b2[i - 128];
        else if (i < 384) v = p.dec.b3[1 + i - 256];
        else if (i < 512) v = p.dec.b4[i - 384];
        else if (i == 512) v = p.dec.b3[0];
        else v = p.dec.b5[i - 513];
        sBias[i] = i < 512 ? kScale * v : v;
    }
    for (int i = threadIdx.x; i < 3 * 128; i += kPPThreads) sW5[i] = p.dec.W5[i];
    for (int i = threadIdx.x; i < 128; i += kPPThreads) sW30[i] = p.dec.W3[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    cluster_sync();
    const uint32_t tmem = *tmem_ptr;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
        if (warp == 0) {
            // ===================== TMA producer: each phase's chunks once per active group, in the issuer's order =====================
            int stage = 0, phase = 0;
            auto emit = [&](int l, int ng) {
                const int nch = (l == 7) ? 4 : n_chunks(l);
                for (int g = 0; g < ng; ++g) {
                    for (int c = 0; c < nch; ++c) {
                        const uint32_t bytes = (uint32_t)chunk_bytes(l, c), part = bytes / kCluster;
                        const unsigned char *src = wstream + chunk_offset(l, c);
                        mbar_wait(empty + stage, phase ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(full + stage, bytes);
                            bulk_g2s_mcast(smem + stage * kStageBytes + crank * part, src + crank * part, part, full + stage, kClusterMask);
                        }
                        __syncwarp();
                        if (++stage == kPPStages) { stage = 0; phase ^= 1; }
                    }
                }
            };
            for (int it = 0; it < iters; ++it) {
                const int ng = (ntiles - it * 2 * G > G) ? 2 : 1;
                if (kIsFwd) { emit(0, ng); emit(1, ng); emit(2, ng); emit(3, ng); }
                else { emit(10, ng); emit(6, ng); emit(7, ng); emit(8, ng); emit(9, ng); }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer: one thread, the two groups' phases interleaved =====================
            {
                PPIssuer mi{smem, full, empty, a_ready, mma_done, tmem, 0, 0, {0u, 0u}};
                for (int it = 0; it < iters; ++it) {
                    const int ng = (ntiles - it * 2 * G > G) ? 2 : 1;
                    if constexpr (kIsFwd) {
                        for (int g = 0; g < ng; ++g) mi.template layer<0, 128, 0>(g, it, 0);
                        for (int g = 0; g < ng; ++g) mi.template layer<1, 128, 0>(g, it, 1);
                        for (int g = 0; g < ng; ++g) mi.template layer<2, 128, 0>(g, it, 2);
                        for (int g = 0; g < ng; ++g) mi.template layer<3, 128, 0>(g, it, 3);
                    } else {
                        for (int g = 0; g < ng; ++g) mi.template layer<10, 16, 0>(g, it, 0);
                        for (int g = 0; g < ng; ++g) mi.template layer<6, 128, 0>(g, it, 1);
                        for (int g = 0; g < ng; ++g) mi.template layer<7, 128, 0>(g, it, 2);
                        for (int g = 0; g < ng; ++g) mi.template layer<8, 128, 0>(g, it, 3);
                        for (int g = 0; g < ng; ++g) mi.template layer<9, 16, 0>(g, it, 4);
                    }
                }
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsWorker));
        // ===================== workers: group g = one tile, thread = one sample row =====================
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;                  // row of the tile
        const uint32_t tg = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * cGroup;
        const int rowoff_big = (m >> 6) * 32768 + ((m >> 3) & 7) * 2048 + (m & 7) * 16;
        const int rowoff_small = (m >> 6) * 4096 + ((m >> 3) & 7) * 256 + (m & 7) * 16;
        unsigned char *sF = smem + oFeat + g * 2 * kFPlane + m * 16;   // this row's 16 B of k-chunk 0, hi plane
        uint32_t done_uses = 0;
        float ymax = 0.0f;
        const float Sg = kIsFwd ? 1.0f : grad_scale(p.gscale), invSg = 1.0f / Sg;
        if (p.finish_zero) {                           // clear the reduction block of the weight-gradient kernels that follow
            for (int i = (int)blockIdx.x * (kPPThreads - 128) + (int)threadIdx.x - 128; i < (int)(kFinishFloats / 4); i += G * (kPPThreads - 128))
                reinterpret_cast<float4 *>(p.finish_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        int tr_it = 0, tr_slot = 0;                    // trace bookkeeping (thread 0 of each group stamps)
        auto layer_done = [&]() {
            mbar_wait(mma_done + g, done_uses & 1u);
            ++done_uses;
            fence_after_sync();
            if (m == 0) PP_TRACE(tr_it, g, tr_slot++);
        };
        auto a_is_ready = [&]() {                      // tensor-memory writes (and accumulator reads) of this thread are complete
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(a_ready + g);
            if (m == 0) PP_TRACE(tr_it, g, tr_slot++);
        };
        using M0 = std::integral_constant<int, 0>; using M1 = std::integral_constant<int, 1>;
        using M2 = std::integral_constant<int, 2>; using M3 = std::integral_constant<int, 3>;
        using X0 = std::integral_constant<int, 0>; using X1 = std::integral_constant<int, 1>;
        using X2 = std::integral_constant<int, 2>; using X3 = std::integral_constant<int, 3>;
        auto epilogue = [&](auto mode, auto extra, const float *bias, uint32_t (&mask)[4], unsigned char *stg, const float *wx, float *acc, float r1) __attribute__((always_inline)) {
            constexpr int MODE = decltype(mode)::value, EXTRA = decltype(extra)::value;
            // 32 columns (one mask word) per trip; not unrolled further: with all 128 columns in one basic block the scheduler
            // hoists the shared-memory operands of every batch and the workers spill
#pragma unroll 1
            for (int w = 0; w < 4; ++w) {
                uint32_t mw = w == 0 ? mask[0] : (w == 1 ? mask[1] : (w == 2 ? mask[2] : mask[3]));
                pp_epi16<MODE, EXTRA>(tg, 32 * w, bias, mw, 0, stg, ymax, wx, acc, r1);
                pp_epi16<MODE, EXTRA>(tg, 32 * w + 16, bias, mw, 16, stg, ymax, wx, acc, r1);
                if (MODE == 0) {
                    mask[0] = w == 0 ? mw : mask[0]; mask[1] = w == 1 ? mw : mask[1];
                    mask[2] = w == 2 ? mw : mask[2]; mask[3] = w == 3 ? mw : mask[3];
                }
            }
        };
        uint32_t nomask[4] = {0u, 0u, 0u, 0u};
        // next tile's per-row inputs: the forward fetches its feature rows a few layers before the tile ends, the backward loads the
        // next tile's ReLU masks / outputs / upstream gradient straight into their registers once the last masked epilogue is done
        float4 pf[4];
        uint32_t m1[4], m2[4], mc[4];
        float4 po = make_float4(0.f, 0.f, 0.f, 0.f), pgo = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 4; ++e) { pf[e] = make_float4(0.f, 0.f, 0.f, 0.f); m1[e] = 0u; m2[e] = 0u; mc[e] = 0u; }
        auto prefetch_tile = [&](int tn) __attribute__((always_inline)) {
            const int sn = tn * 128 + m;
            const bool in = tn < ntiles && sn < nsamp;
            if (kIsFwd) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    pf[e] = in ? __ldg(reinterpret_cast<const float4 *>(p.feat + (size_t)sn * 16) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                const uint32_t *mk = p.act_masks + (size_t)(tn < ntiles ? tn : 0) * (kMaskBytes / 4) + m;
#pragma unroll
                for (int w = 0; w < 4; ++w) { m1[w] = mk[w * 128]; m2[w] = mk[512 + w * 128]; mc[w] = mk[1024 + w * 128]; }
                po = in ? __ldg(reinterpret_cast<const float4 *>(p.out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                pgo = in ? __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch_tile(g * G + (int)blockIdx.x);
        if (g_pp_trace && blockIdx.x == 0 && threadIdx.x == 128) {
            long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            g_pp_trace[(3 * 4 + 3) * 16 + 12] = clock64();
            g_pp_trace[(3 * 4 + 3) * 16 + 14] = ns;
        }
        for (int it = 0; it < iters; ++it) {
            const int ng = (ntiles - it * 2 * G > G) ? 2 : 1;
            if (g >= ng) continue;
            const int tile = it * 2 * G + g * G + (int)blockIdx.x;
            const bool real_tile = tile < ntiles;
            tr_it = it; tr_slot = 0;
            if (m == 0) PP_TRACE(tr_it, g, tr_slot++);
            const int s = real_tile ? tile * 128 + m : nsamp;
            const bool valid = s < nsamp;
            unsigned char *scr = (spill && real_tile) ? p.wg_scratch + (size_t)tile * kTileBytes : nullptr;
            auto stg_of = [&](int op) -> unsigned char * { return scr ? scr + (size_t)op * kOpBytes + rowoff_big : nullptr; };
            if constexpr (kIsFwd) {
                // ---- features -> shared-memory A operand (x16, hi / lo) ----
                {
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { f[4 * e] = pf[e].x; f[4 * e + 1] = pf[e].y; f[4 * e + 2] = pf[e].z; f[4 * e + 3] = pf[e].w; }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 16; ++e) ymax = fmaxf(ymax, fabsf(kScale * f[e]));
#pragma unroll
                    for (int e = 0; e < 8; ++e) h16_split2(kScale * f[2 * e], kScale * f[2 * e + 1], hi[e], lo[e]);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        *reinterpret_cast<uint4 *>(sF + j * 2048) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        *reinterpret_cast<uint4 *>(sF + kFPlane + j * 2048) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                    if (scr) {
                        unsigned char *dst = scr + oF + rowoff_small;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            *reinterpret_cast<uint4 *>(dst + j * 128) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                            *reinterpret_cast<uint4 *>(dst + 2048 + j * 128) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> the tensor core's reads
                    fence_before_sync();
                    mbar_arrive(a_ready + g);
                    if (m == 0) PP_TRACE(tr_it, g, tr_slot++);
                }
                float sdf_acc[1] = {0.0f}, head[3] = {0.f, 0.f, 0.f};
                layer_done();
                epilogue(M0{}, X0{}, sBias, m1, stg_of(oH1), nullptr, nullptr, 0.f);                    // h1
                a_is_ready();
                layer_done();
                epilogue(M0{}, X1{}, sBias + 128, m2, stg_of(oH2), sW30, sdf_acc, 0.f);                 // h2 (+ sdf head)
                prefetch_tile(tile + 2 * G);
                a_is_ready();
                layer_done();
                epilogue(M1{}, X0{}, sBias + 256, nomask, nullptr, nullptr, nullptr, 0.f);              // t
                a_is_ready();
                layer_done();
                epilogue(M0{}, X2{}, sBias + 384, mc, stg_of(oHC), sW5, head, 0.f);                     // hc + colour head
                if (valid) {
                    const float r = sigmoid_f(fmaf(head[0], kInvScale, sBias[513]));
                    const float gg = sigmoid_f(fmaf(head[1], kInvScale, sBias[514]));
                    const float b = sigmoid_f(fmaf(head[2], kInvScale, sBias[515]));
                    const float sdf = fmaf(sdf_acc[0], kInvScale, sBias[512]);
                    *reinterpret_cast<float4 *>(p.out + (size_t)s * 4) = make_float4(r, gg, b, sdf);
                }
                if (KIND == kFwdSave && real_tile) {
                    uint32_t *mk = p.act_masks + (size_t)tile * (kMaskBytes / 4) + m;
#pragma unroll
                    for (int w = 0; w < 4; ++w) { mk[w * 128] = m1[w]; mk[512 + w * 128] = m2[w]; mk[1024 + w * 128] = mc[w]; }
                }
                // the next tile's feature rows overwrite sF: layer 4's SS-mode MMAs have completed (layer_done above)
            } else {
                float4 go = valid ? pgo : make_float4(0.f, 0.f, 0.f, 0.f);
                const float r = valid ? po.x : 0.f, gg = valid ? po.y : 0.f, b = valid ? po.z : 0.f;
                go.x *= Sg; go.y *= Sg; go.z *= Sg; go.w *= Sg;
                const float g5[4] = {go.x * (1.0f - r) * r, go.y * (1.0f - gg) * gg, go.z * (1.0f - b) * b, go.w};
                if (scr) {
                    uint32_t h0, l0, h1w, l1w;
                    h16_split2(g5[0], g5[1], h0, l0);
                    h16_split2(g5[2], g5[3], h1w, l1w);
                    unsigned char *dst = scr + oG5 + rowoff_small;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048) = make_uint4(l0, l1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048 + 128) = make_uint4(0u, 0u, 0u, 0u);
                    const float s0 = warp_sum(g5[0]), s1 = warp_sum(g5[1]), s2 = warp_sum(g5[2]), s3 = warp_sum(g5[3]);
                    if (lane == 0) {
                        atomicAdd(p.g_dec.b5 + 0, s0 * invSg); atomicAdd(p.g_dec.b5 + 1, s1 * invSg); atomicAdd(p.g_dec.b5 + 2, s2 * invSg);
                        atomicAdd(p.g_dec.b3, s3 * invSg);
                    }
                }
                {
                    // g_hc = mask_hc . (W5^T g5) on the CUDA cores -> A
                    unsigned char *stg = stg_of(oG4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c0 = 16 * j;
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
                        tmem_st8(tg + cHi + c0 / 2, hi);
                        tmem_st8(tg + cLo + c0 / 2, lo);
                    }
                }
                a_is_ready();
                layer_done();                                                                           // g_f part through W4's feature columns
                float gf[16];
                {
                    uint32_t v[16];
                    tmem_ld16(tg + cAcc, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) gf[e] = __uint_as_float(v[e]);
                }
                fence_before_sync();
                mbar_arrive(a_ready + g);                                                               // accumulator columns free again, A unchanged
                if (m == 0) PP_TRACE(tr_it, g, tr_slot++);
                layer_done();
                epilogue(M3{}, X0{}, nullptr, nomask, nullptr, nullptr, nullptr, 0.f);                  // g_t
                a_is_ready();
                layer_done();
                epilogue(M2{}, X3{}, nullptr, m2, stg_of(oG2), sW30, nullptr, go.w);                    // g_h2 (+ sdf head's rank-1 term)
                a_is_ready();
                layer_done();
                epilogue(M2{}, X0{}, nullptr, m1, stg_of(oG1), nullptr, nullptr, 0.f);                  // g_h1
                a_is_ready();
                prefetch_tile(tile + 2 * G);                                                            // (its masks are dead now)
                layer_done();
                {
                    uint32_t v[16];
                    tmem_ld16(tg + cAcc, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) gf[e] = (gf[e] + __uint_as_float(v[e])) * (kInvScale * invSg);
                }
                if (valid && p.g_feat) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4 *>(p.g_feat + (size_t)s * 16 + 4 * j) = make_float4(gf[4 * j], gf[4 * j + 1], gf[4 * j + 2], gf[4 * j + 3]);
                }
                // the next tile's g_hc overwrites A and its first small layer D[0,16): this tile's last MMAs have completed and been read
            }
        }
        if (p.range_flag && ymax >= 32752.0f) atomicOr(p.range_flag, 4);
        if (g_pp_trace && blockIdx.x == 0 && threadIdx.x == 128) {   // kernel span of CTA 0 in clocks and in nanoseconds: [3][3][12..15]
            long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            g_pp_trace[(3 * 4 + 3) * 16 + 13] = clock64();
            g_pp_trace[(3 * 4 + 3) * 16 + 15] = ns;
        }
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc(tmem, bf::kTmemCols);
}

static int g_pp_enabled = 1;
int pp_enabled() { return g_pp_enabled; }
void pp_set_enabled(int on) { g_pp_enabled = on ? 1 : 0; }

struct PPDeviceState { bool configured[4]; int max_clusters[4]; };
static PPDeviceState g_pp_state[64] = {};

template <int KIND>
static int launch_pp(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("field_pp: cudaGetDevice failed"); return PSLAM_E_ARG; }
    PPDeviceState &ds = g_pp_state[dev];
    if (!ds.configured[KIND]) {      // per device: the attribute belongs to the function on the current device
        cudaError_t e = cudaFuncSetAttribute(k_field_pp<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, pp::kSmemBytes);
        if (e != cudaSuccess) { set_error("field_pp: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / bf::kCluster * bf::kCluster);
        cfg.blockDim = dim3(pp::kPPThreads);
        cfg.dynamicSmemBytes = pp::kSmemBytes;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = bf::kCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, k_field_pp<KIND>, &cfg);
        if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / bf::kCluster; }
        ds.max_clusters[KIND] = n < num_sms() / bf::kCluster ? n : num_sms() / bf::kCluster;
        ds.configured[KIND] = true;
    }
    // one tile per CTA until every SM has one, two in flight beyond that
    const int tiles = ceil_div(max_samples > 0 ? max_samples : 1, 128);
    int grid = ceil_div(tiles, bf::kCluster) * bf::kCluster;
    if (grid > ds.max_clusters[KIND] * bf::kCluster) grid = ds.max_clusters[KIND] * bf::kCluster;
    launch_chain(k_field_pp<KIND>, dim3(grid), dim3(pp::kPPThreads), pp::kSmemBytes, st, fp, reinterpret_cast<const unsigned char *>(fp.ws_tc));
    static const char *names[4] = {"field_pp_forward", "", "field_pp_forward_save", "field_pp_backward_saved"};
    PSLAM_CHECK_LAUNCH(names[KIND]);
    return 0;
}

int pp_launch(int kind, const FieldParams &fp, int max_samples, cudaStream_t st)
{
    if (kind == bf::kFwd) return launch_pp<bf::kFwd>(fp, max_samples, st);
    if (kind == bf::kFwdSave) return launch_pp<bf::kFwdSave>(fp, max_samples, st);
    if (kind == bf::kBwdSaved) return launch_pp<bf::kBwdSaved>(fp, max_samples, st);
    set_error("field_pp: unsupported kernel kind %d", kind);
    return PSLAM_E_ARG;
}

}  // namespace pslam

extern "C" int pslam_debug_pp_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(pslam::g_pp_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { pslam::set_error("pp_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
