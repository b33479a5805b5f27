// Width-256 decoder (the reference's ScanNet / ARKit configuration, configs/scannet/scannet.yaml:15-21 with
// src/variations/nrgbd.py:106-135: 16 -> 256 -> 256 -> 129 -> [128 + 16] -> 256 -> 3) on the 5th-generation tensor cores,
// 3xF16 like the width-128 build (field_pp.cu): forward, dgrad chain and weight gradients.
//
// One 128-sample tile per CTA, because tensor memory holds exactly one: the A operand of a 256-feature layer is 128 + 128
// columns (f16 hi / lo pairs) and a 256-column fp32 accumulator fills the other half of the 512 columns.  What makes that fit:
//   * the 16 input features never go to tensor memory (K-major shared-memory operand, SS-mode MMAs), the sdf head and the
//     colour head run on the CUDA cores inside the h2 / hc epilogues, the two N = 16 products of dL/dfeatures go to D[0,16)
//     at the ends of the backward chain -- as in field_pp.cu;
//   * the 256 worker threads are two per sample row, each owning 128 of the 256 columns (= one 128-feature operand block).
// N = 256 MMAs run 128 clocks each, so the issue overhead that bounds the width-128 chain is halved per FLOP and a single tile
// per CTA already keeps the tensor pipe busier than two 128-wide tiles did.
// Weight gradients: the forward spills H1, H2, HC, T, F and the chain G4, Gt, G2, G1, G5 -- already split into f16 hi / lo planes
// and in the MN-major core-matrix order of tcgen05.mma (field_bf.cuh), 7.1 kB per sample -- and k_wgrad_w256 reduces over the
// samples with both operands from shared memory.  A 256 x 256 weight gradient alone is a full tensor memory of accumulators,
// so the CTAs of that kernel take ROLES (role 0: dW2; role 1: dW3[1:], dW4[:, :128]; role 2: the N = 16 products and the bias
// column sums), each role walking its share of the tiles with its accumulators resident and flushing once with red.v4.
#include "field_bf.cuh"
#include "kernels.h"
#include <mutex>
#include <type_traits>

namespace pslam {

using namespace umma;

namespace w2 {
using namespace bf;

constexpr int kWThreads = 384;            // warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle, warps 4-11 workers
constexpr int kWStages = 5;
constexpr int kWStage = 32768;            // one chunk: 256 rows x 32 k x (hi, lo) x 2 B
constexpr int cHi = 0, cLo = 128, cAcc = 256;
// packed stream: layer -> (N, K)
//   0 L1 (256,16)  1 L2 (256,256)  2 L3 features (128,256)  3 L4 (256,144: t then the 16 features)
//   4 B0 (16,256) = W4[:,128:]^T   5 B1 (128,256) = W4[:, :128]^T   6 B2 (256,128) = W3[1:]^T   7 B3 (256,256) = W2^T   8 B4 (16,256) = W1^T
constexpr int kWLayers = 9;
__host__ __device__ constexpr int wN(int l) { return (l == 4 || l == 8) ? 16 : ((l == 2 || l == 5) ? 128 : 256); }
__host__ __device__ constexpr int wK(int l) { return l == 0 ? 16 : (l == 3 ? 144 : (l == 6 ? 128 : 256)); }
__host__ __device__ constexpr int w_layer_offset(int l) { int o = 0; for (int i = 0; i < l; ++i) o += wN(i) * wK(i) * 4; return o; }
__host__ __device__ constexpr int w_chunks(int l) { return (wK(l) + 31) / 32; }
__host__ __device__ constexpr int w_chunk_kk(int l, int c) { return (wK(l) - 32 * c) >= 32 ? 32 : 16; }
__host__ __device__ constexpr int w_chunk_bytes(int l, int c) { return wN(l) * w_chunk_kk(l, c) * 4; }
__host__ __device__ constexpr int w_chunk_offset(int l, int c) { return w_layer_offset(l) + c * wN(l) * 32 * 4; }
constexpr int kWStreamBytes = w_layer_offset(kWLayers);
static_assert(kWStreamBytes == 4 * 278528, "stream size");
// shared memory
constexpr int kFPlane = 4096;             // features, K-major [2 k-chunks][128 rows][16 B]
constexpr int oFeat = kWStages * kWStage; // [hi | lo]
constexpr int oBars = oFeat + 2 * kFPlane;                     // full[5] empty[5] a_ready mma_done
constexpr int oTmemPtr = oBars + 8 * (2 * kWStages + 2);
constexpr int oBias = oTmemPtr + 16;                           // b1[256] b2[256] b3[1:129] b4[256] (x16) | b3[0] b5[3]
constexpr int oW5 = oBias + 4 * (3 * 256 + 128 + 4);           // W5 [3][256]
constexpr int oW30 = oW5 + 4 * 3 * 256;                        // W3 row 0 [256]
constexpr int oHead = oW30 + 4 * 256;                          // [128 rows][4]: partial heads of the second column half
constexpr int kWSmem = oHead + 4 * 128 * 4;
static_assert(kWSmem <= 232448, "shared memory budget");
// scratch of one tile: 128-feature operand blocks of 64 kB in the order below, then F and G5 (8 kB each)
constexpr int bH1 = 0, bH2 = 2, bHC = 4, bT = 6, bG1 = 7, bG2 = 9, bG4 = 11, bGt = 13, kBlocks = 14;
constexpr size_t oFsm = (size_t)kBlocks * kOpBytes, oG5sm = oFsm + kSmallBytes, kWTile = oG5sm + kSmallBytes;
constexpr int kWMaskBytes = 3 * 2 * 4 * 128 * 4;               // [layer h1, h2, hc][column half][32-column word][row]
constexpr size_t kFeatTile = 128 * 16 * sizeof(float);
}  // namespace w2

// source element of packed layer l at (output row n, reduction index k)
static __device__ __forceinline__ float w2_weight(const pslam_decoder_t &d, int l, int n, int k)
{
    switch (l) {
        case 0: return d.W1[n * 16 + k];
        case 1: return d.W2[n * 256 + k];
        case 2: return d.W3[(1 + n) * 256 + k];
        case 3: return d.W4[n * 144 + k];
        case 4: return d.W4[k * 144 + 128 + n];
        case 5: return d.W4[k * 144 + n];
        case 6: return d.W3[(1 + k) * 256 + n];
        case 7: return d.W2[k * 256 + n];
        default: return d.W1[k * 16 + n];
    }
}

// weights x16 as f16 hi / lo planes in the chunk order of the kernels: chunk c of a layer = reduction elements [32c, 32c + 32),
// [hi: kk/8 k-chunks x N rows x 8 halves | lo: same] -- K-major, no swizzle
__global__ void k_w2_pack(pslam_decoder_t d, uint16_t *__restrict__ out, int *__restrict__ range_flag)
{
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int base = 0;   // uint16 offset of the layer
#pragma unroll 1
    for (int l = 0; l < w2::kWLayers; ++l) {
        const int N = w2::wN(l), K = w2::wK(l);
        if (i < N * K) {
            const int n = i / K, k = i % K;
            const int c = k >> 5, kk = w2::w_chunk_kk(l, c), kr = k & 31;
            uint32_t hi, lo;
            const float ws = bf::kScale * w2_weight(d, l, n, k);
            if (range_flag && fabsf(ws) >= 32752.0f) atomicOr(range_flag, 4);
            h16_split2(ws, 0.0f, hi, lo);
            uint16_t *chunk = out + base + c * (N * 32 * 2);
            const int off = (kr >> 3) * (N * 8) + n * 8 + (kr & 7);
            chunk[off] = (uint16_t)(hi & 0xffffu);
            chunk[N * kk + off] = (uint16_t)(lo & 0xffffu);
            return;
        }
        i -= N * K;
        base += 2 * N * K;
    }
}

// 16 accumulator columns D[dcol, dcol + 16) of this thread's row = features [f0, f0 + 16) of the layer's output:
// accumulators -> (bias / activation / mask) -> f16 hi / lo -> next A operand (tensor memory) and, optionally, the scratch.
//   MODE 0: y = relu(D/16 + bias[f])   MODE 1: y = D/16 + bias[f]   MODE 2: y = mask ? D/16 (+ r1 * wx[f]) : 0   MODE 3: y = D/16
//   EXTRA 1: acc[0] += wx[f] * y     EXTRA 2: acc[0..2] += W5[.][f] * y, no A operand     EXTRA 3: rank-1 term r1 * wx[f] before the mask
template <int MODE, int EXTRA>
__device__ __forceinline__ void w2_epi16(uint32_t tg, int dcol, int f0, const float *bias, uint32_t &mask, int shift, unsigned char *stg,
                                         int stg_block, float &ymax, const float *wx, float *acc, float r1)
{
    using namespace w2;
    uint32_t v[16];
    tmem_ld16(tg + cAcc + dcol, v);
    tmem_wait_ld();
    uint32_t bits = 0u;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        float y = __uint_as_float(v[e]);
        if (MODE == 0) { y = fmaxf(fmaf(y, kInvScale, bias[f0 + e]), 0.0f); bits |= (y > 0.0f ? 1u : 0u) << e; }
        if (MODE == 1) y = fmaf(y, kInvScale, bias[f0 + e]);
        if (MODE == 2) {
            y = (EXTRA == 3) ? fmaf(r1, wx[f0 + e], y * kInvScale) : y * kInvScale;
            y = ((mask >> (shift + e)) & 1u) ? y : 0.0f;
        }
        if (MODE == 3) y = y * kInvScale;
        ymax = fmaxf(ymax, fabsf(y));
        if (EXTRA == 1) acc[0] = fmaf(wx[f0 + e], y, acc[0]);
        if (EXTRA == 2) {
            acc[0] = fmaf(wx[f0 + e], y, acc[0]);
            acc[1] = fmaf(wx[256 + f0 + e], y, acc[1]);
            acc[2] = fmaf(wx[512 + f0 + e], y, acc[2]);
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
            unsigned char *dst = stg + (size_t)(stg_block / 8 + j) * 128;     // stg_block: feature index inside the 128-feature block
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4 *>(dst + 16384) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
    }
    if (EXTRA != 2) {
        tmem_st8(tg + cHi + f0 / 2, hi);
        tmem_st8(tg + cLo + f0 / 2, lo);
    }
}

// The MMA-issuing warp: ring position, and the phases issued so far
struct W2Issuer {
    unsigned char *smem;
    uint64_t *full, *empty, *a_ready, *mma_done;
    uint32_t tmem;
    int stage, phase;
    uint32_t uses;

    __device__ __forceinline__ void wait_operand()
    {
        mbar_wait(a_ready, uses & 1u);
        fence_after_sync();
        ++uses;
    }
    // all chunks of packed layer L into D[DCOL, DCOL + N); FRESH: the first MMA overwrites the accumulator
    template <int L, int DCOL>
    __device__ __forceinline__ void layer()
    {
        using namespace w2;
        constexpr int N = wN(L), K = wK(L), NCH = w_chunks(L);
        const uint32_t idesc = idesc_h16(128, N);
        const uint32_t a_hi = tmem + cHi, a_lo = tmem + cLo, d = tmem + cAcc + DCOL;
        const uint32_t fbase = smem_u32(smem + oFeat);
        const uint64_t f_hi = sdesc(fbase, 2048, 128), f_lo = sdesc(fbase + kFPlane, 2048, 128);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            constexpr bool small = K < 32;
            const int kk = (small || (K == 144 && c == 4)) ? 16 : 32;
            const bool feat = small || (K == 144 && c == 4);           // the reduction runs over the 16 input features (shared-memory A)
            mbar_wait(full + stage, phase);
            fence_after_sync();
            const uint64_t b0 = sdesc(smem_u32(smem + stage * kWStage), N * 16, 128);
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (s * 16 < kk) {
                        const uint64_t b_hi = b0 + (uint64_t)((s * (2 * N * 16)) >> 4);
                        const uint64_t b_lo = b0 + (uint64_t)((N * kk * 2 + s * (2 * N * 16)) >> 4);
                        const uint32_t acc0 = (c == 0 && s == 0) ? 0u : 1u;
                        if (feat) {
                            mma_h16_ss(d, f_lo, b_hi, idesc, acc0);
                            mma_h16_ss(d, f_hi, b_lo, idesc, 1u);
                            mma_h16_ss(d, f_hi, b_hi, idesc, 1u);
                        } else {
                            const uint32_t acol = (uint32_t)(16 * (2 * c + s)) >> 1;
                            mma_h16_ts(d, a_lo + acol, b_hi, idesc, acc0);
                            mma_h16_ts(d, a_hi + acol, b_lo, idesc, 1u);
                            mma_h16_ts(d, a_hi + acol, b_hi, idesc, 1u);
                        }
                    }
                }
                mma_commit_mcast(empty + stage, kClusterMask);
            }
            __syncwarp();
            if (++stage == kWStages) { stage = 0; phase ^= 1; }
        }
    }
    __device__ __forceinline__ void done()
    {
        if (elect_one()) mma_commit(mma_done);
        __syncwarp();
    }
};

// KIND: bf::kFwd (plain forward), bf::kFwdSave (+ ReLU masks, + wgrad operands when p.spill_ops), bf::kBwdSaved (dgrad chain)
template <int KIND>
__global__ void __cluster_dims__(bf::kCluster, 1, 1) __launch_bounds__(w2::kWThreads, 1)
k_field_w256(FieldParams p, const unsigned char *__restrict__ wstream)
{
    pdl_enter();
    using namespace w2;
    constexpr bool kIsFwd = KIND != kBwdSaved;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *empty = full + kWStages;
    uint64_t *a_ready = empty + kWStages;    // the next A operand is complete (and the accumulator columns may be overwritten)
    uint64_t *mma_done = a_ready + 1;        // the phase's accumulators are complete
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oTmemPtr);
    float *sBias = reinterpret_cast<float *>(smem + oBias);
    float *sW5 = reinterpret_cast<float *>(smem + oW5), *sW30 = reinterpret_cast<float *>(smem + oW30);
    float *sHead = reinterpret_cast<float *>(smem + oHead);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    const int G = (int)gridDim.x;
    const int iters = (ntiles + G - 1) / G;            // the CTAs of a cluster share one weight stream: same iterations everywhere
    const uint32_t crank = cluster_ctarank();

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kCluster); }
        mbar_init(a_ready, kWorkers);
        mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    const bool spill = p.wg_scratch != nullptr && p.spill_ops != 0;
    for (int i = threadIdx.x; i < 3 * 256 + 128 + 4; i += kWThreads) {
        float v;
        if (i < 256) v = p.dec.b1[i];
        else if (i < 512) v = p.dec.b2[i - 256];
        else if (i < 640) v = p.dec.b3[1 + i - 512];
        else if (i < 896) v = p.dec.b4[i - 640];
        else if (i == 896) v = p.dec.b3[0];
        else v = p.dec.b5[i - 897];
        sBias[i] = i < 896 ? kScale * v : v;
    }
    for (int i = threadIdx.x; i < 3 * 256; i += kWThreads) sW5[i] = p.dec.W5[i];
    for (int i = threadIdx.x; i < 256; i += kWThreads) sW30[i] = p.dec.W3[i];
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    cluster_sync();
    const uint32_t tmem = *tmem_ptr;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
        if (warp == 0) {
            // ===================== TMA producer: the chunks in the issuer's order, multicast over the cluster =====================
            int stage = 0, phase = 0;
            auto emit = [&](int l) {
                for (int c = 0; c < w_chunks(l); ++c) {
                    const uint32_t bytes = (uint32_t)w_chunk_bytes(l, c), part = bytes / kCluster;
                    const unsigned char *src = wstream + w_chunk_offset(l, c);
                    mbar_wait(empty + stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + stage, bytes);
                        bulk_g2s_mcast(smem + stage * kWStage + crank * part, src + crank * part, part, full + stage, kClusterMask);
                    }
                    __syncwarp();
                    if (++stage == kWStages) { stage = 0; phase ^= 1; }
                }
            };
            for (int it = 0; it < iters; ++it) {
                if (kIsFwd) { emit(0); emit(1); emit(2); emit(3); }
                else { emit(4); emit(5); emit(6); emit(7); emit(8); }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            W2Issuer mi{smem, full, empty, a_ready, mma_done, tmem, 0, 0, 0u};
            for (int it = 0; it < iters; ++it) {
                if constexpr (kIsFwd) {
                    mi.wait_operand(); mi.template layer<0, 0>(); mi.done();     // features -> h1
                    mi.wait_operand(); mi.template layer<1, 0>(); mi.done();     // h1 -> h2
                    mi.wait_operand(); mi.template layer<2, 0>(); mi.done();     // h2 -> t (D[0,128))
                    mi.wait_operand(); mi.template layer<3, 0>(); mi.done();     // [t; features] -> hc
                } else {
                    mi.wait_operand(); mi.template layer<4, 0>(); mi.template layer<5, 128>(); mi.done();   // g_hc -> g_f part D[0,16), g_t D[128,256)
                    mi.wait_operand(); mi.template layer<6, 0>(); mi.done();     // g_t -> g_h2
                    mi.wait_operand(); mi.template layer<7, 0>(); mi.done();     // g_h2 -> g_h1
                    mi.wait_operand(); mi.template layer<8, 0>(); mi.done();     // g_h1 -> g_f part D[0,16)
                }
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsWorker));
        // ===================== workers: thread = (sample row m, column half h) =====================
        const int h = (warp - 4) >> 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;
        const bool lead = h == 0;
        const uint32_t tg = tmem + ((uint32_t)(q * 32) << 16);
        const int rowoff_big = (m >> 6) * 32768 + ((m >> 3) & 7) * 2048 + (m & 7) * 16;
        const int rowoff_small = (m >> 6) * 4096 + ((m >> 3) & 7) * 256 + (m & 7) * 16;
        unsigned char *sF = smem + oFeat + m * 16;    // this row's 16 B of k-chunk 0, hi plane
        uint32_t done_uses = 0;
        float ymax = 0.0f;
        const float Sg = kIsFwd ? 1.0f : grad_scale(p.gscale), invSg = 1.0f / Sg;
        auto layer_done = [&]() {
            mbar_wait(mma_done, done_uses & 1u);
            ++done_uses;
            fence_after_sync();
        };
        auto a_is_ready = [&]() {                      // this thread's tensor-memory writes (and accumulator reads) are complete
            tmem_wait_st();
            fence_before_sync();
            mbar_arrive(a_ready);
        };
        uint32_t nomask = 0u;
        for (int it = 0; it < iters; ++it) {
            const int tile = it * G + (int)blockIdx.x;
            const bool real_tile = tile < ntiles;
            const int s = real_tile ? tile * 128 + m : nsamp;
            const bool valid = s < nsamp;
            unsigned char *scr = (spill && real_tile) ? p.wg_scratch + (size_t)tile * kWTile : nullptr;
            // operand block `op + h` for the 256-feature operands, `op` for the 128-feature ones
            auto stg_of = [&](int op) -> unsigned char * { return scr ? scr + (size_t)op * kOpBytes + rowoff_big : nullptr; };
            uint32_t *mk = (p.act_masks && real_tile) ? p.act_masks + (size_t)tile * (kWMaskBytes / 4) + h * 512 + m : nullptr;   // [layer][half][word][row]
            if constexpr (kIsFwd) {
                uint32_t m1[4] = {0u, 0u, 0u, 0u}, m2[4] = {0u, 0u, 0u, 0u}, mc[4] = {0u, 0u, 0u, 0u};
                // ---- features -> shared-memory A operand (x16, hi / lo); the lead thread of a row does it ----
                if (lead) {
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 v = valid ? __ldg(reinterpret_cast<const float4 *>(p.feat + (size_t)s * 16) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                        f[4 * e] = v.x; f[4 * e + 1] = v.y; f[4 * e + 2] = v.z; f[4 * e + 3] = v.w;
                    }
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
                        unsigned char *dst = scr + oFsm + rowoff_small;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            *reinterpret_cast<uint4 *>(dst + j * 128) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                            *reinterpret_cast<uint4 *>(dst + 2048 + j * 128) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> the tensor core's reads
                }
                fence_before_sync();
                mbar_arrive(a_ready);
                float sdf_acc[1] = {0.0f}, head[3] = {0.f, 0.f, 0.f};
                // ---- h1 ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 4; ++w) {
                    uint32_t mw = 0u;
                    const int f0 = 128 * h + 32 * w;
                    w2_epi16<0, 0>(tg, f0, f0, sBias, mw, 0, stg_of(bH1 + h), 32 * w, ymax, nullptr, nullptr, 0.f);
                    w2_epi16<0, 0>(tg, f0 + 16, f0 + 16, sBias, mw, 16, stg_of(bH1 + h), 32 * w + 16, ymax, nullptr, nullptr, 0.f);
                    m1[0] = w == 0 ? mw : m1[0]; m1[1] = w == 1 ? mw : m1[1]; m1[2] = w == 2 ? mw : m1[2]; m1[3] = w == 3 ? mw : m1[3];
                }
                a_is_ready();
                // ---- h2 (+ this thread's share of the sdf head) ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 4; ++w) {
                    uint32_t mw = 0u;
                    const int f0 = 128 * h + 32 * w;
                    w2_epi16<0, 1>(tg, f0, f0, sBias + 256, mw, 0, stg_of(bH2 + h), 32 * w, ymax, sW30, sdf_acc, 0.f);
                    w2_epi16<0, 1>(tg, f0 + 16, f0 + 16, sBias + 256, mw, 16, stg_of(bH2 + h), 32 * w + 16, ymax, sW30, sdf_acc, 0.f);
                    m2[0] = w == 0 ? mw : m2[0]; m2[1] = w == 1 ? mw : m2[1]; m2[2] = w == 2 ? mw : m2[2]; m2[3] = w == 3 ? mw : m2[3];
                }
                a_is_ready();
                // ---- t: 128 features, 64 per thread ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 2; ++w) {
                    const int f0 = 64 * h + 32 * w;
                    w2_epi16<1, 0>(tg, f0, f0, sBias + 512, nomask, 0, stg_of(bT), f0, ymax, nullptr, nullptr, 0.f);
                    w2_epi16<1, 0>(tg, f0 + 16, f0 + 16, sBias + 512, nomask, 0, stg_of(bT), f0 + 16, ymax, nullptr, nullptr, 0.f);
                }
                a_is_ready();
                // ---- hc + colour head ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 4; ++w) {
                    uint32_t mw = 0u;
                    const int f0 = 128 * h + 32 * w;
                    w2_epi16<0, 2>(tg, f0, f0, sBias + 640, mw, 0, stg_of(bHC + h), 32 * w, ymax, sW5, head, 0.f);
                    w2_epi16<0, 2>(tg, f0 + 16, f0 + 16, sBias + 640, mw, 16, stg_of(bHC + h), 32 * w + 16, ymax, sW5, head, 0.f);
                    mc[0] = w == 0 ? mw : mc[0]; mc[1] = w == 1 ? mw : mc[1]; mc[2] = w == 2 ? mw : mc[2]; mc[3] = w == 3 ? mw : mc[3];
                }
                // the two threads of a row combine their head sums
                if (!lead) *reinterpret_cast<float4 *>(sHead + m * 4) = make_float4(head[0], head[1], head[2], sdf_acc[0]);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (lead && valid) {
                    const float4 o = *reinterpret_cast<const float4 *>(sHead + m * 4);
                    const float r = sigmoid_f(fmaf(head[0] + o.x, kInvScale, sBias[897]));
                    const float gg = sigmoid_f(fmaf(head[1] + o.y, kInvScale, sBias[898]));
                    const float b = sigmoid_f(fmaf(head[2] + o.z, kInvScale, sBias[899]));
                    const float sdf = fmaf(sdf_acc[0] + o.w, kInvScale, sBias[896]);
                    *reinterpret_cast<float4 *>(p.out + (size_t)s * 4) = make_float4(r, gg, b, sdf);
                }
                if (KIND == kFwdSave && mk) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) { mk[w * 128] = m1[w]; mk[1024 + w * 128] = m2[w]; mk[2048 + w * 128] = mc[w]; }
                }
                // (the next tile's features overwrite sF and its heads sHead: layer 4's MMAs have completed, and no thread passes the
                //  next tile's first layer_done before every lead thread has arrived on a_ready after reading sHead)
            } else {
                uint32_t m1[4] = {0u, 0u, 0u, 0u}, m2[4] = {0u, 0u, 0u, 0u}, mc[4] = {0u, 0u, 0u, 0u};
                if (mk) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) { m1[w] = mk[w * 128]; m2[w] = mk[1024 + w * 128]; mc[w] = mk[2048 + w * 128]; }
                }
                const float4 po = valid ? __ldg(reinterpret_cast<const float4 *>(p.out + (size_t)s * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 go = valid ? __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)s * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                go.x *= Sg; go.y *= Sg; go.z *= Sg; go.w *= Sg;
                const float g5[4] = {go.x * (1.0f - po.x) * po.x, go.y * (1.0f - po.y) * po.y, go.z * (1.0f - po.z) * po.z, go.w};
                if (lead && scr) {
                    uint32_t h0, l0, h1w, l1w;
                    h16_split2(g5[0], g5[1], h0, l0);
                    h16_split2(g5[2], g5[3], h1w, l1w);
                    unsigned char *dst = scr + oG5sm + rowoff_small;
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048) = make_uint4(l0, l1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(dst + 2048 + 128) = make_uint4(0u, 0u, 0u, 0u);
                }
                if (lead && p.grad_dec) {   // bias gradients of the two heads: column sums of G5
                    const float s0 = warp_sum(g5[0]), s1 = warp_sum(g5[1]), s2 = warp_sum(g5[2]), s3 = warp_sum(g5[3]);
                    if (lane == 0 && real_tile) {
                        atomicAdd(p.g_dec.b5 + 0, s0 * invSg); atomicAdd(p.g_dec.b5 + 1, s1 * invSg); atomicAdd(p.g_dec.b5 + 2, s2 * invSg);
                        atomicAdd(p.g_dec.b3, s3 * invSg);
                    }
                }
                {
                    // g_hc = mask_hc . (W5^T g5) on the CUDA cores -> A (+ G4 block h)
                    unsigned char *stg = stg_of(bG4 + h);
#pragma unroll 1
                    for (int j = 0; j < 8; ++j) {
                        const int c0 = 128 * h + 16 * j;
                        const uint32_t bits = mc[j >> 1] >> ((j & 1) * 16);
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float y0 = fmaf(g5[2], sW5[512 + c0 + 2 * e], fmaf(g5[1], sW5[256 + c0 + 2 * e], g5[0] * sW5[c0 + 2 * e]));
                            float y1 = fmaf(g5[2], sW5[512 + c0 + 2 * e + 1], fmaf(g5[1], sW5[256 + c0 + 2 * e + 1], g5[0] * sW5[c0 + 2 * e + 1]));
                            y0 = ((bits >> (2 * e)) & 1u) ? y0 : 0.0f;
                            y1 = ((bits >> (2 * e + 1)) & 1u) ? y1 : 0.0f;
                            ymax = fmaxf(ymax, fmaxf(fabsf(y0), fabsf(y1)));
                            h16_split2(y0, y1, hi[e], lo[e]);
                        }
                        if (stg) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                unsigned char *dst = stg + (size_t)(2 * j + k) * 128;
                                *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                                *reinterpret_cast<uint4 *>(dst + 16384) = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
                            }
                        }
                        tmem_st8(tg + cHi + c0 / 2, hi);
                        tmem_st8(tg + cLo + c0 / 2, lo);
                    }
                }
                a_is_ready();
                // ---- g_f part through W4's feature columns (D[0,16), lead) and g_t (D[128,256): 64 features per thread) ----
                layer_done();
                float gf[16];
                if (lead) {
                    uint32_t v[16];
                    tmem_ld16(tg + cAcc, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) gf[e] = __uint_as_float(v[e]);
                }
#pragma unroll 1
                for (int w = 0; w < 2; ++w) {
                    const int f0 = 64 * h + 32 * w;
                    w2_epi16<3, 0>(tg, 128 + f0, f0, nullptr, nomask, 0, stg_of(bGt), f0, ymax, nullptr, nullptr, 0.f);
                    w2_epi16<3, 0>(tg, 128 + f0 + 16, f0 + 16, nullptr, nomask, 0, stg_of(bGt), f0 + 16, ymax, nullptr, nullptr, 0.f);
                }
                a_is_ready();
                // ---- g_h2 (+ the sdf head's rank-1 term) ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 4; ++w) {
                    uint32_t mw = m2[0];
                    mw = w == 1 ? m2[1] : mw; mw = w == 2 ? m2[2] : mw; mw = w == 3 ? m2[3] : mw;
                    const int f0 = 128 * h + 32 * w;
                    w2_epi16<2, 3>(tg, f0, f0, nullptr, mw, 0, stg_of(bG2 + h), 32 * w, ymax, sW30, nullptr, go.w);
                    w2_epi16<2, 3>(tg, f0 + 16, f0 + 16, nullptr, mw, 16, stg_of(bG2 + h), 32 * w + 16, ymax, sW30, nullptr, go.w);
                }
                a_is_ready();
                // ---- g_h1 ----
                layer_done();
#pragma unroll 1
                for (int w = 0; w < 4; ++w) {
                    uint32_t mw = m1[0];
                    mw = w == 1 ? m1[1] : mw; mw = w == 2 ? m1[2] : mw; mw = w == 3 ? m1[3] : mw;
                    const int f0 = 128 * h + 32 * w;
                    w2_epi16<2, 0>(tg, f0, f0, nullptr, mw, 0, stg_of(bG1 + h), 32 * w, ymax, nullptr, nullptr, 0.f);
                    w2_epi16<2, 0>(tg, f0 + 16, f0 + 16, nullptr, mw, 16, stg_of(bG1 + h), 32 * w + 16, ymax, nullptr, nullptr, 0.f);
                }
                a_is_ready();
                // ---- g_f: the part through W1 joins the part through W4's feature columns ----
                layer_done();
                if (lead) {
                    uint32_t v[16];
                    tmem_ld16(tg + cAcc, v);
                    tmem_wait_ld();
                    if (valid && p.g_feat) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) gf[e] = (gf[e] + __uint_as_float(v[e])) * (kInvScale * invSg);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4 *>(p.g_feat + (size_t)s * 16 + 4 * j) = make_float4(gf[4 * j], gf[4 * j + 1], gf[4 * j + 2], gf[4 * j + 3]);
                    }
                }
                // (the next tile's g_hc overwrites A, and its first layer D: this tile's last MMAs have completed and been read)
            }
        }
        if (p.range_flag && ymax >= 32752.0f) atomicOr(p.range_flag, 4);
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc(tmem, bf::kTmemCols);
}

// ------------------------------------------------------------------------------------------
// Weight gradients of the width-256 decoder: dW[n][k] = sum over samples of G[s][n] * A[s][k], MMAs whose reduction dimension is
// the SAMPLE index, both operands MN-major from shared memory straight out of the scratch (128-feature operand blocks, one
// 64-sample half = 32 kB).  A CTA has a ROLE = a list of sub-steps per (tile, half); a sub-step loads up to three operand blocks
// (+ F and G5) into a stage and issues up to four products into accumulators that stay in tensor memory for all the CTA's tiles:
//   role 0  dW2 = G2^T H1                       (4 accumulators of 128 columns)
//   role 1  dW3[1:] = Gt^T H2 ; dW4[:, :128] = G4^T T
//   role 2  dW4[:, 128:] = G4^T F ; dW1 = G1^T F ; dW5 = (HC^T G5)[:, 0:3] ; dW3[0] = (H2^T G5)[:, 3] ; bias column sums of
//           G4, G1, Gt, G2 (MMAs against a ones operand)
// ------------------------------------------------------------------------------------------
namespace wg2 {
constexpr int kThreads = 192;                   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 drain
constexpr int kStages = 2;
constexpr int kHalf = 32768, kSmallHalf = 4096;
constexpr int kStage = 3 * kHalf + 2 * kSmallHalf;          // 106 496
constexpr int oOnes = kStages * kStage;         // 16 samples x 16 features of f16 1.0
constexpr int oBars = oOnes + 512;              // full[2] free[2] all_done, tmem ptr
constexpr int kSmem = oBars + 128;
constexpr int kMaxProd = 6;
// a product: accumulator columns [col, col + (n16 ? 16 : 128)) += X^T Y, X = stage slot x (always a 128-feature block),
// Y = stage slot y (0..2), 3 = F, 4 = G5, 5 = ones
struct Prod { short x, y, col, n16; };
struct Sub { short op[3]; short small; short nprod; Prod prod[kMaxProd]; };   // op: scratch block per stage slot (-1: unused)
struct Role { int nsub; Sub sub[5]; };
using w2::bH1; using w2::bH2; using w2::bHC; using w2::bT; using w2::bG1; using w2::bG2; using w2::bG4; using w2::bGt;
__device__ __constant__ Role cRoles[3] = {
    {2, {{{bG2, bH1, bH1 + 1}, 0, 2, {{0, 1, 0, 0}, {0, 2, 128, 0}}},
         {{bG2 + 1, bH1, bH1 + 1}, 0, 2, {{0, 1, 256, 0}, {0, 2, 384, 0}}}}},
    {2, {{{bGt, bH2, bH2 + 1}, 0, 2, {{0, 1, 0, 0}, {0, 2, 128, 0}}},
         {{bT, bG4, bG4 + 1}, 0, 2, {{1, 0, 256, 0}, {2, 0, 384, 0}}}}},
    {5, {{{bG4, bG4 + 1, -1}, 1, 4, {{0, 3, 0, 1}, {1, 3, 16, 1}, {0, 5, 128, 1}, {1, 5, 144, 1}}},
         {{bG1, bG1 + 1, -1}, 1, 4, {{0, 3, 32, 1}, {1, 3, 48, 1}, {0, 5, 160, 1}, {1, 5, 176, 1}}},
         {{bHC, bHC + 1, -1}, 1, 2, {{0, 4, 64, 1}, {1, 4, 80, 1}}},
         {{bH2, bH2 + 1, bGt}, 1, 3, {{0, 4, 96, 1}, {1, 4, 112, 1}, {2, 5, 192, 1}}},
         {{bG2, bG2 + 1, -1}, 0, 2, {{0, 5, 208, 1}, {1, 5, 224, 1}}}}}};
}  // namespace wg2

__global__ void __launch_bounds__(wg2::kThreads, 1) k_wgrad_w256(FieldParams p, int n0, int n1)
{
    pdl_enter();
    using namespace wg2;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *freeb = full + kStages, *all_done = freeb + kStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oBars + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    // CTAs [0, n0) take role 0, [n0, n0 + n1) role 1, the rest role 2; inside a role, CTA j takes tiles j, j + group, ...
    const int b = (int)blockIdx.x;
    const int role = b < n0 ? 0 : (b < n0 + n1 ? 1 : 2);
    const int j = role == 0 ? b : (role == 1 ? b - n0 : b - n0 - n1);
    const int group = role == 0 ? n0 : (role == 1 ? n1 : (int)gridDim.x - n0 - n1);
    const Role &R = cRoles[role];
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
    const int my_tiles = (j < ntiles) ? (ntiles - 1 - j) / group + 1 : 0;
    const int nsteps = my_tiles * 2 * R.nsub;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int tl = 0, h = 0, su = 0;
        for (int g = 0; g < nsteps; ++g) {
            const int rs = g % kStages, use = g / kStages;
            const Sub &S = R.sub[su];
            const unsigned char *tile = p.wg_scratch + (size_t)(j + (size_t)tl * group) * w2::kWTile;
            if (use >= 1) mbar_wait(freeb + rs, (use - 1) & 1);
            if (elect_one()) {
                unsigned char *dst = smem + rs * kStage;
                uint32_t bytes = 0;
                for (int k = 0; k < 3; ++k) bytes += S.op[k] >= 0 ? kHalf : 0;
                bytes += S.small ? 2 * kSmallHalf : 0;
                mbar_arrive_expect_tx(full + rs, bytes);
                for (int k = 0; k < 3; ++k)
                    if (S.op[k] >= 0) bulk_g2s(dst + k * kHalf, tile + (size_t)S.op[k] * bf::kOpBytes + (size_t)h * kHalf, kHalf, full + rs);
                if (S.small) {
                    bulk_g2s(dst + 3 * kHalf, tile + w2::oFsm + (size_t)h * kSmallHalf, kSmallHalf, full + rs);
                    bulk_g2s(dst + 3 * kHalf + kSmallHalf, tile + w2::oG5sm + (size_t)h * kSmallHalf, kSmallHalf, full + rs);
                }
            }
            __syncwarp();
            if (++su == R.nsub) { su = 0; if (++h == 2) { h = 0; ++tl; } }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t id128 = idesc_h16(128, 128, 1, 1), id16 = idesc_h16(128, 16, 1, 1);
        const uint64_t ones = sdesc(smem_u32(smem + oOnes), 256, 128);
        int su = 0;
        for (int g = 0; g < nsteps; ++g) {
            const int rs = g % kStages;
            const Sub &S = R.sub[su];
            mbar_wait(full + rs, (g / kStages) & 1);
            fence_after_sync();
            const uint32_t base = smem_u32(smem + rs * kStage);
            if (elect_one()) {
                for (int ks = 0; ks < 4; ++ks) {            // 16 samples = 2 kb blocks per MMA
                    const uint32_t fresh = (g < R.nsub && ks == 0) ? 0u : 1u;   // first touch of this sub-step's accumulators
                    for (int q = 0; q < S.nprod; ++q) {
                        const Prod P = S.prod[q];
                        const uint32_t xb = base + P.x * kHalf + ks * 4096;
                        const uint64_t x_hi = sdesc(xb, 2048, 128), x_lo = sdesc(xb + 16384, 2048, 128);
                        const uint32_t idesc = P.n16 ? id16 : id128;
                        if (P.y == 5) {                     // column sums: X^T ones
                            mma_h16_ss(tmem + P.col, x_hi, ones, id16, fresh);
                            mma_h16_ss(tmem + P.col, x_lo, ones, id16, 1u);
                            continue;
                        }
                        uint64_t y_hi, y_lo;
                        if (P.y < 3) {
                            const uint32_t yb = base + P.y * kHalf + ks * 4096;
                            y_hi = sdesc(yb, 2048, 128); y_lo = sdesc(yb + 16384, 2048, 128);
                        } else {
                            const uint32_t yb = base + 3 * kHalf + (P.y - 3) * kSmallHalf + ks * 512;
                            y_hi = sdesc(yb, 256, 128); y_lo = sdesc(yb + 2048, 256, 128);
                        }
                        mma_h16_ss(tmem + P.col, x_lo, y_hi, idesc, fresh);
                        mma_h16_ss(tmem + P.col, x_hi, y_lo, idesc, 1u);
                        mma_h16_ss(tmem + P.col, x_hi, y_hi, idesc, 1u);
                    }
                }
                mma_commit(freeb + rs);
                if (g == nsteps - 1) mma_commit(all_done);
            }
            __syncwarp();
            if (++su == R.nsub) su = 0;
        }
    } else if (nsteps > 0) {
        // ===================== drain: accumulators -> global gradients (warp & 3 = TMEM lane quarter) =====================
        mbar_wait(all_done, 0);
        fence_after_sync();
        const int qd = warp & 3;
        const int n = qd * 32 + lane;                            // accumulator row = TMEM lane = feature of the X block
        const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
        // accumulators hold 16 x Sg x (sum of products), the column sums Sg x (sum): both factors are powers of two
        const float invSg = 1.0f / grad_scale(p.gscale), cW = bf::kInvScale * invSg;
        auto flush = [&](int col0, int ncols, float *dst_row) {   // dst_row: &dW[row][0], ncols % 16 == 0
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + col0 + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    red_add_v4(dst_row + c0 + 4 * k, cW * __uint_as_float(v[4 * k]), cW * __uint_as_float(v[4 * k + 1]),
                               cW * __uint_as_float(v[4 * k + 2]), cW * __uint_as_float(v[4 * k + 3]));
            }
        };
        auto col0_of = [&](int col) { uint32_t v[8]; tmem_ld8(trow + col, v); tmem_wait_ld(); return __uint_as_float(v[0]); };
        if (role == 0) {
            flush(0, 128, p.g_dec.W2 + (size_t)n * 256);                  // G2 block 0 x H1 block 0
            flush(128, 128, p.g_dec.W2 + (size_t)n * 256 + 128);
            flush(256, 128, p.g_dec.W2 + (size_t)(128 + n) * 256);
            flush(384, 128, p.g_dec.W2 + (size_t)(128 + n) * 256 + 128);
        } else if (role == 1) {
            flush(0, 128, p.g_dec.W3 + (size_t)(1 + n) * 256);            // Gt x H2 block 0 / 1
            flush(128, 128, p.g_dec.W3 + (size_t)(1 + n) * 256 + 128);
            flush(256, 128, p.g_dec.W4 + (size_t)n * 144);                // G4 block 0 / 1 x T
            flush(384, 128, p.g_dec.W4 + (size_t)(128 + n) * 144);
        } else {
            flush(0, 16, p.g_dec.W4 + (size_t)n * 144 + 128);             // G4 x F
            flush(16, 16, p.g_dec.W4 + (size_t)(128 + n) * 144 + 128);
            flush(32, 16, p.g_dec.W1 + (size_t)n * 16);                   // G1 x F
            flush(48, 16, p.g_dec.W1 + (size_t)(128 + n) * 16);
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t v[8], w[8];
                tmem_ld8(trow + 64 + 16 * blk, v);                        // HC^T G5: columns 0..2 = dW5 rows
                tmem_ld8(trow + 96 + 16 * blk, w);                        // H2^T G5: column 3 = dW3 row 0
                tmem_wait_ld();
                atomicAdd(p.g_dec.W5 + 128 * blk + n, cW * __uint_as_float(v[0]));
                atomicAdd(p.g_dec.W5 + 256 + 128 * blk + n, cW * __uint_as_float(v[1]));
                atomicAdd(p.g_dec.W5 + 512 + 128 * blk + n, cW * __uint_as_float(v[2]));
                atomicAdd(p.g_dec.W3 + 128 * blk + n, cW * __uint_as_float(w[3]));
                atomicAdd(p.g_dec.b4 + 128 * blk + n, invSg * col0_of(128 + 16 * blk));
                atomicAdd(p.g_dec.b1 + 128 * blk + n, invSg * col0_of(160 + 16 * blk));
                atomicAdd(p.g_dec.b2 + 128 * blk + n, invSg * col0_of(208 + 16 * blk));
            }
            atomicAdd(p.g_dec.b3 + 1 + n, invSg * col0_of(192));
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int w2_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st, int *range_flag)
{
    launch_chain(k_w2_pack, dim3(ceil_div(w2::kWStreamBytes / 4, 256)), dim3(256), 0, st, d, reinterpret_cast<uint16_t *>(ws_tc), range_flag);
    PSLAM_CHECK_LAUNCH("w2_pack");
    return 0;
}

// scratch = [tiles x kWTile operands][tiles x kWMaskBytes ReLU masks][tiles x 8 kB features][tiles x 8 kB feature gradients][gradient scale]
size_t w2_scratch_bytes(int max_samples)
{
    return (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128) * (w2::kWTile + w2::kWMaskBytes + 2 * w2::kFeatTile) + 64;
}
static unsigned char *w2_masks(const FieldParams &fp, int max_samples)
{
    return fp.wg_scratch + (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128) * w2::kWTile;
}
static float *w2_feat(const FieldParams &fp, int max_samples, int which)
{
    const size_t tiles = (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128);
    return reinterpret_cast<float *>(fp.wg_scratch + tiles * (w2::kWTile + w2::kWMaskBytes) + (size_t)which * tiles * w2::kFeatTile);
}
static uint32_t *w2_gscale(const FieldParams &fp, int max_samples)
{
    const size_t tiles = (size_t)ceil_div(max_samples > 0 ? max_samples : 1, 128);
    return reinterpret_cast<uint32_t *>(fp.wg_scratch + tiles * (w2::kWTile + w2::kWMaskBytes + 2 * w2::kFeatTile));
}

struct W2DeviceState { bool configured[4]; int max_clusters[4]; bool wg_configured; };
static W2DeviceState g_w2_state[64] = {};

// Which workspace holds the masks (and operands) of the most recent saving forward, and for which sample outputs: a backward
// may start from them only if nothing rewrote either since (same rule as field_bf.cu: the record is a host-side decision,
// the data itself is stream-ordered).
struct W2Saved { const void *scratch, *out; int spilled; };
static W2Saved g_w2_saved[64] = {};
static std::mutex g_w2_mutex;
static void w2_saved_set(const void *scratch, const void *out, int spilled)
{
    std::lock_guard<std::mutex> lock(g_w2_mutex);
    g_w2_saved[current_device()] = W2Saved{scratch, out, spilled};
}
static void w2_saved_invalidate(const void *scratch, const void *out)
{
    std::lock_guard<std::mutex> lock(g_w2_mutex);
    W2Saved &r = g_w2_saved[current_device()];
    if ((scratch && scratch == r.scratch) || (out && out == r.out)) r = W2Saved{nullptr, nullptr, 0};
}
static bool w2_saved_matches(const void *scratch, const void *out, int need_spill)
{
    std::lock_guard<std::mutex> lock(g_w2_mutex);
    const W2Saved &r = g_w2_saved[current_device()];
    return scratch && scratch == r.scratch && out == r.out && (!need_spill || r.spilled);
}

template <int KIND>
static int launch_w2(const FieldParams &fp, int max_samples, cudaStream_t st)
{
    W2DeviceState &ds = g_w2_state[current_device()];
    if (!ds.configured[KIND]) {      // per device: the attribute belongs to the function on the current device
        cudaError_t e = cudaFuncSetAttribute(k_field_w256<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, w2::kWSmem);
        if (e != cudaSuccess) { set_error("field_w256: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / bf::kCluster * bf::kCluster);
        cfg.blockDim = dim3(w2::kWThreads);
        cfg.dynamicSmemBytes = w2::kWSmem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = bf::kCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, k_field_w256<KIND>, &cfg);
        if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / bf::kCluster; }
        ds.max_clusters[KIND] = n < num_sms() / bf::kCluster ? n : num_sms() / bf::kCluster;
        ds.configured[KIND] = true;
    }
    const int tiles = ceil_div(max_samples > 0 ? max_samples : 1, 128);
    int grid = ceil_div(tiles, bf::kCluster) * bf::kCluster;
    if (grid > ds.max_clusters[KIND] * bf::kCluster) grid = ds.max_clusters[KIND] * bf::kCluster;
    launch_chain(k_field_w256<KIND>, dim3(grid), dim3(w2::kWThreads), w2::kWSmem, st, fp, reinterpret_cast<const unsigned char *>(fp.ws_tc));
    PSLAM_CHECK_LAUNCH("field_w256");
    return 0;
}

// Can this call run on the tensor-core build?  The forward needs feature rows: given (stand-alone decoder) or gathered into
// the workspace (fused pipeline); a backward needs the masks its forward saved there.
bool w2_usable(const FieldParams &fp, int max_samples, bool bwd)
{
    const bool ws = fp.wg_scratch && fp.wg_scratch_bytes >= w2_scratch_bytes(max_samples);
    if (!bwd) return fp.feat != nullptr || (fp.paired && ws);
    return fp.paired && ws && w2_saved_matches(fp.wg_scratch, fp.out, fp.grad_dec);
}

int w2_launch_field_forward(const FieldParams &fp_in, int max_samples, cudaStream_t st, int part)
{
    FieldParams fp = fp_in;
    const bool ws = fp.wg_scratch && fp.wg_scratch_bytes >= w2_scratch_bytes(max_samples);
    if (!fp.feat) {
        float *feat = w2_feat(fp, max_samples, 0);
        if (part != 3) { if (int rc = launch_tri_gather_rows(fp, feat, max_samples, st)) return rc; }   // part 3 (profiling): the rows of the previous full forward
        if (part == 5) return 0;
        fp.feat = feat;
    }
    const bool save = fp.paired && ws && (fp.grad_dec || fp.grad_emb || fp.grad_rays);
    fp.spill_ops = save && fp.grad_dec;
    w2_saved_invalidate(fp.wg_scratch, fp.out);     // every forward rewrites its outputs (and, with a workspace, the feature rows)
    if (!save) { fp.wg_scratch = nullptr; fp.act_masks = nullptr; return launch_w2<bf::kFwd>(fp, max_samples, st); }
    fp.act_masks = reinterpret_cast<uint32_t *>(w2_masks(fp, max_samples));
    if (int rc = launch_w2<bf::kFwdSave>(fp, max_samples, st)) return rc;
    w2_saved_set(fp.wg_scratch, fp.out, fp.spill_ops);
    return 0;
}

int w2_launch_field_backward(const FieldParams &fp_in, int max_samples, cudaStream_t st, int part)
{
    FieldParams fp = fp_in;
    fp.spill_ops = fp.grad_dec;
    fp.gscale = fp.gmax_ready ? fp.gmax_ready : w2_gscale(fp, max_samples);
    float *g_feat = w2_feat(fp, max_samples, 1);
    if (part == 4) {   // profiling: the trilinear scatter alone
        if (fp.grad_emb || fp.grad_rays) return launch_tri_scatter_rows(fp, g_feat, max_samples, st);
        return 0;
    }
    if (part != 2) {
        if (!fp.gmax_ready && part != 3) { if (int rc = launch_grad_scale(fp, fp.gscale, st)) return rc; }
        FieldParams fps = fp;                           // the scatter kernel wants the sample tables, not the feature rows
        fp.act_masks = reinterpret_cast<uint32_t *>(w2_masks(fp, max_samples));
        fp.feat = w2_feat(fp, max_samples, 0);
        fp.g_feat = g_feat;
        if (int rc = launch_w2<bf::kBwdSaved>(fp, max_samples, st)) return rc;
        if ((fp.grad_emb || fp.grad_rays) && part != 3) { if (int rc = launch_tri_scatter_rows(fps, g_feat, max_samples, st)) return rc; }
    }
    if (!fp.grad_dec || part == 1 || part == 3) return 0;
    W2DeviceState &ds = g_w2_state[current_device()];
    if (!ds.wg_configured) {
        cudaError_t e = cudaFuncSetAttribute(k_wgrad_w256, cudaFuncAttributeMaxDynamicSharedMemorySize, wg2::kSmem);
        if (e != cudaSuccess) { set_error("wgrad_w256: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        ds.wg_configured = true;
    }
    // CTAs per role in proportion to the bytes a role reads per sample (3 : 3 : 5.6), never more roles' CTAs than tiles
    const int tiles = ceil_div(max_samples > 0 ? max_samples : 1, 128);
    const int sms = num_sms();
    int n0 = sms * 3 / 12, n1 = n0, n2 = sms - n0 - n1;
    if (n0 > tiles) n0 = tiles;
    if (n1 > tiles) n1 = tiles;
    if (n2 > tiles) n2 = tiles;
    launch_chain(k_wgrad_w256, dim3(n0 + n1 + n2), dim3(wg2::kThreads), wg2::kSmem, st, fp, n0, n1);
    PSLAM_CHECK_LAUNCH("wgrad_w256");
    return 0;
}

}  // namespace pslam
