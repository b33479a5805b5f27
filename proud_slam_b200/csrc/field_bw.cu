// Mapping backward of the width-128 decoder in ONE kernel: dgrad chain + weight gradients (3xF16, tcgen05).
// Reference: the autograd backward of src/variations/nrgbd.py:116-135 (decoder) inside loss.backward(),
// src/variations/render_helpers.py:671.
//
// field_bf.cu / field_pp.cu run the chain and leave the gradient operands G4, G2, G1, G5 in an HBM scratch (1.6 kB per
// sample written, and 3.2 kB per sample read back by k_wgrad_bf together with the forward's activations): 890 MB of the
// 1.19 GB of DRAM traffic of a mapping iteration.  Here the weight-gradient MMAs run inside the chain kernel:
//   * the chain keeps the tensor-memory plan of field_pp.cu for ONE tile (A_hi 64 + A_lo 64 + D 128 columns; features and
//     sdf head off the tensor-memory path), which leaves 256 columns for the two big accumulators dW2 = G2^T H1 and
//     M = G4^T H2, resident across all tiles of the CTA;
//   * every gradient operand is written once more by its epilogue into a 64 kB shared-memory buffer in the MN-major
//     (sample-major) core-matrix order, where the weight-gradient MMAs read it in place; the forward's activations
//     H2, H1, HC (and F) of the tile arrive from the forward's scratch by bulk TMA, one 64-sample half at a time;
//   * the big products are issued right behind the chain layer that consumes the same operand, i.e. they execute while
//     the workers run that layer's epilogue -- the tensor pipe idled there before;
//   * the small products (16 / 8 output columns: dW4[:,128:], dW1, dW5, dW3[0]) go to spare accumulator columns next to
//     the two N = 16 layers at the ends of the chain, fresh for every tile; 128 worker threads read them with the
//     chain's own results and keep the running sums in registers.  Every one of these N = 16 MMAs re-reads a 4 kB A
//     operand from shared memory (~45 clocks whatever N is) and they sit on the critical path, so the bias gradients
//     db4 / db2 / db1 are NOT ones-operand products but butterfly column sums of the epilogues that produce G4 / G2 / G1;
//   * the tile's first epilogue (g_hc, CUDA cores only) is computed one tile ahead, while the workers wait for phase 4.
// DRAM traffic of the decoder backward: the forward's 1.6 kB per sample read once, nothing written but the feature
// gradients (64 B per sample).  k_wgrad_finish (field_bf.cu) still turns M into dW3 / dW4[:, :128].
#include "field_bf.cuh"
#include "kernels.h"
#include "trilinear.cuh"
#include <type_traits>

namespace pslam {

using namespace umma;

namespace bw {
using namespace bf;
constexpr int kBWThreads = 384;           // warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle, warps 4-11 workers (two per sample row)
constexpr int kBWWorkers = 256;
constexpr int kBWStages = 4;              // weight ring
// tensor memory
constexpr int cHi = 0, cLo = 64, cAcc = 128, cW2 = 256, cM = 384;
// spare accumulator columns of the small products (inside D, next to the N = 16 layers that use D[0,16))
constexpr int cS0 = cAcc + 16;            // at G4 time: [0,16) dW4[:,128:] = G4^T F   [16,32) H2^T G5
                                          // at G1 time: [0,16) dW1 = G1^T F          [16,32) HC^T G5
                                          // (the bias gradients = column sums of G4 / G2 / G1 are butterfly sums of the epilogues)
// shared memory
constexpr int kPlane = 32768;             // one f16 plane of a 128 x 128 operand: [kb 16][fb 16][8 samples][8 features x 2 B]
constexpr int oG = kBWStages * kStageBytes;                    // gradient operand of the current layer (hi plane | lo plane)
constexpr int oH = oG + 2 * kPlane;                            // forward activation, two 64-sample halves of 32 kB ([plane 2][kb 8][fb 16][128 B])
constexpr int oFs = oH + 2 * 32768;                            // features of the tile, two halves of 4 kB ([plane 2][kb 8][fb 2][128 B])
constexpr int oG5s = oFs + 2 * 4096;                           // G5 of the tile: [plane 2][kb 16][fb 2][128 B]
constexpr int oBars = oG5s + 2 * 4096;                            // full[4] empty[4] a_ready mma_done h_full[2] h_free[2] all_done
constexpr int oTmemPtr = oBars + 8 * (2 * kBWStages + 7);
constexpr int oW5 = (oTmemPtr + 16 + 15) & ~15;                           // W5 [3][128] fp32
constexpr int oW30 = oW5 + 4 * 3 * 128;                        // W3 row 0 [128] fp32
constexpr int oScat = (oW30 + 4 * 128 + 15) & ~15;                        // two warps x (weights [32][9] + gradient rows [32][16] + corner ids [32][8]) of the fused scatter
constexpr int oScatDone = oScat + 2 * 4 * kScatWarpFloats;     // tiles whose feature-gradient rows are complete (x 128 lead threads)
constexpr int kSmemBytes = oScatDone + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget");
__host__ __device__ constexpr int n_chunks(int l) { return 4; }                       // layers 10, 6, 7 (without its sdf chunk), 8, 9
__host__ __device__ constexpr int chunk_bytes(int l) { return hN(l) * 32 * 4; }
__host__ __device__ constexpr int chunk_offset(int l, int c) { return layer_offset(l) + c * hN(l) * 32 * 4; }
}  // namespace bw

// Butterfly column sums over the 32 rows of a warp: 16 values per lane -> 1 (lane pairs hold the same sum); the column a lane
// ends up with is c0 + 8 b4 + 4 b3 + 2 b2 + b1 (b = bits of the lane index).  Clobbers y.
__device__ __forceinline__ float colsum16(float (&y)[16], int lane)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = (lane & 16) ? y[i] : y[i + 8], keep = (lane & 16) ? y[i + 8] : y[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = (lane & 8) ? y[i] : y[i + 4], keep = (lane & 8) ? y[i + 4] : y[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = (lane & 4) ? y[i] : y[i + 2], keep = (lane & 4) ? y[i + 2] : y[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float send = (lane & 2) ? y[0] : y[1], keep = (lane & 2) ? y[1] : y[0];
        y[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    y[0] += __shfl_xor_sync(0xffffffffu, y[0], 1);
    return y[0];
}

// 16 accumulator columns of this thread's row: D -> (scale / mask / rank-1 term) -> f16 hi / lo -> next A operand in tensor
// memory and, with `stg`, the same packed words into the shared-memory gradient operand (MN-major core-matrix order).
//   MODE 2: y = mask ? D/16 (+ r1 * wx[c]) : 0     MODE 3: y = D/16
template <int MODE, bool RANK1, bool COLSUM>
__device__ __forceinline__ void bw_epi16(uint32_t trow, int c0, uint32_t mask, int shift, unsigned char *stg, float &ymax, const float *wx,
                                         float r1, float &colsum, int lane)
{
    using namespace bw;
    uint32_t v[16];
    tmem_ld16(trow + cAcc + c0, v);
    tmem_wait_ld();
    float y[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        float t = __uint_as_float(v[e]);
        if (MODE == 2) {
            t = RANK1 ? fmaf(r1, wx[c0 + e], t * kInvScale) : t * kInvScale;
            t = ((mask >> (shift + e)) & 1u) ? t : 0.0f;
        } else {
            t = t * kInvScale;
        }
        ymax = fmaxf(ymax, fabsf(t));
        y[e] = t;
    }
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h16_split2(y[2 * e], y[2 * e + 1], hi[e], lo[e]);
    if (stg) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned char *dst = stg + (c0 / 8 + j) * 128;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4 *>(dst + kPlane) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
    }
    tmem_st8(trow + cHi + c0 / 2, hi);
    tmem_st8(trow + cLo + c0 / 2, lo);
    if (COLSUM) colsum += colsum16(y, lane);
}

// optional timeline of CTA 0 (pslam_debug_bw_trace): [tile < 4][worker | issuer][16] clock64 stamps
//   worker: 0 tile start, 1 g_hc published, 2 + 2i phase i accumulators seen, 3 + 2i its epilogue done and published
//   issuer: 3i operand of phase i seen, 3i + 1 its chain layer issued, 3i + 2 everything of the phase issued
__device__ long long *g_bw_trace = nullptr;
#define BW_TRACE(it_, who_, slot_)                                                                          \
    do {                                                                                                    \
        if (g_bw_trace && blockIdx.x == 0 && (it_) < 4) g_bw_trace[((it_) * 2 + (who_)) * 16 + (slot_)] = clock64(); \
    } while (0)

template <bool SCATTER>     // SCATTER: the trilinear backward rides along in warps 2-3 (PSLAM_OPT_FUSED_SCATTER)
__global__ void __cluster_dims__(bf::kCluster, 1, 1) __launch_bounds__(bw::kBWThreads, 1)
k_field_bw(FieldParams p, const unsigned char *__restrict__ wstream, float *__restrict__ finish, FieldParams ps)
{
    pdl_enter();
    using namespace bw;
    if (threadIdx.x == 128) BW_TRACE(3, 0, 11);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + oBars);
    uint64_t *empty = full + kBWStages;
    uint64_t *a_ready = empty + kBWStages;   // the next A operand (and what rides with it in shared memory) is complete, D may be overwritten
    uint64_t *mma_done = a_ready + 1;        // the phase's accumulators are complete
    uint64_t *h_full = mma_done + 1;         // [2]: a 64-sample half of a forward activation has landed
    uint64_t *h_free = h_full + 2;           // [2]: the MMAs that read it have completed
    uint64_t *all_done = h_free + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + oTmemPtr);
    float *sW5 = reinterpret_cast<float *>(smem + oW5), *sW30 = reinterpret_cast<float *>(smem + oW30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nsamp = p.nsamp_dev ? *p.nsamp_dev : p.nsamp;
    const int ntiles = (nsamp + 127) / 128;
    const int G = (int)gridDim.x;
    // the CTAs of a CLUSTER share one weight stream: same number of iterations for both, an out-of-range tile is a dummy (no
    // valid rows, no weight-gradient MMAs, nothing loaded or stored).  Clusters without a tile in the last round stop one
    // round early: their drain (128 kB of red.v4 per CTA, ~16 k clocks when all 148 CTAs flush at once) then runs under the
    // last round of the others.
    const uint32_t crank = cluster_ctarank();
    const int cbase = (int)blockIdx.x - (int)crank;
    const int iters = cbase < ntiles ? (ntiles - 1 - cbase) / G + 1 : 1;

    volatile int *scat_done = reinterpret_cast<volatile int *>(smem + oScatDone);
    if (threadIdx.x == 0) {
        *scat_done = 0;
        for (int i = 0; i < kBWStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, kCluster); }
        mbar_init(a_ready, kBWWorkers);
        mbar_init(mma_done, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(h_full + i, 1); mbar_init(h_free + i, 1); }
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, kTmemCols);
    for (int i = threadIdx.x; i < 3 * 128; i += kBWThreads) sW5[i] = p.dec.W5[i];
    for (int i = threadIdx.x; i < 128; i += kBWThreads) sW30[i] = p.dec.W3[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    cluster_sync();
    const uint32_t tmem = *tmem_ptr;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssue));
        if (warp == 0) {
            // ===================== TMA producer: the weight chunks in the issuer's order (multicast over the cluster) and, per real
            // tile, the forward's activations in the order the weight-gradient MMAs consume them =====================
            int stage = 0, phase = 0;
            uint32_t hload = 0;   // activation halves loaded so far (buffer = hload & 1)
            auto emit = [&](int l) {
                for (int c = 0; c < n_chunks(l); ++c) {
                    const uint32_t bytes = (uint32_t)chunk_bytes(l), part = bytes / kCluster;
                    const unsigned char *src = wstream + chunk_offset(l, c);
                    mbar_wait(empty + stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full + stage, bytes);
                        bulk_g2s_mcast(smem + stage * kStageBytes + crank * part, src + crank * part, part, full + stage, kClusterMask);
                    }
                    __syncwarp();
                    if (++stage == kBWStages) { stage = 0; phase ^= 1; }
                }
            };
            auto load_half = [&](const unsigned char *tile, int op, int h, bool with_f) {
                const int b = hload & 1;
                if (hload >= 2) mbar_wait(h_free + b, ((hload >> 1) - 1) & 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(h_full + b, (uint32_t)(32768 + (with_f ? 2 * 4096 : 0)));
                    bulk_g2s(smem + oH + b * 32768, tile + (size_t)op * kOpBytes + (size_t)h * 32768, 32768, h_full + b);
                    if (with_f) {
                        bulk_g2s(smem + oFs, tile + oF, 4096, h_full + b);
                        bulk_g2s(smem + oFs + 4096, tile + oF + 4096, 4096, h_full + b);
                    }
                }
                __syncwarp();
                ++hload;
            };
            for (int it = 0; it < iters; ++it) {
                const int tile_i = it * G + (int)blockIdx.x;
                const bool real = tile_i < ntiles;
                const unsigned char *tile = p.wg_scratch + (size_t)tile_i * kTileBytes;
                emit(10);
                if (real) { load_half(tile, oH2, 0, true); load_half(tile, oH2, 1, false); }
                emit(6);
                emit(7);
                if (real) { load_half(tile, oH1, 0, false); load_half(tile, oH1, 1, false); }
                emit(8);
                if (real) { load_half(tile, oHC, 0, false); load_half(tile, oHC, 1, false); }
                emit(9);
            }
        } else if (warp == 1) {
            // ===================== MMA issuer: the whole warp walks the schedule, one elected lane issues =====================
            int stage = 0, phase = 0;
            uint32_t uses = 0, huse = 0;
            const uint32_t a_hi = tmem + cHi, a_lo = tmem + cLo, d = tmem + cAcc;
            const uint32_t id128 = idesc_h16(128, 128, 1, 1), id16 = idesc_h16(128, 16, 1, 1);
            const uint32_t sG = smem_u32(smem + oG), sH = smem_u32(smem + oH), sF = smem_u32(smem + oFs), sG5 = smem_u32(smem + oG5s);
            // one chain layer (packed layer L, N output columns) from the A operand in tensor memory
            auto chain = [&](auto lc, auto nc, bool first_acc_fresh) {
                constexpr int L = decltype(lc)::value, N = decltype(nc)::value;
                constexpr int NP = hN(L);
                const uint32_t idesc = idesc_h16(128, N);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    mbar_wait(full + stage, phase);
                    fence_after_sync();
                    const uint64_t b0 = sdesc(smem_u32(smem + stage * kStageBytes), NP * 16, 128);
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const uint64_t b_hi = b0 + (uint64_t)((s * (2 * NP * 16)) >> 4);
                            const uint64_t b_lo = b0 + (uint64_t)((NP * 32 * 2 + s * (2 * NP * 16)) >> 4);
                            const uint32_t acol = (uint32_t)(16 * (c + 4 * s)) >> 1;
                            mma_h16_ts(d, a_lo + acol, b_hi, idesc, (first_acc_fresh && c == 0 && s == 0) ? 0u : 1u);
                            mma_h16_ts(d, a_hi + acol, b_lo, idesc, 1u);
                            mma_h16_ts(d, a_hi + acol, b_hi, idesc, 1u);
                        }
                        mma_commit_mcast(empty + stage, kClusterMask);
                    }
                    __syncwarp();
                    if (++stage == kBWStages) { stage = 0; phase ^= 1; }
                }
            };
            // operand descriptors of k-step ks (16 samples) = a base descriptor + a multiple of 16 B in its address field: gradient
            // operand (all 16 kb contiguous), activation half buffers, features (two halves), G5
            const uint64_t gd0 = sdesc(sG, 2048, 128), gd1 = sdesc(sG + kPlane, 2048, 128);
            const uint64_t hd0 = sdesc(sH, 2048, 128), hd1 = sdesc(sH + 16384, 2048, 128);                  // (+ 32768 B for the second half)
            const uint64_t fd0 = sdesc(sF, 256, 128), fd1 = sdesc(sF + 2048, 256, 128);                      // (+ 4096 B for the second half)
            const uint64_t g5d0 = sdesc(sG5, 256, 128), g5d1 = sdesc(sG5 + 4096, 256, 128);
            auto g_desc = [&](int ks, int plane) { return (plane ? gd1 : gd0) + (uint64_t)(ks * (4096 >> 4)); };
            auto h_desc = [&](int ks, int plane) { return (plane ? hd1 : hd0) + (uint64_t)((ks >> 2) * (32768 >> 4) + (ks & 3) * (4096 >> 4)); };
            auto f_desc = [&](int ks, int plane) { return (plane ? fd1 : fd0) + (uint64_t)((ks >> 2) * (4096 >> 4) + (ks & 3) * (512 >> 4)); };
            auto g5_desc = [&](int ks, int plane) { return (plane ? g5d1 : g5d0) + (uint64_t)(ks * (512 >> 4)); };
            auto prod3 = [&](uint32_t dcol, uint64_t xh, uint64_t xl, uint64_t yh, uint64_t yl, uint32_t idesc, uint32_t fresh) {
                mma_h16_ss(tmem + dcol, xl, yh, idesc, fresh);
                mma_h16_ss(tmem + dcol, xh, yl, idesc, 1u);
                mma_h16_ss(tmem + dcol, xh, yh, idesc, 1u);
            };
            auto wait_half = [&]() {          // the next activation half has landed
                mbar_wait(h_full + (huse & 1), (huse >> 1) & 1);
                fence_after_sync();
            };
            auto release_half = [&]() {       // (elected lane) the MMAs issued so far are the last readers of that half
                mma_commit(h_free + (huse & 1));
            };
            bool big_fresh = true;            // first real tile of this CTA: the big accumulators start from zero
            for (int it = 0; it < iters; ++it) {
                const bool real = it * G + (int)blockIdx.x < ntiles;
                // ---- phase 0: g_hc is in A / the G buffer, G5 in its buffer.  D[0,16) = W4[:,128:]^T g_hc; small products of G4 ----
                mbar_wait(a_ready, uses & 1u); ++uses;
                if (lane == 0) BW_TRACE(it, 1, 0);
                fence_after_sync();
                chain(std::integral_constant<int, 10>{}, std::integral_constant<int, 16>{}, true);
                if (lane == 0) BW_TRACE(it, 1, 1);
                if (real) {
                    wait_half();              // H2 half 0 (+ F, both halves)
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t fresh = ks == 0 ? 0u : 1u;
                            prod3(cS0, g_desc(ks, 0), g_desc(ks, 1), f_desc(ks, 0), f_desc(ks, 1), id16, fresh);        // dW4[:,128:] = G4^T F
                            if (ks < 4) prod3(cS0 + 16, h_desc(ks, 0), h_desc(ks, 1), g5_desc(ks, 0), g5_desc(ks, 1), id16, fresh);   // H2^T G5
                        }
                    }
                    __syncwarp();
                    ++huse;                   // (half 0 stays in use: the big product below reads it again, released there)
                    wait_half();              // H2 half 1
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 4; ks < 8; ++ks) prod3(cS0 + 16, h_desc(ks, 0), h_desc(ks, 1), g5_desc(ks, 0), g5_desc(ks, 1), id16, 1u);
                    }
                    __syncwarp();
                    --huse;
                }
                if (elect_one()) mma_commit(mma_done);
                __syncwarp();
                // ---- phase 1: D = g_t, then M += G4^T H2 behind it (runs under the g_t epilogue) ----
                mbar_wait(a_ready, uses & 1u); ++uses;
                if (lane == 0) BW_TRACE(it, 1, 3);
                fence_after_sync();
                chain(std::integral_constant<int, 6>{}, std::integral_constant<int, 128>{}, true);
                if (lane == 0) BW_TRACE(it, 1, 4);
                if (elect_one()) mma_commit(mma_done);
                __syncwarp();
                if (real) {
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < 4; ++ks) prod3(cM, g_desc(ks, 0), g_desc(ks, 1), h_desc(ks, 0), h_desc(ks, 1), id128, (big_fresh && ks == 0) ? 0u : 1u);
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 4; ks < 8; ++ks) prod3(cM, g_desc(ks, 0), g_desc(ks, 1), h_desc(ks, 0), h_desc(ks, 1), id128, 1u);
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                }
                // ---- phase 2: D = g_h2 (its epilogue overwrites the G buffer: every MMA above is ahead of this layer's commit) ----
                mbar_wait(a_ready, uses & 1u); ++uses;
                if (lane == 0) BW_TRACE(it, 1, 6);
                fence_after_sync();
                chain(std::integral_constant<int, 7>{}, std::integral_constant<int, 128>{}, true);
                if (lane == 0) BW_TRACE(it, 1, 7);
                if (elect_one()) mma_commit(mma_done);
                __syncwarp();
                // ---- phase 3: dW2 += G2^T H1 first (the g_h1 epilogue overwrites the G buffer), then D = g_h1 ----
                mbar_wait(a_ready, uses & 1u); ++uses;
                if (lane == 0) BW_TRACE(it, 1, 9);
                fence_after_sync();
                if (real) {
                    wait_half();
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < 4; ++ks) prod3(cW2, g_desc(ks, 0), g_desc(ks, 1), h_desc(ks, 0), h_desc(ks, 1), id128, (big_fresh && ks == 0) ? 0u : 1u);
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                    wait_half();
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 4; ks < 8; ++ks) prod3(cW2, g_desc(ks, 0), g_desc(ks, 1), h_desc(ks, 0), h_desc(ks, 1), id128, 1u);
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                    big_fresh = false;
                }
                chain(std::integral_constant<int, 8>{}, std::integral_constant<int, 128>{}, true);
                if (lane == 0) BW_TRACE(it, 1, 10);
                if (elect_one()) mma_commit(mma_done);
                __syncwarp();
                // ---- phase 4: D[0,16) = W1^T g_h1; small products of G1 ----
                mbar_wait(a_ready, uses & 1u); ++uses;
                if (lane == 0) BW_TRACE(it, 1, 12);
                fence_after_sync();
                chain(std::integral_constant<int, 9>{}, std::integral_constant<int, 16>{}, true);
                if (lane == 0) BW_TRACE(it, 1, 13);
                if (real) {
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t fresh = ks == 0 ? 0u : 1u;
                            prod3(cS0, g_desc(ks, 0), g_desc(ks, 1), f_desc(ks, 0), f_desc(ks, 1), id16, fresh);        // dW1 = G1^T F
                        }
                    }
                    __syncwarp();
                    wait_half();              // HC half 0
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < 4; ++ks) prod3(cS0 + 16, h_desc(ks, 0), h_desc(ks, 1), g5_desc(ks, 0), g5_desc(ks, 1), id16, ks == 0 ? 0u : 1u);   // HC^T G5
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                    wait_half();              // HC half 1
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 4; ks < 8; ++ks) prod3(cS0 + 16, h_desc(ks, 0), h_desc(ks, 1), g5_desc(ks, 0), g5_desc(ks, 1), id16, 1u);
                        release_half();
                    }
                    __syncwarp();
                    ++huse;
                }
                if (elect_one()) { mma_commit(mma_done); if (it == iters - 1) mma_commit(all_done); }
                __syncwarp();
                if (lane == 0) BW_TRACE(it, 1, 14);
            }
        } else if constexpr (SCATTER) {
            // ===================== warps 2-3: the trilinear backward of every finished tile (embedding scatter, ray gradients) ============
            // (PSLAM_OPT_FUSED_SCATTER, off by default.)  The chain leaves these warps and most issue slots idle and the stand-alone
            // scatter kernel is the tail of the iteration (25 us after this kernel) -- but measured: two warps without an L1 to
            // speak of (the kernel's 225 kB of shared memory leave none) take ~22 us for a tile's 128 samples, longer than the
            // chain needs for the tile: the kernel went from 184 to 247 us.  Kept for batches where the chain is slower per tile.  The lead workers store a tile's feature-gradient rows, fence, and count themselves in;
            // each of the two warps then takes 64 of the tile's samples in two passes of tri_scatter_warp (trilinear.cuh).
            float *sw = reinterpret_cast<float *>(smem + oScat) + (warp - 2) * kScatWarpFloats;
            int done = 0;
            for (int it = 0; it < iters; ++it) {
                const int tile_i = it * G + (int)blockIdx.x;
                if (tile_i >= ntiles) break;
                ++done;
                while (*scat_done < 128 * done) __nanosleep(200);
                __threadfence();
                for (int pass = 0; pass < 2; ++pass) {
                    const int s0 = tile_i * 128 + (warp - 2) * 64 + pass * 32;
                    if (s0 < nsamp) tri_scatter_warp<true>(ps, p.g_feat, s0, nsamp, sw);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsWorker));
        // ===================== workers: two threads per sample row (64 accumulator columns each) =====================
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int col0 = half * 64;
        const bool lead = half == 0;
        const int m = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        unsigned char *sGrow = smem + oG + (m >> 3) * 2048 + (m & 7) * 16;            // this row's 16 B of feature block 0, hi plane
        unsigned char *sG5row = smem + oG5s + (m >> 3) * 256 + (m & 7) * 16;
        uint32_t done_uses = 0;
        float ymax = 0.0f;
        const float Sg = grad_scale(p.gscale), invSg = 1.0f / Sg;
        int tr_it = 0, tr_slot = 0;
        auto layer_done = [&]() {
            mbar_wait(mma_done, done_uses & 1u);
            ++done_uses;
            fence_after_sync();
            if (threadIdx.x == 128) BW_TRACE(tr_it, 0, tr_slot++);
        };
        auto a_is_ready = [&]() {
            tmem_wait_st();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's rows of the G buffer -> the tensor core's reads
            fence_before_sync();
            mbar_arrive(a_ready);
            if (threadIdx.x == 128) BW_TRACE(tr_it, 0, tr_slot++);
        };
        // running sums of the small products of this thread's accumulator row n = m (threads of the second column half hold them)
        float acc_w4f[16], acc_w1[16], acc_h2g5 = 0.f, acc_hcg5[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 16; ++e) { acc_w4f[e] = 0.f; acc_w1[e] = 0.f; }
        float db2[4] = {0.f, 0.f, 0.f, 0.f};              // butterfly column sums of g_h2: one column per 16-column batch and lane pair
        float db4[4] = {0.f, 0.f, 0.f, 0.f}, db1[4] = {0.f, 0.f, 0.f, 0.f};   // the same of g_hc (compute_next) and g_h1 (phase 3)
        // next tile's per-row inputs
        uint32_t pm[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        float4 po = make_float4(0.f, 0.f, 0.f, 0.f), pgo = make_float4(0.f, 0.f, 0.f, 0.f);
        auto prefetch_tile = [&](int tn) {
            const int sn = tn * 128 + m;
            const bool in = tn < ntiles && sn < nsamp;
            const uint32_t *mk = p.act_masks + (size_t)(tn < ntiles ? tn : 0) * (kMaskBytes / 4) + half * 256 + m;
            pm[0] = mk[0]; pm[1] = mk[128]; pm[2] = mk[512]; pm[3] = mk[512 + 128]; pm[4] = mk[1024]; pm[5] = mk[1024 + 128];
            po = in ? __ldg(reinterpret_cast<const float4 *>(p.out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            pgo = in ? __ldg(reinterpret_cast<const float4 *>(p.g_out + (size_t)sn * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // The first epilogue of a tile (g_hc = mask_hc . (W5^T g5), on the CUDA cores: ~3.6k clocks of issue for the eight worker
        // warps) is computed one tile AHEAD, while the workers would otherwise wait ~3k clocks for the small products of
        // phase 4: the packed operand words wait in registers until the tile's last MMAs have completed, then only the stores
        // remain at the top of the next tile.
        uint32_t nhi[4][8] = {}, nlo[4][8] = {};          // next tile: g_hc, f16 hi / lo words of this thread's 64 columns
        float ng5[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t nm1[2] = {0u, 0u}, nm2[2] = {0u, 0u};
        auto compute_next = [&](int tn) {
            const bool real_n = tn < ntiles;
            nm1[0] = pm[0]; nm1[1] = pm[1]; nm2[0] = pm[2]; nm2[1] = pm[3];
            const uint32_t mc[2] = {pm[4], pm[5]};
            const float r = po.x, gg = po.y, b = po.z;    // (zeros outside the batch: prefetch_tile)
            const float gx = pgo.x * Sg, gy = pgo.y * Sg, gz = pgo.z * Sg;
            ng5[0] = gx * (1.0f - r) * r; ng5[1] = gy * (1.0f - gg) * gg; ng5[2] = gz * (1.0f - b) * b; ng5[3] = pgo.w * Sg;
            if (lead) {
                // bias gradients of the two heads: column sums of G5
                const float s0 = warp_sum(ng5[0]), s1 = warp_sum(ng5[1]), s2 = warp_sum(ng5[2]), s3 = warp_sum(ng5[3]);
                if (lane == 0 && real_n) {
                    atomicAdd(p.g_dec.b5 + 0, s0 * invSg); atomicAdd(p.g_dec.b5 + 1, s1 * invSg); atomicAdd(p.g_dec.b5 + 2, s2 * invSg);
                    atomicAdd(p.g_dec.b3, s3 * invSg);
                }
            }
            // (rolled: the kernel's code is refetched from L2 -- after an L2 flush from DRAM -- in every round, the instruction
            // caches hold a third of it; the four 16-column batches share one body and the packed words rotate through the
            // register arrays, 64 moves per batch)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int c0 = col0 + 16 * j;
                const uint32_t bits = ((j & 2) ? mc[1] : mc[0]) >> ((j & 1) * 16);
                float y[16];
                uint32_t thi[8], tlo[8];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {          // (128-bit broadcast loads: the tensor core's operand reads want the shared-memory cycles)
                    const float4 w0 = *reinterpret_cast<const float4 *>(sW5 + c0 + 4 * qd);
                    const float4 w1 = *reinterpret_cast<const float4 *>(sW5 + 128 + c0 + 4 * qd);
                    const float4 w2 = *reinterpret_cast<const float4 *>(sW5 + 256 + c0 + 4 * qd);
                    float y0 = fmaf(ng5[2], w2.x, fmaf(ng5[1], w1.x, ng5[0] * w0.x));
                    float y1 = fmaf(ng5[2], w2.y, fmaf(ng5[1], w1.y, ng5[0] * w0.y));
                    float y2 = fmaf(ng5[2], w2.z, fmaf(ng5[1], w1.z, ng5[0] * w0.z));
                    float y3 = fmaf(ng5[2], w2.w, fmaf(ng5[1], w1.w, ng5[0] * w0.w));
                    y0 = ((bits >> (4 * qd)) & 1u) ? y0 : 0.0f;
                    y1 = ((bits >> (4 * qd + 1)) & 1u) ? y1 : 0.0f;
                    y2 = ((bits >> (4 * qd + 2)) & 1u) ? y2 : 0.0f;
                    y3 = ((bits >> (4 * qd + 3)) & 1u) ? y3 : 0.0f;
                    ymax = fmaxf(ymax, fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fmaxf(fabsf(y2), fabsf(y3))));
                    h16_split2(y0, y1, thi[2 * qd], tlo[2 * qd]);
                    h16_split2(y2, y3, thi[2 * qd + 1], tlo[2 * qd + 1]);
                    y[4 * qd] = y0; y[4 * qd + 1] = y1; y[4 * qd + 2] = y2; y[4 * qd + 3] = y3;
                }
                const float cs = colsum16(y, lane);       // bias gradient of layer 4 (here it costs nothing: the workers are waiting)
                db4[0] += (j == 0) ? cs : 0.f; db4[1] += (j == 1) ? cs : 0.f; db4[2] += (j == 2) ? cs : 0.f; db4[3] += (j == 3) ? cs : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {             // batch j ends up in n??[j] after the four rotations
                    nhi[0][e] = nhi[1][e]; nhi[1][e] = nhi[2][e]; nhi[2][e] = nhi[3][e]; nhi[3][e] = thi[e];
                    nlo[0][e] = nlo[1][e]; nlo[1][e] = nlo[2][e]; nlo[2][e] = nlo[3][e]; nlo[3][e] = tlo[e];
                }
            }
        };
        if (threadIdx.x == 128) BW_TRACE(3, 0, 12);
        // round -1 is the look-ahead for the first tile alone (one call site: compute_next is a fifth of the worker code)
        for (int it = -1; it < iters; ++it) {
            const int tile = it * G + (int)blockIdx.x;
            const bool real = it >= 0 && tile < ntiles;
            const int s = real ? tile * 128 + m : nsamp;
            const bool valid = s < nsamp;
            if (it < 0) {
                prefetch_tile((int)blockIdx.x);
            } else {
                tr_it = it; tr_slot = 0;
                if (threadIdx.x == 128) BW_TRACE(tr_it, 0, tr_slot++);
                const uint32_t m1[2] = {nm1[0], nm1[1]}, m2[2] = {nm2[0], nm2[1]};
                const float gow = ng5[3];
                if (lead) {
                    // G5 = (g5 r, g, b, g_sdf, 0 ...) as a 16-feature operand (second feature block zero: written once below)
                    uint32_t h0, l0, h1w, l1w;
                    h16_split2(ng5[0], ng5[1], h0, l0);
                    h16_split2(ng5[2], ng5[3], h1w, l1w);
                    *reinterpret_cast<uint4 *>(sG5row) = make_uint4(h0, h1w, 0u, 0u);
                    *reinterpret_cast<uint4 *>(sG5row + 4096) = make_uint4(l0, l1w, 0u, 0u);
                    if (it == 0) {
                        *reinterpret_cast<uint4 *>(sG5row + 128) = make_uint4(0u, 0u, 0u, 0u);
                        *reinterpret_cast<uint4 *>(sG5row + 4096 + 128) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                // g_hc (computed during the previous tile) -> A and the G buffer
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c0 = col0 + 16 * j;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        unsigned char *dst = sGrow + (c0 / 8 + k) * 128;
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(nhi[j][4 * k], nhi[j][4 * k + 1], nhi[j][4 * k + 2], nhi[j][4 * k + 3]);
                        *reinterpret_cast<uint4 *>(dst + kPlane) = make_uint4(nlo[j][4 * k], nlo[j][4 * k + 1], nlo[j][4 * k + 2], nlo[j][4 * k + 3]);
                    }
                    tmem_st8(trow + cHi + c0 / 2, nhi[j]);
                    tmem_st8(trow + cLo + c0 / 2, nlo[j]);
                }
                a_is_ready();
                // ---- phase 0 results: g_f part (lead) and the small products of G4 (second column half) ----
                layer_done();
                if (lead) {                                   // (the lead threads keep no running sums: acc_w4f doubles as this tile's g_f part)
                    uint32_t v[16];
                    tmem_ld16(trow + cAcc, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc_w4f[e] = __uint_as_float(v[e]);
                } else if (real) {
                    uint32_t v[16], w[8];
                    tmem_ld16(trow + cS0, v);
                    tmem_ld8(trow + cS0 + 16, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) acc_w4f[e] += __uint_as_float(v[e]);
                    acc_h2g5 += __uint_as_float(w[3]);        // H2^T G5, column 3 (g_sdf) -> dW3 row 0
                }
                fence_before_sync();
                mbar_arrive(a_ready);
                if (threadIdx.x == 128) BW_TRACE(tr_it, 0, tr_slot++);
                prefetch_tile(tile + G);                      // per-row inputs of the next tile: in flight during phases 1-3
                // ---- phase 1: g_t -> A (not needed by the weight gradients: the G buffer keeps G4 for M += G4^T H2) ----
                layer_done();
                float nocs = 0.f;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) bw_epi16<3, false, false>(trow, col0 + 16 * j, 0u, 0, nullptr, ymax, nullptr, 0.f, nocs, lane);
                a_is_ready();
                // ---- phase 2: g_h2 -> A and the G buffer (+ rank-1 sdf term, + db2) ----
                layer_done();
#pragma unroll                                    // (unrolled: on the critical path, and the batches overlap each other's TMEM / shuffle latencies)
                for (int j = 0; j < 4; ++j) {
                    float cs = 0.f;
                    bw_epi16<2, true, true>(trow, col0 + 16 * j, (j & 2) ? m2[1] : m2[0], (j & 1) * 16, sGrow, ymax, sW30, gow, cs, lane);
                    db2[0] += (j == 0) ? cs : 0.f; db2[1] += (j == 1) ? cs : 0.f; db2[2] += (j == 2) ? cs : 0.f; db2[3] += (j == 3) ? cs : 0.f;
                }
                a_is_ready();
                // ---- phase 3: g_h1 -> A and the G buffer ----
                layer_done();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float cs = 0.f;
                    bw_epi16<2, false, true>(trow, col0 + 16 * j, (j & 2) ? m1[1] : m1[0], (j & 1) * 16, sGrow, ymax, nullptr, 0.f, cs, lane);
                    db1[0] += (j == 0) ? cs : 0.f; db1[1] += (j == 1) ? cs : 0.f; db1[2] += (j == 2) ? cs : 0.f; db1[3] += (j == 3) ? cs : 0.f;
                }
                a_is_ready();
            }
            compute_next(tile + G);                       // (under the MMAs of phase 4; its inputs were requested after phase 0)
            if (it < 0) continue;
            // ---- phase 4 results: g_f (lead) and the small products of G1 ----
            layer_done();
            if (lead) {
                uint32_t v[16];
                tmem_ld16(trow + cAcc, v);
                tmem_wait_ld();
                if (valid && p.g_feat) {
                    float o[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = (acc_w4f[e] + __uint_as_float(v[e])) * (kInvScale * invSg);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4 *>(p.g_feat + (size_t)s * 16 + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                }
                if (SCATTER && real) {                      // this row of the tile is in place for warps 2-3
                    __threadfence();
                    atomicAdd(const_cast<int *>(scat_done), 1);
                }
            } else if (real) {
                uint32_t v[16], w[8];
                tmem_ld16(trow + cS0, v);
                tmem_ld8(trow + cS0 + 16, w);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) acc_w1[e] += __uint_as_float(v[e]);
                acc_hcg5[0] += __uint_as_float(w[0]); acc_hcg5[1] += __uint_as_float(w[1]); acc_hcg5[2] += __uint_as_float(w[2]);
            }
            // (the next tile's first epilogue writes A, the G buffer and G5: every MMA of this tile has completed -- mma_done above)
        }
        if (p.range_flag && ymax >= 32752.0f) atomicOr(p.range_flag, 4);
        // ===================== drain: the big accumulators and this thread's running sums -> global gradients =====================
        mbar_wait(all_done, 0);
        fence_after_sync();
        if (threadIdx.x == 128) BW_TRACE(3, 0, 13);
        const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / G + 1 : 0;
        if (my_tiles > 0) {
            const float cW = kInvScale * invSg;           // accumulators hold 16 x Sg x (sum of products), column sums Sg x (sum)
            const int n = m;                              // accumulator row = TMEM lane
            float *dst = lead ? p.g_dec.W2 + (size_t)n * 128 : finish + (size_t)n * 128;
            const uint32_t col = lead ? cW2 : cM;
            for (int c0 = 0; c0 < 128; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(trow + col + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    red_add_v4(dst + c0 + 4 * j, cW * __uint_as_float(v[4 * j]), cW * __uint_as_float(v[4 * j + 1]), cW * __uint_as_float(v[4 * j + 2]),
                               cW * __uint_as_float(v[4 * j + 3]));
            }
            if (!lead) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    red_add_v4(p.g_dec.W4 + (size_t)n * 144 + 128 + 4 * j, cW * acc_w4f[4 * j], cW * acc_w4f[4 * j + 1], cW * acc_w4f[4 * j + 2], cW * acc_w4f[4 * j + 3]);
                    red_add_v4(p.g_dec.W1 + (size_t)n * 16 + 4 * j, cW * acc_w1[4 * j], cW * acc_w1[4 * j + 1], cW * acc_w1[4 * j + 2], cW * acc_w1[4 * j + 3]);
                }
                atomicAdd(p.g_dec.W5 + n, cW * acc_hcg5[0]);
                atomicAdd(p.g_dec.W5 + 128 + n, cW * acc_hcg5[1]);
                atomicAdd(p.g_dec.W5 + 256 + n, cW * acc_hcg5[2]);
                atomicAdd(p.g_dec.W3 + n, cW * acc_h2g5);
            }
            if ((lane & 1) == 0) {                        // bias gradients: lane pairs hold the same column sums
                const int cb = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(p.g_dec.b2 + col0 + 16 * j + cb, invSg * db2[j]);
                    atomicAdd(p.g_dec.b1 + col0 + 16 * j + cb, invSg * db1[j]);
                    atomicAdd(finish + 128 * 128 + col0 + 16 * j + cb, invSg * db4[j]);
                }
            }
        }
    }
    if (threadIdx.x == 128) BW_TRACE(3, 0, 14);
    fence_before_sync();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc(tmem, bf::kTmemCols);
    if (threadIdx.x == 128) BW_TRACE(3, 0, 15);
}

static int g_bw_enabled = 1, g_bw_scatter = 0;   // fused scatter: measured slower (see the role's comment), off by default
int bw_fused_scatter() { return g_bw_scatter; }
void bw_set_fused_scatter(int on) { g_bw_scatter = on ? 1 : 0; }
int bw_enabled() { return g_bw_enabled; }
void bw_set_enabled(int on) { g_bw_enabled = on ? 1 : 0; }

struct BWDeviceState { bool configured; int max_clusters; };
template <bool SCATTER>
static cudaError_t bw_configure(int *max_clusters)
{
    cudaError_t e = cudaFuncSetAttribute(k_field_bw<SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw::kSmemBytes);
    if (e != cudaSuccess || !max_clusters) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / bf::kCluster * bf::kCluster);
    cfg.blockDim = dim3(bw::kBWThreads);
    cfg.dynamicSmemBytes = bw::kSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = bf::kCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, k_field_bw<SCATTER>, &cfg);
    if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / bf::kCluster; }
    *max_clusters = n < num_sms() / bf::kCluster ? n : num_sms() / bf::kCluster;
    return cudaSuccess;
}
static BWDeviceState g_bw_state[64] = {};

// dgrad chain + weight gradients of the tiles whose forward saved masks and activations into fp.wg_scratch; `finish` = the
// reduction block k_wgrad_finish consumes (cleared by this kernel's predecessor through fp.finish_zero or by the caller)
int bw_launch(const FieldParams &fp, int max_samples, float *finish, cudaStream_t st, const FieldParams *scatter)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("field_bw: cudaGetDevice failed"); return PSLAM_E_ARG; }
    BWDeviceState &ds = g_bw_state[dev];
    if (!ds.configured) {
        cudaError_t e = bw_configure<false>(&ds.max_clusters);
        if (e == cudaSuccess) e = bw_configure<true>(nullptr);
        if (e != cudaSuccess) { set_error("field_bw: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        ds.configured = true;
    }
    const int tiles = ceil_div(max_samples > 0 ? max_samples : 1, 128);
    int grid = ceil_div(tiles, bf::kCluster) * bf::kCluster;
    if (grid > ds.max_clusters * bf::kCluster) grid = ds.max_clusters * bf::kCluster;
    // scatter != NULL: the trilinear backward runs inside the kernel (warps 2-3) on the feature-gradient rows fp.g_feat;
    // *scatter = the parameters the stand-alone scatter kernel would get (sample tables, gradient targets)
    if (scatter) launch_chain(k_field_bw<true>, dim3(grid), dim3(bw::kBWThreads), bw::kSmemBytes, st, fp, reinterpret_cast<const unsigned char *>(fp.ws_tc), finish, *scatter);
    else launch_chain(k_field_bw<false>, dim3(grid), dim3(bw::kBWThreads), bw::kSmemBytes, st, fp, reinterpret_cast<const unsigned char *>(fp.ws_tc), finish, fp);
    PSLAM_CHECK_LAUNCH("field_bw");
    return 0;
}

}  // namespace pslam

extern "C" int pslam_debug_bw_trace(long long *dev_buf)
{
    cudaError_t e = cudaMemcpyToSymbol(pslam::g_bw_trace, &dev_buf, sizeof(dev_buf));
    if (e != cudaSuccess) { pslam::set_error("bw_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
