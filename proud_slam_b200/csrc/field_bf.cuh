// Constants shared by the 3xF16 decoder kernels (field_bf.cu: one tile in flight per CTA; field_pp.cu: two tiles in flight):
// operand scales, weight-stream chunking, ReLU-mask and wgrad-scratch layouts.
#pragma once
#include "decoder_layers.cuh"
#include "field.cuh"
#include "umma.cuh"

namespace pslam {

namespace bf {
// warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer, warp 2 scratch store, warp 3 idle; warpgroups 1-2: workers.
// The split is by warpgroup so that setmaxnreg can move registers from the three single-lane roles to the
// workers (64 vs 216 per thread): at the launch-time 168 the backward workers spilled inside the epilogues.
constexpr int kThreads = 384;
constexpr int kRegsIssue = 64, kRegsWorker = 216;   // 128 x (168 - 64) registers released >= 256 x (216 - 168) acquired, else setmaxnreg.inc never returns
constexpr int kWorkers = 256;
constexpr int kStages = 8;        // weight ring depth of the forward kernel
constexpr int kStagesBwd = 4;     // ... of the backward kernel (the rest of shared memory stages the wgrad scratch)
constexpr int kChunkK = 32;       // reduction elements per weight chunk (two MMA k-steps; the tail chunk of K = 144 / 16 has one)
constexpr int kStageBytes = 18432;   // 144 rows x 32 k x 2 B x (hi, lo)
constexpr int kStagingBytes = 65536; // one 128-feature operand of one tile in scratch order
constexpr int kCluster = 2;
constexpr uint16_t kClusterMask = (1u << kCluster) - 1u;
constexpr int kTmemCols = 512;
constexpr int cAHI = 0, cALO = 72, cD = 144, kDCols = 144;   // two accumulator buffers D0 [144,288), D1 [288,432) alternate per layer
// Power-of-two operand scales that keep the f16 halves in their precise window (umma.cuh: h16_split2).  Weights and
// forward activations are stored x16 (|value| < 4094 representable; absolute resolution 2^-29): an accumulator then
// holds 256 x the product and the epilogue folds 1/16 into its bias FMA, which leaves the next layer's operand
// scaled x16 again.  Gradients carry a per-launch scale Sg = 2^k taken from max |g_out| (k_grad_scale) so that
// the chain stays near 2^8; gradient accumulators hold 16 x Sg x the product.
constexpr float kScale = 16.0f, kInvScale = 1.0f / 16.0f;
// Spill route of the wgrad operands: true = the workers store their packed words straight to the scratch (16 B per thread,
// 8 lanes = one 128-byte line); false = through a shared-memory staging buffer and one 64 kB bulk TMA store per operand.
constexpr bool kDirectSpill = true;
// Weight chunks follow the order in which the A operand becomes available.  An epilogue writes the next A in four
// batches of 16 columns per thread; each row has two worker threads (columns [0,64) and [64,128)), so batch j
// completes the k-steps j and j + 4 (k in [16j, 16j+16) and [64+16j, 64+16j+16)).  Chunk j of a layer therefore holds
// exactly those two k-steps and its MMAs are issued as soon as a_ready[j] fires, while the epilogue is still
// producing the later batches; chunk 4 (K = 144) holds k-step 8, K = 16 layers have the single chunk 0.
__host__ __device__ constexpr int kstep_chunk(int K, int ks) { return (K < 64 || ks >= 8) ? (K < 64 ? 0 : 4) : (ks & 3); }
__host__ __device__ constexpr int kstep_slot(int K, int ks) { return (K < 64 || ks >= 8) ? 0 : (ks >> 2); }
__host__ __device__ constexpr int chunk_kk(int K, int c) { return (K < 64 || c == 4) ? 16 : 32; }
using declayers::kLayersAll;
using declayers::kLayersFwd;
// the packed stream has an 11th layer behind the 10 of the chain: W4's feature columns transposed on their own
// (g_f part = W4[:, 128:]^T g_hc, N = 16, K = 128), which field_pp.cu issues as a separate small layer
constexpr int kLayersPack = kLayersAll + 1;
static __device__ __constant__ int cN[kLayersPack] = {128, 128, 144, 128, 16, 128, 144, 128, 128, 16, 16};
static __device__ __constant__ int cK[kLayersPack] = {16, 128, 128, 144, 128, 16, 128, 144, 128, 128, 128};
static __device__ __constant__ int cAcol[kLayersAll] = {64, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // packed A column of the layer's k = 0
// the same tables as functions (usable with a run-time layer index in device code)
__host__ __device__ constexpr int hN(int l) { return (l == 2 || l == 6) ? 144 : ((l == 4 || l == 9 || l == 10) ? 16 : 128); }
__host__ __device__ constexpr int hK(int l) { return (l == 0 || l == 5) ? 16 : ((l == 3 || l == 7) ? 144 : 128); }
// byte offset of layer l in the packed stream (hi + lo halves: 4 B per weight)
__host__ __device__ constexpr int layer_offset(int l) { int o = 0; for (int i = 0; i < l; ++i) o += hN(i) * hK(i) * 4; return o; }
// Kernel kinds.  The mapping iteration runs kFwdSave + kBwdSaved: the forward spills its activations (H1, H2, HC, F)
// and ReLU masks, so the backward is the dgrad chain alone.  kBwdRecompute (stand-alone backward: forward recompute +
// dgrad, spills everything itself) serves calls whose forward did not save (pslam_decoder_bwd, tracking).
constexpr int kFwd = 0, kBwdRecompute = 1, kFwdSave = 2, kBwdSaved = 3;
constexpr int kMaskBytes = 3 * 2 * 2 * 128 * 4;   // per tile: [layer h1, h2, hc][column half][32-column batch][row] ReLU bits
constexpr int kFwdStreamBytes = 4 * (128 * 16 + 128 * 128 + 144 * 128 + 128 * 144 + 16 * 128);   // weight stream of layers 0-4
template <int KIND>
struct Smem {
    static constexpr bool BWD = KIND != kFwd && !kDirectSpill;   // has the two staging buffers of the wgrad scratch
    static constexpr int nStages = BWD ? kStagesBwd : kStages;
    static constexpr int oStaging = nStages * kStageBytes;
    static constexpr int oBars = oStaging + (BWD ? 2 * kStagingBytes : 0);   // full[8], empty[8], a_ready, mma_done, st_full[2], st_free[2]
    static constexpr int oTmemPtr = oBars + 8 * (2 * kStages + 5 + 4);   // ... a_ready[4], mma_done, st_full[2], st_free[2]
    static constexpr int oBias = oTmemPtr + 16;                     // b1[128] b2[128] b3f[128] b4[128] b3_0 b5[3]
    static constexpr int oW5 = oBias + 4 * (4 * 128 + 4);           // W5 [3][128] fp32: the colour head runs on the CUDA cores
    static constexpr int oHead = oW5 + 4 * 3 * 128;                 // [128 rows][2 threads][4]: each thread's partial dot products of the colour head
    static constexpr int bytes = oHead + 4 * 128 * 8;
};

// wgrad scratch, per 128-sample tile: six 128-feature operands and two 16-feature operands, each already split
// (hi / lo bf16 planes) and in the MN-major no-swizzle core-matrix layout of tcgen05.mma (umma.cuh: sdesc):
//     [half h = sample / 64][plane hi, lo][kb = (sample / 8) % 8][fb = feature / 8][sample % 8][8 features x 2 B]
// so that (operand, half) = one contiguous bulk copy whose 16-sample k-steps are 2 kb blocks apart.
// T = W3 h2 + b3 and G3 = g_t = W4t^T G4 are linear images of spilled operands, so neither is spilled: with
// M = G4^T H2 (accumulated by k_wgrad_bf) the two gradients that would need them are tiny post-products,
//     dW3[1+j][k] = sum_n W4[n][j] M[n][k],     dW4[n][j] = sum_k M[n][k] W3[1+j][k] + db4[n] b3[1+j]   (k_wgrad_finish)
// which takes a quarter off the scratch traffic that bounds both backward kernels.
constexpr size_t kOpBytes = 65536, kSmallBytes = 8192;
constexpr int oH1 = 0, oH2 = 1, oHC = 2, oG1 = 3, oG2 = 4, oG4 = 5, kBigOps = 6;   // x kOpBytes
constexpr size_t oF = kBigOps * kOpBytes, oG5 = oF + kSmallBytes, kTileBytes = oG5 + kSmallBytes;   // 409 600 B = 3.2 kB / sample
constexpr size_t kFinishFloats = 128 * 128 + 128;   // after the tiles: M = G4^T H2 and the column sums of G4 (zeroed per launch)
}  // namespace bf

// Gradient scale of one backward launch: gscale[0] holds the bit pattern of max |g_out| (k_grad_scale); Sg = 2^(8 - e)
// brings that maximum to [256, 512).  Powers of two: scaling and unscaling are exact.
__device__ __forceinline__ float grad_scale(const uint32_t *gscale)
{
    const uint32_t bits = gscale ? *gscale : 0u;
    const int e = (int)((bits >> 23) & 0xffu);          // biased exponent of the maximum
    if (e == 0 || e == 255) return 1.0f;                // all zero / not finite: leave alone
    int k = 127 + 8 - (e - 127);                        // biased exponent of Sg
    k = k < 1 ? 1 : (k > 254 ? 254 : k);
    return __uint_as_float((uint32_t)k << 23);
}

// launchers of the two-tiles-in-flight build (field_pp.cu); KIND = bf::kFwd / kBwdRecompute / kFwdSave / kBwdSaved
int pp_launch(int kind, const FieldParams &fp, int max_samples, cudaStream_t st);
// field_bw.cu: dgrad chain + weight gradients in one kernel (the forward must have saved masks and activations)
int bw_launch(const FieldParams &fp, int max_samples, float *finish, cudaStream_t st, const FieldParams *scatter = nullptr);
int bw_fused_scatter();              // PSLAM_OPT_FUSED_SCATTER
void bw_set_fused_scatter(int on);
int bw_enabled();
void bw_set_enabled(int on);
int pp_enabled();
void pp_set_enabled(int on);

}  // namespace pslam
