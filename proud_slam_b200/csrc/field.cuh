// Argument block shared by the decoder/field kernels (SIMT build in field.cu, tcgen05 build in
// field_tc.cu).
#pragma once
#include "common.cuh"

namespace pslam {

struct FieldParams {
    int nsamp;               // static sample count, or
    const int *nsamp_dev;    // device-side count (pipeline)
    // gather source (feat == nullptr)
    const float *rays_o, *rays_d;   // [R,3]
    const int *hit_ray;             // rank -> ray id
    const int *samp_ray, *samp_vox; // [P]
    const float *samp_z;            // [P]
    const float *centres;           // [N,3]
    const int *vertex_idx;          // [N,8]
    const float *emb;               // [E,16]
    float voxel_size;
    // direct feature source (standalone decoder)
    const float *feat;              // [P,16]
    // decoder
    pslam_decoder_t dec;
    const float *ws;                // packed transposed weights (SIMT build)
    const float *ws_tc;             // weight stream in UMMA operand order (tcgen05 build)
    float *out;                     // [P,4]
    // backward
    const float *g_out;             // [P,4]
    float *g_feat;                  // [P,16] (standalone) or nullptr
    unsigned char *wg_scratch;      // tcgen05 wgrad scratch (tc_wgrad_scratch_bytes) or nullptr
    uint32_t *act_masks;            // ReLU masks saved by the forward (3xBF16 build, inside wg_scratch; set by its launcher)
    uint32_t *gscale;               // bit pattern of max |g_out| of this backward launch (3xF16 build; set by its launcher)
    int spill_ops;                  // 3xF16 build: write the wgrad operands to wg_scratch (decoder gradients wanted)
    float *finish_zero;             // 3xF16 backward: the wgrad kernel's reduction block (kFinishFloats), cleared by the chain kernel that precedes it
    int *range_flag;                // fused pipeline: counters[PSLAM_C_OVERFLOW]; bit 4 = an operand left the 3xF16 window
    uint32_t *gmax_ready;           // fused pipeline: the compositing backward already left that maximum here (counters[PSLAM_C_TILE])
    int paired;                     // forward and backward come from one pslam_render_t (fused pipeline): the forward may save
                                    // its activations for the backward; stand-alone calls (pslam_decoder_*) always recompute
    size_t wg_scratch_bytes;
    pslam_decoder_grad_t g_dec;
    float *g_emb;                   // [E,16] +=
    float *g_rays_o, *g_rays_d;     // [R,3] += (pipeline zeroes them first)
    int grad_dec, grad_emb, grad_rays;
};


// field_tc.cu: tcgen05 build of the width-128 decoder (3xTF32, fp32-equivalent accuracy)
constexpr int kTcPackFloats = 229376;   // forward + dgrad weight streams in UMMA operand order (hi, lo)
// decoder build selection (PSLAM_OPT_DECODER), width 128: 0 = tcgen05 3xTF32 (fp32-equivalent), 1 = fp32 SIMT build,
// 2 = tcgen05 3xF16 with power-of-two operand scales (fp32-equivalent; field_pp.cu / field_bw.cu / field_bf.cu); width 256
// always runs the SIMT build
int decoder_mode();
int tc_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st);
int tc_launch_field_forward(const FieldParams &fp, int max_samples, cudaStream_t st);
// part: 0 = dgrad kernel + wgrad kernel, 1 = dgrad kernel only, 2 = wgrad kernel only (profiling)
int tc_launch_field_backward(const FieldParams &fp, int max_samples, cudaStream_t st, int part = 0);
size_t tc_wgrad_scratch_bytes(int max_samples);
// field_bf.cu: 3xF16 build, same entry points; its weight stream (f16 hi/lo) fits in the first half of the tc region
int bf_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st, int *range_flag = nullptr);
int bf_launch_field_forward(const FieldParams &fp, int max_samples, cudaStream_t st, int part = 0);
int bf_launch_field_backward(const FieldParams &fp, int max_samples, cudaStream_t st, int part = 0);
size_t bf_wgrad_scratch_bytes(int max_samples);
void bf_set_save_activations(int on);   // PSLAM_OPT_SAVE_ACT
// stand-alone trilinear stages over feature rows [P,16] (field_bf.cu), shared with the width-256 build
int launch_tri_gather_rows(const FieldParams &fp, float *feat, int max_samples, cudaStream_t st);
int launch_tri_scatter_rows(const FieldParams &fp, const float *g_feat, int max_samples, cudaStream_t st);
int launch_grad_scale(const FieldParams &fp, uint32_t *gscale, cudaStream_t st);
// field_w256.cu: 3xF16 tcgen05 build of the width-256 decoder (forward, dgrad chain, weight gradients)
constexpr int kSimt256Floats = 256 * 256 + 288 * 256;   // floats of the SIMT pack of the width-256 decoder (FieldCfg<256>::WS); the stream follows it
constexpr int kW2PackFloats = 278528 + 16;              // f16 hi / lo weight stream of field_w256.cu
int w2_pack_decoder(const pslam_decoder_t &d, float *ws_tc, cudaStream_t st, int *range_flag = nullptr);
bool w2_usable(const FieldParams &fp, int max_samples, bool bwd);
int w2_launch_field_forward(const FieldParams &fp, int max_samples, cudaStream_t st, int part = 0);
int w2_launch_field_backward(const FieldParams &fp, int max_samples, cudaStream_t st, int part = 0);
size_t w2_scratch_bytes(int max_samples);
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
SideStream *side_stream();            // one non-blocking side stream + fork / join events per device (field_bf.cu)

}  // namespace pslam
