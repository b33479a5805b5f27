// libproud_b200.so: error reporting, device queries and the three-stage driver
// of the fused render path (include/proud_slam_b200.h).
//
// The driver replaces the Python control flow of the reference's render_rays
// (src/variations/render_helpers.py:351-556), Criterion.forward (src/criterion.py:16-68) and
// loss.backward() (render_helpers.py:671): the reference synchronises with the host at least
// seven times per call to size its padded tensors (voxel_helpers.py:582, 320, 359;
// render_helpers.py:388-433); here every data-dependent size is a device-side counter and the
// launches are enqueued back to back.  Errors are returned, never exit()ed
// (sparse_voxels/include/cuda_utils.h:37-48 exits).
#include <stdarg.h>
#include <stddef.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "field_bf.cuh"
#include "peer.cuh"

namespace pslam {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int num_sms()
{
    static int cached[64] = {0};  // per device, written idempotently
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static int g_decoder_mode = 2;   // tcgen05 3xF16 (field_pp.cu / field_bw.cu / field_bf.cu)
int decoder_mode() { return g_decoder_mode; }
static int g_pdl = 1;             // programmatic dependent launch between the kernels of the fused step
int pdl_enabled() { return g_pdl; }

static int check_render(const pslam_render_t *p)
{
    PSLAM_CHECK_ARG(p, PSLAM_E_ARG, "null pslam_render_t");
    PSLAM_CHECK_ARG(p->R > 0 && p->N > 0 && p->E > 0 && p->n_max > 0 && p->sample_cap > 0, PSLAM_E_ARG,
                    "sizes must be positive (R=%d N=%d E=%d n_max=%d sample_cap=%d)", p->R, p->N, p->E, p->n_max, p->sample_cap);
    PSLAM_CHECK_ARG(p->R <= (1 << 26), PSLAM_E_RANGE, "R=%d too large", p->R);
    PSLAM_CHECK_ARG(p->N <= (1 << 26), PSLAM_E_RANGE, "N=%d octree rows exceed the traversal's 26-bit row ids", p->N);
    PSLAM_CHECK_ARG(p->voxel_size > 0.0f && p->step_size > 0.0f && p->truncation > 0.0f, PSLAM_E_ARG, "voxel_size, step_size and truncation must be > 0");
    PSLAM_CHECK_ARG(p->rays_o && p->rays_d && p->centres && p->structure && p->vertex_idx && p->emb, PSLAM_E_ARG, "null input pointer");
    PSLAM_CHECK_ARG(p->hit_idx && p->hit_min && p->hit_max && p->hit_count && p->hit_ray && p->ray_rank && p->samp_off && p->samp_vox &&
                        p->samp_ray && p->samp_z && p->samp_dist && p->scratch_i && p->scratch_f && p->counters,
                    PSLAM_E_ARG, "null intermediate buffer");
    PSLAM_CHECK_ARG(p->noise == nullptr || p->noise_stride > 0, PSLAM_E_ARG, "noise_stride must be > 0 with an explicit noise tensor");
    if (p->peer.world > 1) {
        PSLAM_CHECK_ARG(p->peer.world <= PSLAM_MAX_PEERS && p->peer.rank >= 0 && p->peer.rank < p->peer.world, PSLAM_E_ARG,
                        "peer table: world %d / rank %d", p->peer.world, p->peer.rank);
        for (int q = 0; q < p->peer.world; ++q) PSLAM_CHECK_ARG(p->peer.sync[q], PSLAM_E_ARG, "peer table: null exchange area of rank %d", q);
    }
    return 0;
}

static int check_render_field(const pslam_render_t *p, bool backward, bool need_targets = true)
{
    PSLAM_CHECK_ARG(p->dec.width == 128 || p->dec.width == 256, PSLAM_E_RANGE, "decoder width %d not supported (128 or 256)", p->dec.width);
    PSLAM_CHECK_ARG(p->dec.W1 && p->dec.b1 && p->dec.W2 && p->dec.b2 && p->dec.W3 && p->dec.b3 && p->dec.W4 && p->dec.b4 && p->dec.W5 && p->dec.b5,
                    PSLAM_E_ARG, "null decoder parameter");
    PSLAM_CHECK_ARG(p->dec_ws && p->samp_out && p->ray_out && p->loss && p->loss_raw, PSLAM_E_ARG, "null forward buffer");
    PSLAM_CHECK_ARG((p->target_rgb == nullptr) == (p->target_depth == nullptr), PSLAM_E_ARG, "target_rgb and target_depth go together");
    if (backward && need_targets) PSLAM_CHECK_ARG(p->target_rgb && p->target_depth, PSLAM_E_ARG, "backward needs the Criterion targets (use pslam_render_backward_ext otherwise)");
    PSLAM_CHECK_ARG(((uintptr_t)p->dec.W1 | (uintptr_t)p->dec.W2 | (uintptr_t)p->dec.W3 | (uintptr_t)p->dec.W4 | (uintptr_t)p->dec_ws |
                     (uintptr_t)p->samp_out | (uintptr_t)p->emb) % 16 == 0,
                    PSLAM_E_ALIGN, "decoder weights, dec_ws, samp_out and emb must be 16-byte aligned");
    if (backward) {
        PSLAM_CHECK_ARG(p->samp_gout && ((uintptr_t)p->samp_gout % 16 == 0), PSLAM_E_ARG, "samp_gout missing or misaligned");
        if (p->flags & PSLAM_F_GRAD_EMB) PSLAM_CHECK_ARG(p->g_emb && ((uintptr_t)p->g_emb % 16 == 0), PSLAM_E_ARG, "g_emb missing or misaligned");
        if (p->flags & PSLAM_F_GRAD_RAYS) PSLAM_CHECK_ARG(p->g_rays_o && p->g_rays_d, PSLAM_E_ARG, "g_rays_o/g_rays_d missing");
        if (p->flags & PSLAM_F_GRAD_DEC)
            PSLAM_CHECK_ARG(p->g_dec.W1 && p->g_dec.b1 && p->g_dec.W2 && p->g_dec.b2 && p->g_dec.W3 && p->g_dec.b3 && p->g_dec.W4 &&
                                p->g_dec.b4 && p->g_dec.W5 && p->g_dec.b5,
                            PSLAM_E_ARG, "null decoder gradient pointer");
        if (p->flags & PSLAM_F_GRAD_DEC)
            PSLAM_CHECK_ARG(((uintptr_t)p->g_dec.W1 | (uintptr_t)p->g_dec.W2 | (uintptr_t)p->g_dec.W3 | (uintptr_t)p->g_dec.W4) % 16 == 0,
                            PSLAM_E_ALIGN, "decoder weight gradients must be 16-byte aligned");
    }
    return 0;
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_abi_version(void) { return PSLAM_ABI_VERSION; }

extern "C" int pslam_set_option(int key, int value)
{
    if (key == PSLAM_OPT_DECODER && (value == 0 || value == 1 || value == 2)) { g_decoder_mode = value; return 0; }
    if (key == PSLAM_OPT_SAVE_ACT && (value == 0 || value == 1)) { bf_set_save_activations(value); return 0; }
    if (key == PSLAM_OPT_PDL && (value == 0 || value == 1)) { g_pdl = value; return 0; }
    if (key == PSLAM_OPT_TILES && (value == 1 || value == 2)) { pp_set_enabled(value == 2); return 0; }
    if (key == PSLAM_OPT_FUSED_WGRAD && (value == 0 || value == 1)) { bw_set_enabled(value); return 0; }
    if (key == PSLAM_OPT_FUSED_SCATTER && (value == 0 || value == 1)) { bw_set_fused_scatter(value); return 0; }
    if (key == PSLAM_OPT_WALK && (value == 0 || value == 1)) { set_walk_mode(value); return 0; }
    set_error("unknown option %d=%d", key, value);
    return PSLAM_E_ARG;
}
extern "C" const char *pslam_last_error(void) { return g_error; }

extern "C" int pslam_device_info(int *out3)
{
    PSLAM_CHECK_ARG(out3, PSLAM_E_ARG, "null output");
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    int major = 0, minor = 0, smem = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    out3[0] = num_sms(); out3[1] = major * 10 + minor; out3[2] = smem;
    return 0;
}

extern "C" int64_t pslam_render_scratch_i_count(int R)
{
    return (int64_t)scratch_i_composite_off(R) + 8 * (int64_t)ceil_div(R, 8) + 64;
}
extern "C" int64_t pslam_render_scratch_f_count(int R) { return 8 * (int64_t)ceil_div(R, 8) + 64; }

extern "C" int pslam_build_node_cache(int N, const float *centres, const int *structure, void *node_cache, pslam_stream_t stream)
{
    PSLAM_CHECK_ARG(N > 0 && centres && structure && node_cache && ((uintptr_t)node_cache % 16 == 0), PSLAM_E_ARG, "build_node_cache: bad argument");
    return launch_build_node_cache(N, centres, structure, node_cache, (cudaStream_t)stream);
}

extern "C" int pslam_render_sample(const pslam_render_t *p, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_intersect_fused(p, st)) return rc;   // (clears the step's counters first)
    return launch_sample_fused(p, st);
}

extern "C" int pslam_render_forward(const pslam_render_t *p, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    if (int rc = check_render_field(p, false)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_field_forward(p, st)) return rc;
    return launch_composite_forward(p, st);
}

extern "C" int pslam_loss_finalize(const pslam_render_t *p, const double *rows, int nrows, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    PSLAM_CHECK_ARG(rows && nrows > 0 && p->loss, PSLAM_E_ARG, "loss_finalize: bad argument");
    return launch_loss_coeffs(p, rows, nrows, (cudaStream_t)stream);
}

extern "C" int pslam_render_backward(const pslam_render_t *p, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    if (int rc = check_render_field(p, true)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_composite_backward(p, st)) return rc;
    return launch_field_backward(p, st);
}

extern "C" int pslam_render_backward_ext(const pslam_render_t *p, const float *g_color, const float *g_depth, const float *g_sdf,
                                         const float *g_weight, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    if (int rc = check_render_field(p, true, false)) return rc;   // decoder, alignment, gradient targets; no Criterion targets needed
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_composite_backward_ext(p, g_color, g_depth, g_sdf, g_weight, st)) return rc;
    return launch_field_backward(p, st);
}

extern "C" int pslam_render_step(const pslam_render_t *p, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    if (int rc = check_render_field(p, false)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = launch_intersect_fused(p, st)) return rc;
    cudaEvent_t packed = nullptr;                               // the decoder re-pack runs underneath the sampling kernel
    if (int rc = fork_decoder_pack(p, st, &packed)) return rc;
    if (int rc = launch_sample_fused(p, st)) return rc;
    if (packed && cudaStreamWaitEvent(st, packed, 0) != cudaSuccess) {
        set_error("render_step: cudaStreamWaitEvent: %s", cudaGetErrorString(cudaGetLastError()));
        return PSLAM_E_ARG;
    }
    if (int rc = launch_field_forward(p, st, packed ? 4 : 0)) return rc;
    const bool backward = !(p->flags & (PSLAM_F_FORWARD_ONLY | PSLAM_F_DEFER_LOSS));
    const bool fold = backward && p->target_rgb && p->target_depth;   // (the loss kernel runs: it takes the backward's prologue along)
    if (int rc = launch_composite_forward(p, st, fold ? 1 : 0)) return rc;
    if (!backward) return 0;
    if (int rc = check_render_field(p, true)) return rc;
    if (int rc = launch_composite_backward(p, st, fold ? 1 : 0)) return rc;
    if (int rc = launch_field_backward(p, st)) return rc;
    if (p->peer.world > 1 && p->peer.flat[0]) return launch_peer_allreduce(&p->peer, p->counters + PSLAM_C_OVERFLOW, st);
    return 0;
}

/* sizeof / field offsets so that language bindings can verify their struct mirrors */
extern "C" int pslam_render_sizeof(void) { return (int)sizeof(pslam_render_t); }
extern "C" int pslam_render_offsetof_loss(void) { return (int)offsetof(pslam_render_t, loss); }

/* Profiling hook: launches ONE stage of the step so that a benchmark can bracket a single kernel
 * with events.  0 intersect(+compaction) 1 sampling 2 field fwd 3 composite fwd(+loss) 4 composite bwd
 * 5 field bwd (6 / 7: only its dgrad / wgrad stage in the tcgen05 builds; 8 / 9: the forward / backward decoder kernel
 * alone, without the pack / gather / scale / scatter launches around it, 3xF16 build; 10 / 11: the trilinear gather / scatter
 * kernel alone).  The preceding stages must have run on the same argument block. */
extern "C" int pslam_render_stage(const pslam_render_t *p, int stage, pslam_stream_t stream)
{
    if (int rc = check_render(p)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (stage) {
        case 0: return launch_intersect_fused(p, st);
        case 1: return launch_sample_fused(p, st);
        case 2: if (int rc = check_render_field(p, false)) return rc; return launch_field_forward(p, st);
        case 3: if (int rc = check_render_field(p, false)) return rc; return launch_composite_forward(p, st);
        case 4: if (int rc = check_render_field(p, true)) return rc; return launch_composite_backward(p, st);
        case 5: if (int rc = check_render_field(p, true)) return rc; return launch_field_backward(p, st);
        case 6: if (int rc = check_render_field(p, true)) return rc; return launch_field_backward(p, st, 1);
        case 7: if (int rc = check_render_field(p, true)) return rc; return launch_field_backward(p, st, 2);
        case 8: if (int rc = check_render_field(p, false)) return rc; return launch_field_forward(p, st, 3);
        case 9: if (int rc = check_render_field(p, true)) return rc; return launch_field_backward(p, st, 3);
        case 10: if (int rc = check_render_field(p, false)) return rc; return launch_field_forward(p, st, 5);
        case 11: if (int rc = check_render_field(p, true)) return rc; return launch_field_backward(p, st, 4);
    }
    set_error("unknown stage %d", stage);
    return PSLAM_E_ARG;
}
