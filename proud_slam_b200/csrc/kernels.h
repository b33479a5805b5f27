// Internal launch entry points shared between the .cu files of libproud_b200.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/proud_slam_b200.h"

namespace pslam {

// scratch_i layout: [intersect block hits: R/32 + 8][sample look-back state (64-bit per block of >= 8 rays): 2 x (R/8 + 8)][composite partials: 8 x R/8 + 64]
static inline int scratch_i_sample_off(int R) { return (R + 31) / 32 + 8; }
static inline int scratch_i_sample_len(int R) { return 2 * ((R + 7) / 8 + 8); }
static inline int scratch_i_composite_off(int R) { return scratch_i_sample_off(R) + scratch_i_sample_len(R); }

// intersect.cu
int scan_partials(int *partials, int nb, int *total_out, cudaStream_t st);
int launch_intersect_fused(const pslam_render_t *p, cudaStream_t st);
int walk_mode();
void set_walk_mode(int m);   // PSLAM_OPT_WALK
int launch_build_node_cache(int N, const float *centres, const int *structure, void *node_cache, cudaStream_t st);
// sample.cu
int launch_sample_fused(const pslam_render_t *p, cudaStream_t st);
// field.cu (trilinear lookup + decoder MLP)
int launch_field_forward(const pslam_render_t *p, cudaStream_t st, int part = 0);   // part 3: the decoder kernel only (profiling)
int fork_decoder_pack(const pslam_render_t *p, cudaStream_t st, cudaEvent_t *packed);   // then launch_field_forward(..., part 4) after waiting for *packed
int launch_field_backward(const pslam_render_t *p, cudaStream_t st, int part = 0);  // part 1/2: dgrad / wgrad stage only, 3: the dgrad kernel only
// composite.cu (SDF->weights compositing + loss)
int launch_composite_forward(const pslam_render_t *p, cudaStream_t st, int fold_prologue = 0);   // fold_prologue: the loss kernel also does the backward's prologue
int launch_composite_backward(const pslam_render_t *p, cudaStream_t st, int prologue_done = 0);
int launch_composite_backward_ext(const pslam_render_t *p, const float *g_color, const float *g_depth, const float *g_sdf,
                                  const float *g_weight, cudaStream_t st);
int launch_loss_coeffs(const pslam_render_t *p, const double *rows, int nrows, cudaStream_t st);

}  // namespace pslam
