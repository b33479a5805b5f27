/*
 * proud_slam_b200.h -- C ABI of libproud_b200.so (sm_100a).
 *
 * Drop-in boundary for the Proud-SLAM mapping/tracking render path.  Every
 * entry point takes plain device pointers, sizes and a CUDA stream
 * (cudaStream_t passed as void*); no torch types cross this boundary.  Each
 * declaration cites the reference interface it replaces (paths relative to the
 * reference repository).  Conventions:
 *
 *   - return value: 0 = ok, <0 = bad argument (PSLAM_E_*), >0 = cudaError_t of
 *     the failing launch.  Nothing aborts or exit()s (the reference's
 *     CUDA_CHECK_ERRORS does, sparse_voxels/include/cuda_utils.h:37-48);
 *     pslam_last_error() returns a thread-local message for the last failure.
 *   - all launches are asynchronous on `stream`; the library retains no caller
 *     pointer after return.  What it does keep, per process: the options of
 *     pslam_set_option (plain ints), the debug-trace pointers, and per DEVICE a
 *     side stream + two events (forked work of one step), the "shared-memory
 *     attribute already set" flags, and a mutex-guarded record of which
 *     workspace holds the activations saved by the last forward.  Calls on
 *     different devices are independent; two host threads may drive two
 *     pipelines of one device, but one pslam_render_t / workspace belongs to
 *     one stream at a time.
 *   - outputs and workspaces are caller-allocated (the reference allocates
 *     inside C++, intersect.cpp:98-106 / sample.cpp:80-89; the Python shim
 *     proud_slam_b200/grid.py does that allocation instead).
 *   - tensors are dense row-major fp32 / int32 unless stated.
 */
#ifndef PROUD_SLAM_B200_H
#define PROUD_SLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSLAM_ABI_VERSION 2
#define PSLAM_E_ARG (-1)      /* null pointer / non-positive size */
#define PSLAM_E_RANGE (-2)    /* size outside what the kernels support */
#define PSLAM_E_ALIGN (-3)    /* pointer not 16-byte aligned where required */

typedef void *pslam_stream_t; /* cudaStream_t */

int pslam_abi_version(void);
const char *pslam_last_error(void);
/* Device properties the host side sizes grids with: out[0]=SM count,
 * out[1]=compute capability*10, out[2]=max opt-in smem per block. */
int pslam_device_info(int *out3);
/* Process-wide configuration.  PSLAM_OPT_DECODER selects the build of the width-128 decoder:
 * 0 = tcgen05 tensor cores with 3xTF32 operand splitting (fp32-equivalent), 1 = fp32 SIMT build,
 * 2 = tcgen05 tensor cores with 3xF16 operand splitting and power-of-two operand scales (22 significant bits per
 * operand, fp32-equivalent; default).  Width 256 always runs the SIMT build. */
#define PSLAM_OPT_DECODER 1
/* PSLAM_OPT_SAVE_ACT (3xF16 build): 1 (default) = a forward that will be followed by a decoder-gradient
 * backward (PSLAM_F_GRAD_DEC with a wgrad workspace) spills its activations and ReLU masks, and that backward
 * runs the gradient chain only; 0 = the backward always recomputes the forward. */
#define PSLAM_OPT_SAVE_ACT 2
/* PSLAM_OPT_PDL: 1 (default) = the kernels of the fused step are launched with programmatic stream serialization
 * (each starts with griddepcontrol.wait, so only launch latency overlaps, never data); 0 = plain stream order. */
#define PSLAM_OPT_PDL 3
/* PSLAM_OPT_TILES (3xF16 build): 2 (default) = two 128-sample tiles in flight per CTA, the epilogue of one under the MMAs of
 * the other (csrc/field_pp.cu); 1 = one tile per CTA (csrc/field_bf.cu).  Results are identical up to the fp32 order of
 * the sdf / colour heads. */
#define PSLAM_OPT_TILES 4
/* PSLAM_OPT_FUSED_WGRAD (3xF16 build, with PSLAM_OPT_TILES = 2): 1 (default) = the mapping backward runs the dgrad chain and the
 * weight-gradient MMAs in one kernel (csrc/field_bw.cu; no gradient operand leaves the SM); 0 = chain kernel + k_wgrad_bf through
 * the HBM scratch. */
#define PSLAM_OPT_FUSED_WGRAD 5
/* PSLAM_OPT_FUSED_SCATTER (with PSLAM_OPT_FUSED_WGRAD): 1 = the trilinear backward (embedding scatter, ray gradients) of every
 * finished tile runs in two otherwise idle warps of that kernel; 0 (default: measured faster) = as a kernel of its own behind it. */
#define PSLAM_OPT_FUSED_SCATTER 6
/* PSLAM_OPT_WALK: octree walk of the fused pipeline.  0 (default) = a warp per ray (k_intersect_warp), 1 = block-cooperative,
 * one trip per octree level over a shared queue of slab tests (k_intersect_bfs).  Identical results. */
#define PSLAM_OPT_WALK 7
int pslam_set_option(int key, int value);

/* ------------------------------------------------------------------------
 * `grid` module, third_party/sparse_voxels/src/binding.cpp:12-20
 * ---------------------------------------------------------------------- */

/* svo_intersect, include/intersect.h:14-15, src/intersect.cpp:83-112,
 * kernel src/intersect_gpu.cu:191-270.  ray_start/ray_dir [b,m,3], points
 * [b,n,3], children [b,n,9] -> idx [b,m,n_max] (-1 padded), min_depth /
 * max_depth [b,m,n_max] (0 padded).  Hits are in the reference's DFS order. */
int pslam_svo_intersect(int b, int n, int m, float voxelsize, int n_max,
                        const float *ray_start, const float *ray_dir,
                        const float *points, const int *children,
                        int *idx, float *min_depth, float *max_depth,
                        pslam_stream_t stream);

/* aabb_intersect, include/intersect.h:12-13, src/intersect.cpp:49-76, kernel :142-189. */
int pslam_aabb_intersect(int b, int n, int m, float voxelsize, int n_max,
                         const float *ray_start, const float *ray_dir,
                         const float *points,
                         int *idx, float *min_depth, float *max_depth,
                         pslam_stream_t stream);

/* ball_intersect, include/intersect.h:10-11, src/intersect.cpp:15-42, kernel :13-73. */
int pslam_ball_intersect(int b, int n, int m, float radius, int n_max,
                         const float *ray_start, const float *ray_dir,
                         const float *points,
                         int *idx, float *min_depth, float *max_depth,
                         pslam_stream_t stream);

/* triangle_intersect, include/intersect.h:16-17, src/intersect.cpp:119-146,
 * kernel :272-387.  face_points [b,n,9] -> idx [b,m,n_max], depth
 * [b,m,n_max,3], uv [b,m,n_max,2]. */
int pslam_triangle_intersect(int b, int n, int m, float cagesize, float blur, int n_max,
                             const float *ray_start, const float *ray_dir,
                             const float *face_points,
                             int *idx, float *depth, float *uv,
                             pslam_stream_t stream);

/* inverse_cdf_sampling, include/sample.h:13-15, src/sample.cpp:56-95, kernel
 * src/sample_gpu.cu:133-239.  pts_idx/min_depth/max_depth/probs
 * [b,num_rays,max_hits], steps [b,num_rays], uniform_noise and the three
 * outputs [b,num_rays,max_steps] (outputs are fully initialised: idx=-1,
 * depth=dists=0).  Reproduces the reference's tail-loop behaviour (SURVEY
 * Appendix A-Q7) bit for bit. */
int pslam_inverse_cdf_sampling(int b, int num_rays, int max_hits, int max_steps,
                               float fixed_step_size,
                               const int *pts_idx, const float *min_depth,
                               const float *max_depth, const float *uniform_noise,
                               const float *probs, const float *steps,
                               int *sampled_idx, float *sampled_depth,
                               float *sampled_dists, pslam_stream_t stream);

/* uniform_ray_sampling, include/sample.h:10-12, src/sample.cpp:21-54, kernel :13-131. */
int pslam_uniform_ray_sampling(int b, int num_rays, int max_hits, int max_steps,
                               float step_size,
                               const int *pts_idx, const float *min_depth,
                               const float *max_depth, const float *uniform_noise,
                               int *sampled_idx, float *sampled_depth,
                               float *sampled_dists, pslam_stream_t stream);

/* Test helper: out[i] = __fdividef(1.0f, in[i]) -- the device reciprocal the
 * slab test uses (intersect_gpu.cu:91-101), so a CPU oracle can be compared
 * bit for bit. */
int pslam_debug_rcp(const float *in, float *out, int n, pslam_stream_t stream);
/* Test helper: D[128,N] = A[128,K] * B[N,K]^T on one CTA through the same tcgen05 / tensor-memory
 * primitives the decoder uses (split3 != 0: 3xTF32).  N in 16..144 step 16, K in 8..144 step 8. */
/* Test helper: timeline trace of CTA 0 of the tcgen05 decoder kernels into dev_buf[4*10*8] (clock64), NULL = off. */
int pslam_debug_tc_trace(long long *dev_buf);
int pslam_debug_umma_gemm(const float *A, const float *B, float *D, int N, int K, int split3, pslam_stream_t stream);
/* Same for the 3xBF16 build (kind::f16).  mode 0: D = A[128,K] * B[N,K]^T with A packed in tensor memory and B
 * K-major in shared memory; mode 1: D = At[K,128]^T * Bt[K,N], both operands MN-major in shared memory (the
 * weight-gradient form; K <= 64).  N in 16..144 step 16, K a multiple of 16. */
int pslam_debug_umma_gemm_bf(const float *A, const float *B, float *D, int N, int K, int mode, pslam_stream_t stream);
int pslam_debug_bf_trace(long long *dev_buf);
/* Same for the two-tiles-in-flight kernels (csrc/field_pp.cu): dev_buf[4 iterations][worker g0, worker g1, issuer g0, issuer g1][16]. */
int pslam_debug_pp_trace(long long *dev_buf);
/* Same for the fused backward (csrc/field_bw.cu): dev_buf[4 tiles][worker, issuer][16]. */
int pslam_debug_bw_trace(long long *dev_buf);
/* Per-warp timeline of the sampling kernel (k_sample_warp, 8 warps per block): [block][warp][8] (globaltimer at entry, clock64 at
 * entry / hits staged / rays sampled / offsets known / copied out, globaltimer at exit, the warp's largest sample count).
 * NULL switches it off. */
int pslam_debug_sample_trace(long long *dev_buf);
/* Same for the octree walk (k_intersect_warp, 32 warps = 32 rays per block): [block][warp][8] (globaltimer at entry, clock64 at
 * entry / after the walk / after ranking + write-out / at exit, globaltimer at exit, trips of the walk, the ray's hit count). */
int pslam_debug_intersect_trace(long long *dev_buf);

/* ------------------------------------------------------------------------
 * Torch-level stages of render_rays (src/variations/render_helpers.py)
 * ---------------------------------------------------------------------- */

/* Decoder parameters in the reference's own layout (nn.Linear weight [out,in],
 * src/variations/nrgbd.py:106-113; depth=2, skips=[], embedder "none",
 * sdf_dim=128, in_dim=16).  width is 128 (Replica) or 256 (ScanNet/ARKit). */
typedef struct {
    int width;
    const float *W1, *b1;   /* [w,16], [w]        */
    const float *W2, *b2;   /* [w,w], [w]         */
    const float *W3, *b3;   /* [129,w], [129]     */
    const float *W4, *b4;   /* [w,144], [w]       */
    const float *W5, *b5;   /* [3,w], [3]         */
} pslam_decoder_t;

/* Gradients, same shapes; accumulated into (+=).  Any pointer may be NULL only
 * if all are (decoder gradients disabled). */
typedef struct {
    float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4, *W5, *b5;
} pslam_decoder_grad_t;

/* get_features_vox + get_embeddings_vox + trilinear_interp,
 * render_helpers.py:105-156, 87-99, 47-59.  xyz [p,3], vox_idx [p] (row ids
 * into centres/vertex_idx), centres [N,3], vertex_idx [N,8], emb [E,16]
 * -> feat [p,16]. */
int pslam_trilinear_fwd(int p, const float *xyz, const int *vox_idx,
                        const float *centres, const int *vertex_idx, const float *emb,
                        float voxel_size, float *feat, pslam_stream_t stream);
/* Backward of the above: g_feat [p,16] -> g_emb [E,16] (+=, atomics) and
 * g_xyz [p,3] (either may be NULL). */
int pslam_trilinear_bwd(int p, const float *xyz, const int *vox_idx,
                        const float *centres, const int *vertex_idx, const float *emb,
                        float voxel_size, const float *g_feat,
                        float *g_emb, float *g_xyz, pslam_stream_t stream);

/* Floats of workspace the decoder kernels need for `width` (repacked weights;
 * rewritten by every call, so it may be shared by calls on one stream). */
int64_t pslam_decoder_ws_count(int width);

/* Decoder.get_values, nrgbd.py:116-135: feat [p,16] -> out [p,4] = (r,g,b,sdf).
 * ws: pslam_decoder_ws_count(width) floats, 16-byte aligned. */
int pslam_decoder_fwd(int p, const pslam_decoder_t *dec, const float *feat,
                      float *ws, float *out, pslam_stream_t stream);
/* Backward: g_out [p,4] -> g_feat [p,16] (may be NULL) and parameter
 * gradients (grad may be NULL).  Activations are recomputed, not stored. */
/* wgrad_ws: pslam_wgrad_ws_bytes(p) bytes of scratch for the tensor-core weight-gradient kernel
 * (width 128), or NULL to run the SIMT build when parameter gradients are requested. */
int64_t pslam_wgrad_ws_bytes(int max_samples);
/* ... for a given decoder width (256: the workspace of csrc/field_w256.cu, 7.3 kB per sample) */
int64_t pslam_wgrad_ws_bytes_w(int max_samples, int width);
int pslam_decoder_bwd(int p, const pslam_decoder_t *dec, const float *feat,
                      float *ws, const float *g_out, float *g_feat,
                      const pslam_decoder_grad_t *grad, void *wgrad_ws, int64_t wgrad_ws_bytes,
                      pslam_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused render + loss + backward (one mapping / tracking iteration):
 * render_rays (render_helpers.py:351-556) + Criterion.forward
 * (src/criterion.py:16-116) + loss.backward(), with zero host syncs.
 * ---------------------------------------------------------------------- */
#define PSLAM_F_TRACKING   1   /* median-gated depth loss (criterion.py:45-49) */
#define PSLAM_F_GRAD_EMB   2
#define PSLAM_F_GRAD_DEC   4
#define PSLAM_F_GRAD_RAYS  8
#define PSLAM_F_FORWARD_ONLY 16
#define PSLAM_F_DEFER_LOSS 32   /* forward stops at this rank's raw loss sums (loss_raw); the
                                  caller exchanges them and calls pslam_loss_finalize */
#define PSLAM_F_NODE_CACHE_VALID 64   /* node_cache already holds the child records of the bound map (built by an earlier step
                                         on the same centres / structure, or by pslam_build_node_cache): the step does not
                                         rebuild it.  The octree only changes when a keyframe is inserted
                                         (src/mapping.py:258-295), the BA loop renders the same map ~10-50 times in between. */

/* device-side counters written by the pipeline (int32 each) */
enum {
    PSLAM_C_RH = 0,       /* rays that hit (R_h) */
    PSLAM_C_P = 1,        /* max hits per ray after trimming (P / H) */
    PSLAM_C_NSAMP = 2,    /* total valid samples */
    PSLAM_C_S = 3,        /* max samples per ray (S) */
    PSLAM_C_OVERFLOW = 4, /* bit 0: sample_cap too small, bit 1: DFS stack overflow, bit 2: decoder operand outside the 3xF16 range,
                             bit 4 (16): a peer rank did not arrive at a cross-GPU exchange */
    PSLAM_C_TILE = 5,     /* internal work counters */
    PSLAM_C_TILE2 = 6,
    PSLAM_C_STICKY = 7,   /* OR of PSLAM_C_OVERFLOW over all steps since the caller last cleared it, | 8 if a step had no hit ray.
                             The next step folds the previous step's flags in before it clears the counters, so a loop of
                             pslam_render_step calls can be checked once at its end without a host sync per step. */
    PSLAM_C_STEPS = 8,    /* steps started since the caller last cleared it */
    PSLAM_C_COUNT = 16
};

/* loss block written by the pipeline (fp32 each) */
enum {
    PSLAM_L_TOTAL = 0, PSLAM_L_COLOR = 1, PSLAM_L_DEPTH = 2, PSLAM_L_FS = 3, PSLAM_L_SDF = 4,
    PSLAM_L_COUNT = 16
};

/* Data-parallel mapping over the GPUs of one box (rays / keyframes sharded, map and decoder replicated; the
 * reference is single GPU).  Every rank passes the same table: device pointers, valid on THIS device, to every rank's
 * peer-mapped exchange area (`sync`, pslam_peer_sync_bytes() bytes, zeroed once by its owner) and flat gradient buffer
 * (`flat`, [flat_count] floats, flat_count % 4 == 0; g_emb / g_dec of the step point into this rank's own copy).  With
 * world >= 2, pslam_render_step closes the loss over all ranks' raw sums (peer stores + flags inside the loss kernel: no
 * host-side collective between forward and backward) and, when `flat` is set, ends with a two-shot sum all-reduce of the
 * flat buffers over NVLink peer memory (csrc/peer.cu).  Every rank must run the same sequence of steps. */
#define PSLAM_MAX_PEERS 8
typedef struct {
    int world, rank;                     /* world <= 1: single GPU, everything below ignored */
    void *sync[PSLAM_MAX_PEERS];
    float *flat[PSLAM_MAX_PEERS];        /* all NULL: no gradient all-reduce inside the step */
    int64_t flat_count;
    void *stage[PSLAM_MAX_PEERS];        /* optional: every rank's staging area of pslam_peer_stage_bytes(flat_count, world) bytes
                                            (peer-mapped, zeroed once).  With it, buffers up to 4 MB take the low-latency
                                            all-reduce: data and flag travel in the same 16-byte store, no barrier round trips. */
} pslam_peer_t;
int64_t pslam_peer_sync_bytes(void);
int64_t pslam_peer_stage_bytes(int64_t flat_count, int world);
/* The all-reduce alone (sum over ranks, in place in every rank's `flat`); fail_flag: optional device int that gets bit 4
 * (value 16) if a peer did not arrive within ~2 s. */
int pslam_peer_allreduce(const pslam_peer_t *peer, int *fail_flag, pslam_stream_t stream);

typedef struct {
    /* sizes */
    int R;              /* rays in the batch */
    int N;              /* octree rows */
    int E;              /* embedding rows */
    int n_max;          /* hit cap per ray (reference hard-codes 50, voxel_helpers.py:561) */
    int sample_cap;     /* capacity (samples) of the CSR sample arrays */
    int flags;          /* PSLAM_F_* */
    float voxel_size, step_size, truncation, max_distance, max_depth;
    float w_rgb, w_depth, w_fs, w_sdf;
    /* inputs */
    const float *rays_o, *rays_d;          /* [R,3] */
    const float *target_rgb;               /* [R,3] */
    const float *target_depth;             /* [R]   */
    const float *centres;                  /* [N,3] voxel_center_xyz */
    const int *structure;                  /* [N,9] voxel_structure  */
    const int *vertex_idx;                 /* [N,8] voxel_vertex_idx */
    const float *emb;                      /* [E,16] voxel_vertex_emb */
    pslam_decoder_t dec;
    float *dec_ws;                         /* [pslam_decoder_ws_count(dec.width)] repacked weights */
    void *wgrad_ws;                        /* pslam_wgrad_ws_bytes_w(sample_cap, dec.width) bytes (wgrad operands, ReLU masks, feature rows),
                                              or NULL: the fp32 SIMT build then serves any call that needs a backward */
    int64_t wgrad_ws_bytes;
    const float *noise;                    /* [>=R_h, noise_stride] uniform(0.001,0.999) or NULL */
    int noise_stride;
    uint64_t seed;                         /* counter-based noise when noise==NULL */
    const uint64_t *seed_dev;              /* optional device-side addend to `seed` (lets a CUDA graph replay draw fresh noise) */
    /* intermediates (caller-allocated; readable by tests / the drop-in API) */
    int *hit_idx;                          /* [n_max,R] slot-major, sorted by entry depth */
    float *hit_min, *hit_max;              /* [n_max,R] */
    int *hit_count;                        /* [R] valid hits per ray after trimming */
    int *hit_ray;                          /* [R] rank -> ray id (first R_h valid) */
    int *ray_rank;                         /* [R] ray id -> rank or -1 */
    int *samp_off;                         /* [R+1] CSR offsets by rank */
    int *samp_vox;                         /* [sample_cap] voxel row id */
    int *samp_ray;                         /* [sample_cap] rank of the owning ray */
    float *samp_z;                         /* [sample_cap] depth (segment mid-point) */
    float *samp_dist;                      /* [sample_cap] segment length */
    float *samp_out;                       /* [sample_cap,4] (r,g,b,sdf) from the decoder */
    float *samp_w;                         /* [sample_cap] compositing weight per sample (may be NULL) */
    float *samp_gout;                      /* [sample_cap,4] dL/d(r,g,b,sdf) */
    float *ray_out;                        /* [R,8] by rank: r,g,b,depth,z_min,U,|dd|/sqrt(var),gate */
    int *scratch_i;                        /* [pslam_render_scratch_i_count(R)] block partials (scans, loss counts) */
    float *scratch_f;                      /* [pslam_render_scratch_f_count(R)] block partials (loss sums) */
    int *counters;                         /* [PSLAM_C_COUNT] */
    /* outputs */
    double *loss_raw;                      /* [16] this rank's raw loss sums (see composite.cu RAW_*) */
    float *loss;                           /* [PSLAM_L_COUNT] */
    float *g_emb;                          /* [E,16] += */
    pslam_decoder_grad_t g_dec;            /* += */
    float *g_rays_o, *g_rays_d;            /* [R,3] by ray id, overwritten */
    /* optional traversal cache: [N,8] x 16 B child records (child row id + child centre), built by pslam_render_sample /
     * pslam_render_step from `structure` / `centres` unless PSLAM_F_NODE_CACHE_VALID is set; halves the dependent-load
     * chain of the octree walk.  NULL or node_cache_bytes < 128*N: the walk reads the two arrays directly. */
    void *node_cache;
    int64_t node_cache_bytes;
    pslam_peer_t peer;                     /* multi-GPU exchange tables (world <= 1: unused) */
} pslam_render_t;

/* sizeof(pslam_render_t) and offsetof(.., loss): lets a binding check its mirror of the struct. */
int pslam_render_sizeof(void);
int pslam_render_offsetof_loss(void);

/* Elements the caller must provide for scratch_i / scratch_f for a batch of R rays. */
int64_t pslam_render_scratch_i_count(int R);
int64_t pslam_render_scratch_f_count(int R);

/* The traversal cache alone: node_cache [N,8] x 16 B from centres [N,3] / structure [N,9] (once per map generation). */
int pslam_build_node_cache(int N, const float *centres, const int *structure, void *node_cache, pslam_stream_t stream);

/* Stage 1: intersection, sort/trim, compaction, sampling (kernels 1-2 + a4-a6).
 * Fills hit_*, samp_* and counters.  No host sync. */
int pslam_render_sample(const pslam_render_t *p, pslam_stream_t stream);
/* Stage 2: trilinear lookup + decoder + compositing + loss (kernels 3-5 forward). */
int pslam_render_forward(const pslam_render_t *p, pslam_stream_t stream);
/* Multi-GPU: closes the loss from the raw sums of all ranks (rows [nrows,16],
 * e.g. an all-gather of every rank's loss_raw): sums add, S is the maximum
 * (criterion.py:70-112 couples all rays through global means and counts). */
int pslam_loss_finalize(const pslam_render_t *p, const double *rows, int nrows, pslam_stream_t stream);
/* Stage 3: backward of stage 2 into g_emb / g_dec / g_rays_*. */
int pslam_render_backward(const pslam_render_t *p, pslam_stream_t stream);
/* Stage 3 for a caller-side loss: backward of stage 2 from arbitrary upstream gradients w.r.t.
 * the render_rays outputs -- g_color [R_h,3], g_depth [R_h] (by hit-ray rank), g_sdf and g_weight
 * per sample in CSR order; any of them may be NULL.  This is what the autograd wrapper of the
 * drop-in render_rays (render_helpers.py:351-556) calls. */
int pslam_render_backward_ext(const pslam_render_t *p, const float *g_color, const float *g_depth,
                              const float *g_sdf, const float *g_weight, pslam_stream_t stream);
/* All three stages back to back. */
int pslam_render_step(const pslam_render_t *p, pslam_stream_t stream);
/* Profiling hook: one stage of the step (0 intersect, 1 sampling, 2 field fwd, 3 composite fwd + loss,
 * 4 composite bwd, 5 field bwd) so a benchmark can bracket a single kernel with events. */
int pslam_render_stage(const pslam_render_t *p, int stage, pslam_stream_t stream);

/* ------------------------------------------------------------------------
 * SE(3) pose of the tracking loop (src/se3pose.py:24-34, 62-91: 6-vector (t, w), Rodrigues' formula with the
 * reference's 11-term series; render_helpers.py:679-761: per iteration the sampled camera-frame ray directions
 * are rotated by the current pose, and the pose takes one Adam step from the rendered loss).
 * pslam_track_assemble: rays_o[i] = t, rays_d[i] = R(w) rays_d_cam[idx[i]], rgb / depth gathered with the same
 * indices (NULL outputs are skipped).  pslam_track_pose_step: dL/dpose from dL/d(rays_o, rays_d) (the analytic
 * derivative of the same series) followed by torch.optim.Adam's update, in place on pose6 and on the optimizer's
 * exp_avg [6] / exp_avg_sq [6] / step [1] (float, the capturable form; step == NULL: the count lives on the host
 * and step_value is its new value); grad_out [6] optional.
 * ---------------------------------------------------------------------- */
/* n distinct pixel indices out of hw, uniform (frame.sample_rays / src/utils/sample_util.py:4-20: Gumbel top-k over uniform
 * weights = uniform sampling without replacement), from a keyed permutation of [0, hw) evaluated at 0..n-1; seed_dev is an
 * optional device-side addend to `seed` (a CUDA-graph replay then draws a fresh set). */
int pslam_sample_pixels(int n, long long hw, unsigned long long seed, const unsigned long long *seed_dev, long long *idx,
                        pslam_stream_t stream);
int pslam_track_assemble(int n, const float *pose6, const long long *idx, const float *rays_d_cam, const float *rgb_all,
                         const float *depth_all, float *rays_o, float *rays_d, float *rgb, float *depth, pslam_stream_t stream);
/* pslam_sample_pixels + pslam_track_assemble in one launch (a captured tracking iteration): idx [n] receives the selection. */
int pslam_track_sample_assemble(int n, long long hw, unsigned long long seed, const unsigned long long *seed_dev, const float *pose6,
                                const float *rays_d_cam, const float *rgb_all, const float *depth_all, long long *idx, float *rays_o,
                                float *rays_d, float *rgb, float *depth, pslam_stream_t stream);
int pslam_track_pose_step(int n, float *pose6, const long long *idx, const float *rays_d_cam, const float *g_rays_o,
                          const float *g_rays_d, float *exp_avg, float *exp_avg_sq, float *step, double step_value, double lr,
                          double beta1, double beta2, double eps, float *grad_out, pslam_stream_t stream);
/* The same step as the last kernel of a CAPTURED tracking iteration (device-side Adam count only), which also does the
 * iteration's bookkeeping: hit_mask[i] = hit_count[i] > 0 (track_frame's third return value, render_helpers.py:741; both or
 * neither) and *iter_counter += 1 (the device-side addend of the next iteration's pixel selection / sampling-noise seeds). */
int pslam_track_pose_step_iter(int n, float *pose6, const long long *idx, const float *rays_d_cam, const float *g_rays_o,
                               const float *g_rays_d, float *exp_avg, float *exp_avg_sq, float *step, double lr, double beta1,
                               double beta2, double eps, unsigned long long *iter_counter, const int *hit_count,
                               unsigned char *hit_mask, pslam_stream_t stream);

/* ------------------------------------------------------------------------
 * Optimizer step of the mapping loop (src/mapping.py:81-82: torch.optim.Adam over the embedding table and over the decoder,
 * stepped at src/variations/render_helpers.py:667-676).  One launch for up to 16 tensors, in place on the tensors
 * torch.optim.Adam owns; torch's arithmetic without weight decay / amsgrad.  row = 16 with row_active ([n / 16] bytes, zeroed
 * by the caller once): rows that never received a gradient are skipped (exact: zero state and zero gradient give a zero
 * update).  step: the optimizer's device-side count (capturable form), incremented here; NULL: the count lives on the host
 * and step_value is its NEW value.  zero_grad != 0 clears the gradients in the same pass.
 * ---------------------------------------------------------------------- */
typedef struct {
    float *param, *grad, *exp_avg, *exp_avg_sq;   /* [n], 16-byte aligned */
    float *step;                                  /* [1] or NULL */
    unsigned char *row_active;                    /* [n / row] or NULL (dense) */
    int64_t n;
    int row;                                      /* 0 = dense, 16 = embedding rows */
    float lr;
} pslam_adam_tensor_t;
int pslam_adam_step(const pslam_adam_tensor_t *tensors, int count, double step_value, double beta1, double beta2, double eps,
                    int zero_grad, pslam_stream_t stream);

/* ------------------------------------------------------------------------
 * Host-side octree: torch.classes.svo.Octree,
 * third_party/sparse_octree/src/bindings.cpp:11-35 (init / insert /
 * get_centres_and_children / has_voxel / count_nodes / count_leaf_nodes /
 * get_leaf_voxels).  Plain host pointers; the producer of `map_states`.
 * ---------------------------------------------------------------------- */
void *pslam_octree_new(int grid_dim);                       /* Octree::init, octree.cpp:46-60 */
void pslam_octree_free(void *tree);
int pslam_octree_count(void *tree);                         /* count_nodes */
int pslam_octree_count_leaves(void *tree);                  /* count_leaf_nodes (SURFACE voxels) */
int pslam_octree_insert(void *tree, const int *vox, int m); /* Octree::insert, octree.cpp:104-294; vox [m,3] */
int pslam_octree_has_voxel(void *tree, int x, int y, int z);
/* get_centres_and_children, octree.cpp:561-687: voxels [N,4] f32, children [N,8] f32, features [N,8] i32 */
int pslam_octree_flatten(void *tree, float *voxels, float *children, int *features);
int pslam_octree_leaf_voxels(void *tree, int *out_xyz, int cap);

/* ------------------------------------------------------------------------
 * Device-side octree (csrc/octree_dev.cu): the same insert / get_centres_and_children with the same row ids (creation order
 * of the sequential insertion, octree.h:41), built on the GPU from device tensors and flattened into device tensors -- the
 * producer of `map_states` where the map is consumed (src/mapping.py:258-295, 301-406 re-flatten a host tree and upload it per
 * keyframe).  vox [m,3] int32 ON THE DEVICE, 0 <= v < grid_dim - 1.  insert synchronises the stream once (array growth).
 * ---------------------------------------------------------------------- */
void *pslam_doctree_new(int grid_dim, int capacity_hint);
void pslam_doctree_free(void *tree);
int pslam_doctree_count(void *tree);
int pslam_doctree_insert(void *tree, const int *vox_dev, int m, pslam_stream_t stream);
int pslam_doctree_flatten(void *tree, float *voxels_dev, float *children_dev, int *features_dev, pslam_stream_t stream);
int pslam_doctree_types(void *tree, int *types_dev_out, pslam_stream_t stream);   /* per row: -1 internal, 0 SURFACE voxel, 1 FEATURE corner */

#ifdef __cplusplus
}
#endif
#endif /* PROUD_SLAM_B200_H */
