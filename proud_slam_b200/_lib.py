"""ctypes binding of ``libproud_b200.so`` (the C ABI in ``include/proud_slam_b200.h``).

There is no fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised (the reference's pybind module raises
``RuntimeError`` from ``TORCH_CHECK`` the same way, sparse_voxels/include/utils.h:10-34).
PyTorch is used only for device memory and streams; no torch type crosses the
boundary -- every argument is a raw pointer or a size.
"""
import ctypes as C
import os

import torch

from . import _build

c_float_p = C.c_void_p   # raw device pointers
c_int_p = C.c_void_p


class DecoderT(C.Structure):
    _fields_ = [("width", C.c_int)] + [(n, C.c_void_p) for n in
                                      ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5")]


class DecoderGradT(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5")]


class AdamTensorT(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("param", "grad", "exp_avg", "exp_avg_sq", "step", "row_active")] + [("n", C.c_int64), ("row", C.c_int),
                                                                                                         ("lr", C.c_float)]


MAX_PEERS = 8


class PeerT(C.Structure):
    """pslam_peer_t: every rank's peer-mapped exchange area and flat gradient buffer (include/proud_slam_b200.h)."""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("sync", C.c_void_p * MAX_PEERS), ("flat", C.c_void_p * MAX_PEERS),
                ("flat_count", C.c_int64), ("stage", C.c_void_p * MAX_PEERS)]


class RenderT(C.Structure):
    _fields_ = (
        [(n, C.c_int) for n in ("R", "N", "E", "n_max", "sample_cap", "flags")]
        + [(n, C.c_float) for n in ("voxel_size", "step_size", "truncation", "max_distance", "max_depth",
                                    "w_rgb", "w_depth", "w_fs", "w_sdf")]
        + [(n, C.c_void_p) for n in ("rays_o", "rays_d", "target_rgb", "target_depth", "centres", "structure",
                                     "vertex_idx", "emb")]
        + [("dec", DecoderT), ("dec_ws", C.c_void_p), ("wgrad_ws", C.c_void_p), ("wgrad_ws_bytes", C.c_int64),
           ("noise", C.c_void_p), ("noise_stride", C.c_int),
           ("seed", C.c_uint64), ("seed_dev", C.c_void_p)]
        + [(n, C.c_void_p) for n in ("hit_idx", "hit_min", "hit_max", "hit_count", "hit_ray", "ray_rank",
                                     "samp_off", "samp_vox", "samp_ray", "samp_z", "samp_dist", "samp_out",
                                     "samp_w", "samp_gout", "ray_out", "scratch_i", "scratch_f", "counters")]
        + [("loss_raw", C.c_void_p), ("loss", C.c_void_p), ("g_emb", C.c_void_p), ("g_dec", DecoderGradT),
           ("g_rays_o", C.c_void_p), ("g_rays_d", C.c_void_p), ("node_cache", C.c_void_p), ("node_cache_bytes", C.c_int64),
           ("peer", PeerT)]
    )


# flags / counter slots / loss slots (include/proud_slam_b200.h)
F_TRACKING, F_GRAD_EMB, F_GRAD_DEC, F_GRAD_RAYS, F_FORWARD_ONLY, F_DEFER_LOSS, F_NODE_CACHE_VALID = 1, 2, 4, 8, 16, 32, 64
C_RH, C_P, C_NSAMP, C_S, C_OVERFLOW, C_STICKY, C_STEPS, C_COUNT = 0, 1, 2, 3, 4, 7, 8, 16
L_TOTAL, L_COLOR, L_DEPTH, L_FS, L_SDF, L_COUNT = 0, 1, 2, 3, 4, 16

_I, _F, _P, _S = C.c_int, C.c_float, C.c_void_p, C.c_void_p
_PROTOTYPES = {
    "pslam_abi_version": (C.c_int, []),
    "pslam_last_error": (C.c_char_p, []),
    "pslam_device_info": (C.c_int, [_P]),
    "pslam_svo_intersect": (C.c_int, [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_aabb_intersect": (C.c_int, [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_ball_intersect": (C.c_int, [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_triangle_intersect": (C.c_int, [_I, _I, _I, _F, _F, _I, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_inverse_cdf_sampling": (C.c_int, [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_uniform_ray_sampling": (C.c_int, [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_debug_rcp": (C.c_int, [_P, _P, _I, _S]),
    "pslam_debug_tc_trace": (C.c_int, [_P]),
    "pslam_debug_umma_gemm": (C.c_int, [_P, _P, _P, _I, _I, _I, _S]),
    "pslam_debug_umma_gemm_bf": (C.c_int, [_P, _P, _P, _I, _I, _I, _S]),
    "pslam_sample_pixels": (C.c_int, [_I, C.c_longlong, C.c_uint64, _P, _P, _S]),
    "pslam_track_assemble": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_track_sample_assemble": (C.c_int, [_I, C.c_longlong, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _S]),
    "pslam_track_pose_step": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P, _S]),
    "pslam_track_pose_step_iter": (C.c_int, [_I, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P, _S]),
    "pslam_adam_step": (C.c_int, [_P, _I, C.c_double, C.c_double, C.c_double, C.c_double, _I, _S]),
    "pslam_debug_bf_trace": (C.c_int, [_P]),
    "pslam_debug_pp_trace": (C.c_int, [_P]),
    "pslam_debug_bw_trace": (C.c_int, [_P]),
    "pslam_debug_sample_trace": (C.c_int, [_P]),
    "pslam_debug_intersect_trace": (C.c_int, [_P]),
    "pslam_set_option": (C.c_int, [_I, _I]),
    "pslam_decoder_ws_count": (C.c_int64, [_I]),
    "pslam_trilinear_fwd": (C.c_int, [_I, _P, _P, _P, _P, _P, _F, _P, _S]),
    "pslam_trilinear_bwd": (C.c_int, [_I, _P, _P, _P, _P, _P, _F, _P, _P, _P, _S]),
    "pslam_decoder_fwd": (C.c_int, [_I, C.POINTER(DecoderT), _P, _P, _P, _S]),
    "pslam_decoder_bwd": (C.c_int, [_I, C.POINTER(DecoderT), _P, _P, _P, _P, C.POINTER(DecoderGradT), _P, C.c_int64, _S]),
    "pslam_wgrad_ws_bytes": (C.c_int64, [_I]),
    "pslam_wgrad_ws_bytes_w": (C.c_int64, [_I, _I]),
    "pslam_render_sizeof": (C.c_int, []),
    "pslam_render_offsetof_loss": (C.c_int, []),
    "pslam_render_scratch_i_count": (C.c_int64, [_I]),
    "pslam_render_scratch_f_count": (C.c_int64, [_I]),
    "pslam_build_node_cache": (C.c_int, [_I, _P, _P, _P, _S]),
    "pslam_peer_sync_bytes": (C.c_int64, []),
    "pslam_peer_stage_bytes": (C.c_int64, [C.c_int64, _I]),
    "pslam_peer_allreduce": (C.c_int, [C.POINTER(PeerT), _P, _S]),
    "pslam_render_sample": (C.c_int, [C.POINTER(RenderT), _S]),
    "pslam_render_forward": (C.c_int, [C.POINTER(RenderT), _S]),
    "pslam_render_backward": (C.c_int, [C.POINTER(RenderT), _S]),
    "pslam_loss_finalize": (C.c_int, [C.POINTER(RenderT), _P, _I, _S]),
    "pslam_render_backward_ext": (C.c_int, [C.POINTER(RenderT), _P, _P, _P, _P, _S]),
    "pslam_render_step": (C.c_int, [C.POINTER(RenderT), _S]),
    "pslam_render_stage": (C.c_int, [C.POINTER(RenderT), _I, _S]),
    "pslam_octree_new": (C.c_void_p, [_I]),
    "pslam_octree_free": (None, [_P]),
    "pslam_octree_count": (C.c_int, [_P]),
    "pslam_octree_count_leaves": (C.c_int, [_P]),
    "pslam_octree_insert": (C.c_int, [_P, _P, _I]),
    "pslam_octree_has_voxel": (C.c_int, [_P, _I, _I, _I]),
    "pslam_octree_flatten": (C.c_int, [_P, _P, _P, _P]),
    "pslam_octree_leaf_voxels": (C.c_int, [_P, _P, _I]),
    "pslam_doctree_new": (C.c_void_p, [_I, _I]),
    "pslam_doctree_free": (None, [_P]),
    "pslam_doctree_count": (C.c_int, [_P]),
    "pslam_doctree_insert": (C.c_int, [_P, _P, _I, _S]),
    "pslam_doctree_flatten": (C.c_int, [_P, _P, _P, _P, _S]),
    "pslam_doctree_types": (C.c_int, [_P, _P, _S]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def lib():
    """Loads the library (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.is_stale() and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        _build.build_library()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the render path)")
    handle = C.CDLL(path)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    if handle.pslam_render_sizeof() != C.sizeof(RenderT):
        raise RuntimeError("pslam_render_t layout mismatch between the library and its Python mirror "
                           f"({handle.pslam_render_sizeof()} vs {C.sizeof(RenderT)})")
    if handle.pslam_render_offsetof_loss() != RenderT.loss.offset:
        raise RuntimeError("pslam_render_t field offsets differ between the library and its Python mirror")
    mode = os.environ.get("PSLAM_DECODER")   # decoder build override (include/proud_slam_b200.h: PSLAM_OPT_DECODER)
    if mode is not None and handle.pslam_set_option(1, int(mode)) != 0:
        raise RuntimeError(f"PSLAM_DECODER={mode}: " + handle.pslam_last_error().decode(errors="replace"))
    pdl = os.environ.get("PSLAM_PDL")        # programmatic dependent launch on / off (PSLAM_OPT_PDL)
    if pdl is not None and handle.pslam_set_option(3, int(pdl)) != 0:
        raise RuntimeError(f"PSLAM_PDL={pdl}: " + handle.pslam_last_error().decode(errors="replace"))
    tiles = os.environ.get("PSLAM_TILES")    # tiles in flight per CTA of the 3xF16 decoder (PSLAM_OPT_TILES)
    if tiles is not None and handle.pslam_set_option(4, int(tiles)) != 0:
        raise RuntimeError(f"PSLAM_TILES={tiles}: " + handle.pslam_last_error().decode(errors="replace"))
    fw = os.environ.get("PSLAM_FUSED_WGRAD")  # dgrad chain + weight gradients in one kernel (PSLAM_OPT_FUSED_WGRAD)
    if fw is not None and handle.pslam_set_option(5, int(fw)) != 0:
        raise RuntimeError(f"PSLAM_FUSED_WGRAD={fw}: " + handle.pslam_last_error().decode(errors="replace"))
    _lib = handle
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(rc, what):
    if rc != 0:
        msg = lib().pslam_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(t, name, dtype=None, contiguous=True):
    """Argument checks with the reference's error wording (sparse_voxels/include/utils.h:10-34)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if dtype is torch.float32 and t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")
    if dtype is torch.int32 and t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")
    return t
