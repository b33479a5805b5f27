"""Drop-in for the reference's pybind module ``grid``
(``third_party/sparse_voxels/src/binding.cpp:10-21``; imported by the reference as
``import grid as _ext``, ``src/variations/voxel_helpers.py:22``).

Same function names, argument order, dtypes, shapes and output allocation
semantics (outputs are created here, like ``intersect.cpp:98-106`` /
``sample.cpp:80-89`` create them in C++).  Argument errors raise
``RuntimeError`` with the reference's wording; launch errors raise
``RuntimeError`` too (the reference prints and ``exit(-1)``s,
``include/cuda_utils.h:37-48``).  Launches go to torch's current stream.
"""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

F32, I32 = torch.float32, torch.int32


def _rays_points(ray_start, ray_dir, points):
    require_cuda(ray_start, "ray_start", F32)
    require_cuda(ray_dir, "ray_dir", F32)
    require_cuda(points, "points", F32)
    return ray_start.size(0), points.size(1), ray_start.size(1)


def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max):
    """intersect.h:14-15 / intersect.cpp:83-112.  ray_* f32[B,K,3], points f32[B,N,3], children
    i32[B,N,9] -> (idx i32[B,K,n_max] (-1 padded), min_depth, max_depth f32[B,K,n_max])."""
    b, n, m = _rays_points(ray_start, ray_dir, points)
    require_cuda(children, "children", I32)
    idx = torch.empty(b, m, n_max, dtype=I32, device=ray_start.device)
    min_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    max_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    check(_lib.lib().pslam_svo_intersect(b, n, m, float(voxelsize), int(n_max), ptr(ray_start), ptr(ray_dir),
                                         ptr(points), ptr(children), ptr(idx), ptr(min_depth), ptr(max_depth),
                                         stream_ptr(ray_start.device)), "svo_intersect")
    return idx, min_depth, max_depth


def aabb_intersect(ray_start, ray_dir, points, voxelsize, n_max):
    """intersect.h:12-13 / intersect.cpp:49-76."""
    b, n, m = _rays_points(ray_start, ray_dir, points)
    idx = torch.empty(b, m, n_max, dtype=I32, device=ray_start.device)
    min_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    max_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    check(_lib.lib().pslam_aabb_intersect(b, n, m, float(voxelsize), int(n_max), ptr(ray_start), ptr(ray_dir),
                                          ptr(points), ptr(idx), ptr(min_depth), ptr(max_depth),
                                          stream_ptr(ray_start.device)), "aabb_intersect")
    return idx, min_depth, max_depth


def ball_intersect(ray_start, ray_dir, points, radius, n_max):
    """intersect.h:10-11 / intersect.cpp:15-42."""
    b, n, m = _rays_points(ray_start, ray_dir, points)
    idx = torch.empty(b, m, n_max, dtype=I32, device=ray_start.device)
    min_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    max_depth = torch.empty(b, m, n_max, dtype=F32, device=ray_start.device)
    check(_lib.lib().pslam_ball_intersect(b, n, m, float(radius), int(n_max), ptr(ray_start), ptr(ray_dir),
                                          ptr(points), ptr(idx), ptr(min_depth), ptr(max_depth),
                                          stream_ptr(ray_start.device)), "ball_intersect")
    return idx, min_depth, max_depth


def triangle_intersect(ray_start, ray_dir, face_points, cagesize, blur, n_max):
    """intersect.h:16-17 / intersect.cpp:119-146: face_points f32[B,N,9] -> (idx i32[B,K,n_max],
    depth f32[B,K,n_max,3], uv f32[B,K,n_max,2])."""
    require_cuda(ray_start, "ray_start", F32)
    require_cuda(ray_dir, "ray_dir", F32)
    require_cuda(face_points, "face_points", F32)
    b, n, m = ray_start.size(0), face_points.size(1), ray_start.size(1)
    idx = torch.empty(b, m, n_max, dtype=I32, device=ray_start.device)
    depth = torch.empty(b, m, n_max * 3, dtype=F32, device=ray_start.device)
    uv = torch.empty(b, m, n_max * 2, dtype=F32, device=ray_start.device)
    check(_lib.lib().pslam_triangle_intersect(b, n, m, float(cagesize), float(blur), int(n_max), ptr(ray_start),
                                              ptr(ray_dir), ptr(face_points), ptr(idx), ptr(depth), ptr(uv),
                                              stream_ptr(ray_start.device)), "triangle_intersect")
    return idx, depth, uv


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, uniform_noise, probs, steps, fixed_step_size):
    """sample.h:13-15 / sample.cpp:56-95.  pts_idx i32[G,n,P], min/max_depth, probs f32[G,n,P],
    uniform_noise f32[G,n,M], steps f32[G,n] -> (sampled_idx i32[G,n,M] (-1), sampled_depth,
    sampled_dists f32[G,n,M] (0))."""
    require_cuda(pts_idx, "pts_idx", I32)
    require_cuda(min_depth, "min_depth", F32)
    require_cuda(max_depth, "max_depth", F32)
    require_cuda(uniform_noise, "uniform_noise", F32)
    require_cuda(probs, "probs", F32)
    require_cuda(steps, "steps", F32)
    g, n, P = pts_idx.size(0), pts_idx.size(1), pts_idx.size(2)
    M = uniform_noise.size(2)
    dev = pts_idx.device
    sidx = torch.empty(g, n, M, dtype=I32, device=dev)
    sdepth = torch.empty(g, n, M, dtype=F32, device=dev)
    sdist = torch.empty(g, n, M, dtype=F32, device=dev)
    check(_lib.lib().pslam_inverse_cdf_sampling(g, n, P, M, float(fixed_step_size), ptr(pts_idx), ptr(min_depth),
                                                ptr(max_depth), ptr(uniform_noise), ptr(probs), ptr(steps),
                                                ptr(sidx), ptr(sdepth), ptr(sdist), stream_ptr(dev)),
          "inverse_cdf_sampling")
    return sidx, sdepth, sdist


def uniform_ray_sampling(pts_idx, min_depth, max_depth, uniform_noise, step_size, max_steps):
    """sample.h:10-12 / sample.cpp:21-54: pts_idx i32[G,n,P], min/max_depth f32[G,n,P],
    uniform_noise f32[G,n,max_steps] -> outputs [G,n,max_steps]."""
    require_cuda(pts_idx, "pts_idx", I32)
    require_cuda(min_depth, "min_depth", F32)
    require_cuda(max_depth, "max_depth", F32)
    require_cuda(uniform_noise, "uniform_noise", F32)
    g, n, P = pts_idx.size(0), pts_idx.size(1), pts_idx.size(2)
    dev = pts_idx.device
    M = int(max_steps)
    sidx = torch.empty(g, n, M, dtype=I32, device=dev)
    sdepth = torch.empty(g, n, M, dtype=F32, device=dev)
    sdist = torch.empty(g, n, M, dtype=F32, device=dev)
    check(_lib.lib().pslam_uniform_ray_sampling(g, n, P, M, float(step_size), ptr(pts_idx), ptr(min_depth),
                                                ptr(max_depth), ptr(uniform_noise), ptr(sidx), ptr(sdepth),
                                                ptr(sdist), stream_ptr(dev)), "uniform_ray_sampling")
    return sidx, sdepth, sdist


def build_octree(center, points, depth):
    """octree.h:10 / sparse_voxels/src/octree.cpp:12-160 (EasyOctree, CPU tensors).  Not on the
    SLAM path (`build_easy_octree`, voxel_helpers.py:494, has no caller)."""
    from .easy_octree import build_octree as _impl
    return _impl(center, points, depth)
