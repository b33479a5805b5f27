"""Drop-in for the render-path part of ``src/variations/voxel_helpers.py``.

``svo_ray_intersect`` / ``inverse_cdf_sampling`` wrap the CUDA kernels of ``proud_slam_b200.grid``
with the reference's pre/post-processing (``voxel_helpers.py:110-159, 288-367``), minus its batching
hacks: the octree is not replicated 256x (``:132-135``) and rays are not padded to a multiple of
G -- both are result-neutral.  The sampling kernel still receives the reference's ``[200, n, P]``
grouping because the tail-loop quirk of ``sample_gpu.cu:224-237`` depends on it (SURVEY A-Q7).
``ray_intersect_vox`` and ``ray_sample`` are ``voxel_helpers.py:557-595, 637-663``.

These are the modular entry points (same tensors in and out as the reference, host syncs at the
same places).  The SLAM loops use the fused device-side path in ``proud_slam_b200.pipeline``.
"""
import math

import torch

from .. import grid as _ext

MAX_DEPTH = 10.0   # voxel_helpers.py:24


@torch.no_grad()
def svo_ray_intersect(voxelsize, n_max, points, children, ray_start, ray_dir):
    """SparseVoxelOctreeRayIntersect.forward, voxel_helpers.py:110-159.  points [N,3], children [N,9],
    ray_start / ray_dir [S,R,3] -> inds, min_depth, max_depth [S,R,n_max] (DFS order, -1 / 0 padded)."""
    S, R = ray_start.shape[:2]
    pts = points.float().reshape(1, -1, 3).expand(S, -1, 3).contiguous()
    ch = children.int().reshape(1, -1, 9).expand(S, -1, 9).contiguous()
    inds, min_depth, max_depth = _ext.svo_intersect(ray_start.float().contiguous(), ray_dir.float().contiguous(),
                                                    pts, ch, voxelsize, n_max)
    return inds, min_depth.type_as(ray_start), max_depth.type_as(ray_start)


@torch.no_grad()
def inverse_cdf_sampling(pts_idx, min_depth, max_depth, probs, steps, fixed_step_size=-1, deterministic=False,
                         noise=None):
    """InverseCDFRaySampling.forward, voxel_helpers.py:288-367.  Inputs [N,P] / [N]; returns
    (sampled_idx, sampled_depth, sampled_dists) [N, S] trimmed to the longest row.  ``noise``
    ([200, n, max_steps]) replays a recorded draw; otherwise it is drawn from torch's CUDA generator
    with the reference's call and shape (same seed -> same samples as the reference)."""
    G, N, P = 200, pts_idx.size(0), pts_idx.size(1)
    H = int(math.ceil(N / G)) * G
    if H > N:   # pad with copies of ray 0, :302-311
        pts_idx = torch.cat([pts_idx, pts_idx[:1].expand(H - N, P)], 0)
        min_depth = torch.cat([min_depth, min_depth[:1].expand(H - N, P)], 0)
        max_depth = torch.cat([max_depth, max_depth[:1].expand(H - N, P)], 0)
        probs = torch.cat([probs, probs[:1].expand(H - N, P)], 0)
        steps = torch.cat([steps, steps[:1].expand(H - N)], 0)
    pts_idx = pts_idx.reshape(G, -1, P)
    min_depth = min_depth.reshape(G, -1, P)
    max_depth = max_depth.reshape(G, -1, P)
    probs = probs.reshape(G, -1, P)
    steps = steps.reshape(G, -1)
    max_steps = int(steps.ceil().long().max()) + P
    if noise is None:
        noise = min_depth.new_zeros(*min_depth.size()[:-1], max_steps)
        if deterministic:
            noise += 0.5
        else:
            noise = noise.uniform_().clamp(min=0.001, max=0.999)
    chunk = 4 * G   # the grouping is part of the kernel's observable behaviour (A-Q7), keep it
    results = [
        _ext.inverse_cdf_sampling(
            pts_idx[:, i: i + chunk].int().contiguous(), min_depth.float()[:, i: i + chunk].contiguous(),
            max_depth.float()[:, i: i + chunk].contiguous(), noise.float()[:, i: i + chunk].contiguous(),
            probs.float()[:, i: i + chunk].contiguous(), steps.float()[:, i: i + chunk].contiguous(), fixed_step_size)
        for i in range(0, min_depth.size(1), chunk)
    ]
    sampled_idx, sampled_depth, sampled_dists = [torch.cat([r[i] for r in results], 1) for i in range(3)]
    sampled_idx = sampled_idx.reshape(H, -1)[:N]
    sampled_depth = sampled_depth.type_as(min_depth).reshape(H, -1)[:N]
    sampled_dists = sampled_dists.type_as(min_depth).reshape(H, -1)[:N]
    max_len = int(sampled_idx.ne(-1).sum(-1).max())
    return sampled_idx[:, :max_len], sampled_depth[:, :max_len], sampled_dists[:, :max_len]


@torch.no_grad()
def ray_intersect_vox(ray_start, ray_dir, flatten_centers, flatten_children, voxel_size, max_hits, max_distance=10.0):
    """voxel_helpers.py:557-595.  ``max_hits`` is ignored exactly as in the reference (the kernel cap
    is 50 and the result is trimmed to the longest hit list, SURVEY A-Q1).  Equal entry depths keep
    DFS order (the reference's ``torch.sort`` leaves ties unspecified, A-Q3)."""
    pts_idx, min_depth, max_depth = svo_ray_intersect(voxel_size, 50, flatten_centers, flatten_children, ray_start, ray_dir)
    miss = pts_idx.eq(-1)
    min_depth.masked_fill_(miss, max_distance)
    max_depth.masked_fill_(miss, max_distance)
    min_depth, sorted_idx = min_depth.sort(dim=-1, stable=True)
    max_depth = max_depth.gather(-1, sorted_idx)
    pts_idx = pts_idx.gather(-1, sorted_idx)
    pts_idx[min_depth > max_distance] = -1
    miss = pts_idx.eq(-1)
    min_depth.masked_fill_(miss, max_distance)
    max_depth.masked_fill_(miss, max_distance)
    width = int(torch.max(pts_idx.ne(-1).sum(-1)))
    out = {"min_depth": min_depth[..., :width], "max_depth": max_depth[..., :width], "intersected_voxel_idx": pts_idx[..., :width]}
    return out, out["intersected_voxel_idx"].ne(-1).any(-1)


@torch.no_grad()
def ray_intersect_vox_AABB(ray_start, ray_dir, flatten_centers, voxel_size, max_hits, max_distance=10.0):
    """voxel_helpers.py:598-635: brute force over the given voxel centres (the reference's cross-check)."""
    S = ray_start.shape[0]
    pts = flatten_centers.float().reshape(1, -1, 3).expand(S, -1, 3).contiguous()
    pts_idx, min_depth, max_depth = _ext.aabb_intersect(ray_start.float().contiguous(), ray_dir.float().contiguous(), pts, voxel_size, 50)
    miss = pts_idx.eq(-1)
    min_depth.masked_fill_(miss, max_distance)
    max_depth.masked_fill_(miss, max_distance)
    min_depth, sorted_idx = min_depth.sort(dim=-1, stable=True)
    max_depth = max_depth.gather(-1, sorted_idx)
    pts_idx = pts_idx.gather(-1, sorted_idx)
    pts_idx[min_depth > max_distance] = -1
    width = int(torch.max(pts_idx.ne(-1).sum(-1)))
    out = {"min_depth": min_depth[..., :width], "max_depth": max_depth[..., :width], "intersected_voxel_idx": pts_idx[..., :width]}
    return out, out["intersected_voxel_idx"].ne(-1).any(-1)


@torch.no_grad()
def ray_sample(intersection_outputs, step_size=0.01, fixed=False, noise=None):
    """voxel_helpers.py:637-663 (adds ``probs`` / ``steps`` to the input dict like the reference)."""
    idx = intersection_outputs["intersected_voxel_idx"]
    dists = (intersection_outputs["max_depth"] - intersection_outputs["min_depth"]).masked_fill(idx.eq(-1), 0)
    intersection_outputs["probs"] = dists / dists.sum(dim=-1, keepdim=True)
    intersection_outputs["steps"] = dists.sum(-1) / step_size
    sampled_idx, sampled_depth, sampled_dists = inverse_cdf_sampling(
        idx, intersection_outputs["min_depth"], intersection_outputs["max_depth"], intersection_outputs["probs"],
        intersection_outputs["steps"], -1, fixed, noise=noise)
    sampled_dists = sampled_dists.clamp(min=0.0)
    sampled_depth.masked_fill_(sampled_idx.eq(-1), MAX_DEPTH)
    sampled_dists.masked_fill_(sampled_idx.eq(-1), 0.0)
    return {"sampled_point_depth": sampled_depth, "sampled_point_distance": sampled_dists, "sampled_point_voxel_idx": sampled_idx}
