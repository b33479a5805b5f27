"""Synthetic Replica/ScanNet-shaped RGB-D scenes.

There is no dataset access, so tests and ``bench.py`` use an axis-aligned box
room with analytic z-depth.  Conventions follow the reference so that the
tensors have the shapes and value ranges the render path sees in production:

* per-pixel camera rays ``((ix-cx)/fx, (iy-cy)/fy, 1)``, NOT normalised
  (reference ``src/frame.py:43-58``) -- "depth" along a ray is z-depth;
* the world is offset by +10 m on every axis (``src/frame.py:24``) so voxel
  coordinates stay inside the 256^3 grid of ``src/mapping.py:87``;
* voxels are allocated as ``floor(point / voxel_size)`` over the back-projected
  depth image (``src/mapping.py:258-264``).

Pure torch/numpy; nothing here touches CUDA or the oracle.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
import torch

REPLICA_CAM = dict(H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5)   # src/dataset/replica.py:20-26
SCANNET_CAM = dict(H=480, W=640, fx=577.87, fy=577.87, cx=319.5, cy=239.5)  # stand-in intrinsics
WORLD_OFFSET = 10.0


def camera_rays(H, W, fx, fy, cx, cy):
    """[H,W,3] float32 camera-frame ray directions with z == 1."""
    ix, iy = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    return torch.stack([(ix - cx) / fx, (iy - cy) / fy, torch.ones_like(ix)], -1).float()


def look_pose(position, yaw, pitch=0.0):
    """4x4 camera-to-world matrix: camera z forward, y down; world y is 'up'."""
    cy_, sy_ = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = position
    return torch.tensor(T, dtype=torch.float32)


def box_depth(rays_cam, pose, room_min, room_max):
    """Analytic z-depth of a camera inside an axis-aligned box, [H,W] float32."""
    R, t = pose[:3, :3].double(), pose[:3, 3].double()
    d = rays_cam.double() @ R.T
    lo = torch.as_tensor(room_min, dtype=torch.float64)
    hi = torch.as_tensor(room_max, dtype=torch.float64)
    wall = torch.where(d > 0, hi, lo)
    with np.errstate(divide="ignore"):
        tt = (wall - t) / d
    tt = torch.where(d.abs() < 1e-12, torch.full_like(tt, float("inf")), tt)
    return tt.min(dim=-1).values.float()


def unique_first(vox):
    """Unique rows keeping first-occurrence order (duplicate inserts are no-ops
    in the reference octree, so this yields the identical tree)."""
    vox = np.ascontiguousarray(vox, dtype=np.int64)
    key = (vox[:, 0] << 42) | (vox[:, 1] << 21) | vox[:, 2]
    _, first = np.unique(key, return_index=True)
    return vox[np.sort(first)].astype(np.int32)


@dataclass
class Frame:
    pose: torch.Tensor          # [4,4] camera-to-world
    depth: torch.Tensor         # [H,W] z-depth, metres
    rgb: torch.Tensor           # [H,W,3] in [0,1]


@dataclass
class Scene:
    cam: dict
    voxel_size: float
    room_min: Tuple[float, float, float]
    room_max: Tuple[float, float, float]
    rays_cam: torch.Tensor      # [H,W,3]
    frames: List[Frame] = field(default_factory=list)
    voxels: np.ndarray = None   # [M,3] int32 unique voxel coordinates, insertion order
    grid_dim: int = 256         # octree root side in voxels (src/mapping.py:87)


def make_scene(kind="replica_small", seed=0, pixel_stride=1, relief=0.25):
    """Named scenes of SURVEY 8(d):

    replica_small  7x3x5 m room, 0.2 m voxels, 8 yaw poses      (cfg A/C, ~9k octants)
    replica_20k    10.5x3x9 m room, 0.2 m, 12 poses             (cfg B, ~19k octants)
    scannet_large  20x4x16 m hall, 0.1 m voxels, 24 poses       (cfg D, ~400k octants)
    tiny           2.4x1.6x2 m, 0.2 m, 2 poses, 60x80 camera    (unit tests)
    """
    g = torch.Generator().manual_seed(seed)
    if kind == "tiny":
        cam = dict(H=60, W=80, fx=50.0, fy=50.0, cx=39.5, cy=29.5)
        size, vs, nposes = (2.4, 1.6, 2.0), 0.2, 2
    elif kind == "replica_small":
        cam, size, vs, nposes = dict(REPLICA_CAM), (7.0, 3.0, 5.0), 0.2, 8
    elif kind == "replica_20k":
        cam, size, vs, nposes = dict(REPLICA_CAM), (10.5, 3.0, 9.0), 0.2, 12
    elif kind == "scannet_large":
        cam, size, vs, nposes = dict(SCANNET_CAM), (20.0, 4.0, 16.0), 0.1, 24
    else:
        raise ValueError(kind)
    # a slightly off-grid room origin so walls do not sit exactly on voxel faces
    lo = np.array([WORLD_OFFSET + 0.03, WORLD_OFFSET + 0.05, WORLD_OFFSET + 0.07])
    hi = lo + np.array(size)
    rays_cam = camera_rays(**cam)
    scene = Scene(cam, vs, tuple(lo), tuple(hi), rays_cam)
    scene.grid_dim = 512 if kind == "scannet_large" else 256   # 30 m / 0.1 m exceeds 256
    vox_all = []
    for i in range(nposes):
        if nposes <= 12:
            pos = (lo + hi) / 2 + (torch.rand(3, generator=g).numpy() - 0.5) * 0.2
        else:  # lattice over the hall
            gx, gz = i % 6, i // 6
            pos = np.array([lo[0] + (gx + 0.5) * size[0] / 6, (lo[1] + hi[1]) / 2,
                            lo[2] + (gz + 0.5) * size[2] / 4])
        yaw = 2 * np.pi * i / min(nposes, 8) + float(torch.rand(1, generator=g)) * 0.1
        pitch = (float(torch.rand(1, generator=g)) - 0.5) * 0.2
        pose = look_pose(pos, yaw, pitch)
        depth = box_depth(rays_cam, pose, lo, hi)
        if relief > 0:   # wall relief: makes the allocated shell ~2 voxels thick like real scans
            pw = (rays_cam * depth[..., None]) @ pose[:3, :3].T + pose[:3, 3]
            bump = 0.5 + 0.5 * torch.sin(3.1 * pw[..., 0] + 1.7 * pw[..., 1]) * torch.sin(2.3 * pw[..., 2] + 0.5 * pw[..., 1])
            depth = depth * (1.0 - relief * bump / depth.clamp(min=0.5))
        # smooth synthetic colour: function of the hit position
        pts_w = (rays_cam * depth[..., None]) @ pose[:3, :3].T + pose[:3, 3]
        rgb = 0.5 + 0.5 * torch.sin(pts_w * torch.tensor([1.3, 2.1, 1.7]) + torch.tensor([0.0, 1.0, 2.0]))
        scene.frames.append(Frame(pose, depth, rgb.float()))
        sub = pts_w[::pixel_stride, ::pixel_stride].reshape(-1, 3)
        vox = torch.div(sub, vs, rounding_mode="floor").int().numpy()   # mapping.py:264
        vox_all.append(unique_first(vox))
    scene.voxels = unique_first(np.concatenate(vox_all, 0))
    return scene


def map_states_from_flat(voxels, children, features, voxel_size, num_embeddings=None, seed=0,
                         device="cpu"):
    """Build the ``map_states`` dict exactly as ``Mapping.update_grid_pcd_features``
    (reference src/mapping.py:301-377) does from ``get_centres_and_children()``."""
    voxels = torch.as_tensor(voxels, dtype=torch.float32)
    children = torch.as_tensor(children, dtype=torch.float32)
    features = torch.as_tensor(features, dtype=torch.int32)
    centres = (voxels[:, :3] + voxels[:, -1:] / 2) * voxel_size
    structure = torch.cat([children, voxels[:, -1:]], -1).int()
    n = voxels.shape[0]
    if num_embeddings is None:
        num_embeddings = max(20000, n)
    assert num_embeddings >= n, "F.embedding would fault (SURVEY A-Q14)"
    g = torch.Generator().manual_seed(seed)
    emb = torch.zeros(num_embeddings, 16).normal_(0.0, 0.01, generator=g)   # mapping.py:71-80
    return {
        "voxel_vertex_idx": features.to(device),
        "voxel_center_xyz": centres.float().to(device),
        "voxel_structure": structure.to(device),
        "voxel_vertex_emb": emb.to(device).requires_grad_(True),
    }


def sample_batch(scene, frame_ids, rays_per_frame, seed=0):
    """Ray batch assembled like ``bundle_adjust_frames`` (render_helpers.py:620-646):
    for each frame, a sorted random pixel subset (mask-indexing order), rays
    rotated into the world, origin = camera position.  Returns CPU tensors
    rays_o [1,R,3], rays_d [1,R,3], rgb [1,R,3], depth [1,R]."""
    g = torch.Generator().manual_seed(seed)
    H, W = scene.cam["H"], scene.cam["W"]
    ro, rd, cs, ds = [], [], [], []
    for fid in frame_ids:
        f = scene.frames[fid % len(scene.frames)]
        if rays_per_frame <= H * W:
            pix = torch.randperm(H * W, generator=g)[:rays_per_frame].sort().values
        else:  # sweep E: with replacement
            pix = torch.randint(0, H * W, (rays_per_frame,), generator=g).sort().values
        d_cam = scene.rays_cam.reshape(-1, 3)[pix]
        d = d_cam @ f.pose[:3, :3].T
        ro.append(f.pose[:3, 3].reshape(1, 3).expand_as(d))
        rd.append(d)
        cs.append(f.rgb.reshape(-1, 3)[pix])
        ds.append(f.depth.reshape(-1)[pix])
    cat = lambda xs: torch.cat(xs, 0).unsqueeze(0).contiguous()
    return cat(ro), cat(rd), cat(cs), cat(ds)
