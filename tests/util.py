"""Shared helpers for the test-suite (fixtures, scene construction, error metrics)."""
import os

import numpy as np
import torch

import oracle
from oracle import render_oracle as ro
from proud_slam_b200 import scene as sc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CRIT = dict(rgb_weight=0.5, depth_weight=1.0, sdf_weight=5000.0, fs_weight=10.0, truncation=0.1, max_depth=10.0)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_map_states(g, device="cpu", requires_grad=True):
    ms = {
        "voxel_vertex_idx": torch.from_numpy(g["vertex_idx"]).to(device),
        "voxel_center_xyz": torch.from_numpy(g["centres"]).to(device),
        "voxel_structure": torch.from_numpy(g["structure"]).to(device),
        "voxel_vertex_emb": torch.from_numpy(g["emb"]).to(device).requires_grad_(requires_grad),
    }
    return ms


def golden_decoder(g, device="cpu", requires_grad=True):
    return [torch.from_numpy(g[f"dec_{i}"]).to(device).requires_grad_(requires_grad) for i in range(10)]


def build_scene(kind="tiny", num_embeddings=None, seed=0, emb_scale=None):
    """(scene, map_states on CPU) through the ORACLE octree."""
    s = sc.make_scene(kind, seed=seed)
    oc = oracle.Octree(s.grid_dim)
    oc.insert(s.voxels)
    v, c, f = oc.get_centres_and_children()
    ms = sc.map_states_from_flat(v, c, f, s.voxel_size, num_embeddings=num_embeddings, seed=seed)
    if emb_scale is not None:
        with torch.no_grad():
            ms["voxel_vertex_emb"].mul_(emb_scale / 0.01)
    return s, ms


def to_device(ms, device):
    out = {}
    for k, v in ms.items():
        t = v.detach().to(device)
        out[k] = t.requires_grad_(v.requires_grad) if v.is_floating_point() else t
    return out


def rel_err(a, b):
    """max |a-b| / max |b| (norm-wise relative error; the 1e-4 bound of BASELINE.json is read this way)."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)


def oracle_step(rays_o, rays_d, rgb, depth, ms, dec, *, voxel_size, noise=None, tracking=False, inv_dir=None,
                generator=None):
    """Oracle forward + loss + backward on CPU tensors.  Returns (outputs, loss, parts)."""
    for t in [ms["voxel_vertex_emb"], rays_o, rays_d] + list(dec):
        if t.grad is not None:
            t.grad = None
    out = ro.render_rays(rays_o, rays_d, ms, dec, 0.1 * voxel_size, voxel_size, CRIT["truncation"], 10, 10.0,
                         noise=noise, generator=generator, inv_dir=inv_dir)
    kw = {k: CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
    if tracking:
        out2 = dict(out)
        out2["ray_mask"] = out["ray_mask"].view(-1)
        loss, parts = ro.criterion(out2, (rgb[0], depth[0]), weight_depth_loss=True, **kw)
    else:
        loss, parts = ro.criterion(out, (rgb, depth), **kw)
    loss.backward()
    return out, loss, parts


def device_rcp(x, device):
    """__fdividef(1, x) evaluated on the device (the reciprocal the slab test uses)."""
    from proud_slam_b200 import _lib
    xin = torch.as_tensor(x, dtype=torch.float32, device=device).contiguous()
    out = torch.empty_like(xin)
    _lib.check(_lib.lib().pslam_debug_rcp(_lib.ptr(xin), _lib.ptr(out), xin.numel(), _lib.stream_ptr(device)), "rcp")
    return out.cpu().numpy()
