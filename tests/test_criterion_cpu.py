"""The drop-in Criterion object (proud_slam_b200/criterion.py: loss as ratios of raw sums, the CUDA path's formulation)
against the oracle's restatement of src/criterion.py:16-116 on CPU tensors: values and gradients, mapping and tracking."""
import types

import pytest
import torch

from oracle import render_oracle as ro
from proud_slam_b200 import scene as sc
from proud_slam_b200.criterion import Criterion
from tests import util


@pytest.mark.parametrize("tracking", [False, True])
def test_criterion_matches_oracle(tracking):
    s, ms = util.build_scene("tiny")
    dec = util.test_decoder(width=128, seed=2)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0], 200, seed=3)
    depth = depth * (1.0 + 0.02 * torch.randn(depth.shape, generator=torch.Generator().manual_seed(1)))
    out = ro.render_rays(rays_o, rays_d, ms, dec, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0, generator=torch.Generator().manual_seed(5))
    kw = {k: util.CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
    leaves = {k: out[k].detach().clone().requires_grad_(True) for k in ("sdf", "color", "depth", "weights")}

    def outputs():
        o = {k: v for k, v in out.items() if k != "_dbg"}
        o.update(leaves)
        o["ray_mask"] = out["ray_mask"].view(-1) if tracking else out["ray_mask"]
        return o

    obs = (rgb[0], depth[0]) if tracking else (rgb, depth)
    ref_loss, ref_parts = ro.criterion(outputs(), obs, weight_depth_loss=tracking, **kw)
    ref_grads = torch.autograd.grad(ref_loss, [leaves["sdf"], leaves["color"], leaves["depth"]])
    args = types.SimpleNamespace(criteria=dict(rgb_weight=kw["rgb_weight"], depth_weight=kw["depth_weight"], sdf_weight=kw["sdf_weight"],
                                               fs_weight=kw["fs_weight"], sdf_truncation=kw["truncation"]),
                                 data_specs=dict(max_depth=kw["max_depth"]))
    crit = Criterion(args)
    loss, parts = crit(outputs(), obs, weight_depth_loss=tracking)
    grads = torch.autograd.grad(loss, [leaves["sdf"], leaves["color"], leaves["depth"]])
    assert abs(float(loss) - float(ref_loss)) <= 1e-6 * abs(float(ref_loss))
    for k in ("color_loss", "depth_loss", "fs_loss", "sdf_loss"):
        assert abs(parts[k] - float(ref_parts[k])) <= 1e-6 * max(abs(float(ref_parts[k])), 1e-12), k
    for a, b in zip(grads, ref_grads):
        assert util.rel_err(a, b) < 1e-5
    # the two helpers of the object
    fs, sdf = crit.get_sdf_loss(out["z_vals"], obs[1][outputs()["ray_mask"]], leaves["sdf"], crit.truncation)
    assert abs(float(fs) - float(ref_parts["fs_loss"])) <= 1e-6 * max(abs(float(ref_parts["fs_loss"])), 1e-12)
    assert abs(float(sdf) - float(ref_parts["sdf_loss"])) <= 1e-6 * max(abs(float(ref_parts["sdf_loss"])), 1e-12)
