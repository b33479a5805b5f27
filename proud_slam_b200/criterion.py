"""Drop-in for the reference's ``Criterion`` object (``src/criterion.py:8-116``) on the dict the drop-in ``render_rays`` returns.

This is the *modular* route (``render_rays`` -> ``Criterion`` -> ``loss.backward()``); ``bundle_adjust_frames`` /
``track_frame`` use the fused CUDA loss of ``pslam_render_step`` instead.  It is written the way the CUDA path computes the
loss (``csrc/composite.cu``: raw sums first, then one closing step -- the same split that makes the loss shardable over
GPUs, ``pslam_loss_finalize``), not as a transcription of the reference module: every term is a ratio of two sums over the
padded ``[R_h, S]`` grid,

    colour   = sum |c_gt - c| / (3 R_h)                                   (criterion.py:37-40)
    depth    = sum_valid |d_gt - d| / #valid                              (:42-52; tracking: valid also needs
                                                                             |d_gt - d| / sqrt(var) < 10 median, :45-49)
    fs       = (1 - n_fs / (n_fs + n_sdf)) * sum_front (s - 1)^2 / (R_h S)          (:78-100)
    sdf      = (1 - n_sdf / (n_fs + n_sdf)) * sum_band (z + tau s - d_gt)^2 / (R_h S) (:102-116)

with front = z < d_gt - tau, band = not front, not (z > d_gt + tau), 0 < d_gt < max_depth.  Names of the constructor
argument, the attributes (``max_dpeth`` sic, criterion.py:14) and the returned ``(loss, dict of floats)`` are the reference's.
"""
import torch
import torch.nn as nn


def raw_loss_sums(outputs, obs, truncation, max_depth, weight_depth_loss=False):
    """The sums every loss term is made of (tensors, differentiable where the reference's terms are)."""
    img, depth = obs
    hit = outputs["ray_mask"]
    d_gt, c_gt = depth[hit], img[hit]
    z, s, d, c = outputs["z_vals"], outputs["sdf"], outputs["depth"], outputs["color"]
    err_d = (d_gt - d).abs()
    ok = (d_gt > 0.01) & (d_gt < max_depth)
    if weight_depth_loss:
        # median gate: a boolean mask only, no gradient flows through the variance (criterion.py:45-49)
        var = (outputs["weights"] * (d[:, None] - z) ** 2).sum(-1)
        ratio = err_d / (var + 1e-10).sqrt()
        ok = ok & (ratio < 10 * ratio.median())
    dg = d_gt[:, None]
    front = z < dg - truncation
    band = ~front & ~(z > dg + truncation) & ((dg > 0.0) & (dg < max_depth))
    return {
        "abs_color": (c_gt - c).abs().sum(), "n_color": float(c.numel()),
        "abs_depth": err_d[ok].sum(), "n_depth": ok.sum(),
        "sq_fs": ((s - 1.0) ** 2)[front].sum(), "n_fs": front.sum().float(),
        "sq_sdf": ((z + truncation * s - dg) ** 2)[band].sum(), "n_sdf": band.sum().float(),
        "n_elems": float(z.numel()),
    }


class Criterion(nn.Module):
    def __init__(self, args) -> None:
        super().__init__()
        self.args = args
        c = args.criteria
        self.rgb_weight, self.depth_weight = c["rgb_weight"], c["depth_weight"]
        self.sdf_weight, self.fs_weight = c["sdf_weight"], c["fs_weight"]
        self.truncation = c["sdf_truncation"]
        self.max_dpeth = args.data_specs["max_depth"]

    def weights(self):
        """(rgb, depth, fs, sdf) in the order of the C ABI."""
        return (self.rgb_weight, self.depth_weight, self.fs_weight, self.sdf_weight)

    def forward(self, outputs, obs, use_color_loss=True, use_depth_loss=True, compute_sdf_loss=True, weight_depth_loss=False):
        r = raw_loss_sums(outputs, obs, self.truncation, self.max_dpeth, weight_depth_loss)
        terms = {}
        if use_color_loss:
            terms["color_loss"] = (self.rgb_weight, r["abs_color"] / r["n_color"])
        if use_depth_loss:
            terms["depth_loss"] = (self.depth_weight, r["abs_depth"] / r["n_depth"])
        if compute_sdf_loss:
            n_both = r["n_fs"] + r["n_sdf"]
            terms["fs_loss"] = (self.fs_weight, (1.0 - r["n_fs"] / n_both) * r["sq_fs"] / r["n_elems"])
            terms["sdf_loss"] = (self.sdf_weight, (1.0 - r["n_sdf"] / n_both) * r["sq_sdf"] / r["n_elems"])
        loss = sum(w * t for w, t in terms.values())
        report = {k: float(t.detach()) for k, (_, t) in terms.items()}
        report["loss"] = float(loss.detach())
        return loss, report

    # the two helpers the reference exposes on the object (criterion.py:70-116), in terms of the same sums
    def get_masks(self, z_vals, depth, epsilon):
        front = z_vals < depth - epsilon
        band = ~front & ~(z_vals > depth + epsilon) & ((depth > 0.0) & (depth < self.max_dpeth))
        n_fs, n_sdf = front.sum().float(), band.sum().float()
        return front.to(z_vals.dtype), band.to(z_vals.dtype), 1.0 - n_fs / (n_fs + n_sdf), 1.0 - n_sdf / (n_fs + n_sdf)

    def get_sdf_loss(self, z_vals, depth, predicted_sdf, truncation, loss_type="l2"):
        if loss_type != "l2":
            raise NotImplementedError("only the l2 form is on the SLAM path (criterion.py:78)")
        dg = depth[:, None].expand_as(z_vals)
        front, band, w_fs, w_sdf = self.get_masks(z_vals, dg, truncation)
        n = float(z_vals.numel())
        return (w_fs * (front * (predicted_sdf - 1.0) ** 2).sum() / n,
                w_sdf * (band * (z_vals + truncation * predicted_sdf - dg) ** 2).sum() / n)
