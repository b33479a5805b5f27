"""Drop-in for ``src/criterion.py``: colour L1 + depth L1 (optionally median-gated) + free-space and
SDF L2 terms, on the dict ``render_rays`` returns.  Plain torch ops in the reference's order (this is
the modular route; ``bundle_adjust_frames`` / ``track_frame`` use the fused CUDA loss instead)."""
import torch
import torch.nn as nn


class Criterion(nn.Module):
    def __init__(self, args) -> None:
        super().__init__()
        self.args = args
        self.rgb_weight = args.criteria["rgb_weight"]
        self.depth_weight = args.criteria["depth_weight"]
        self.sdf_weight = args.criteria["sdf_weight"]
        self.fs_weight = args.criteria["fs_weight"]
        self.truncation = args.criteria["sdf_truncation"]
        self.max_dpeth = args.data_specs["max_depth"]   # (sic) attribute name of the reference, criterion.py:14

    def weights(self):
        """(rgb, depth, fs, sdf) in the order of the C ABI."""
        return (self.rgb_weight, self.depth_weight, self.fs_weight, self.sdf_weight)

    def forward(self, outputs, obs, use_color_loss=True, use_depth_loss=True, compute_sdf_loss=True,
                weight_depth_loss=False):
        img, depth = obs
        loss, loss_dict = 0, {}
        pred_depth, pred_color, pred_sdf = outputs["depth"], outputs["color"], outputs["sdf"]
        z_vals, ray_mask, weights = outputs["z_vals"], outputs["ray_mask"], outputs["weights"]
        gt_depth, gt_color = depth[ray_mask], img[ray_mask]
        if use_color_loss:
            color_loss = (gt_color - pred_color).abs().mean()
            loss += self.rgb_weight * color_loss
            loss_dict["color_loss"] = color_loss.item()
        if use_depth_loss:
            valid_depth = (gt_depth > 0.01) & (gt_depth < self.max_dpeth)
            depth_loss = (gt_depth - pred_depth).abs()
            if weight_depth_loss:
                depth_var = torch.sum(weights * ((pred_depth.unsqueeze(-1) - z_vals) ** 2), -1)
                tmp = depth_loss / torch.sqrt(depth_var + 1e-10)
                valid_depth = (tmp < 10 * tmp.median()) & valid_depth
            depth_loss = depth_loss[valid_depth].mean()
            loss += self.depth_weight * depth_loss
            loss_dict["depth_loss"] = depth_loss.item()
        if compute_sdf_loss:
            fs_loss, sdf_loss = self.get_sdf_loss(z_vals, gt_depth, pred_sdf, truncation=self.truncation)
            loss += self.fs_weight * fs_loss
            loss += self.sdf_weight * sdf_loss
            loss_dict["fs_loss"] = fs_loss.item()
            loss_dict["sdf_loss"] = sdf_loss.item()
        loss_dict["loss"] = loss.item()
        return loss, loss_dict

    def get_masks(self, z_vals, depth, epsilon):
        front_mask = torch.where(z_vals < (depth - epsilon), torch.ones_like(z_vals), torch.zeros_like(z_vals))
        back_mask = torch.where(z_vals > (depth + epsilon), torch.ones_like(z_vals), torch.zeros_like(z_vals))
        depth_mask = torch.where((depth > 0.0) & (depth < self.max_dpeth), torch.ones_like(depth), torch.zeros_like(depth))
        sdf_mask = (1.0 - front_mask) * (1.0 - back_mask) * depth_mask
        num_fs_samples = torch.count_nonzero(front_mask).float()
        num_sdf_samples = torch.count_nonzero(sdf_mask).float()
        num_samples = num_sdf_samples + num_fs_samples
        return front_mask, sdf_mask, 1.0 - num_fs_samples / num_samples, 1.0 - num_sdf_samples / num_samples

    def get_sdf_loss(self, z_vals, depth, predicted_sdf, truncation, loss_type="l2"):
        d = depth.unsqueeze(-1).expand(*z_vals.shape)
        front_mask, sdf_mask, fs_weight, sdf_weight = self.get_masks(z_vals, d, truncation)
        fs_loss = torch.mean(torch.square(predicted_sdf * front_mask - torch.ones_like(predicted_sdf) * front_mask)) * fs_weight
        sdf_loss = torch.mean(torch.square((z_vals + predicted_sdf * truncation) * sdf_mask - d * sdf_mask)) * sdf_weight
        return fs_loss, sdf_loss
