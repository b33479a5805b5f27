"""oracle/render_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch) restatement of the Python half of the reference render path, using
``oracle.svo_intersect`` / ``oracle.inverse_cdf_sampling`` (C) for the two
native kernels.  Each function cites the reference lines it follows.  The
reference's own Python cannot travel to the GPU box, so this module is what
the ``-m gpu`` parity tests compare against; it is itself pinned against the
real reference (imported from /root/reference in the build container) by
``tests/test_oracle_vs_reference.py`` and the fixtures under ``tests/golden/``.

All tensors are CPU float32 unless stated; shapes follow the reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import inverse_cdf_sampling as _c_sample
from . import svo_intersect as _c_intersect

MAX_DEPTH = 10.0          # voxel_helpers.py:24
N_MAX_HITS = 50           # voxel_helpers.py:561 (max_voxel_hit is ignored, SURVEY A-Q1)
G_SAMPLE = 200            # voxel_helpers.py:300


# ---------------------------------------------------------------- kernel 1 + a4
def ray_intersect_vox(ray_start, ray_dir, centres, children, voxel_size, max_hits,
                      max_distance=10.0, inv_dir=None):
    """voxel_helpers.py:110-159 (batching shim, result-neutral so not replayed)
    + :558-595 (sort by entry depth, drop beyond max_distance, trim)."""
    R = ray_start.shape[1]
    idx, tmin, tmax = _c_intersect(
        ray_start.detach().reshape(1, R, 3).numpy(), ray_dir.detach().reshape(1, R, 3).numpy(),
        centres.detach().reshape(1, -1, 3).numpy(), children.reshape(1, -1, 9).numpy(),
        float(voxel_size), N_MAX_HITS,
        None if inv_dir is None else np.asarray(inv_dir).reshape(1, R, 3))
    pts_idx = torch.from_numpy(idx)
    min_depth = torch.from_numpy(tmin)
    max_depth = torch.from_numpy(tmax)
    miss = pts_idx.eq(-1)
    min_depth.masked_fill_(miss, max_distance)
    max_depth.masked_fill_(miss, max_distance)
    # reference: unstable torch.sort; ours is defined as (depth, DFS order), A-Q3
    min_depth, order = min_depth.sort(dim=-1, stable=True)
    max_depth = max_depth.gather(-1, order)
    pts_idx = pts_idx.gather(-1, order)
    pts_idx[min_depth > max_distance] = -1
    miss = pts_idx.eq(-1)
    min_depth.masked_fill_(miss, max_distance)
    max_depth.masked_fill_(miss, max_distance)
    width = int(pts_idx.ne(-1).sum(-1).max())
    out = {
        "min_depth": min_depth[..., :width],
        "max_depth": max_depth[..., :width],
        "intersected_voxel_idx": pts_idx[..., :width],
    }
    return out, out["intersected_voxel_idx"].ne(-1).any(-1)


# ---------------------------------------------------------------- a5 + a6 + kernel 2
def sample_noise_shape(num_rays, steps, P):
    """Shape of the noise tensor the reference draws (voxel_helpers.py:300-328)."""
    n = int(math.ceil(num_rays / G_SAMPLE))
    return (G_SAMPLE, n, int(steps.ceil().long().max()) + P)


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, probs, steps, noise=None,
                         fixed_step_size=-1.0, deterministic=False, generator=None):
    """InverseCDFRaySampling.forward, voxel_helpers.py:288-367.  ``noise``
    (shape ``sample_noise_shape``) may be passed in to replay a recorded draw."""
    G, N, P = G_SAMPLE, pts_idx.size(0), pts_idx.size(1)
    H = int(math.ceil(N / G)) * G
    if H > N:  # pad with copies of ray 0, :302-311
        rep = lambda t: torch.cat([t, t[:1].expand(H - N, *t.shape[1:])], 0)
        pts_idx, min_depth, max_depth, probs, steps = map(rep, (pts_idx, min_depth, max_depth, probs, steps))
    pts_idx = pts_idx.reshape(G, -1, P)
    min_depth = min_depth.reshape(G, -1, P)
    max_depth = max_depth.reshape(G, -1, P)
    probs = probs.reshape(G, -1, P)
    steps = steps.reshape(G, -1)
    max_steps = int(steps.ceil().long().max()) + P
    if noise is None:
        noise = min_depth.new_zeros(G, min_depth.size(1), max_steps)
        if deterministic:
            noise += 0.5
        else:
            noise = noise.uniform_(generator=generator).clamp(min=0.001, max=0.999)
    assert tuple(noise.shape) == (G, min_depth.size(1), max_steps), (noise.shape, max_steps)
    chunk = 4 * G   # :331
    outs = []
    for i in range(0, min_depth.size(1), chunk):
        sl = slice(i, i + chunk)
        outs.append(_c_sample(
            pts_idx[:, sl].contiguous().numpy(), min_depth[:, sl].contiguous().numpy(),
            max_depth[:, sl].contiguous().numpy(), noise[:, sl].contiguous().numpy(),
            probs[:, sl].contiguous().numpy(), steps[:, sl].contiguous().numpy(),
            float(fixed_step_size)))
    sidx, sdepth, sdist = [torch.from_numpy(np.concatenate([o[k] for o in outs], 1)) for k in range(3)]
    sidx, sdepth, sdist = sidx.reshape(H, -1)[:N], sdepth.reshape(H, -1)[:N], sdist.reshape(H, -1)[:N]
    width = int(sidx.ne(-1).sum(-1).max())   # :359 (assumes front-contiguous rows, A-Q9)
    return sidx[:, :width], sdepth[:, :width], sdist[:, :width], noise


def ray_sample(intersections, step_size, noise=None, generator=None):
    """voxel_helpers.py:637-663.  Returns (samples dict, noise used)."""
    idx = intersections["intersected_voxel_idx"]
    dists = (intersections["max_depth"] - intersections["min_depth"]).masked_fill(idx.eq(-1), 0)
    probs = dists / dists.sum(dim=-1, keepdim=True)
    steps = dists.sum(-1) / step_size
    sidx, sdepth, sdist, noise = inverse_cdf_sampling(
        idx, intersections["min_depth"], intersections["max_depth"], probs, steps,
        noise=noise, generator=generator)
    sdist = sdist.clamp(min=0.0)
    sdepth = sdepth.masked_fill(sidx.eq(-1), MAX_DEPTH)
    sdist = sdist.masked_fill(sidx.eq(-1), 0.0)
    return {
        "sampled_point_depth": sdepth,
        "sampled_point_distance": sdist,
        "sampled_point_voxel_idx": sidx,
        "probs": probs, "steps": steps,
    }, noise


# ---------------------------------------------------------------- a8 trilinear
_CORNERS = torch.tensor([[(i >> 2) & 1, (i >> 1) & 1, i & 1] for i in range(8)], dtype=torch.float32)


def get_features_vox(xyz, vox_idx, map_states, voxel_size):
    """render_helpers.py:105-156, 87-99, 47-59, 67-83: p = (x-c)/vs + .5,
    w_i = prod_a (p_a q_ia + (1-p_a)(1-q_ia)), feat = sum_i w_i E[vidx[j,i]].
    offset_points(bits=2) yields corners in the order (i>>2&1, i>>1&1, i&1)."""
    centres = map_states["voxel_center_xyz"]
    vidx = map_states["voxel_vertex_idx"]
    emb = map_states["voxel_vertex_emb"]
    j = vox_idx.long()
    c = F.embedding(j, centres)
    feats = F.embedding(F.embedding(j, vidx).long(), emb)            # [p,8,16]
    p = ((xyz - c) / voxel_size + 0.5).unsqueeze(1)                   # [p,1,3]
    q = _CORNERS.to(xyz).unsqueeze(0)                                 # [1,8,3]
    w = (p * q + (1 - p) * (1 - q)).prod(dim=-1, keepdim=True)        # [p,8,1]
    return (w * feats).sum(1)


# ---------------------------------------------------------------- a9 decoder
def decoder_params(width=128, in_dim=16, sdf_dim=128, seed=0, dtype=torch.float32):
    """Parameters with nn.Linear's default init, in the reference's
    state_dict order (nrgbd.py:106-113, depth=2, skips=[], embedder none):
    W1[w,16] b1, W2[w,w] b2, W3[1+sdf_dim,w] b3, W4[w,sdf_dim+16] b4, W5[3,w] b5."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(width, in_dim), (width, width), (1 + sdf_dim, width), (width, sdf_dim + in_dim), (3, width)]
    params = []
    for (o, i) in shapes:
        bound = 1.0 / math.sqrt(i)
        W = (torch.rand(o, i, generator=g, dtype=dtype) * 2 - 1) * bound
        b = (torch.rand(o, generator=g, dtype=dtype) * 2 - 1) * bound
        params += [W.requires_grad_(True), b.requires_grad_(True)]
    return params


def decoder_forward(params, f):
    """nrgbd.py:116-146: returns (rgb [p,3], sdf [p])."""
    W1, b1, W2, b2, W3, b3, W4, b4, W5, b5 = params
    h = F.relu(F.linear(f, W1, b1))
    h = F.relu(F.linear(h, W2, b2))
    o = F.linear(h, W3, b3)
    sdf, feat = o[:, 0], o[:, 1:]
    hc = F.relu(F.linear(torch.cat([feat, f], -1), W4, b4))
    rgb = torch.sigmoid(F.linear(hc, W5, b5))
    return rgb, sdf


# ---------------------------------------------------------------- a7 + a10
def sdf2weights(sdf, z_vals, valid, trunc):
    """render_helpers.py:521-539 on padded [R_h,S] tensors."""
    w = torch.sigmoid(sdf / trunc) * torch.sigmoid(-sdf / trunc)
    signs = sdf[:, 1:] * sdf[:, :-1]
    mask = torch.where(signs < 0.0, torch.ones_like(signs), torch.zeros_like(signs))
    inds = torch.argmax(mask, dim=1)[..., None]
    z_min = torch.gather(z_vals, 1, inds)
    mask = torch.where(z_vals < z_min + trunc, torch.ones_like(z_vals), torch.zeros_like(z_vals))
    w = w * mask * valid
    return w / (torch.sum(w, dim=-1, keepdim=True) + 1e-8), z_min


def render_rays(rays_o, rays_d, map_states, dec_params, step_size, voxel_size, truncation,
                max_voxel_hit, max_distance, noise=None, generator=None, inv_dir=None):
    """render_helpers.py:351-556 (chunking and file dumps are result-neutral
    and omitted).  Returns the reference's dict plus a few intermediates used by
    the tests (under "_dbg")."""
    inter, hits = ray_intersect_vox(rays_o, rays_d, map_states["voxel_center_xyz"].detach(),
                                    map_states["voxel_structure"], voxel_size, max_voxel_hit,
                                    max_distance, inv_dir=inv_dir)
    assert hits.sum() > 0
    ray_mask = hits.view(1, -1)
    inter = {k: v[ray_mask].reshape(-1, v.size(-1)) for k, v in inter.items()}
    ro = rays_o[ray_mask].reshape(-1, 3)
    rd = rays_d[ray_mask].reshape(-1, 3)
    samples, noise = ray_sample(inter, step_size, noise=noise, generator=generator)
    z = samples["sampled_point_depth"]
    sidx = samples["sampled_point_voxel_idx"].long()
    smask = sidx.ne(-1)
    if smask.sum() == 0:
        return None, 0
    xyz = ro.unsqueeze(1) + rd.unsqueeze(1) * z.unsqueeze(2)          # :436-437
    f = get_features_vox(xyz[smask], sidx[smask], map_states, voxel_size)
    rgb_p, sdf_p = decoder_forward(dec_params, f)
    sdf = torch.ones_like(z).masked_scatter(smask, sdf_p)            # pad 1, :510
    colour = z.new_zeros(*z.shape, 3).masked_scatter(smask.unsqueeze(-1).expand(*z.shape, 3), rgb_p)
    weights, z_min = sdf2weights(sdf, z, smask.to(z.dtype), truncation)
    rgb = torch.sum(weights[..., None] * colour, dim=-2)
    depth = torch.sum(weights * z, dim=-1)
    return {
        "weights": weights, "color": rgb, "depth": depth, "z_vals": z, "sdf": sdf,
        "ray_mask": ray_mask, "raw": z_min,
        "_dbg": {"intersections": inter, "samples": samples, "noise": noise, "sample_mask": smask,
                 "feat": f, "rgb_p": rgb_p, "sdf_p": sdf_p},
    }


# ---------------------------------------------------------------- a11 loss
def criterion(outputs, obs, *, rgb_weight, depth_weight, sdf_weight, fs_weight, truncation,
              max_depth, weight_depth_loss=False):
    """criterion.py:16-116.  Returns (loss, dict of tensors)."""
    img, depth = obs
    pred_depth, pred_color, pred_sdf = outputs["depth"], outputs["color"], outputs["sdf"]
    z, ray_mask, weights = outputs["z_vals"], outputs["ray_mask"], outputs["weights"]
    gt_depth, gt_color = depth[ray_mask], img[ray_mask]
    parts = {}
    parts["color_loss"] = (gt_color - pred_color).abs().mean()
    valid = (gt_depth > 0.01) & (gt_depth < max_depth)
    dl = (gt_depth - pred_depth).abs()
    if weight_depth_loss:   # :45-49, tracking only
        var = torch.sum(weights * ((pred_depth.unsqueeze(-1) - z) ** 2), -1)
        tmp = dl / torch.sqrt(var + 1e-10)
        valid = (tmp < 10 * tmp.median()) & valid
    parts["depth_loss"] = dl[valid].mean()
    # get_masks / get_sdf_loss, :78-116
    d = gt_depth.unsqueeze(-1).expand(*z.shape)
    front = torch.where(z < (d - truncation), torch.ones_like(z), torch.zeros_like(z))
    back = torch.where(z > (d + truncation), torch.ones_like(z), torch.zeros_like(z))
    dmask = torch.where((d > 0.0) & (d < max_depth), torch.ones_like(d), torch.zeros_like(d))
    smask = (1.0 - front) * (1.0 - back) * dmask
    n_fs = torch.count_nonzero(front).float()
    n_sdf = torch.count_nonzero(smask).float()
    n = n_fs + n_sdf
    fs_w, sdf_w = 1.0 - n_fs / n, 1.0 - n_sdf / n
    parts["fs_loss"] = torch.mean(torch.square(pred_sdf * front - torch.ones_like(pred_sdf) * front)) * fs_w
    parts["sdf_loss"] = torch.mean(torch.square((z + pred_sdf * truncation) * smask - d * smask)) * sdf_w
    loss = (rgb_weight * parts["color_loss"] + depth_weight * parts["depth_loss"]
            + fs_weight * parts["fs_loss"] + sdf_weight * parts["sdf_loss"])
    return loss, parts


# ---------------------------------------------------------------- se3 (boundary)
def se3_rotation(w):
    """se3pose.py:24-32, 62-91: Rodrigues with 10-term Taylor A, B."""
    w0, w1, w2 = w.unbind(-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], -1), torch.stack([w2, O, -w0], -1),
                      torch.stack([-w1, w0, O], -1)], -2)
    theta = w.norm(dim=-1)[..., None, None]
    A = torch.zeros_like(theta)
    B = torch.zeros_like(theta)
    dA, dB = 1.0, 1.0
    for i in range(11):
        if i > 0:
            dA *= (2 * i) * (2 * i + 1)
        A = A + (-1) ** i * theta ** (2 * i) / dA
        dB *= (2 * i + 1) * (2 * i + 2)
        B = B + (-1) ** i * theta ** (2 * i) / dB
    return torch.eye(3, dtype=w.dtype) + A * wx + B * wx @ wx
