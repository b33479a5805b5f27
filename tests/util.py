"""Shared helpers for the test-suite (fixtures, scene construction, error metrics)."""
import contextlib
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


def test_decoder(width=128, seed=1, sdf_gain=50.0):
    """Random decoder whose sdf head is scaled up so that |sdf| is ~1e-2 like a trained field, not
    ~1e-4: with the default init thousands of samples sit within fp32 rounding of sdf = 0 and the
    reference's first-sign-change decision is a coin toss there."""
    from oracle import render_oracle as ro
    dec = ro.decoder_params(width=width, seed=seed)
    with torch.no_grad():
        dec[4][0].mul_(sdf_gain)
        dec[5][0] = dec[5][0] * sdf_gain
    return dec


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


def elem_err(a, b, floor=0.02):
    """Element-wise relative error with an absolute floor: max_i |a_i - b_i| / max(|b_i|, floor * max|b|).  Entries above
    `floor` of the largest magnitude are held to the relative bound one by one; smaller ones (sums that cancel) to an
    absolute bound of floor * tolerance * max|b|."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    denom = b.abs().clamp(min=float(floor) * float(b.abs().max().clamp(min=1e-30)))
    return float(((a - b).abs() / denom).max())


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


def decision_margins(out, rgb, depth, tracking=False, truncation=0.1):
    """The reference's losses are discontinuous where a sign or mask decision flips (first sdf sign
    change, sign(pred - gt) of the L1 terms, the tracking median gate).  Returns how far the oracle's
    values are from each decision boundary, so that tests only compare cases where an fp32 rounding
    difference cannot flip a decision."""
    mask = out["_dbg"]["sample_mask"]
    sdf = out["sdf"].detach()
    ray_mask = out["ray_mask"].view(-1)
    gt_d = depth.reshape(-1)[ray_mask]
    gt_c = rgb.reshape(-1, 3)[ray_mask]
    # only samples up to (and including) the first sign change decide `ind`
    signs = (sdf[:, 1:] * sdf[:, :-1] < 0).float()
    has = signs.sum(-1) > 0
    ind = torch.where(has, signs.argmax(1), torch.full_like(signs.argmax(1), sdf.shape[1] - 1))
    k = torch.arange(sdf.shape[1])[None, :]
    relevant = mask & (k <= ind[:, None] + 1)
    m = {
        "sdf": float(sdf[relevant].abs().min()),
        "depth": float((out["depth"].detach() - gt_d).abs().min()),
        "color": float((out["color"].detach() - gt_c).abs().min()),
    }
    if tracking:
        z, w, pd = out["z_vals"], out["weights"].detach(), out["depth"].detach()
        var = torch.sum(w * (pd.unsqueeze(-1) - z) ** 2, -1)
        tmp = (gt_d - pd).abs() / torch.sqrt(var + 1e-10)
        thr = 10 * tmp.median()
        m["gate"] = float(((tmp - thr).abs() / thr).min())
    return m


def margins_ok(m):
    # a few times the fp32 rounding noise of each quantity (sdf ~1e-3, depth ~1, colour ~0.5)
    return m["sdf"] > 2e-7 and m["depth"] > 2e-6 and m["color"] > 1e-6 and m.get("gate", 1.0) > 1e-4


# ---------------------------------------------------------------- multi-rank loss closure (checker)
RAW = dict(COLOR=0, FS=1, SDF=2, D0=3, D1=4, NFS=5, F0=6, F1=7, NSDF=8, M0=9, M1=10, RH=11, DEPTH=12, NVALID=13, S=14, THRESH=15)


def raw_loss_sums(out, rgb, depth, truncation=0.1, max_depth=10.0):
    """This shard's raw loss sums in the slot layout of csrc/composite.cu (RAW_*), from oracle outputs."""
    mask = out["_dbg"]["sample_mask"]
    ray_mask = out["ray_mask"].view(-1)
    gt = depth.reshape(-1)[ray_mask].double()
    gc = rgb.reshape(-1, 3)[ray_mask].double()
    z, s = out["z_vals"].double(), out["sdf"].detach().double()
    cnt = mask.sum(-1).double()
    d = gt[:, None]
    front = (z < d - truncation) & mask
    back = (z > d + truncation)
    dm = ((d > 0) & (d < max_depth))
    sm = (~(z < d - truncation)) & (~back) & dm & mask
    pad_front = (10.0 < gt - truncation)
    pad_sm = (~pad_front) & (~(10.0 > gt + truncation)) & (gt > 0) & (gt < max_depth)
    pad_d2 = (10.0 + truncation - gt) ** 2
    valid = (gt > 0.01) & (gt < max_depth)
    r = torch.zeros(16, dtype=torch.float64)
    r[RAW["COLOR"]] = (gc - out["color"].detach().double()).abs().sum()
    r[RAW["FS"]] = (((s - 1.0) ** 2) * front).sum()
    r[RAW["SDF"]] = (((z + s * truncation - d) ** 2) * sm).sum()
    r[RAW["D0"]] = (pad_d2 * pad_sm).sum()
    r[RAW["D1"]] = (pad_d2 * pad_sm * cnt).sum()
    r[RAW["NFS"]] = front.sum()
    r[RAW["F0"]] = pad_front.sum()
    r[RAW["F1"]] = (pad_front * cnt).sum()
    r[RAW["NSDF"]] = sm.sum()
    r[RAW["M0"]] = pad_sm.sum()
    r[RAW["M1"]] = (pad_sm * cnt).sum()
    r[RAW["RH"]] = float(ray_mask.sum())
    r[RAW["DEPTH"]] = ((gt - out["depth"].detach().double()).abs() * valid).sum()
    r[RAW["NVALID"]] = valid.sum()
    r[RAW["S"]] = float(z.shape[1])
    r[RAW["THRESH"]] = float("inf")
    return r


def close_loss(rows, weights, truncation=0.1):
    """Restatement of k_loss_coeffs: all ranks' raw sums -> (loss dict, coefficient dict)."""
    t = rows[:, :RAW["S"]].sum(0)
    S = rows[:, RAW["S"]].max()
    w_rgb, w_depth, w_fs, w_sdf = weights
    Rh = t[RAW["RH"]]
    n = Rh * S
    nfs = t[RAW["NFS"]] + S * t[RAW["F0"]] - t[RAW["F1"]]
    nsdf = t[RAW["NSDF"]] + S * t[RAW["M0"]] - t[RAW["M1"]]
    sdf_sum = t[RAW["SDF"]] + S * t[RAW["D0"]] - t[RAW["D1"]]
    fs_w, sdf_w = 1 - nfs / (nfs + nsdf), 1 - nsdf / (nfs + nsdf)
    parts = dict(color_loss=t[RAW["COLOR"]] / (3 * Rh), depth_loss=t[RAW["DEPTH"]] / t[RAW["NVALID"]],
                 fs_loss=t[RAW["FS"]] / n * fs_w, sdf_loss=sdf_sum / n * sdf_w)
    parts["loss"] = (w_rgb * parts["color_loss"] + w_depth * parts["depth_loss"] + w_fs * parts["fs_loss"]
                     + w_sdf * parts["sdf_loss"])
    coef = dict(color=w_rgb / (3 * Rh), depth=w_depth / t[RAW["NVALID"]], fs=w_fs * fs_w / n, sdf=w_sdf * sdf_w / n)
    return {k: float(v) for k, v in parts.items()}, {k: float(v) for k, v in coef.items()}


def concat_outputs(outs):
    """Pads the shards' oracle outputs to the common S and concatenates them (what one big padded
    batch of the reference would hold: z pad 10, sdf pad 1, weight pad 0)."""
    S = max(o["z_vals"].shape[1] for o in outs)
    padc = lambda t, v: torch.cat([t, t.new_full((t.shape[0], S - t.shape[1]), v)], 1)
    return {
        "z_vals": torch.cat([padc(o["z_vals"], 10.0) for o in outs]),
        "sdf": torch.cat([padc(o["sdf"], 1.0) for o in outs]),
        "weights": torch.cat([padc(o["weights"], 0.0) for o in outs]),
        "color": torch.cat([o["color"] for o in outs]),
        "depth": torch.cat([o["depth"] for o in outs]),
        "ray_mask": torch.cat([o["ray_mask"].view(-1) for o in outs]).view(1, -1),
    }


# ---------------------------------------------------------------- staged parity (decision-proof)
def oracle_composite(out, sdf_p, rgb_p, rgb, depth, tracking=False):
    """Compositing + Criterion of the oracle evaluated on GIVEN per-sample decoder outputs
    (sdf_p [P], rgb_p [P,3], leaf tensors).  Used to check kernel 5 on bit-identical inputs: the
    first-sign-change / mask / sign() decisions of the reference are discontinuous, so they are
    only comparable when both sides see the same sdf bits."""
    smask = out["_dbg"]["sample_mask"]
    z = out["z_vals"]
    sdf = torch.ones_like(z).masked_scatter(smask, sdf_p)
    colour = z.new_zeros(*z.shape, 3).masked_scatter(smask.unsqueeze(-1).expand(*z.shape, 3), rgb_p)
    weights, z_min = ro.sdf2weights(sdf, z, smask.to(z.dtype), CRIT["truncation"])
    res = {"weights": weights, "color": torch.sum(weights[..., None] * colour, dim=-2),
           "depth": torch.sum(weights * z, dim=-1), "z_vals": z, "sdf": sdf, "ray_mask": out["ray_mask"], "raw": z_min}
    kw = {k: CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
    if tracking:
        r2 = dict(res)
        r2["ray_mask"] = res["ray_mask"].view(-1)
        loss, parts = ro.criterion(r2, (rgb[0], depth[0]), weight_depth_loss=True, **kw)
    else:
        loss, parts = ro.criterion(res, (rgb, depth), **kw)
    return res, loss, parts


def oracle_field(out, rays_o, rays_d, ms, dec, voxel_size):
    """Trilinear lookup + decoder of the oracle on the oracle's own samples: (rgb_p [P,3], sdf_p [P])
    connected to rays_o / rays_d / embeddings / decoder parameters for autograd."""
    smask = out["_dbg"]["sample_mask"]
    rm = out["ray_mask"]
    ro_ = rays_o[rm].reshape(-1, 3)
    rd_ = rays_d[rm].reshape(-1, 3)
    z = out["z_vals"]
    xyz = ro_.unsqueeze(1) + rd_.unsqueeze(1) * z.unsqueeze(2)
    sidx = out["_dbg"]["samples"]["sampled_point_voxel_idx"].long()
    f = ro.get_features_vox(xyz[smask], sidx[smask], ms, voxel_size)
    return ro.decoder_forward(dec, f)


# ReLU'(0) is a hard decision: a pre-activation within a build's own rounding error of 0 flips one unit's mask between
# two correct implementations and changes that sample's gradient by O(1) (the reference's fp32 has the same kink at its
# own rounding level).  Backward comparisons therefore give such samples zero upstream gradient on both sides.
RELU_MARGIN = {"tf32": 2e-6, "simt": 2e-6, "f16": 4e-6}


def decoder_near(dec, f, eps):
    """bool [P]: rows of the feature matrix f with a decoder pre-activation closer to 0 than eps."""
    with torch.no_grad():
        W1, b1, W2, b2, W3, b3, W4, b4, W5, b5 = [p.detach() for p in dec]
        a1 = f @ W1.t() + b1
        a2 = torch.relu(a1) @ W2.t() + b2
        t = (torch.relu(a2) @ W3.t() + b3)[:, 1:]
        a4 = torch.cat([t, f], 1) @ W4.t() + b4
        return (a1.abs().min(1).values < eps) | (a2.abs().min(1).values < eps) | (a4.abs().min(1).values < eps)


def relu_near_samples(out, rays_o, rays_d, ms, dec, voxel_size, eps):
    """bool [P]: samples (oracle / CSR order) with a decoder pre-activation closer to 0 than eps."""
    with torch.no_grad():
        smask = out["_dbg"]["sample_mask"]
        rm = out["ray_mask"]
        z = out["z_vals"]
        xyz = rays_o[rm].reshape(-1, 3).unsqueeze(1) + rays_d[rm].reshape(-1, 3).unsqueeze(1) * z.unsqueeze(2)
        sidx = out["_dbg"]["samples"]["sampled_point_voxel_idx"].long()
        f = ro.get_features_vox(xyz[smask], sidx[smask], ms, voxel_size)
        W1, b1, W2, b2, W3, b3, W4, b4, W5, b5 = [p.detach() for p in dec]
        a1 = f @ W1.t() + b1
        a2 = torch.relu(a1) @ W2.t() + b2
        t = (torch.relu(a2) @ W3.t() + b3)[:, 1:]
        a4 = torch.cat([t, f], 1) @ W4.t() + b4
        return (a1.abs().min(1).values < eps) | (a2.abs().min(1).values < eps) | (a4.abs().min(1).values < eps)


# ---------------------------------------------------------------- a frame like the reference's RGBDFrame
class TestFrame:
    """Duck-type of reference ``src/frame.py:10-85`` for the loop tests: rays_d per pixel, rgb, depth,
    an OptimizablePose, ``sample_rays`` (uniform without replacement), ``get_pose``."""

    def __init__(self, scene, frame, stamp, device, perturb=None, seed=0):
        from proud_slam_b200.se3pose import OptimizablePose
        self.stamp = stamp
        self.rays_d = scene.rays_cam.reshape(-1, 3).to(device)
        self.rgb = frame.rgb.reshape(-1, 3).to(device)
        self.depth = frame.depth.reshape(-1).to(device)
        pose = frame.pose.clone()
        if perturb is not None:
            pose[:3, 3] += torch.as_tensor(perturb, dtype=torch.float32)
        self.pose = OptimizablePose.from_matrix(pose).to(device)
        self.optim = torch.optim.Adam(self.pose.parameters(), lr=1e-3)
        self.gen = torch.Generator().manual_seed(seed)
        self.sample_mask = None

    def get_pose(self):
        return self.pose.matrix()

    def sample_rays(self, n):
        idx = torch.randperm(self.rays_d.shape[0], generator=self.gen)[:n]
        m = torch.zeros(self.rays_d.shape[0], dtype=torch.bool)
        m[idx] = True
        self.sample_mask = m.to(self.rays_d.device)


# PSLAM_OPT_DECODER values (include/proud_slam_b200.h); the library default is the 3xF16 tensor-core build
DECODER_BUILDS = {"tf32": 0, "simt": 1, "f16": 2}
DEFAULT_DECODER_BUILD = int(os.environ.get("PSLAM_DECODER", "2"))


@contextlib.contextmanager
def decoder_build(build, save_activations=True):
    """Selects the decoder build (name or option value) for a block and restores the defaults afterwards."""
    from proud_slam_b200 import _lib
    lib = _lib.lib()
    mode = DECODER_BUILDS[build] if isinstance(build, str) else int(build)
    _lib.check(lib.pslam_set_option(1, mode), "set_option")
    _lib.check(lib.pslam_set_option(2, 1 if save_activations else 0), "set_option")
    try:
        yield mode
    finally:
        lib.pslam_set_option(1, DEFAULT_DECODER_BUILD)
        lib.pslam_set_option(2, 1)
