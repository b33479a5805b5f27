"""Parity of the fused CUDA render path (through the C ABI) against the golden fixtures
generated from the real reference, and against the oracle on larger seeded scenes.

Tolerances (BASELINE.json north_star): voxel hit lists and sample indices bit-exact;
rendered colour/depth/sdf, losses and gradients within 1e-4 relative (max-norm) in fp32.
"""
import numpy as np
import pytest
import torch

from tests import util
from tests.util import elem_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4        # norm-wise: max |a - b| / max |b|
# Element-wise (tests.util.elem_err: every entry above 2 % of the largest one is held to the relative bound by itself, smaller
# ones -- sums that cancel -- to that bound x 2 % of the maximum).  Measured on the B200: gradients and rendered outputs
# 1e-7 .. 1.4e-4 in every build including the exact-fp32 SIMT one (the remaining differences are summation order against the
# CPU oracle), hence 2e-4.  The per-sample sdf is a 128-term dot product whose value is ~1 % of its terms: against the CPU
# oracle the fp32 SIMT build itself reaches only 6e-4 at a 2 % floor, so it is held to 5e-4 above 20 % of the maximum (measured: SIMT 6e-5, 3xF16 1.1e-4, 3xTF32 2.8e-4).
ELEM_TOL = 2e-4
SDF_ELEM_TOL, SDF_ELEM_FLOOR = 5e-4, 0.2
NEAR_TIE_MAX = 0.02   # largest share of samples whose upstream gradient may be zeroed as "a ReLU pre-activation within rounding of 0"


def _run_pipeline(device, rays_o, rays_d, rgb, depth, ms, dec, *, voxel_size, step_size, truncation, max_distance,
                  max_depth, weights, noise, tracking, grads=True, zero_upstream=None, samples_per_ray=96):
    from proud_slam_b200.pipeline import RenderPipeline
    R = rays_o.reshape(-1, 3).shape[0]
    pipe = RenderPipeline(R, device, samples_per_ray=samples_per_ray)
    g_emb = torch.zeros_like(ms["voxel_vertex_emb"]) if grads else None
    g_dec = [torch.zeros_like(p) for p in dec] if grads else None
    pipe.bind(rays_o, rays_d, ms, dec, voxel_size=voxel_size, step_size=step_size, truncation=truncation,
              max_distance=max_distance, max_depth=max_depth, target_rgb=rgb, target_depth=depth, noise=noise,
              weights=weights, tracking=tracking, g_emb=g_emb, g_dec=g_dec, grad_rays=grads)
    if zero_upstream is None:
        pipe.step()
    else:
        # same launches as step(), with the upstream gradient of the flagged samples zeroed before the field backward
        pipe.sample()
        pipe.forward()
        pipe.stage(4)                                   # compositing backward -> samp_gout
        pipe.g_full = pipe.samp_gout[:zero_upstream.numel()].clone()
        pipe.samp_gout[:zero_upstream.numel()][zero_upstream.to(device)] = 0.0
        pipe.stage(5)                                   # field backward
    torch.cuda.synchronize()
    return pipe, g_emb, g_dec


@pytest.mark.parametrize("build", ["f16", "f16-recompute", "tf32"])
@pytest.mark.parametrize("name", ["mapping_tiny", "tracking_tiny", "mapping_tiny_w256", "mapping_replica_small"])
def test_step_matches_reference_golden(name, build, device):
    with util.decoder_build(build.split("-")[0], save_activations=not build.endswith("recompute")):
        _golden_step(name, device)


def _golden_step(name, device):
    g = util.load_golden(name)
    ms = util.golden_map_states(g, device, requires_grad=False)
    dec = util.golden_decoder(g, device, requires_grad=False)
    dev = lambda k: torch.from_numpy(g[k]).to(device).contiguous()
    noise = dev("noise")
    noise = noise.reshape(-1, noise.shape[-1]).contiguous()
    cw = g["crit_weights"]
    pipe, g_emb, g_dec = _run_pipeline(
        device, dev("rays_o"), dev("rays_d"), dev("rgb"), dev("depth"), ms, dec, voxel_size=float(g["voxel_size"]),
        step_size=float(g["step_size"]), truncation=float(g["truncation"]), max_distance=float(g["max_distance"]),
        max_depth=float(g["max_depth"]), weights=tuple(float(x) for x in cw), noise=noise, tracking=bool(g["tracking"]))
    inter, hits = pipe.intersections()
    # hit lists: bit-exact ids; the fixture's depths come from the CPU oracle (IEEE 1/d instead of the
    # device reciprocal), so depths are compared to a few ulp here and bit-exactly in test_gpu_grid
    assert np.array_equal(hits.cpu().numpy(), g["hits"])
    assert np.array_equal(inter["intersected_voxel_idx"].cpu().numpy(), g["hit_idx"])
    np.testing.assert_allclose(inter["min_depth"].cpu().numpy(), g["hit_min"], rtol=2e-6, atol=1e-6)
    out = pipe.outputs()
    assert np.array_equal(out["ray_mask"].cpu().numpy().reshape(-1), g["out_ray_mask"].reshape(-1))
    assert tuple(out["z_vals"].shape) == g["out_z_vals"].shape
    np.testing.assert_allclose(out["z_vals"].cpu().numpy(), g["out_z_vals"], rtol=1e-5, atol=1e-5)
    assert rel_err(out["sdf"], g["out_sdf"]) < TOL
    assert rel_err(out["color"], g["out_color"]) < TOL
    assert rel_err(out["depth"], g["out_depth"]) < TOL
    assert rel_err(out["weights"], g["out_weights"]) < TOL
    assert rel_err(out["raw"], g["out_raw"]) < TOL
    l = pipe.losses()
    assert abs(l["loss"] - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    for k, ref in zip(("color_loss", "depth_loss", "fs_loss", "sdf_loss"), g["loss_parts"]):
        assert abs(l[k] - float(ref)) <= TOL * max(abs(float(ref)), 1e-12), k
    assert rel_err(g_emb, g["g_emb"]) < TOL
    for i in range(10):
        assert rel_err(g_dec[i], g[f"g_dec_{i}"]) < TOL, f"decoder grad {i}"
    R = g["rays_o"].reshape(-1, 3).shape[0]
    assert rel_err(pipe.g_rays_o[:R], g["g_rays_o"].reshape(-1, 3)) < TOL
    assert rel_err(pipe.g_rays_d[:R], g["g_rays_d"].reshape(-1, 3)) < TOL
    # element-wise as well (entries above 2 % of the largest one by one; the rest to an absolute 2e-6 x max)
    errs = {"sdf": elem_err(out["sdf"], g["out_sdf"]), "color": elem_err(out["color"], g["out_color"]), "depth": elem_err(out["depth"], g["out_depth"]),
            "g_emb": elem_err(g_emb, g["g_emb"]), "g_rays_d": elem_err(pipe.g_rays_d[:R], g["g_rays_d"].reshape(-1, 3))}
    errs.update({f"g_dec_{i}": elem_err(g_dec[i], g[f"g_dec_{i}"]) for i in range(10)})
    print(f"{name}: element-wise errors " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < ELEM_TOL, errs


@pytest.mark.parametrize("decoder_build", ["f16", "f16-recompute", "tf32", "simt"])
@pytest.mark.parametrize("kind,frames,rays,tracking,width", [
    ("tiny", 2, 300, False, 128),
    ("replica_small", 2, 1024, False, 128),       # BASELINE.json configs[0]: 2048 rays on the 0.2 m octree
    ("replica_small", 1, 1024, True, 128),        # configs[2]: one tracking iteration
    ("replica_small", 2, 256, False, 256),
    ("replica_20k", 8, 1024, False, 128),         # configs[1] at full size: 8192 rays, ~190k samples (the bench workload)
    ("scannet_large", 2, 1024, False, 128),       # configs[3]: 0.1 m voxels, 276k octants
])
def test_step_matches_oracle(kind, frames, rays, tracking, width, decoder_build, device):
    """Whole iteration against the oracle, stage by stage on identical inputs.

    The reference's compositing and losses contain hard decisions (first sdf sign change along the
    ray, z < z_min + tau, sign(pred - gt), the tracking median gate).  An fp32 rounding difference in
    one sdf value near zero flips such a decision and legitimately changes that ray's result, so an
    end-to-end comparison is only meaningful where no decision is near a tie (the golden-fixture
    test above).  Here every stage is instead compared on bit-identical inputs:
      1. hit lists, sample ids/depths: bit-exact;
      2. per-sample decoder outputs (continuous in its inputs): 1e-4;
      3. compositing + losses + dL/d(sample outputs) evaluated by the oracle ON THE GPU'S OWN
         per-sample outputs: identical decisions, 1e-4;
      4. field backward (decoder dgrad/wgrad, embedding scatter, ray gradients) of the oracle fed
         with the GPU's dL/d(sample outputs): 1e-4.
    """
    from proud_slam_b200 import _lib, scene as sc
    if width == 256 and decoder_build in ("tf32", "f16-recompute"):
        pytest.skip("width 256 has two builds: tcgen05 3xF16 (f16) and SIMT fp32")
    s, ms = util.build_scene(kind)
    dec = util.test_decoder(width=width, seed=1)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, list(range(frames)), rays, seed=5)
    depth = depth * (1.0 + 0.01 * torch.randn(depth.shape, generator=torch.Generator().manual_seed(4)))
    rays_o.requires_grad_(True)
    rays_d.requires_grad_(True)
    inv = util.device_rcp(rays_d.detach().reshape(-1, 3), device)
    from oracle import render_oracle as ro
    out = ro.render_rays(rays_o, rays_d, ms, dec, 0.1 * s.voxel_size, s.voxel_size, util.CRIT["truncation"], 10, 10.0,
                         generator=torch.Generator().manual_seed(11), inv_dir=inv)
    noise = out["_dbg"]["noise"]
    noise_d = noise.reshape(-1, noise.shape[-1]).to(device).contiguous()
    msd = util.to_device(ms, device)
    msd["voxel_vertex_emb"] = msd["voxel_vertex_emb"].detach()
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    near = util.relu_near_samples(out, rays_o.detach(), rays_d.detach(), ms, dec, s.voxel_size, util.RELU_MARGIN[decoder_build.split("-")[0]])
    near_frac = float(near.float().mean())
    print(f"{kind}/{decoder_build}: {100 * near_frac:.3f} % of the samples masked as ReLU near-ties")
    assert near_frac < NEAR_TIE_MAX * width / 128          # (twice the ReLU units per sample at width 256)
    with util.decoder_build(decoder_build.split("-")[0], save_activations=not decoder_build.endswith("recompute")):
        pipe, g_emb, g_dec = _run_pipeline(
            device, rays_o.detach().to(device), rays_d.detach().to(device), rgb.to(device), depth.to(device), msd, decd,
            voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=util.CRIT["truncation"], max_distance=10.0,
            max_depth=util.CRIT["max_depth"], weights=cw, noise=noise_d, tracking=tracking, zero_upstream=near,
            samples_per_ray=max(96, int(noise.shape[-1])))
    # ---- 1. bit-exact hit lists and samples
    inter, hits = pipe.intersections()
    hit_rows = hits.view(-1).cpu()
    ref_inter = out["_dbg"]["intersections"]
    assert np.array_equal(hit_rows.numpy(), out["ray_mask"].view(-1).numpy())
    for k in ("intersected_voxel_idx", "min_depth", "max_depth"):
        assert torch.equal(inter[k][0].cpu()[hit_rows], ref_inter[k]), k
    smp = pipe.samples()
    ref_s = out["_dbg"]["samples"]
    assert torch.equal(smp["sampled_point_voxel_idx"].cpu(), ref_s["sampled_point_voxel_idx"])
    assert torch.equal(smp["sampled_point_depth"].cpu(), ref_s["sampled_point_depth"])
    assert torch.equal(smp["sampled_point_distance"].cpu(), ref_s["sampled_point_distance"])
    # ---- 2. per-sample decoder outputs
    P = pipe.counts()["n_samples"]
    assert P == int(out["_dbg"]["sample_mask"].sum())
    so = pipe.samp_out[:P].cpu()
    assert rel_err(so[:, :3], out["_dbg"]["rgb_p"].detach()) < TOL
    assert rel_err(so[:, 3], out["_dbg"]["sdf_p"].detach()) < TOL
    # ---- 3. compositing + losses on the GPU's own sample outputs
    sdf_p = so[:, 3].clone().requires_grad_(True)
    rgb_p = so[:, :3].clone().requires_grad_(True)
    res, loss, parts = util.oracle_composite(out, sdf_p, rgb_p, rgb, depth, tracking)
    loss.backward()
    o = pipe.outputs()
    for k in ("sdf", "color", "depth", "weights", "raw"):
        assert rel_err(o[k], res[k].detach()) < TOL, k
    l = pipe.losses()
    assert abs(l["loss"] - float(loss)) <= TOL * abs(float(loss))
    for k in ("color_loss", "depth_loss", "fs_loss", "sdf_loss"):
        assert abs(l[k] - float(parts[k])) <= TOL * max(abs(float(parts[k])), 1e-12), k
    g_so = pipe.g_full.cpu()
    assert rel_err(g_so[:, :3], rgb_p.grad) < TOL
    assert rel_err(g_so[:, 3], sdf_p.grad) < TOL
    # ---- 4. field backward fed with the GPU's upstream gradient (zero where a ReLU decision is within rounding of a tie)
    g_so = pipe.samp_gout[:P].cpu()
    rgb_o, sdf_o = util.oracle_field(out, rays_o, rays_d, ms, dec, s.voxel_size)
    params = [ms["voxel_vertex_emb"], rays_o, rays_d] + list(dec)
    grads = torch.autograd.grad((rgb_o * g_so[:, :3]).sum() + (sdf_o * g_so[:, 3]).sum(), params)
    assert rel_err(g_emb, grads[0]) < TOL
    R = rays_o.shape[1]
    assert rel_err(pipe.g_rays_o[:R], grads[1].reshape(-1, 3)) < TOL
    assert rel_err(pipe.g_rays_d[:R], grads[2].reshape(-1, 3)) < TOL
    for i in range(10):
        assert rel_err(g_dec[i], grads[3 + i]) < TOL, f"decoder grad {i}"
    sdf_elem = elem_err(so[:, 3], out["_dbg"]["sdf_p"].detach(), floor=SDF_ELEM_FLOOR)
    assert sdf_elem < SDF_ELEM_TOL, sdf_elem
    errs = {"rgb_p": elem_err(so[:, :3], out["_dbg"]["rgb_p"].detach()),
            "g_emb": elem_err(g_emb, grads[0]), "g_rays_d": elem_err(pipe.g_rays_d[:R], grads[2].reshape(-1, 3))}
    errs.update({f"g_dec_{i}": elem_err(g_dec[i], grads[3 + i]) for i in range(10)})
    print(f"{kind}/{decoder_build}: element-wise errors sdf_p {sdf_elem:.1e}, " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < ELEM_TOL, errs


def test_saved_activations_equal_recompute(device):
    """3xF16 build: the backward that starts from the activations / ReLU masks spilled by the forward (kFwdSave +
    kBwdSaved) gives the same gradients as the stand-alone backward that recomputes the forward (same arithmetic,
    same decisions; only the order of the floating-point atomics differs)."""
    from proud_slam_b200 import scene as sc
    s, ms = util.build_scene("replica_small")
    dec = util.test_decoder(width=128, seed=3)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 1024, seed=9)
    msd = util.to_device(ms, device)
    msd["voxel_vertex_emb"] = msd["voxel_vertex_emb"].detach()
    decd = [p.detach().to(device) for p in dec]
    noise = torch.rand(2048, 96, generator=torch.Generator().manual_seed(2)).clamp(0.001, 0.999).to(device)
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    res = []
    for save in (True, False):
        with util.decoder_build("f16", save_activations=save):
            pipe, g_emb, g_dec = _run_pipeline(
                device, rays_o.to(device), rays_d.to(device), rgb.to(device), depth.to(device), msd, decd, voxel_size=s.voxel_size,
                step_size=0.1 * s.voxel_size, truncation=util.CRIT["truncation"], max_distance=10.0, max_depth=util.CRIT["max_depth"],
                weights=cw, noise=noise, tracking=False)
        assert pipe.counts()["n_samples"] > 20000          # many tiles per CTA pair
        res.append([g_emb.clone(), pipe.g_rays_o.clone(), pipe.g_rays_d.clone(), pipe.samp_out.clone()] + [g.clone() for g in g_dec])
    assert torch.equal(res[0][3], res[1][3])               # forward outputs: bit-identical
    for a, b in zip(res[0], res[1]):
        assert rel_err(a, b) < 1e-5


def test_programmatic_launch_changes_nothing(device):
    """PSLAM_OPT_PDL: launching the chain with programmatic stream serialization (every kernel starts with
    griddepcontrol.wait) overlaps launch latency only -- samples and forward outputs are bit-identical to plain stream
    order over repeated steps, gradients equal up to the order of the floating-point atomics."""
    from proud_slam_b200 import _lib, scene as sc
    s, ms = util.build_scene("replica_small")
    dec = util.test_decoder(width=128, seed=3)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 1024, seed=9)
    msd = util.to_device(ms, device)
    msd["voxel_vertex_emb"] = msd["voxel_vertex_emb"].detach()
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    res = []
    try:
        for pdl in (1, 0):
            assert _lib.lib().pslam_set_option(3, pdl) == 0
            for rep in range(3):                            # back-to-back steps: the chain also crosses step boundaries
                pipe, g_emb, g_dec = _run_pipeline(
                    device, rays_o.to(device), rays_d.to(device), rgb.to(device), depth.to(device), msd, decd, voxel_size=s.voxel_size,
                    step_size=0.1 * s.voxel_size, truncation=util.CRIT["truncation"], max_distance=10.0, max_depth=util.CRIT["max_depth"],
                    weights=cw, noise=None, tracking=False)
            n = pipe.counts()["n_samples"]
            res.append([pipe.samp_vox[:n].clone(), pipe.samp_z[:n].clone(), pipe.samp_out[:n].clone(), pipe.loss.clone(),
                        g_emb.clone(), pipe.g_rays_o.clone(), pipe.g_rays_d.clone()] + [g.clone() for g in g_dec])
    finally:
        _lib.lib().pslam_set_option(3, 1)
    for a, b in zip(res[0][:4], res[1][:4]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][4:], res[1][4:]):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("option,value", [(7, 1), (6, 1)])
def test_optional_kernel_paths_change_nothing(option, value, device):
    """PSLAM_OPT_WALK = 1 (block-cooperative level-synchronous octree walk) and PSLAM_OPT_FUSED_SCATTER = 1 (trilinear backward in
    the idle warps of the fused backward kernel) against the defaults: hit lists, samples and forward outputs bit-identical,
    gradients equal up to the order of the floating-point atomics."""
    from proud_slam_b200 import _lib, scene as sc
    s, ms = util.build_scene("replica_small")
    dec = util.test_decoder(width=128, seed=3)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1, 2], 1024, seed=13)
    msd = util.to_device(ms, device)
    msd["voxel_vertex_emb"] = msd["voxel_vertex_emb"].detach()
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    res = []
    try:
        for v in (0, value):
            assert _lib.lib().pslam_set_option(option, v) == 0
            pipe, g_emb, g_dec = _run_pipeline(
                device, rays_o.to(device), rays_d.to(device), rgb.to(device), depth.to(device), msd, decd, voxel_size=s.voxel_size,
                step_size=0.1 * s.voxel_size, truncation=util.CRIT["truncation"], max_distance=10.0, max_depth=util.CRIT["max_depth"],
                weights=cw, noise=None, tracking=False)
            n = pipe.counts()["n_samples"]
            R = rays_o.shape[1]
            res.append([pipe.hit_count[:R].clone(), pipe.hit_idx[: 50 * R].clone() * 0 + pipe.hit_idx[: 50 * R].clone(), pipe.samp_vox[:n].clone(),
                        pipe.samp_z[:n].clone(), pipe.samp_out[:n].clone(), pipe.loss.clone(),
                        g_emb.clone(), pipe.g_rays_o.clone(), pipe.g_rays_d.clone()] + [g.clone() for g in g_dec])
    finally:
        _lib.lib().pslam_set_option(option, 0)
    cnt = res[0][0].long()
    assert torch.equal(res[0][0], res[1][0])
    valid = (torch.arange(50, device=device)[:, None] < cnt[None, :]).reshape(-1)      # slot-major [50, R]: only the valid slots are defined
    assert torch.equal(res[0][1][valid], res[1][1][valid])
    for a, b in zip(res[0][2:6], res[1][2:6]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][6:], res[1][6:]):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("tiles", [1, 147, 148, 149, 150, 297, 445])
def test_fused_backward_at_tile_count_boundaries(tiles, device):
    """k_field_bw (dgrad + weight gradients in one kernel): its CTAs work in 2-CTA clusters that share one weight stream, clusters
    run a per-cluster number of rounds and drain early when they have no tile in the last one.  The tile count is pinned through
    the sample capacity (the batch has more samples; the tail is cut identically for both runs): fewer tiles than CTAs, exactly
    one round, one tile more (a cluster with one real and one dummy tile), two, 2 rounds + 1, 3 rounds + 1.  Against the
    unfused kernels (PSLAM_OPT_FUSED_WGRAD = 0: chain kernel + k_wgrad_bf)."""
    from proud_slam_b200 import _lib, scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("replica_small")
    dec = util.test_decoder(width=128, seed=3)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1, 2], 2048, seed=17)
    msd = util.to_device(ms, device)
    msd["voxel_vertex_emb"] = msd["voxel_vertex_emb"].detach()
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    cap = 128 * tiles
    R = min(rays_o.shape[0] * rays_o.shape[1], cap)
    inp = [t.to(device).reshape(-1, 3)[:R].contiguous() for t in (rays_o, rays_d, rgb)] + [depth.to(device).reshape(-1)[:R].contiguous()]
    spr = next(d for d in range(cap // R, 0, -1) if cap % d == 0)          # sample capacity = max_rays x samples_per_ray = cap exactly
    res = []
    try:
        for fused in (1, 0):
            assert _lib.lib().pslam_set_option(5, fused) == 0
            pipe = RenderPipeline(cap // spr, device, samples_per_ray=spr)
            assert pipe.sample_cap == cap
            g_emb = torch.zeros_like(msd["voxel_vertex_emb"])
            g_dec = [torch.zeros_like(p) for p in decd]
            pipe.bind(inp[0], inp[1], msd, decd, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=util.CRIT["truncation"],
                      max_distance=10.0, max_depth=util.CRIT["max_depth"], target_rgb=inp[2], target_depth=inp[3], noise=None, weights=cw,
                      tracking=False, g_emb=g_emb, g_dec=g_dec, grad_rays=True)
            pipe.step()
            torch.cuda.synchronize()
            host = pipe.counters.cpu()
            assert int(host[_lib.C_NSAMP]) >= cap and int(host[_lib.C_OVERFLOW]) & 1, "the batch must overfill the capacity: that pins the tile count"
            res.append([g_emb, pipe.g_rays_o.clone(), pipe.g_rays_d.clone()] + g_dec)
    finally:
        _lib.lib().pslam_set_option(5, 1)
    for a, b in zip(res[0], res[1]):
        assert torch.isfinite(a).all()
        assert rel_err(a, b) < 1e-5


def test_f16_operand_range_is_guarded(device):
    """3xF16 build: operands are kept in f16's window by fixed power-of-two scales; a decoder whose activations leave it
    (|16 x value| >= 32752) must be reported through the overflow counter, not silently clipped."""
    from proud_slam_b200 import scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("tiny")
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 200, seed=5)
    for gain, expect_error in ((1.0, False), (3.0e5, True)):
        dec = [p.detach().clone().to(device) for p in util.test_decoder(width=128, seed=1)]
        dec[0] *= gain                                   # first-layer weights: activations (and the weights themselves) blow up
        pipe = RenderPipeline(400, device)
        pipe.bind(rays_o.to(device), rays_d.to(device), msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size,
                  truncation=0.1, max_distance=10.0, target_rgb=rgb.to(device), target_depth=depth.to(device), seed=1,
                  forward_only=True)
        with util.decoder_build("f16"):
            pipe.step()
            torch.cuda.synchronize()
        if expect_error:
            with pytest.raises(RuntimeError, match="3xF16"):
                pipe.counts()
        else:
            pipe.counts()


def test_hash_noise_is_in_range_and_deterministic(device):
    """Production noise (no tensor): same seed -> identical samples; range like the reference's clamp."""
    from oracle import render_oracle as ro
    from proud_slam_b200 import scene as sc
    s, ms = util.build_scene("tiny")
    dec = [p.detach().to(device) for p in ro.decoder_params(seed=1)]
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 200, seed=5)
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    zs = []
    for seed in (3, 3, 4):
        from proud_slam_b200.pipeline import RenderPipeline
        pipe = RenderPipeline(400, device)
        pipe.bind(rays_o.to(device), rays_d.to(device), msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size,
                  truncation=0.1, max_distance=10.0, target_rgb=rgb.to(device), target_depth=depth.to(device), seed=seed,
                  forward_only=True)
        pipe.step()
        zs.append(pipe.samples()["sampled_point_depth"].cpu())
    assert torch.equal(zs[0], zs[1])
    assert zs[0].shape != zs[2].shape or not torch.equal(zs[0], zs[2])


@pytest.mark.parametrize("R", [1, 5, 63, 65, 299])
def test_ragged_batch_sizes_hit_lists(R, device):
    """Batch sizes that leave warps / blocks of the 8-lanes-per-ray traversal partly empty."""
    from oracle import render_oracle as ro
    from proud_slam_b200 import scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("tiny")
    dec = [p.detach().to(device) for p in ro.decoder_params(seed=1)]
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0], R, seed=R)
    inv = util.device_rcp(rays_d.reshape(-1, 3), device)
    ref, ref_hits = ro.ray_intersect_vox(rays_o, rays_d, ms["voxel_center_xyz"].detach(), ms["voxel_structure"], s.voxel_size, 10, 10.0, inv_dir=inv)
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    pipe = RenderPipeline(R, device)
    pipe.bind(rays_o.to(device), rays_d.to(device), msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1,
              max_distance=10.0, seed=1, forward_only=True)
    pipe.step()
    inter, hits = pipe.intersections()
    assert torch.equal(hits.cpu(), ref_hits)
    for k in ref:
        assert torch.equal(inter[k].cpu(), ref[k]), k
    assert pipe.counts()["R_h"] == int(ref_hits.sum())


@pytest.mark.parametrize("walk", [0, 1])
@pytest.mark.parametrize("n_max", [1, 3, 6])
def test_hit_cap_keeps_the_first_hits_of_the_reference_walk(n_max, walk, device):
    """Rays that hit more leaves than the cap: the reference's DFS stops after n_max emissions (intersect_gpu.cu:233), i.e. it
    keeps the FIRST n_max leaves of its walk order -- not the nearest.  The product's walks run in another order and keep the
    n_max largest DFS keys instead; both walks (PSLAM_OPT_WALK 0 / 1) must reproduce the oracle's DFS, cut included."""
    import oracle
    from oracle import render_oracle as ro
    from proud_slam_b200 import _lib, scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("replica_small")
    dec = [p.detach().to(device) for p in ro.decoder_params(seed=1)]
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 3], 700, seed=21)
    R = rays_o.shape[1]
    inv = util.device_rcp(rays_d.reshape(-1, 3), device)
    idx, tmin, tmax = oracle.svo_intersect(rays_o.reshape(1, R, 3).numpy(), rays_d.reshape(1, R, 3).numpy(),
                                           ms["voxel_center_xyz"].detach().reshape(1, -1, 3).numpy(), ms["voxel_structure"].reshape(1, -1, 9).numpy(),
                                           float(s.voxel_size), n_max, np.asarray(inv).reshape(1, R, 3))
    idx, tmin, tmax = torch.from_numpy(idx)[0], torch.from_numpy(tmin)[0], torch.from_numpy(tmax)[0]
    full, _ = ro.ray_intersect_vox(rays_o, rays_d, ms["voxel_center_xyz"].detach(), ms["voxel_structure"], s.voxel_size, 10, 10.0, inv_dir=inv)
    assert int(full["intersected_voxel_idx"].ne(-1).sum(-1).max()) > n_max          # the cap really cuts
    tmin = tmin.masked_fill(idx.eq(-1), 10.0)
    tmin, order = tmin.sort(dim=-1, stable=True)
    idx, tmax = idx.gather(-1, order), tmax.gather(-1, order)
    idx[tmin > 10.0] = -1
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    try:
        assert _lib.lib().pslam_set_option(7, walk) == 0
        pipe = RenderPipeline(R, device, n_max=n_max)
        pipe.bind(rays_o.to(device), rays_d.to(device), msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1,
                  max_distance=10.0, seed=1, forward_only=True)
        pipe.sample()
        inter, hits = pipe.intersections()
    finally:
        _lib.lib().pslam_set_option(7, 0)
    got, width = inter["intersected_voxel_idx"][0].cpu(), inter["intersected_voxel_idx"].shape[-1]
    assert width == int(idx.ne(-1).sum(-1).max())
    assert torch.equal(got, idx[:, :width])
    valid = got.ne(-1)
    assert torch.equal(inter["min_depth"][0].cpu()[valid], tmin[:, :width][valid])
    assert torch.equal(inter["max_depth"][0].cpu()[valid], tmax[:, :width][valid])


def test_large_ray_batch_properties(device):
    """configs[4] (ray-batch sweep): 2^17 rays in one launch on the configs[1] scene, checked through properties that do
    not need the oracle at that size: a ray's hit list does not depend on the batch it is in (bit-exact against a
    2048-ray launch of a subset), CSR offsets are consistent, samples are ordered along their ray (up to the reference's tail-loop quirk),
    compositing weights are a partition of unity, and nothing overflowed."""
    from proud_slam_b200 import scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    R = 1 << 17
    s, ms = util.build_scene("replica_20k")
    dec = [p.detach().to(device) for p in util.test_decoder(width=128, seed=1)]
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    frames = list(range(min(8, len(s.frames))))
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, frames, R // len(frames), seed=3)
    ro_d, rd_d = rays_o.reshape(-1, 3).to(device), rays_d.reshape(-1, 3).to(device)
    assert ro_d.shape[0] == R
    big = RenderPipeline(R, device, samples_per_ray=48)
    big.bind(ro_d, rd_d, msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
             target_rgb=rgb.reshape(-1, 3).to(device), target_depth=depth.reshape(-1).to(device), seed=5, forward_only=True)
    big.step()
    c = big.counts()                                         # raises on capacity / stack overflow
    assert c["R_h"] > R // 2 and c["n_samples"] > 20 * c["R_h"]
    # CSR consistency
    off = big.samp_off[:c["R_h"] + 1].cpu()
    assert int(off[0]) == 0 and int(off[-1]) == c["n_samples"] and bool((off[1:] >= off[:-1]).all())
    assert int((off[1:] - off[:-1]).max()) == c["S"]
    # samples: ordered along each ray, inside the trimmed range
    z = big.samp_z[:c["n_samples"]]
    ray_of = big.samp_ray[:c["n_samples"]].long()
    same = ray_of[1:] == ray_of[:-1]
    # (not ALL pairs: the reference's tail loop may close a ray with samples of a foreign voxel, SURVEY A-Q7)
    ordered = (z[1:][same] >= z[:-1][same]).float().mean()
    assert float(ordered) > 0.995, float(ordered)
    assert bool((ray_of[1:] >= ray_of[:-1]).all()) and float(z.min()) >= 0.0 and float(z.max()) <= 10.0
    # compositing weights: non-negative, sum to 1 on rays whose weights are not all masked out
    w = big.samp_w[:c["n_samples"]]
    sums = torch.zeros(c["R_h"], device=device).index_add_(0, ray_of, w)
    assert float(w.min()) >= 0.0
    assert bool(((sums - 1.0).abs() < 1e-4).logical_or(sums == 0.0).all()) and float((sums > 0).float().mean()) > 0.9
    out = big.ray_out[:c["R_h"]]
    assert bool(torch.isfinite(out).all())
    # a ray's hit list is independent of its batch: rerun a strided subset alone
    sub = torch.arange(0, R, R // 2048, device=device)[:2048]
    small = RenderPipeline(2048, device, samples_per_ray=48)
    small.bind(ro_d[sub].contiguous(), rd_d[sub].contiguous(), msd, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size,
               truncation=0.1, max_distance=10.0, target_rgb=rgb.reshape(-1, 3).to(device)[sub].contiguous(),
               target_depth=depth.reshape(-1).to(device)[sub].contiguous(), seed=5, forward_only=True)
    small.step()
    nb, ns = big.hit_count[:R][sub], small.hit_count[:2048]
    assert torch.equal(nb, ns)
    width = int(ns.max())
    for name in ("hit_idx", "hit_min", "hit_max"):
        a = getattr(big, name).view(-1, R)[:width][:, sub]
        b = getattr(small, name).view(-1, 2048)[:width]
        valid = torch.arange(width, device=device)[:, None] < ns[None, :]
        assert torch.equal(a[valid], b[valid]), name
