"""The reference-facing Python surface (same names, arguments, tensors in and out as
src/variations/{voxel_helpers,render_helpers,nrgbd}.py and src/criterion.py) on the GPU."""
import types

import numpy as np
import pytest
import torch

from oracle import render_oracle as ro
from tests import util
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _setup(device, kind="tiny", rays=200, frames=(0, 1), width=128):
    from proud_slam_b200 import scene as sc
    s, ms = util.build_scene(kind)
    dec = util.test_decoder(width=width, seed=1)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, list(frames), rays, seed=5)
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    return s, ms, msd, dec, rays_o, rays_d, rgb, depth


def test_ray_intersect_vox_and_ray_sample_match_oracle(device):
    from proud_slam_b200.variations import voxel_helpers as vh
    s, ms, msd, dec, rays_o, rays_d, rgb, depth = _setup(device, "replica_small", 500)
    inv = util.device_rcp(rays_d.reshape(-1, 3), device)
    ref, ref_hits = ro.ray_intersect_vox(rays_o, rays_d, ms["voxel_center_xyz"].detach(), ms["voxel_structure"], s.voxel_size, 10, 10.0, inv_dir=inv)
    out, hits = vh.ray_intersect_vox(rays_o.to(device), rays_d.to(device), msd["voxel_center_xyz"], msd["voxel_structure"], s.voxel_size, 10, 10.0)
    assert torch.equal(hits.cpu(), ref_hits)
    for k in ref:
        assert torch.equal(out[k].cpu(), ref[k]), k
    # sampling on the hit rays, replaying the oracle's noise
    mask = ref_hits.view(-1)
    inter_ref = {k: v[0][mask] for k, v in ref.items()}
    smp_ref, noise = ro.ray_sample(inter_ref, 0.1 * s.voxel_size, generator=torch.Generator().manual_seed(3))
    inter = {k: v[0][mask.to(device)] for k, v in out.items()}
    smp = vh.ray_sample(inter, 0.1 * s.voxel_size, noise=noise.to(device))
    # ids bit-exact; depths only to rounding here because probs/steps come from torch.sum, whose
    # reduction order differs between CPU and CUDA (SURVEY A-Q10) -- the kernel itself is compared
    # bit for bit on identical inputs in test_gpu_grid.py
    assert torch.equal(smp["sampled_point_voxel_idx"].cpu(), smp_ref["sampled_point_voxel_idx"])
    for k in ("sampled_point_depth", "sampled_point_distance"):
        assert torch.allclose(smp[k].cpu(), smp_ref[k], rtol=1e-5, atol=1e-6), k
    assert "probs" in inter and "steps" in inter   # the reference adds them to the dict too


def test_get_features_vox_autograd(device):
    from proud_slam_b200.variations import render_helpers as rh
    s, ms = util.build_scene("tiny", emb_scale=0.5)
    leaf = torch.nonzero((ms["voxel_vertex_idx"] >= 0).all(-1)).view(-1)
    g = torch.Generator().manual_seed(0)
    vox = leaf[torch.randint(0, leaf.numel(), (3000,), generator=g)]
    xyz = (ms["voxel_center_xyz"].detach()[vox] + (torch.rand(3000, 3, generator=g) - 0.5) * s.voxel_size).requires_grad_(True)
    f = ro.get_features_vox(xyz, vox, ms, s.voxel_size)
    gf = torch.randn(3000, 16, generator=g)
    (f * gf).sum().backward()
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    msd["voxel_vertex_emb"].requires_grad_(True)
    xd = xyz.detach().to(device).requires_grad_(True)
    out = rh.get_features_vox({"sampled_point_xyz": xd, "sampled_point_voxel_idx": vox.to(device),
                               "sampled_point_distance": torch.zeros(3000, device=device)}, msd, s.voxel_size)
    assert set(out) == {"dists", "emb"}
    (out["emb"] * gf.to(device)).sum().backward()
    assert rel_err(out["emb"], f.detach()) < TOL
    assert rel_err(xd.grad, xyz.grad) < TOL
    assert rel_err(msd["voxel_vertex_emb"].grad, ms["voxel_vertex_emb"].grad) < TOL


@pytest.mark.parametrize("width", [128, 256])
def test_decoder_module_is_state_dict_compatible(width, device):
    from proud_slam_b200.variations.nrgbd import Decoder
    torch.manual_seed(7 + width)
    dec = Decoder(depth=2, width=width, in_dim=16, skips=[], embedder="none").to(device)
    assert set(dec.state_dict()) == {f"pts_linears.{i}.{p}" for i in (0, 1) for p in ("weight", "bias")} | {
        "sdf_out.weight", "sdf_out.bias", "color_out.0.weight", "color_out.0.bias", "color_out.2.weight", "color_out.2.bias"}
    x = (torch.randn(700, 16, generator=torch.Generator().manual_seed(1)) * 0.05)
    params = [p.detach().cpu().clone().requires_grad_(True) for p in dec.param_list()]
    xc = x.clone().requires_grad_(True)
    rgb, sdf = ro.decoder_forward(params, xc)
    # rows with a ReLU decision within rounding of a tie get zero weight on both sides (tests/util.py: RELU_MARGIN)
    keep = (~util.decoder_near(params, x, 4e-6)).float()
    assert float(keep.mean()) > 0.9
    ((rgb.sum(-1) * 0.3 + sdf * sdf) * keep).sum().backward()
    xd = x.to(device).requires_grad_(True)
    out = dec({"emb": xd})
    assert set(out) == {"color", "sdf"} and out["color"].shape == (700, 3)
    ((out["color"].sum(-1) * 0.3 + out["sdf"] * out["sdf"]) * keep.to(device)).sum().backward()
    assert rel_err(out["color"], rgb.detach()) < TOL and rel_err(out["sdf"], sdf.detach()) < TOL
    assert rel_err(xd.grad, xc.grad) < TOL
    for p, q in zip(dec.param_list(), params):
        assert rel_err(p.grad, q.grad) < TOL
    vals = dec.get_values(xd)
    assert vals.shape == (700, 4) and torch.equal(vals[:, 3], dec.get_sdf({"emb": xd}))
    with pytest.raises(NotImplementedError):
        Decoder()   # the reference's defaults (depth 8, nerf embedder) are not on the SLAM path


def test_render_rays_with_criterion_equals_fused_step(device):
    """Modular route (render_rays -> Criterion -> autograd) and fused route (pslam_render_step) agree."""
    from proud_slam_b200.criterion import Criterion
    from proud_slam_b200.pipeline import RenderPipeline
    from proud_slam_b200.variations import render_helpers as rh
    s, ms, msd, dec, rays_o, rays_d, rgb, depth = _setup(device, "replica_small", 400)
    decd = [p.detach().to(device).requires_grad_(True) for p in dec]
    msd["voxel_vertex_emb"].requires_grad_(True)
    ro_d, rd_d = rays_o.to(device).requires_grad_(True), rays_d.to(device).requires_grad_(True)
    args = types.SimpleNamespace(criteria=dict(rgb_weight=0.5, depth_weight=1.0, sdf_weight=5000.0, fs_weight=10.0, sdf_truncation=0.1),
                                 data_specs=dict(max_depth=10.0))
    crit = Criterion(args)
    out = rh.render_rays(ro_d, rd_d, msd, decd, None, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0, seed=77, return_raw=True)
    assert set(out) == {"weights", "color", "depth", "z_vals", "sdf", "ray_mask", "raw"}
    loss, parts = crit(out, (rgb.to(device), depth.to(device)))
    loss.backward()
    # fused
    R = rays_o.shape[1]
    pipe = RenderPipeline(R, device)
    g_emb = torch.zeros_like(msd["voxel_vertex_emb"])
    g_dec = [torch.zeros_like(p) for p in decd]
    pipe.bind(rays_o.to(device), rays_d.to(device), {k: v.detach() for k, v in msd.items()}, [p.detach() for p in decd],
              voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
              target_rgb=rgb.to(device), target_depth=depth.to(device), seed=77, weights=crit.weights(), g_emb=g_emb, g_dec=g_dec,
              grad_rays=True)
    pipe.step()
    l = pipe.losses()
    assert abs(l["loss"] - float(loss)) < 1e-5 * abs(float(loss))
    for k in ("color_loss", "depth_loss", "fs_loss", "sdf_loss"):
        assert abs(l[k] - parts[k]) <= 1e-5 * max(abs(parts[k]), 1e-12)
    assert rel_err(msd["voxel_vertex_emb"].grad, g_emb) < 1e-5
    for a, b in zip(decd, g_dec):
        assert rel_err(a.grad, b) < 1e-5
    assert rel_err(ro_d.grad.view(-1, 3), pipe.g_rays_o[:R]) < 1e-5
    assert rel_err(rd_d.grad.view(-1, 3), pipe.g_rays_d[:R]) < 1e-5


def test_render_rays_edge_cases(device):
    from proud_slam_b200.variations import render_helpers as rh
    s, ms, msd, dec, rays_o, rays_d, rgb, depth = _setup(device)
    decd = [p.detach().to(device) for p in dec]
    away = rays_o.clone() + 500.0   # rays that never meet the map: the reference asserts (render_helpers.py:388)
    with pytest.raises(AssertionError):
        rh.render_rays(away.to(device), rays_d.to(device), msd, decd, None, 0.02, s.voxel_size, 0.1, 10, 10.0)
    out = rh.render_rays(rays_o.to(device), rays_d.to(device), msd, decd, None, 0.02, s.voxel_size, 0.1, 10, 10.0)
    assert out["raw"] is None and out["ray_mask"].shape == (1, rays_o.shape[1])
    Rh = int(out["ray_mask"].sum())
    assert out["color"].shape == (Rh, 3) and out["depth"].shape == (Rh,) and out["weights"].shape == out["z_vals"].shape == out["sdf"].shape
    assert torch.all(out["sdf"][out["z_vals"] == 10.0] == 1.0)            # pads: z = 10, sdf = 1
    assert torch.allclose(out["weights"].sum(-1), torch.ones(Rh, device=device), atol=1e-4)


def test_tracking_and_mapping_loops_run_and_improve(device):
    """bundle_adjust_frames lowers the loss on a fixed batch; track_frame pulls a perturbed pose back."""
    from proud_slam_b200 import scene as sc, svo
    from proud_slam_b200.criterion import Criterion
    from proud_slam_b200.variations import render_helpers as rh
    from proud_slam_b200.variations.nrgbd import Decoder
    torch.manual_seed(0)
    s = sc.make_scene("tiny")
    tree = svo.Octree()
    tree.init(s.grid_dim, 16, s.voxel_size, 8)
    tree.insert(torch.from_numpy(s.voxels))
    ms = svo.build_map_states(tree, s.voxel_size, num_embeddings=2048, device=device, seed=0)
    ms["voxel_vertex_emb"].requires_grad_(True)
    dec = Decoder(depth=2, width=128, in_dim=16, skips=[], embedder="none").to(device)
    args = types.SimpleNamespace(criteria=dict(rgb_weight=0.5, depth_weight=1.0, sdf_weight=5000.0, fs_weight=10.0, sdf_truncation=0.1),
                                 data_specs=dict(max_depth=10.0))
    crit = Criterion(args)
    frames = [util.TestFrame(s, s.frames[i], stamp=i, device=device, seed=i) for i in range(2)]
    embed_optim = torch.optim.Adam([ms["voxel_vertex_emb"]], lr=5e-3)
    model_optim = torch.optim.Adam(dec.parameters(), lr=5e-3)

    def eval_loss():
        g = torch.Generator().manual_seed(123)
        f = frames[0]
        idx = torch.randperm(f.rays_d.shape[0], generator=g)[:512].to(device)
        pose = f.get_pose().detach()
        rd = (f.rays_d[idx] @ pose[:3, :3].t()).unsqueeze(0)
        ro_ = pose[:3, 3].view(1, 1, 3).expand_as(rd).contiguous()
        with torch.no_grad():
            out = rh.render_rays(ro_, rd, ms, dec, None, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0, seed=5)
            loss, _ = crit(out, (f.rgb[idx].unsqueeze(0), f.depth[idx].unsqueeze(0)))
        return float(loss)

    before = eval_loss()
    rh.bundle_adjust_frames(frames, ms, dec, None, crit, s.voxel_size, 0.1 * s.voxel_size, N_rays=512, num_iterations=100,
                            truncation=0.1, max_voxel_hit=10, max_distance=10.0, embed_optim=embed_optim, model_optim=model_optim,
                            update_pose=True)
    after = eval_loss()
    assert after < 0.7 * before, (before, after)
    assert float(embed_optim.state[ms["voxel_vertex_emb"]]["step"]) == 100.0      # the fused Adam advanced the caller's optimizer state
    # the reference's own route for ray selection (frame.sample_rays + mask gathers) still works and keeps improving
    rh.bundle_adjust_frames(frames, ms, dec, None, crit, s.voxel_size, 0.1 * s.voxel_size, N_rays=512, num_iterations=3,
                            truncation=0.1, max_voxel_hit=10, max_distance=10.0, embed_optim=embed_optim, model_optim=model_optim,
                            update_pose=False, device_sampling=False)
    assert eval_loss() < 0.7 * before
    # tracking: start 3 cm off, expect to end closer to the true pose
    true_t = frames[1].pose.translation().detach().clone()
    start = util.TestFrame(s, s.frames[1], stamp=1, device=device, perturb=(0.03, -0.02, 0.02), seed=9)
    e0 = float((start.pose.translation().detach() - true_t).norm())
    pose, optim, hit_mask = rh.track_frame(start.pose, start, {k: v.detach() for k, v in ms.items()}, dec, None, crit, s.voxel_size,
                                           N_rays=1024, step_size=0.1 * s.voxel_size, num_iterations=40, truncation=0.1,
                                           learning_rate=0.01, max_voxel_hit=10, max_distance=10.0, depth_variance=True)
    e1 = float((pose.translation().detach() - true_t).norm())
    drift = float((true_t.cpu() - s.frames[1].pose[:3, 3]).norm())
    print(f"BA loss {before:.4f} -> {after:.4f}; BA moved frame 1 by {drift:.4f} m; tracking error {e0:.4f} -> {e1:.4f} m")
    assert hit_mask.shape == (1024,) and hit_mask.dtype == torch.bool
    assert e1 < e0, (e0, e1)
    # the same optimisation as one CUDA graph per iteration
    start2 = util.TestFrame(s, s.frames[1], stamp=1, device=device, perturb=(0.03, -0.02, 0.02), seed=9)
    tracker = rh.GraphTracker(start2.rays_d.shape[0], {k: v.detach() for k, v in ms.items()}, dec, crit, s.voxel_size, N_rays=1024,
                              step_size=0.1 * s.voxel_size, truncation=0.1, learning_rate=0.01, max_distance=10.0, depth_variance=True,
                              device=device)
    pose2, _, hm2 = tracker.track(start2.pose, start2, 40)
    e2 = float((pose2.translation().detach() - true_t).norm())
    assert e2 < e0, (e0, e2)
    pose3, _, _ = tracker.track(start2.pose, start2, 40)      # the captured graph is reused for the next frame
    assert float((pose3.translation().detach() - true_t).norm()) < e0


def test_fused_pose_kernels_match_torch_autograd_and_adam(device):
    """pslam_track_assemble / pslam_track_pose_step (csrc/pose.cu) against the torch route of the reference
    (se3pose.OptimizablePose: Rodrigues with the 11-term series, autograd, torch.optim.Adam) over a few steps."""
    from proud_slam_b200 import _lib
    from proud_slam_b200.se3pose import OptimizablePose
    lib = _lib.lib()
    g = torch.Generator().manual_seed(3)
    HW, N = 5000, 1024
    dirs = torch.nn.functional.normalize(torch.randn(HW, 3, generator=g), dim=-1).to(device)
    rgb_all, depth_all = torch.rand(HW, 3, generator=g).to(device), torch.rand(HW, generator=g).to(device)
    init = torch.tensor([0.3, -0.2, 1.1, 0.4, -0.7, 0.25])
    ref = OptimizablePose(init).to(device)
    opt = torch.optim.Adam(ref.parameters(), lr=0.01, capturable=True)
    pose = init.clone().to(device)
    m, v, step = torch.zeros(6, device=device), torch.zeros(6, device=device), torch.zeros((), device=device)
    pose_h, m_h, v_h = init.clone().to(device), torch.zeros(6, device=device), torch.zeros(6, device=device)
    rays_o, rays_d = torch.empty(N, 3, device=device), torch.empty(N, 3, device=device)
    rgb, depth = torch.empty(N, 3, device=device), torch.empty(N, device=device)
    grad = torch.empty(6, device=device)
    for it in range(4):
        idx = torch.randint(0, HW, (N,), generator=g).to(device)
        g_o = (torch.randn(N, 3, generator=g) * 1e-3).to(device)
        g_d = (torch.randn(N, 3, generator=g) * 1e-3).to(device)
        # torch route
        rd = dirs[idx] @ ref.rotation().transpose(-1, -2)
        ro_ = ref.translation().reshape(1, -1).expand_as(rd)
        ro_ref = ro_.detach().clone()                       # (a view of the parameter: Adam updates it in place below)
        opt.zero_grad()
        torch.autograd.backward([ro_, rd], [g_o, g_d])
        ref_grad = ref.data.grad.clone()
        opt.step()
        # fused route
        _lib.check(lib.pslam_track_assemble(N, _lib.ptr(pose), _lib.ptr(idx), _lib.ptr(dirs), _lib.ptr(rgb_all), _lib.ptr(depth_all),
                                            _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(rgb), _lib.ptr(depth), _lib.stream_ptr(device)), "assemble")
        if it == 0:
            assert rel_err(rays_d, rd.detach()) < 1e-6 and rel_err(rays_o, ro_ref) < 1e-6
            assert torch.equal(rgb, rgb_all[idx]) and torch.equal(depth, depth_all[idx])
        _lib.check(lib.pslam_track_pose_step(N, _lib.ptr(pose), _lib.ptr(idx), _lib.ptr(dirs), _lib.ptr(g_o), _lib.ptr(g_d), _lib.ptr(m),
                                             _lib.ptr(v), _lib.ptr(step), 0.0, 0.01, 0.9, 0.999, 1e-8, _lib.ptr(grad), _lib.stream_ptr(device)), "pose step")
        # host-side step count (non-capturable Adam): step == NULL, the new count passed by value
        _lib.check(lib.pslam_track_pose_step(N, _lib.ptr(pose_h), _lib.ptr(idx), _lib.ptr(dirs), _lib.ptr(g_o), _lib.ptr(g_d), _lib.ptr(m_h),
                                             _lib.ptr(v_h), None, float(it + 1), 0.01, 0.9, 0.999, 1e-8, None, _lib.stream_ptr(device)), "pose step")
        torch.cuda.synchronize()
        assert rel_err(grad, ref_grad) < 1e-5, it
        assert rel_err(pose, ref.data.detach()) < 1e-5, it
        assert torch.equal(pose_h, pose) and torch.equal(m_h, m) and torch.equal(v_h, v)
    st = opt.state[ref.data]
    assert rel_err(m, st["exp_avg"]) < 1e-5 and rel_err(v, st["exp_avg_sq"]) < 1e-5 and float(step) == float(st["step"])


@pytest.mark.parametrize("hw,n", [(1200 * 680, 1024), (4800, 4800), (5, 3), (1 << 20, 4096)])
def test_sample_pixels_distinct_uniform(hw, n, device):
    """pslam_sample_pixels: n distinct indices in [0, hw) (sampling without replacement like frame.sample_rays), a fresh
    set per value of the device-side counter, spread over the whole range."""
    from proud_slam_b200 import _lib
    lib = _lib.lib()
    idx = torch.empty(n, dtype=torch.int64, device=device)
    counter = torch.zeros(1, dtype=torch.int64, device=device)
    sets = []
    for it in range(3):
        counter.fill_(it)
        _lib.check(lib.pslam_sample_pixels(n, hw, 12345, _lib.ptr(counter), _lib.ptr(idx), _lib.stream_ptr(device)), "sample_pixels")
        torch.cuda.synchronize()
        v = idx.cpu()
        assert int(v.min()) >= 0 and int(v.max()) < hw
        assert v.unique().numel() == n                         # no replacement
        sets.append(v)
    if n < hw:
        assert not torch.equal(sets[0], sets[1]) and not torch.equal(sets[1], sets[2])
    if n >= 1024 and n < hw:
        assert abs(float(sets[0].double().mean()) / hw - 0.5) < 0.05      # uniform over the range
    assert lib.pslam_sample_pixels(10, 5, 0, None, _lib.ptr(idx), _lib.stream_ptr(device)) != 0   # more pixels than there are


def test_loop_flags_are_sticky_across_steps(device):
    """A loop of pslam_render_step calls is checked ONCE at its end (RenderPipeline.check): each step folds the previous
    step's overflow flags into counters[PSLAM_C_STICKY] before it clears the per-step counters, so a dropped batch tail
    or a step without hit rays several iterations ago is still reported (the reference asserts / syncs in every call,
    render_helpers.py:388)."""
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms, msd, dec, rays_o, rays_d, rgb, depth = _setup(device, "replica_small", 300)
    decd = [p.detach().to(device) for p in dec]
    msd = {k: v.detach() for k, v in msd.items()}
    kw = dict(voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0, target_rgb=rgb.to(device),
              target_depth=depth.to(device), grad_rays=True)
    R = rays_o.shape[1]
    # (1) a clean loop reports nothing
    pipe = RenderPipeline(R, device, samples_per_ray=96)
    pipe.bind(rays_o.to(device), rays_d.to(device), msd, decd, seed=1, **kw)
    for _ in range(3):
        pipe.step()
    assert pipe.check() is None
    # (2) capacity exceeded in the FIRST of three steps (the later ones fit: fewer rays)
    small = RenderPipeline(R, device, samples_per_ray=8)
    small.bind(rays_o.to(device), rays_d.to(device), msd, decd, seed=1, **kw)
    small.step()
    few = slice(0, 20)
    small.bind(rays_o[:, few].to(device).contiguous(), rays_d[:, few].to(device).contiguous(), msd, decd, seed=2,
               **{**kw, "target_rgb": rgb[:, few].to(device).contiguous(), "target_depth": depth[:, few].to(device).contiguous()})
    small.step()
    small.step()
    with pytest.raises(RuntimeError, match="capacity"):
        small.check()
    assert small.check() is None                                   # cleared by the check
    # (3) a step whose rays all miss the map, followed by a good one
    pipe.bind((rays_o + 500.0).to(device), rays_d.to(device), msd, decd, seed=3, **kw)
    pipe.step()
    pipe.bind(rays_o.to(device), rays_d.to(device), msd, decd, seed=4, **kw)
    pipe.step()
    with pytest.raises(AssertionError, match="no ray hits"):
        pipe.check()


@pytest.mark.parametrize("capturable", [True, False])
def test_fused_adam_matches_torch_adam(capturable, device):
    """pslam_adam_step (csrc/optim.cu) on the optimizers' own state against torch.optim.Adam over 6 steps: an embedding
    table whose rows receive gradients sparsely (different rows per step) plus the ten decoder tensors, two learning rates."""
    from proud_slam_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    emb0 = torch.randn(3000, 16, generator=g) * 0.01
    dec0 = [torch.randn(*s, generator=g) * 0.1 for s in [(128, 16), (128,), (128, 128), (128,), (129, 128), (129,), (128, 144), (128,), (3, 128), (3,)]]

    def make():
        emb = emb0.clone().to(device).requires_grad_(True)
        dec = [p.clone().to(device).requires_grad_(True) for p in dec0]
        return emb, dec, torch.optim.Adam([emb], lr=5e-3, capturable=capturable), torch.optim.Adam(dec, lr=1e-3, capturable=capturable)

    emb_a, dec_a, oe_a, od_a = make()
    emb_b, dec_b, oe_b, od_b = make()
    fused = FusedAdam([oe_b, od_b], row_tensors=[emb_b])
    assert fused.fused
    for it in range(6):
        rows = torch.randperm(3000, generator=g)[:300]
        ge = torch.zeros(3000, 16)
        ge[rows] = torch.randn(300, 16, generator=g)
        gd = [torch.randn(*p.shape, generator=g) for p in dec0]
        emb_a.grad = ge.to(device)
        for p, q in zip(dec_a, gd):
            p.grad = q.to(device)
        oe_a.step(); od_a.step()
        gb = {emb_b: ge.to(device).contiguous()}
        gb.update({p: q.to(device).contiguous() for p, q in zip(dec_b, gd)})
        assert fused.step(grads=gb, zero_grad=True)
        assert all(float(t.abs().max()) == 0.0 for t in gb.values())          # cleared in the same pass
    torch.cuda.synchronize()
    # (the update itself is compared, not the parameter it is added to: 1e-4 of the largest total change)
    assert util.elem_err(emb_b.detach() - emb0.to(device), emb_a.detach() - emb0.to(device)) < 1e-4
    for p, q, p0 in zip(dec_b, dec_a, dec0):
        assert rel_err(p.detach() - p0.to(device), q.detach() - p0.to(device)) < 1e-4
    # the state the torch optimizer owns was advanced: its own next step continues from it
    assert float(oe_b.state[emb_b]["step"]) == 6.0
    untouched = (oe_a.state[emb_a]["exp_avg_sq"].amax(1) == 0)
    assert torch.equal(emb_b[untouched], emb0.to(device)[untouched])            # never-touched rows: bit-identical, skipped


def test_get_scores_eval_points_and_shared_map(device):
    """Meshing queries (render_helpers.py:243-328) = trilinear lookup + decoder on a lattice inside the surface voxels,
    against the oracle; and the device-side map hand-off (mapping.py:236-247 / tracking.py:116-125 without the host trip)."""
    from proud_slam_b200.variations import render_helpers as rh
    s, ms, msd, dec, rays_o, rays_d, rgb, depth = _setup(device, "tiny", 64)
    surf = torch.nonzero((ms["voxel_vertex_idx"] >= 0).all(-1)).view(-1)
    sub = {"voxel_vertex_idx": msd["voxel_vertex_idx"][surf.to(device)], "voxel_center_xyz": msd["voxel_center_xyz"][surf.to(device)],
           "voxel_vertex_emb": msd["voxel_vertex_emb"]}
    decd = [p.detach().to(device) for p in dec]
    res = 4
    scores = rh.get_scores(decd, sub, s.voxel_size, bits=res)
    assert scores.shape == (surf.numel(), res, res, res, 4) and not scores.is_cuda
    lin = torch.linspace(-0.5, 0.5, res)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3) * s.voxel_size
    xyz = (grid + ms["voxel_center_xyz"].detach()[surf].unsqueeze(1)).reshape(-1, 3)
    idx = surf.repeat_interleave(res ** 3)
    f = ro.get_features_vox(xyz, idx, ms, s.voxel_size)
    c, sd = ro.decoder_forward(dec, f)
    ref = torch.cat([c, sd[:, None]], 1).detach().view(-1, res, res, res, 4)
    assert rel_err(scores, ref) < 1e-4
    cols = rh.eval_points(decd, msd, xyz[:1000].to(device), idx[:1000].to(device), s.voxel_size)
    assert cols.shape == (1000, 3) and rel_err(cols, c.detach()[:1000]) < 1e-4
    assert rh.eval_points(decd, msd, xyz[:0].to(device), idx[:0].to(device), s.voxel_size) is None
    # hand-off: two publishes, the reader always sees a complete version; a later edit of the source does not leak in
    shared = rh.SharedMap(device)
    assert shared.acquire() is None
    src = {k: v.clone() for k, v in msd.items()}
    shared.publish(src, decd)
    m1, d1, v1 = shared.acquire()
    src["voxel_vertex_emb"].add_(1.0)
    assert v1 == 1 and torch.equal(m1["voxel_vertex_emb"], msd["voxel_vertex_emb"].detach()) and torch.equal(d1[0], decd[0])
    shared.publish(src, decd)
    m2, d2, v2 = shared.acquire()
    assert v2 == 2 and torch.equal(m2["voxel_vertex_emb"], src["voxel_vertex_emb"]) and m2["voxel_vertex_emb"].data_ptr() != m1["voxel_vertex_emb"].data_ptr()
    assert torch.equal(m1["voxel_vertex_emb"], msd["voxel_vertex_emb"].detach())      # the first version is still intact
