"""Multi-rank step on the device.  With one GPU the ranks are emulated as two pipelines whose raw loss
rows are exchanged by hand (no kernels wait on each other); the closed loss and the summed gradients
must equal ONE padded batch evaluated by the oracle on the GPU's own per-sample outputs."""
import pytest
import torch

from oracle import render_oracle as ro
from tests import util
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_two_shards_close_to_one_padded_batch(device):
    from proud_slam_b200 import parallel, scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("replica_small")
    dec = util.test_decoder(seed=1)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 300, seed=5)
    batch = [rays_o[0], rays_d[0], rgb[0], depth[0]]
    batch = [t[:599] for t in batch]                    # uneven shards: 300 + 299
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    fg = parallel.FlatGrads(msd["voxel_vertex_emb"], decd)     # both "ranks" accumulate into it = the all-reduce
    pipes, outs, shards = [], [], []
    rows = torch.zeros(2, 16, dtype=torch.float64, device=device)
    for r in range(2):
        sh = parallel.shard_rays(batch, r, 2)
        shards.append(sh)
        inv = util.device_rcp(sh[1], device)
        out = ro.render_rays(sh[0][None], sh[1][None], ms, dec, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0,
                             generator=torch.Generator().manual_seed(100 + r), inv_dir=inv)
        outs.append(out)
        noise = out["_dbg"]["noise"]
        pipe = RenderPipeline(sh[0].shape[0], device, samples_per_ray=96)
        pipe.bind(sh[0].to(device), sh[1].to(device), msd, decd, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size,
                  truncation=0.1, max_distance=10.0, target_rgb=sh[2].to(device), target_depth=sh[3].to(device),
                  noise=noise.reshape(-1, noise.shape[-1]).to(device).contiguous(), weights=cw, g_emb=fg.g_emb, g_dec=fg.g_dec,
                  grad_rays=True, defer_loss=True)
        pipe.sample()
        pipe.forward()
        rows[r].copy_(pipe.loss_raw)
        pipes.append(pipe)
    g_full = []
    for pipe, out, sh in zip(pipes, outs, shards):
        pipe.finalize_loss(rows)
        # = pipe.backward(), with the upstream gradient zeroed where a ReLU decision is within rounding of a tie (tests/util.py)
        pipe.stage(4)
        P = int(out["_dbg"]["sample_mask"].sum())
        near = util.relu_near_samples(out, sh[0][None], sh[1][None], ms, dec, s.voxel_size, util.RELU_MARGIN["f16"])
        g_full.append(pipe.samp_gout[:P].clone())
        assert float(near.float().mean()) < 0.02       # share of samples masked as ReLU near-ties
        pipe.samp_gout[:P][near.to(device)] = 0.0
        pipe.stage(5)
    torch.cuda.synchronize()
    assert pipes[0].losses() == pipes[1].losses()            # every rank closes the loss identically
    # one padded batch, oracle compositing on the GPU's own per-sample outputs (decision-proof, see test_gpu_pipeline)
    sdf_ps, rgb_ps, ress = [], [], []
    for pipe, out in zip(pipes, outs):
        P = pipe.counts()["n_samples"]
        so = pipe.samp_out[:P].cpu()
        sdf_p, rgb_p = so[:, 3].clone().requires_grad_(True), so[:, :3].clone().requires_grad_(True)
        smask, z = out["_dbg"]["sample_mask"], out["z_vals"]
        sdf = torch.ones_like(z).masked_scatter(smask, sdf_p)
        colour = z.new_zeros(*z.shape, 3).masked_scatter(smask.unsqueeze(-1).expand(*z.shape, 3), rgb_p)
        w, _ = ro.sdf2weights(sdf, z, smask.to(z.dtype), 0.1)
        ress.append({"weights": w, "color": torch.sum(w[..., None] * colour, -2), "depth": torch.sum(w * z, -1), "z_vals": z,
                     "sdf": sdf, "ray_mask": out["ray_mask"]})
        sdf_ps.append(sdf_p)
        rgb_ps.append(rgb_p)
    big = util.concat_outputs(ress)
    kw = {k: util.CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
    loss, parts = ro.criterion(big, (batch[2][None], batch[3][None]), **kw)
    loss.backward()
    l = pipes[0].losses()
    assert abs(l["loss"] - float(loss)) <= TOL * abs(float(loss))
    for k in ("color_loss", "depth_loss", "fs_loss", "sdf_loss"):
        assert abs(l[k] - float(parts[k])) <= TOL * max(abs(float(parts[k])), 1e-12), k
    for g, sdf_p, rgb_p in zip(g_full, sdf_ps, rgb_ps):
        g = g.cpu()
        assert rel_err(g[:, 3], sdf_p.grad) < TOL and rel_err(g[:, :3], rgb_p.grad) < TOL
    # summed parameter gradients = oracle field backward fed with each shard's upstream gradient
    tot = None
    for pipe, out, sh in zip(pipes, outs, shards):
        P = pipe.counts()["n_samples"]
        g = pipe.samp_gout[:P].cpu()
        rgb_o, sdf_o = util.oracle_field(out, sh[0][None], sh[1][None], ms, dec, s.voxel_size)
        grads = torch.autograd.grad((rgb_o * g[:, :3]).sum() + (sdf_o * g[:, 3]).sum(), [ms["voxel_vertex_emb"]] + list(dec))
        tot = grads if tot is None else [a + b for a, b in zip(tot, grads)]
    assert rel_err(fg.g_emb, tot[0]) < TOL
    for i in range(10):
        assert rel_err(fg.g_dec[i], tot[1 + i]) < TOL, f"decoder grad {i}"


def test_chunked_step_matches_separate_pipelines(device):
    """parallel.ChunkedStep (a batch larger than one launch's workspace, chunks as virtual ranks of one pipeline, each
    rendered twice) against the same chunks on separate pipelines that keep their state between the two passes."""
    from proud_slam_b200 import parallel, scene as sc
    from proud_slam_b200.pipeline import RenderPipeline
    s, ms = util.build_scene("tiny")
    dec = util.test_decoder(width=128, seed=1)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 300, seed=5)
    batch = [t[:599].to(device) for t in (rays_o[0], rays_d[0], rgb[0], depth[0])]
    msd = {k: v.detach().to(device) for k, v in ms.items()}
    decd = [p.detach().to(device) for p in dec]
    cw = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])
    chunks = [parallel.shard_rays(batch, k, 2) for k in range(2)]

    def bind(pipe, fg, k):
        sh = chunks[k]
        pipe.bind(sh[0], sh[1], msd, decd, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
                  target_rgb=sh[2], target_depth=sh[3], seed=50 + k, weights=cw, g_emb=fg.g_emb, g_dec=fg.g_dec, grad_rays=True,
                  defer_loss=True)

    # reference: one pipeline per chunk, state kept between the loss exchange and the backward
    fg_ref = parallel.FlatGrads(msd["voxel_vertex_emb"], decd)
    rows = torch.zeros(2, 16, dtype=torch.float64, device=device)
    pipes = []
    for k in range(2):
        pipe = RenderPipeline(300, device, samples_per_ray=96)
        bind(pipe, fg_ref, k)
        pipe.sample()
        pipe.forward()
        rows[k].copy_(pipe.loss_raw)
        pipes.append(pipe)
    for pipe in pipes:
        pipe.finalize_loss(rows)
        pipe.backward()
    # chunked: ONE pipeline, every chunk rendered twice
    fg = parallel.FlatGrads(msd["voxel_vertex_emb"], decd)
    one = RenderPipeline(300, device, samples_per_ray=96)
    ray_grads = {}
    step = parallel.ChunkedStep(one, fg, 2, lambda k: bind(one, fg, k),
                                after_backward=lambda k: ray_grads.__setitem__(k, one.g_rays_d[: chunks[k][0].shape[0]].clone()))
    step()
    torch.cuda.synchronize()
    assert one.losses() == pipes[1].losses()
    assert rel_err(fg.flat, fg_ref.flat) < 1e-5
    for k in range(2):
        assert rel_err(ray_grads[k], pipes[k].g_rays_d[: chunks[k][0].shape[0]]) < 1e-5
