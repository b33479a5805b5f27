"""N > 1 host logic on CPU: 2 gloo processes shard the rays, exchange raw loss sums, close the loss
identically, and sum-all-reduce one flat gradient buffer.  The kernels are replaced by the oracle
(FakePipe) so this checks proud_slam_b200.parallel and the closure algebra (S*x0 - x1 pads),
not the CUDA code (tests/test_gpu_parallel.py does that on the device)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util

WEIGHTS = (util.CRIT["rgb_weight"], util.CRIT["depth_weight"], util.CRIT["fs_weight"], util.CRIT["sdf_weight"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakePipe:
    """CPU stand-in for RenderPipeline(defer_loss=True) built on the oracle."""

    def __init__(self, shard, ms, dec, voxel_size, seed):
        self.rays_o, self.rays_d, self.rgb, self.depth = shard
        self.ms, self.dec, self.vs, self.seed = ms, dec, voxel_size, seed
        self.loss_raw = torch.zeros(16, dtype=torch.float64)
        self.grads = None

    def sample(self):
        pass

    def forward(self):
        from oracle import render_oracle as ro
        self.out = ro.render_rays(self.rays_o[None], self.rays_d[None], self.ms, self.dec, 0.1 * self.vs, self.vs, 0.1, 10, 10.0,
                                  generator=torch.Generator().manual_seed(self.seed))
        self.loss_raw.copy_(util.raw_loss_sums(self.out, self.rgb, self.depth))

    def finalize_loss(self, rows):
        self.parts, self.coef = util.close_loss(rows, WEIGHTS)

    def backward(self):
        o, c, tau = self.out, self.coef, 0.1
        mask = o["_dbg"]["sample_mask"]
        rm = o["ray_mask"].view(-1)
        gt, gc = self.depth[rm], self.rgb[rm]
        d = gt[:, None]
        z, s = o["z_vals"], o["sdf"]
        front = ((z < d - tau) & mask).float()
        sm = ((~(z < d - tau)) & (~(z > d + tau)) & (d > 0) & (d < 10.0) & mask).float()
        valid = ((gt > 0.01) & (gt < 10.0)).float()
        obj = (c["color"] * (gc - o["color"]).abs().sum() + c["depth"] * ((gt - o["depth"]).abs() * valid).sum()
               + c["fs"] * (((s - 1.0) ** 2) * front).sum() + c["sdf"] * (((z + s * tau - d) ** 2) * sm).sum())
        params = [self.ms["voxel_vertex_emb"]] + list(self.dec)
        g = torch.autograd.grad(obj, params)
        self.grads.g_emb.copy_(g[0])
        for dst, src in zip(self.grads.g_dec, g[1:]):
            dst.copy_(src)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import render_oracle as ro
        from proud_slam_b200 import parallel, scene as sc
        s, ms = util.build_scene("tiny")
        dec = ro.decoder_params(seed=1)
        rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 150, seed=5)
        batch = [rays_o[0], rays_d[0], rgb[0], depth[0]]
        # uneven shards: 2 ranks over 300 rays -> (150,150); then make it uneven by dropping a ray
        batch = [t[:299] for t in batch]
        shard = parallel.shard_rays(batch, rank, world)
        pipe = FakePipe(shard, ms, dec, s.voxel_size, seed=100 + rank)
        grads = parallel.FlatGrads(ms["voxel_vertex_emb"], dec)
        pipe.grads = grads
        step = parallel.DataParallelStep(pipe, grads)
        step()
        # the same thing as ONE padded batch through the oracle's Criterion
        outs = []
        for r in range(world):
            sh = parallel.shard_rays(batch, r, world)
            outs.append(ro.render_rays(sh[0][None], sh[1][None], ms, dec, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0,
                                       generator=torch.Generator().manual_seed(100 + r)))
        big = util.concat_outputs(outs)
        kw = {k: util.CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
        loss, parts = ro.criterion(big, (batch[2][None], batch[3][None]), **kw)
        ref = torch.autograd.grad(loss, [ms["voxel_vertex_emb"]] + list(dec))
        ok = abs(pipe.parts["loss"] - float(loss)) < 1e-5 * abs(float(loss))
        for k in ("color_loss", "depth_loss", "fs_loss", "sdf_loss"):
            ok = ok and abs(pipe.parts[k] - float(parts[k])) <= 1e-5 * max(abs(float(parts[k])), 1e-12)
        errs = [util.rel_err(grads.g_emb, ref[0])] + [util.rel_err(a, b) for a, b in zip(grads.g_dec, ref[1:])]
        q.put((rank, bool(ok), max(errs), outs[0]["z_vals"].shape[1], outs[1]["z_vals"].shape[1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_one_padded_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0, "worker failed"
    res = [q.get(timeout=10) for _ in procs]
    for rank, ok, err, s0, s1 in res:
        assert ok, f"rank {rank}: loss closure differs from the single padded batch"
        assert err < 1e-4, f"rank {rank}: all-reduced gradient differs ({err})"


class AccumulatingFakePipe(FakePipe):
    """... whose backward ADDS into the gradient buffers like the CUDA kernels do, and whose ray shard can be re-bound."""

    def bind(self, shard, seed):
        self.rays_o, self.rays_d, self.rgb, self.depth = shard
        self.seed = seed

    def backward(self):
        keep = parallel_clone(self.grads)
        super().backward()
        self.grads.g_emb.add_(keep[0])
        for dst, src in zip(self.grads.g_dec, keep[1:]):
            dst.add_(src)


def parallel_clone(grads):
    return [grads.g_emb.clone()] + [g.clone() for g in grads.g_dec]


def test_chunked_step_equals_one_padded_batch():
    """parallel.ChunkedStep on one device (no process group): two chunks as virtual ranks == one padded batch."""
    from oracle import render_oracle as ro
    from proud_slam_b200 import parallel, scene as sc
    torch.set_num_threads(2)
    s, ms = util.build_scene("tiny")
    dec = ro.decoder_params(seed=1)
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 150, seed=5)
    batch = [t[:299] for t in (rays_o[0], rays_d[0], rgb[0], depth[0])]
    shards = [parallel.shard_rays(batch, k, 2) for k in range(2)]
    pipe = AccumulatingFakePipe(shards[0], ms, dec, s.voxel_size, seed=100)
    grads = parallel.FlatGrads(ms["voxel_vertex_emb"], dec)
    pipe.grads = grads
    seen = []
    step = parallel.ChunkedStep(pipe, grads, 2, lambda k: pipe.bind(shards[k], 100 + k), after_backward=seen.append)
    step()
    assert seen == [0, 1]
    outs = [ro.render_rays(sh[0][None], sh[1][None], ms, dec, 0.1 * s.voxel_size, s.voxel_size, 0.1, 10, 10.0,
                           generator=torch.Generator().manual_seed(100 + k)) for k, sh in enumerate(shards)]
    big = util.concat_outputs(outs)
    kw = {k: util.CRIT[k] for k in ("rgb_weight", "depth_weight", "sdf_weight", "fs_weight", "truncation", "max_depth")}
    loss, parts = ro.criterion(big, (batch[2][None], batch[3][None]), **kw)
    ref = torch.autograd.grad(loss, [ms["voxel_vertex_emb"]] + list(dec))
    assert abs(pipe.parts["loss"] - float(loss)) < 1e-5 * abs(float(loss))
    assert util.rel_err(grads.g_emb, ref[0]) < 1e-4
    for a, b in zip(grads.g_dec, ref[1:]):
        assert util.rel_err(a, b) < 1e-4


def test_shard_bounds_cover_everything():
    from proud_slam_b200.parallel import shard_bounds, shard_keyframes
    for n in (1, 7, 8, 8192, 8191):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert sum((shard_keyframes(list(range(24)), r, 8) for r in range(8)), []) == list(range(24))
