"""Our kernels 1-2 against the REFERENCE'S OWN CUDA kernels (third_party/sparse_voxels compiled
unmodified for sm_100 into oracle/_ref/grid.so by oracle/build_ref.py), on the same GPU, bit for bit.
This is the pin for hit lists and sample ids: the reference has no golden vectors and no CPU build
of these kernels, so the only ground truth is its own binary.  Also pins the C oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle import build_ref
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_grid():
    m = build_ref.load()
    if m is None:
        pytest.skip("oracle/_ref/grid.so not built (needs /root/reference at build time)")
    return m


@pytest.mark.parametrize("kind,frames,rays", [("tiny", 2, 333), ("replica_small", 4, 2048)])
def test_svo_intersect_equals_reference_binary(ref_grid, kind, frames, rays, device):
    """Through the reference's own wrapper shapes: G=256 groups, rays padded by repetition, octree
    expanded per group (voxel_helpers.py:113-135)."""
    from proud_slam_b200 import grid, scene as sc
    s, ms = util.build_scene(kind)
    rays_o, rays_d, _, _ = sc.sample_batch(s, list(range(frames)), rays, seed=9)
    R = rays_o.shape[1]
    G = 256
    K = -(-R // G)
    pad = G * K - R
    ro_ = torch.cat([rays_o[0], rays_o[0][:pad]], 0).reshape(G, K, 3).contiguous().to(device)
    rd_ = torch.cat([rays_d[0], rays_d[0][:pad]], 0).reshape(G, K, 3).contiguous().to(device)
    pts = ms["voxel_center_xyz"].detach()[None].expand(G, -1, 3).contiguous().to(device)
    ch = ms["voxel_structure"][None].expand(G, -1, 9).contiguous().to(device)
    want = ref_grid.svo_intersect(ro_, rd_, pts, ch, s.voxel_size, 50)
    got = grid.svo_intersect(ro_, rd_, pts, ch, s.voxel_size, 50)
    idx_ref = want[0].cpu()
    assert torch.equal(got[0].cpu(), idx_ref)
    hit = idx_ref >= 0   # the reference leaves depth slots beyond the hits at their zero initialisation
    assert torch.equal(got[1].cpu()[hit], want[1].cpu()[hit]) and torch.equal(got[2].cpu()[hit], want[2].cpu()[hit])
    assert torch.equal(got[1].cpu()[~hit], torch.zeros_like(got[1].cpu()[~hit]))
    assert int(hit.sum()) > R
    # and the C oracle, fed the device reciprocal, is the same function
    inv = util.device_rcp(rd_.reshape(-1, 3).cpu(), device).reshape(G, K, 3)
    oi, omin, omax = oracle.svo_intersect(ro_.cpu().numpy(), rd_.cpu().numpy(), pts.cpu().numpy(), ch.cpu().numpy(), s.voxel_size, 50, inv)
    assert np.array_equal(oi, idx_ref.numpy()) and np.array_equal(omin[hit.numpy()], want[1].cpu().numpy()[hit.numpy()])


@pytest.mark.parametrize("n_rays,P", [(1000, 9), (37, 5), (4000, 12)])
def test_inverse_cdf_sampling_equals_reference_binary(ref_grid, n_rays, P, device):
    from proud_slam_b200 import grid
    gen = torch.Generator().manual_seed(n_rays)
    G = 200
    n = -(-n_rays // G)
    cnt = torch.randint(1, P + 1, (G * n,), generator=gen)
    cnt[::17] = P     # rays that fill every slot: the foreign-voxel tail case (SURVEY A-Q7)
    seg = torch.rand(G * n, P, generator=gen) * 0.3 + 0.02
    gap = torch.rand(G * n, P, generator=gen) * 0.2
    start = torch.rand(G * n, 1, generator=gen) * 2 + 0.3
    tmin = start + torch.cumsum(seg + gap, 1) - seg
    tmax = tmin + seg
    valid = torch.arange(P)[None] < cnt[:, None]
    idx = torch.where(valid, torch.randint(0, 5000, (G * n, P), generator=gen), torch.full((G * n, P), -1)).int()
    tmin = torch.where(valid, tmin, torch.full_like(tmin, 10.0))
    tmax = torch.where(valid, tmax, torch.full_like(tmax, 10.0))
    d = (tmax - tmin).masked_fill(~valid, 0)
    probs = d / d.sum(-1, keepdim=True)
    steps = d.sum(-1) / 0.02
    M = int(steps.ceil().max()) + P
    noise = torch.rand(G * n, M, generator=gen).clamp(0.001, 0.999)
    shp = lambda t, last: t.reshape(G, n, last).contiguous().to(device)
    args = [shp(idx, P), shp(tmin, P), shp(tmax, P), shp(noise, M), shp(probs, P), steps.reshape(G, n).contiguous().to(device)]
    want = ref_grid.inverse_cdf_sampling(*args, -1.0)
    got = grid.inverse_cdf_sampling(*args, -1.0)
    for a, b, name in zip(got, want, ("sampled_idx", "sampled_depth", "sampled_dists")):
        assert torch.equal(a.cpu(), b.cpu()), name
    oi, od, os_ = oracle.inverse_cdf_sampling(*[a.cpu().numpy() for a in args], -1.0)
    assert np.array_equal(oi, want[0].cpu().numpy()) and np.array_equal(od, want[1].cpu().numpy())


def test_aabb_intersect_equals_reference_binary(ref_grid, device):
    from proud_slam_b200 import grid, scene as sc
    s, ms = util.build_scene("tiny")
    rays_o, rays_d, _, _ = sc.sample_batch(s, [0, 1], 200, seed=9)
    leaf = ms["voxel_structure"][:, 8] == 1
    pts = ms["voxel_center_xyz"].detach()[leaf][None].contiguous().to(device)
    want = ref_grid.aabb_intersect(rays_o.to(device), rays_d.to(device), pts, s.voxel_size, 60)
    got = grid.aabb_intersect(rays_o.to(device), rays_d.to(device), pts, s.voxel_size, 60)
    hit = want[0].cpu() >= 0
    assert torch.equal(got[0].cpu(), want[0].cpu())
    assert torch.equal(got[1].cpu()[hit], want[1].cpu()[hit]) and torch.equal(got[2].cpu()[hit], want[2].cpu()[hit])
