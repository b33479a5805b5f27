"""`grid` drop-in (kernels 1-2 in the reference's own layouts) against the C oracle: bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from tests import util

pytestmark = pytest.mark.gpu


def _scene_rays(kind, frames, rays, device):
    from proud_slam_b200 import scene as sc
    s, ms = util.build_scene(kind)
    rays_o, rays_d, _, _ = sc.sample_batch(s, list(range(frames)), rays, seed=9)
    return s, ms, rays_o, rays_d


@pytest.mark.parametrize("kind,frames,rays,B", [("tiny", 2, 333, 1), ("replica_small", 4, 1024, 256), ("replica_small", 1, 100, 7)])
def test_svo_intersect_bit_exact(kind, frames, rays, B, device):
    from proud_slam_b200 import grid
    s, ms, rays_o, rays_d = _scene_rays(kind, frames, rays, device)
    R = rays_o.shape[1]
    K = -(-R // B)
    pad = B * K - R   # the reference wrapper repeats leading rays (voxel_helpers.py:118-122)
    ro_ = torch.cat([rays_o[0], rays_o[0][:pad]], 0).reshape(B, K, 3).contiguous()
    rd_ = torch.cat([rays_d[0], rays_d[0][:pad]], 0).reshape(B, K, 3).contiguous()
    pts = ms["voxel_center_xyz"].detach()[None].expand(B, -1, 3).contiguous()
    ch = ms["voxel_structure"][None].expand(B, -1, 9).contiguous()
    idx, tmin, tmax = grid.svo_intersect(ro_.to(device), rd_.to(device), pts.to(device), ch.to(device), s.voxel_size, 50)
    inv = util.device_rcp(rd_.reshape(-1, 3), device).reshape(B, K, 3)
    oi, omin, omax = oracle.svo_intersect(ro_.numpy(), rd_.numpy(), pts.numpy(), ch.numpy(), s.voxel_size, 50, inv)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(tmin.cpu().numpy(), omin)
    assert np.array_equal(tmax.cpu().numpy(), omax)
    assert (oi >= 0).any()


def test_aabb_intersect_bit_exact_and_agrees_with_svo(device):
    """Independent cross-check the reference's own scratch script uses (src/variations/test_aabb.py):
    brute-force AABB over the leaf voxels finds the same voxels as the octree traversal."""
    from proud_slam_b200 import grid
    s, ms, rays_o, rays_d = _scene_rays("tiny", 2, 200, device)
    leaf = ms["voxel_structure"][:, 8] == 1
    pts = ms["voxel_center_xyz"].detach()[leaf][None].contiguous()
    idx, tmin, tmax = grid.aabb_intersect(rays_o.to(device), rays_d.to(device), pts.to(device), s.voxel_size, 60)
    inv = util.device_rcp(rays_d.reshape(-1, 3), device).reshape(1, -1, 3)
    oi, omin, omax = oracle.aabb_intersect(rays_o.numpy(), rays_d.numpy(), pts.numpy(), s.voxel_size, 60, inv)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(tmin.cpu().numpy(), omin)
    svo_i, _, _ = grid.svo_intersect(rays_o.to(device), rays_d.to(device), ms["voxel_center_xyz"].detach()[None].to(device),
                                     ms["voxel_structure"][None].to(device), s.voxel_size, 60)
    leaf_rows = torch.nonzero(leaf).view(-1)
    for r in range(rays_o.shape[1]):
        a = set(leaf_rows[idx[0, r][idx[0, r] >= 0].cpu().long()].tolist())
        b = set(svo_i[0, r][svo_i[0, r] >= 0].cpu().tolist())
        assert a == b


@pytest.mark.parametrize("n_rays,P", [(1000, 9), (37, 5), (4000, 12)])
def test_inverse_cdf_sampling_bit_exact(n_rays, P, device):
    from proud_slam_b200 import grid
    gen = torch.Generator().manual_seed(n_rays)
    G = 200
    n = -(-n_rays // G)
    cnt = torch.randint(1, P + 1, (G * n,), generator=gen)
    cnt[::17] = P     # some rays fill every slot (the A-Q7 foreign-voxel case)
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
    shp = lambda t, last: t.reshape(G, n, last).contiguous()
    args = [shp(idx, P), shp(tmin, P), shp(tmax, P), shp(noise, M), shp(probs, P), steps.reshape(G, n).contiguous()]
    oi, od, os_ = oracle.inverse_cdf_sampling(*[a.numpy() for a in args], -1.0)
    si, sd, ss = grid.inverse_cdf_sampling(*[a.to(device) for a in args], -1.0)
    assert np.array_equal(si.cpu().numpy(), oi)
    assert np.array_equal(sd.cpu().numpy(), od)
    assert np.array_equal(ss.cpu().numpy(), os_)
    # the tail quirk must actually be exercised by this input
    assert (oi >= 0).sum() > G * n


def test_argument_errors_like_reference(device):
    """TORCH_CHECK wording of sparse_voxels/include/utils.h:10-34."""
    from proud_slam_b200 import grid
    cpu = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        grid.svo_intersect(cpu, cpu, cpu, torch.zeros(1, 4, 9, dtype=torch.int32), 0.2, 10)
    d = cpu.to(device)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        grid.svo_intersect(d, d, d, torch.zeros(1, 4, 9, device=device), 0.2, 10)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        grid.svo_intersect(d.transpose(1, 2).contiguous().transpose(1, 2), d, d,
                           torch.zeros(1, 4, 9, dtype=torch.int32, device=device), 0.2, 10)
