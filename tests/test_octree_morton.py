"""a1 pin: the product's host octree (csrc/octree_host.cpp through svo.Octree) and the C oracle against an independent
restatement of the reference's Morton-key code path (oracle/octree_morton.py <- sparse_octree/src/octree.cpp:104-294, 385-511,
561-687, utils.h:45-124): flattened tensors bit-equal, and the three methods the advisor found deviating."""
import numpy as np
import pytest
import torch

import oracle
from oracle import octree_morton as om
from proud_slam_b200 import scene as sc, svo


def _voxels(seed, n, grid):
    rng = np.random.default_rng(seed)
    # a bent sheet of voxels plus clutter, several insert batches with overlap (later corner-0 visits promote FEATURE leaves)
    u = rng.integers(4, grid - 6, size=(n, 2))
    sheet = np.stack([u[:, 0], u[:, 1], (grid // 2 + (u[:, 0] // 7) - (u[:, 1] // 5)) % (grid - 2)], 1)
    clutter = rng.integers(0, grid - 2, size=(n // 4, 3))
    return [sheet[: n // 2], np.concatenate([sheet[n // 3:], clutter])]


@pytest.mark.parametrize("grid,n,seed", [(16, 40, 0), (64, 300, 1), (256, 600, 2)])
def test_product_octree_matches_morton_restatement(grid, n, seed):
    ref = om.Octree(grid)
    prod = svo.Octree()
    prod.init(grid, 16, 0.2, 8)
    cor = oracle.Octree(grid)
    for batch in _voxels(seed, n, grid):
        ref.insert(batch)
        prod.insert(torch.from_numpy(batch.astype(np.int32)))
        cor.insert(batch.astype(np.int32))
    v, c, f = ref.get_centres_and_children()
    pv, pc, pf, _, _ = prod.get_centres_and_children()
    assert ref.count_nodes() == prod.count_nodes() and ref.count_leaf_nodes() == prod.count_leaf_nodes()
    assert np.array_equal(pv.numpy(), v) and np.array_equal(pc.numpy(), c) and np.array_equal(pf.numpy(), f)
    ov, oc, of = cor.get_centres_and_children()
    assert np.array_equal(np.asarray(ov), v) and np.array_equal(np.asarray(oc), c) and np.array_equal(np.asarray(of), f)
    # get_leaf_voxels: float32, the reference's depth-first order
    lv = prod.get_leaf_voxels()
    assert lv.dtype == torch.float32 and np.array_equal(lv.numpy(), ref.get_leaf_voxels())
    # has_voxel: any leaf (SURFACE or FEATURE corner); try_insert: overlap ratio of corner keys
    rng = np.random.default_rng(seed + 10)
    probes = np.concatenate([_voxels(seed, n, grid)[0][:20] + rng.integers(0, 2, size=(20, 3)), rng.integers(0, grid - 2, size=(20, 3))])
    for p in probes:
        assert prod.has_voxel(torch.tensor(p)) == ref.has_voxel(p), p
    for k in range(4):
        q = np.concatenate([_voxels(seed, n, grid)[k % 2][k:k + 15], rng.integers(0, grid - 2, size=(10, 3))])
        assert abs(prod.try_insert(torch.from_numpy(q.astype(np.int32))) - ref.try_insert(q)) < 1e-12


def test_scene_map_states_agree_on_a_replica_shaped_room():
    """The synthetic room the goldens and the bench are built on: product tree == Morton restatement on the tiny scene."""
    s = sc.make_scene("tiny")
    ref = om.Octree(s.grid_dim)
    ref.insert(s.voxels)
    prod = svo.Octree()
    prod.init(s.grid_dim, 16, s.voxel_size, 8)
    prod.insert(torch.from_numpy(s.voxels))
    v, c, f = ref.get_centres_and_children()
    pv, pc, pf, _, _ = prod.get_centres_and_children()
    assert np.array_equal(pv.numpy(), v) and np.array_equal(pc.numpy(), c) and np.array_equal(pf.numpy(), f)
