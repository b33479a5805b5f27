"""Host octree of the product (csrc/octree_host.cpp via proud_slam_b200.svo) against the oracle's
pointer-tree restatement of the reference, plus structural properties.  CPU only."""
import pickle

import numpy as np
import pytest
import torch

import oracle
from proud_slam_b200 import scene as sc, svo


def _both(vox, grid_dim=256, chunks=1):
    t = svo.Octree()
    t.init(grid_dim, 16, 0.2, 8)
    o = oracle.Octree(grid_dim)
    for part in np.array_split(vox, chunks):
        t.insert(torch.from_numpy(part))
        o.insert(part)
    return t, o


@pytest.mark.parametrize("kind,chunks", [("tiny", 1), ("tiny", 3), ("replica_small", 2)])
def test_matches_oracle(kind, chunks):
    s = sc.make_scene(kind, pixel_stride=4)
    t, o = _both(s.voxels, s.grid_dim, chunks)
    v, c, f, px, pc = t.get_centres_and_children()
    v2, c2, f2 = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), v2) and np.array_equal(c.numpy(), c2) and np.array_equal(f.numpy(), f2)
    assert px.shape == (v.shape[0], 8, 4) and pc.shape == (v.shape[0], 8, 3)


def test_random_voxels_and_duplicates():
    g = np.random.default_rng(0)
    vox = g.integers(0, 200, size=(500, 3)).astype(np.int32)
    vox = np.concatenate([vox, vox[:100]])     # re-inserting is a no-op
    t, o = _both(vox)
    v, c, f, _, _ = t.get_centres_and_children()
    v2, c2, f2 = o.get_centres_and_children()
    assert np.array_equal(v.numpy(), v2) and np.array_equal(c.numpy(), c2) and np.array_equal(f.numpy(), f2)
    assert t.count_leaf_nodes() == len(np.unique(vox, axis=0))
    assert t.has_voxel(torch.tensor(vox[3])) and not t.has_voxel(torch.tensor([255, 255, 255]) - 3)


def test_structure_invariants():
    s = sc.make_scene("tiny")
    t, _ = _both(s.voxels, s.grid_dim)
    ms = svo.build_map_states(t, s.voxel_size, device="cpu", seed=0)
    st, vi, ce = ms["voxel_structure"], ms["voxel_vertex_idx"], ms["voxel_center_xyz"]
    assert st[0, 8] == s.grid_dim                       # root = row 0, side = grid
    leaf = st[:, 8] == 1
    assert int(leaf.sum()) == len(s.voxels)
    assert (st[leaf, :8] == -1).all()                   # leaves have no children
    assert (vi[leaf] >= 0).all() and (vi[~leaf] == -1).all()
    # children are half the parent's size and lie inside it
    for r in torch.nonzero(~leaf & (st[:, 8] > 0)).view(-1).tolist():
        for cidx in st[r, :8].tolist():
            if cidx >= 0:
                assert st[cidx, 8] * 2 == st[r, 8]
                assert ((ce[cidx] - ce[r]).abs() <= st[r, 8] * s.voxel_size / 2).all()
    # corner 0 of a voxel is the voxel's own lower corner leaf... (same x,y,z): its row is the voxel itself
    assert (vi[leaf][:, 0] == torch.nonzero(leaf).view(-1)).all()


def test_pickle_round_trip_reinserts():
    s = sc.make_scene("tiny")
    t, _ = _both(s.voxels, s.grid_dim, chunks=2)
    t2 = pickle.loads(pickle.dumps(t))
    a, b = t.get_centres_and_children(), t2.get_centres_and_children()
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_errors():
    t = svo.Octree()
    with pytest.raises(RuntimeError, match="not initialized"):
        t.count_nodes()
    t.init(256, 16, 0.2, 8)
    with pytest.raises(RuntimeError, match="dimensions mismatch"):
        t.insert(torch.zeros(4, 2, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="outside"):
        t.insert(torch.tensor([[300, 0, 0]], dtype=torch.int32))
    with pytest.raises(RuntimeError, match="num_embeddings"):
        t.insert(torch.tensor([[1, 2, 3]], dtype=torch.int32))
        svo.build_map_states(t, 0.2, num_embeddings=3, device="cpu")
