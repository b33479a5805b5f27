"""grid.build_octree (API surface) against the reference's own compiled EasyOctree when available."""
import pytest
import torch

from proud_slam_b200 import easy_octree


def test_structure():
    pts = torch.tensor([[1, 1, 1], [-1, -1, -1], [3, -3, 1], [1, 3, -1]], dtype=torch.int32)
    centers, children = easy_octree.build_octree(torch.zeros(3, dtype=torch.int32), pts, 1)
    total = centers.shape[0]
    assert children[total - 1, 8] == 4                       # root: 2^(depth+1)
    assert torch.equal(centers[:4], pts)                      # terminals keep their point index
    assert (children[:4, 8] == 1).all() and (children[:4, :8] == -1).all()
    assert sorted(children[:, :8][children[:, :8] >= 0].tolist()) == list(range(total - 1))   # every non-root node has one parent


def test_matches_reference_binary():
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref/grid.so not built")
    g = torch.Generator().manual_seed(0)
    pts = torch.unique(torch.randint(-15, 16, (300, 3), generator=g) * 2 + 1, dim=0).int()
    center = torch.zeros(3, dtype=torch.int32)
    want = ref.build_octree(center, pts, 4)
    got = easy_octree.build_octree(center, pts, 4)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
