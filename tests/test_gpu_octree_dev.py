"""Device-side octree (csrc/octree_dev.cu) against the host octree (csrc/octree_host.cpp, itself pinned by the oracle's pointer
tree and the Morton-key restatement of the reference): identical rows, in one batch and over several insert calls."""
import numpy as np
import pytest
import torch

from proud_slam_b200 import scene as sc, svo


def _both(grid_dim, batches, device):
    host, dev = svo.Octree(), svo.DeviceOctree(device)
    host.init(grid_dim, 16, 0.2, 8)
    dev.init(grid_dim, 16, 0.2, 8)
    for b in batches:
        host.insert(torch.from_numpy(b))
        dev.insert(torch.from_numpy(b).to(device))
    return host, dev


def _assert_equal(host, dev):
    assert host.count_nodes() == dev.count_nodes()
    assert host.count_leaf_nodes() == dev.count_leaf_nodes()
    hv, hc, hf, _, _ = host.get_centres_and_children()
    dv, dc, df, _, _ = dev.get_centres_and_children()
    assert torch.equal(hv, dv.cpu()) and torch.equal(hc, dc.cpu()) and torch.equal(hf, df.cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["tiny", "replica_small", "scannet_large"])
def test_device_octree_equals_host_octree(kind, device):
    s = sc.make_scene(kind, pixel_stride=2 if kind != "tiny" else 1)
    host, dev = _both(s.grid_dim, [s.voxels], device)
    _assert_equal(host, dev)
    ms_h = svo.build_map_states(host, s.voxel_size, num_embeddings=max(20000, host.count_nodes()), device=device, seed=0)
    ms_d = svo.build_map_states(dev, s.voxel_size, num_embeddings=max(20000, dev.count_nodes()), device=device, seed=0)
    for k in ("voxel_vertex_idx", "voxel_center_xyz", "voxel_structure"):
        assert torch.equal(ms_h[k], ms_d[k]), k


@pytest.mark.gpu
def test_device_octree_incremental_inserts_and_duplicates(device):
    """Keyframe after keyframe (overlapping voxel sets, duplicates inside a batch, a corner leaf promoted to a voxel later) and
    array growth from a tiny initial capacity."""
    s = sc.make_scene("replica_small", pixel_stride=4)
    v = s.voxels
    rng = np.random.default_rng(0)
    batches = [v[:700], np.concatenate([v[400:1500], v[100:300]]), v[rng.permutation(len(v))[:2000]], v]
    # a voxel whose corner (+1,+1,+1) is inserted as a voxel of its own in a later batch: FEATURE leaf promoted to SURFACE
    batches.insert(1, (v[:50] + 1).astype(np.int32))
    host, dev = _both(s.grid_dim, batches, device)
    _assert_equal(host, dev)
