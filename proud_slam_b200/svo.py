"""Host-side drop-in for the reference's TorchScript class ``torch.classes.svo.Octree``
(``third_party/sparse_octree/src/bindings.cpp:11-35``, used at ``src/mapping.py:86-87, 292, 302``).

CPU tensors in and out, like the reference.  Backed by ``csrc/octree_host.cpp`` through the C ABI.
The point-cloud payload the reference also returns (``pcd_xyz``, ``pcd_color``) feeds only a branch
that is commented out of ``render_rays``; zeros of the reference's shapes are returned for it.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Octree:
    def __init__(self):
        self._h = None
        self._inserted = []
        self.grid_dim, self.feat_dim, self.voxel_size, self.max_num = 256, 16, 0.2, 8

    def init(self, grid_dim, feat_dim, voxel_size, max_num=8):
        """Octree::init, octree.cpp:46-60."""
        self._free()
        self.grid_dim, self.feat_dim, self.voxel_size, self.max_num = int(grid_dim), int(feat_dim), float(voxel_size), int(max_num)
        self._h = _lib.lib().pslam_octree_new(self.grid_dim)
        if not self._h:
            raise RuntimeError("grid_dim must be >= 2")
        self._inserted = []

    def _free(self):
        if self._h:
            _lib.lib().pslam_octree_free(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("Octree not initialized!")
        return C.c_void_p(self._h)

    def insert(self, pts, color=None, pcd=None):
        """Octree::insert, octree.cpp:104-294: pts int32 [M,3] voxel coordinates (colour / point
        payload accepted and ignored, see module docstring)."""
        pts = torch.as_tensor(pts)
        if pts.dim() != 2 or pts.size(1) != 3:
            raise RuntimeError(f"Point dimensions mismatch: inputs are {tuple(pts.shape)} expect [M,3]")
        v = np.ascontiguousarray(pts.cpu().numpy().astype(np.int32))
        if (v < 0).any() or (v >= self.grid_dim - 1).any():
            raise RuntimeError("voxel coordinates outside the octree grid")
        self._inserted.append(pts.detach().cpu().clone())   # pickling = re-insertion, bindings.cpp:27-35
        _lib.check(_lib.lib().pslam_octree_insert(self._handle(), v.ctypes.data_as(C.c_void_p), int(v.shape[0])), "octree insert")

    def count_nodes(self):
        return int(_lib.lib().pslam_octree_count(self._handle()))

    def count_leaf_nodes(self):
        return int(_lib.lib().pslam_octree_count_leaves(self._handle()))

    def has_voxel(self, pt):
        """octree.cpp:441-473: true when a leaf exists at the coordinate -- a SURFACE voxel or a FEATURE corner."""
        x, y, z = [int(a) for a in torch.as_tensor(pt).view(-1)[:3]]
        return bool(_lib.lib().pslam_octree_has_voxel(self._handle(), x, y, z))

    def try_insert(self, pts):
        """octree.cpp:385-416: the share of the voxels' corner keys (8 per voxel, deduplicated) that are already in the tree,
        as a double in [0, 1] (the reference intersects them with ``all_keys``, the corner keys of everything inserted)."""
        pts = torch.as_tensor(pts)
        if pts.dim() != 2 or pts.size(1) != 3:
            return -1.0
        v = pts.cpu().numpy().astype(np.int64)
        corner = np.array([[(j >> 2) & 1, (j >> 1) & 1, j & 1] for j in range(8)], np.int64)
        keys = np.unique((v[:, None, :] + corner[None]).reshape(-1, 3), axis=0)
        if keys.shape[0] == 0:
            return float("nan")                                   # 0 / 0 in the reference
        present = sum(self.has_voxel(k) for k in keys)
        return float(present) / float(keys.shape[0])

    def get_leaf_voxels(self):
        """octree.cpp:480-511: float32 [n,3] coordinates of the SURFACE voxels in the reference's order (depth-first, children
        0..7, child id = x bit + 2 y bit + 4 z bit per level)."""
        n = self.count_leaf_nodes()
        out = np.empty((max(n, 1), 3), np.int32)
        _lib.lib().pslam_octree_leaf_voxels(self._handle(), out.ctypes.data_as(C.c_void_p), n)
        v = out[:n].astype(np.int64)
        key = np.zeros(n, np.int64)
        for b in range(int(np.log2(self.grid_dim)) - 1, -1, -1):   # most significant level first: (z, y, x) bits
            key = (key << 3) | ((((v[:, 2] >> b) & 1) << 2) | (((v[:, 1] >> b) & 1) << 1) | ((v[:, 0] >> b) & 1))
        return torch.from_numpy(v[np.argsort(key, kind="stable")].astype(np.float32))

    def get_centres_and_children(self):
        """octree.cpp:561-687 -> (voxels f32[N,4], children f32[N,8], features i32[N,8],
        pcd_xyz f32[N,max_num,4], pcd_color f32[N,max_num,3])."""
        n = self.count_nodes()
        voxels = np.empty((n, 4), np.float32)
        children = np.empty((n, 8), np.float32)
        features = np.empty((n, 8), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().pslam_octree_flatten(self._handle(), p(voxels), p(children), p(features)), "octree flatten")
        return (torch.from_numpy(voxels), torch.from_numpy(children), torch.from_numpy(features),
                torch.zeros(n, self.max_num, 4), torch.zeros(n, self.max_num, 3))

    # pickling by re-insertion, as the reference's def_pickle (bindings.cpp:27-35)
    def __getstate__(self):
        return dict(cfg=(self.grid_dim, self.feat_dim, self.voxel_size, self.max_num), pts=self._inserted)

    def __setstate__(self, st):
        self._h = None
        self.init(*st["cfg"])
        for p in st["pts"]:
            self.insert(p)


class DeviceOctree:
    """The same octree on the GPU (``csrc/octree_dev.cu``): CUDA tensors in, CUDA tensors out, identical rows.  ``insert`` takes
    int voxel coordinates [M,3] on the device (what ``Mapping.insert_points`` computes there, src/mapping.py:258-264, before the
    reference copies them to the host tree); ``get_centres_and_children`` returns device tensors, so ``build_map_states`` needs
    no upload.  Row ids equal the host tree's (creation order of the sequential insertion)."""

    def __init__(self, device=None):
        self._h = None
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.grid_dim, self.feat_dim, self.voxel_size, self.max_num = 256, 16, 0.2, 8

    def init(self, grid_dim, feat_dim, voxel_size, max_num=8, capacity_hint=0):
        self._free()
        self.grid_dim, self.feat_dim, self.voxel_size, self.max_num = int(grid_dim), int(feat_dim), float(voxel_size), int(max_num)
        with torch.cuda.device(self.device):
            self._h = _lib.lib().pslam_doctree_new(self.grid_dim, int(capacity_hint))
        if not self._h:
            raise RuntimeError("device octree: " + _lib.lib().pslam_last_error().decode(errors="replace"))

    def _free(self):
        if self._h:
            _lib.lib().pslam_doctree_free(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise RuntimeError("Octree not initialized!")
        return C.c_void_p(self._h)

    def insert(self, pts, color=None, pcd=None):
        pts = torch.as_tensor(pts)
        if pts.dim() != 2 or pts.size(1) != 3:
            raise RuntimeError(f"Point dimensions mismatch: inputs are {tuple(pts.shape)} expect [M,3]")
        v = pts.to(self.device, torch.int32).contiguous()
        if v.numel() and (bool((v < 0).any()) or bool((v >= self.grid_dim - 1).any())):
            raise RuntimeError("voxel coordinates outside the octree grid")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pslam_doctree_insert(self._handle(), _lib.ptr(v), int(v.shape[0]), _lib.stream_ptr(self.device)), "device octree insert")

    def count_nodes(self):
        return int(_lib.lib().pslam_doctree_count(self._handle()))

    def _types(self):
        t = torch.empty(self.count_nodes(), dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().pslam_doctree_types(self._handle(), _lib.ptr(t), _lib.stream_ptr(self.device)), "device octree types")
        return t

    def count_leaf_nodes(self):
        return int((self._types() == 0).sum())

    def get_centres_and_children(self):
        """(voxels f32[N,4], children f32[N,8], features i32[N,8], pcd_xyz, pcd_color) on the device."""
        n = self.count_nodes()
        voxels = torch.empty(n, 4, dtype=torch.float32, device=self.device)
        children = torch.empty(n, 8, dtype=torch.float32, device=self.device)
        features = torch.empty(n, 8, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().pslam_doctree_flatten(self._handle(), _lib.ptr(voxels), _lib.ptr(children), _lib.ptr(features),
                                                        _lib.stream_ptr(self.device)), "device octree flatten")
        return (voxels, children, features, torch.zeros(n, self.max_num, 4, device=self.device), torch.zeros(n, self.max_num, 3, device=self.device))


def build_map_states(octree, voxel_size, num_embeddings=20000, embed_dim=16, device="cuda", seed=None, emb=None):
    """``Mapping.update_grid_pcd_features`` (src/mapping.py:301-377): flatten the octree and build the
    ``map_states`` dict the render path reads."""
    voxels, children, features, _, _ = octree.get_centres_and_children()
    centres = (voxels[:, :3] + voxels[:, -1:] / 2) * voxel_size
    structure = torch.cat([children, voxels[:, -1:]], -1).int()
    n = voxels.shape[0]
    if emb is None:
        if num_embeddings < n:
            raise RuntimeError(f"num_embeddings={num_embeddings} < {n} octants: F.embedding would fault in the reference")
        g = None if seed is None else torch.Generator().manual_seed(seed)
        emb = torch.zeros(num_embeddings, embed_dim).normal_(0.0, 0.01, generator=g)   # mapping.py:71-80
    return {
        "voxel_vertex_idx": features.to(device).contiguous(),
        "voxel_center_xyz": centres.float().to(device).contiguous(),
        "voxel_structure": structure.to(device).contiguous(),
        "voxel_vertex_emb": emb.to(device).contiguous(),
    }
