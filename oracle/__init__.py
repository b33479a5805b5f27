"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the render path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
``proud_slam_b200`` never does (tests/test_layout.py enforces it).

* ``grid_oracle.c``    kernels 1-2 (ray/octree intersection, inverse-CDF sampling)
* ``octree_oracle.c``  octree insert + flatten (producer of ``map_states``)
* ``render_oracle.py`` the torch-level stages (post-processing, trilinear lookup,
                       decoder, compositing, loss) restated on CPU tensors
* ``build_ref.py``     builds the reference's own CUDA ``grid`` extension from
                       /root/reference into ``oracle/_ref`` (GPU-box oracle)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SOURCES = ["grid_oracle.c", "octree_oracle.c"]
_lib = None


def build(force=False):
    """Compile the C restatement with gcc (no CUDA involved)."""
    srcs = [os.path.join(_HERE, s) for s in _SOURCES]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-shared", "-fPIC", "-o", _LIB_PATH] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.octree_new.restype = ctypes.c_void_p
        _lib.octree_new.argtypes = [ctypes.c_int]
        _lib.octree_free.argtypes = [ctypes.c_void_p]
        _lib.octree_count.argtypes = [ctypes.c_void_p]
        _lib.octree_insert.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.octree_flatten.argtypes = [ctypes.c_void_p] * 4
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max, inv_dir=None):
    """numpy in/out, grouped layout of the reference (intersect.cpp:83-112):
    ray_* [B,K,3], points [B,N,3], children [B,N,9] -> idx,min,max [B,K,n_max]."""
    ray_start, ray_dir, points = _f32(ray_start), _f32(ray_dir), _f32(points)
    children = _i32(children)
    B, K = ray_start.shape[:2]
    N = points.shape[1]
    idx = np.empty((B, K, n_max), np.int32)
    tmin = np.empty((B, K, n_max), np.float32)
    tmax = np.empty((B, K, n_max), np.float32)
    inv = None if inv_dir is None else _f32(inv_dir)
    ovf = lib().oracle_svo_intersect(
        ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(K), ctypes.c_float(voxelsize),
        ctypes.c_int(n_max), _p(ray_start), _p(ray_dir), _p(points), _p(children),
        _p(inv), _p(idx), _p(tmin), _p(tmax))
    assert ovf == 0, "DFS stack overflow (reference would assert, intersect_gpu.cu:235)"
    return idx, tmin, tmax


def aabb_intersect(ray_start, ray_dir, points, voxelsize, n_max, inv_dir=None):
    ray_start, ray_dir, points = _f32(ray_start), _f32(ray_dir), _f32(points)
    B, K = ray_start.shape[:2]
    N = points.shape[1]
    idx = np.empty((B, K, n_max), np.int32)
    tmin = np.empty((B, K, n_max), np.float32)
    tmax = np.empty((B, K, n_max), np.float32)
    inv = None if inv_dir is None else _f32(inv_dir)
    lib().oracle_aabb_intersect(
        ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(K), ctypes.c_float(voxelsize),
        ctypes.c_int(n_max), _p(ray_start), _p(ray_dir), _p(points), _p(inv),
        _p(idx), _p(tmin), _p(tmax))
    return idx, tmin, tmax


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, noise, probs, steps,
                         fixed_step_size=-1.0, return_clipped=False):
    """numpy in/out, grouped layout of sample.cpp:56-95: [G,n,P] hits, [G,n]
    steps, [G,n,M] noise -> sampled idx/depth/dists [G,n,M]."""
    pts_idx = _i32(pts_idx)
    min_depth, max_depth, noise = _f32(min_depth), _f32(max_depth), _f32(noise)
    probs, steps = _f32(probs), _f32(steps)
    G, n, P = pts_idx.shape
    M = noise.shape[-1]
    sidx = np.empty((G, n, M), np.int32)
    sdepth = np.empty((G, n, M), np.float32)
    sdist = np.empty((G, n, M), np.float32)
    clipped = lib().oracle_inverse_cdf_sampling(
        ctypes.c_int(G), ctypes.c_int(n), ctypes.c_int(P), ctypes.c_int(M),
        ctypes.c_float(fixed_step_size), _p(pts_idx), _p(min_depth), _p(max_depth),
        _p(noise), _p(probs), _p(steps), _p(sidx), _p(sdepth), _p(sdist))
    if return_clipped:
        return sidx, sdepth, sdist, clipped
    assert clipped == 0, "sample buffer too narrow"
    return sidx, sdepth, sdist


class Octree:
    """Restatement of the subset of torch.classes.svo.Octree the path needs
    (bindings.cpp:11-35): init / insert / get_centres_and_children."""

    def __init__(self, grid_dim=256):
        self._h = lib().octree_new(int(grid_dim))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.octree_free(self._h)
            self._h = None

    def insert(self, vox):
        vox = _i32(vox).reshape(-1, 3)
        lib().octree_insert(self._h, _p(vox), int(vox.shape[0]))

    def count(self):
        return int(lib().octree_count(self._h))

    def get_centres_and_children(self):
        n = self.count()
        voxels = np.empty((n, 4), np.float32)
        children = np.empty((n, 8), np.float32)
        features = np.empty((n, 8), np.int32)
        lib().octree_flatten(self._h, _p(voxels), _p(children), _p(features))
        return voxels, children, features
