"""oracle/ref_import.py -- TEST INFRASTRUCTURE (build container only).

Imports the reference's OWN Python (``/root/reference/src``) unmodified so that
``oracle/render_oracle.py`` and the golden fixtures can be pinned against it.
``/root/reference`` does not exist on the GPU box; everything that calls
``load()`` is skipped there.

What is stubbed (all result-neutral):
* ``open3d``, ``annoy``, ``tensorboardX`` -- imported by render_helpers.py:4,8,9, never used on the path;
* ``grid`` -- the CUDA-only extension; replaced by a module whose
  ``svo_intersect`` / ``inverse_cdf_sampling`` call ``oracle/grid_oracle.c``;
* ``Tensor.cuda`` / ``Module.cuda`` -> identity (hard-coded ``.cuda()`` calls);
* ``numpy.savetxt`` -> no-op inside render_helpers (debug dumps, render_helpers.py:403-404).
"""
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "variations"))


class _Recorder:
    """Keeps the arguments of the last native calls so fixtures can store the
    exact noise tensor the reference drew (SURVEY A-Q6)."""
    noise_chunks = []
    inv_dir = None   # optional [R,3] reciprocal directions (device MUFU.RCP values)


def _fake_grid():
    import oracle
    m = types.ModuleType("grid")

    def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max):
        inv = None
        if _Recorder.inv_dir is not None:
            # reference pads rays by repeating the leading ones (voxel_helpers.py:118-122)
            B, K = ray_start.shape[:2]
            flat = np.asarray(_Recorder.inv_dir, np.float32).reshape(-1, 3)
            reps = int(np.ceil(B * K / flat.shape[0]))
            inv = np.concatenate([flat] + [flat[: B * K - flat.shape[0]]] * (reps > 1), 0).reshape(B, K, 3)
        i, a, b = oracle.svo_intersect(ray_start.numpy(), ray_dir.numpy(), points.numpy(),
                                       children.numpy(), float(voxelsize), int(n_max), inv)
        return torch.from_numpy(i), torch.from_numpy(a), torch.from_numpy(b)

    def inverse_cdf_sampling(pts_idx, min_depth, max_depth, noise, probs, steps, fixed):
        _Recorder.noise_chunks.append(noise.clone())
        i, a, b = oracle.inverse_cdf_sampling(pts_idx.numpy(), min_depth.numpy(), max_depth.numpy(),
                                              noise.numpy(), probs.numpy(), steps.numpy(), float(fixed))
        return torch.from_numpy(i), torch.from_numpy(a), torch.from_numpy(b)

    m.svo_intersect = svo_intersect
    m.inverse_cdf_sampling = inverse_cdf_sampling
    for name in ("ball_intersect", "aabb_intersect", "triangle_intersect", "uniform_ray_sampling", "build_octree"):
        setattr(m, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError(name)))
    return m


_loaded = None


def load():
    """Returns a namespace with the reference modules: render_helpers,
    voxel_helpers, nrgbd, criterion, se3pose, and the recorder."""
    global _loaded
    if _loaded is not None:
        return _loaded
    assert available(), "reference tree not mounted"
    for name in ("open3d", "annoy", "tensorboardX"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.AnnoyIndex = object
            stub.SummaryWriter = object
            sys.modules[name] = stub
    sys.modules["grid"] = _fake_grid()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import importlib
    rh = importlib.import_module("variations.render_helpers")
    vh = importlib.import_module("variations.voxel_helpers")
    nrgbd = importlib.import_module("variations.nrgbd")
    crit = importlib.import_module("criterion")
    se3 = importlib.import_module("se3pose")
    rh.np = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    rh.np.savetxt = lambda *a, **k: None
    _loaded = types.SimpleNamespace(render_helpers=rh, voxel_helpers=vh, nrgbd=nrgbd,
                                    criterion=crit, se3pose=se3, recorder=_Recorder)
    return _loaded
