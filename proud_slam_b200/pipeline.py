"""Fused render path: one mapping / tracking iteration (render_rays + Criterion +
backward of the reference, ``src/variations/render_helpers.py:351-556``,
``src/criterion.py:16-116``) as a short fixed sequence of sm_100a kernels with no
host synchronisation.

``RenderPipeline`` owns the device workspaces and fills the ``pslam_render_t``
argument block of the C ABI (``include/proud_slam_b200.h``).  Intermediates
stay in compact CSR form on the device; ``intersections()``, ``samples()`` and
``outputs()`` materialise the reference's padded ``[R_h, S]`` tensors on demand
(those calls synchronise, like the reference does at the same places).
"""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (C_NSAMP, C_OVERFLOW, C_P, C_RH, C_S, F_FORWARD_ONLY, F_GRAD_DEC, F_GRAD_EMB, F_GRAD_RAYS,
                   F_TRACKING, L_COLOR, L_DEPTH, L_FS, L_SDF, L_TOTAL, DecoderGradT, DecoderT, RenderT, ptr)

MAX_DEPTH = 10.0   # pad value of sampled_point_depth, reference voxel_helpers.py:24
N_MAX_HITS = 50    # reference voxel_helpers.py:561 (max_voxel_hit is ignored there)
DEC_SHAPES = lambda w: [(w, 16), (w,), (w, w), (w,), (129, w), (129,), (w, 144), (w,), (3, w), (3,)]


def _decoder_struct(params: Sequence[torch.Tensor], cls=DecoderT):
    names = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4", "W5", "b5")
    s = cls()
    if cls is DecoderT:
        s.width = int(params[0].shape[0])
    for n, t in zip(names, params):
        setattr(s, n, t.data_ptr())
    return s


def check_decoder_params(params):
    if len(params) != 10:
        raise RuntimeError("decoder must have 10 parameter tensors (depth=2, skips=[], embedder none)")
    w = int(params[0].shape[0])
    if w not in (128, 256):
        raise RuntimeError(f"decoder width {w} not supported (128 or 256)")
    for t, shp in zip(params, DEC_SHAPES(w)):
        if tuple(t.shape) != shp:
            raise RuntimeError(f"decoder parameter shape {tuple(t.shape)} != {shp} "
                               "(reference nrgbd.Decoder with in_dim=16, sdf_dim=128)")
        _lib.require_cuda(t, "decoder parameter", torch.float32)
    return w


class RenderPipeline:
    """Workspaces + launch glue for ``pslam_render_{sample,forward,backward,step}``."""

    def __init__(self, max_rays: int, device=None, samples_per_ray: int = 64, n_max: int = N_MAX_HITS):
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("RenderPipeline needs a CUDA device (there is no CPU path)")
        self.R_cap = int(max_rays)
        self.n_max = int(n_max)
        self.sample_cap = int(max_rays) * int(samples_per_ray)
        d, R, cap = self.device, self.R_cap, self.sample_cap
        i32 = dict(dtype=torch.int32, device=d)
        f32 = dict(dtype=torch.float32, device=d)
        self.hit_idx = torch.empty(self.n_max * R, **i32)
        self.hit_min = torch.empty(self.n_max * R, **f32)
        self.hit_max = torch.empty(self.n_max * R, **f32)
        self.hit_count = torch.zeros(R, **i32)
        self.hit_ray = torch.zeros(R, **i32)
        self.ray_rank = torch.zeros(R, **i32)
        self.samp_off = torch.zeros(R + 1, **i32)
        self.samp_vox = torch.zeros(cap, **i32)
        self.samp_ray = torch.zeros(cap, **i32)
        self.samp_z = torch.zeros(cap, **f32)
        self.samp_dist = torch.zeros(cap, **f32)
        self.samp_out = torch.zeros(cap, 4, **f32)
        self.samp_w = torch.zeros(cap, **f32)
        self.samp_gout = torch.zeros(cap, 4, **f32)
        self.ray_out = torch.zeros(R, 8, **f32)
        self.scratch_i = torch.zeros(int(self.lib.pslam_render_scratch_i_count(R)), **i32)
        self.scratch_f = torch.zeros(int(self.lib.pslam_render_scratch_f_count(R)), **f32)
        self.counters = torch.zeros(_lib.C_COUNT, **i32)
        self.loss = torch.zeros(_lib.L_COUNT, **f32)
        self.loss_raw = torch.zeros(16, dtype=torch.float64, device=d)
        self.g_rays_o = torch.zeros(R, 3, **f32)
        self.g_rays_d = torch.zeros(R, 3, **f32)
        self.dec_ws = {w: torch.empty(int(self.lib.pslam_decoder_ws_count(w)), **f32) for w in (128, 256)}
        self.wgrad_ws = None   # allocated on first use (any backward through the tensor-core path)
        self.node_cache = None
        self._map_key, self._map_built = None, False
        self.args = RenderT()
        self._keep = None      # tensors referenced by self.args
        self.R = 0

    # ------------------------------------------------------------------ argument block
    def bind(self, rays_o, rays_d, map_states, dec_params, *, voxel_size, step_size, truncation,
             max_distance, max_depth=10.0, target_rgb=None, target_depth=None, noise=None, seed=0,
             weights=(0.5, 1.0, 10.0, 5000.0), tracking=False, g_emb=None, g_dec=None, grad_rays=False,
             forward_only=False, defer_loss=False, seed_dev=None):
        """Fills the pslam_render_t block.  Tensors: rays_* [R,3] (or [1,R,3]) f32; map_states as the
        reference's dict (voxel_center_xyz [N,3] f32, voxel_structure [N,9] i32, voxel_vertex_idx [N,8]
        i32, voxel_vertex_emb [E,16] f32); dec_params = the 10 decoder tensors in state_dict order;
        weights = (rgb, depth, fs, sdf) of the Criterion; noise [R_h.., stride] explicit uniform noise or
        None for the counter-based generator; g_emb [E,16] / g_dec (10 tensors) are accumulated into."""
        rays_o = rays_o.reshape(-1, 3)
        rays_d = rays_d.reshape(-1, 3)
        R = rays_o.shape[0]
        if R > self.R_cap or R <= 0:
            raise RuntimeError(f"ray batch of {R} outside this pipeline's capacity (1..{self.R_cap})")
        centres, structure = map_states["voxel_center_xyz"], map_states["voxel_structure"]
        vidx, emb = map_states["voxel_vertex_idx"], map_states["voxel_vertex_emb"]
        for t, n, dt in ((rays_o, "ray_start", torch.float32), (rays_d, "ray_dir", torch.float32),
                         (centres, "points", torch.float32), (structure, "children", torch.int32),
                         (vidx, "voxel_vertex_idx", torch.int32), (emb, "voxel_vertex_emb", torch.float32)):
            _lib.require_cuda(t, n, dt)
        width = check_decoder_params(dec_params)
        a = self.args
        a.R, a.N, a.E, a.n_max, a.sample_cap = R, centres.shape[0], emb.shape[0], self.n_max, self.sample_cap
        flags = 0
        if tracking:
            flags |= F_TRACKING
        if g_emb is not None:
            flags |= F_GRAD_EMB
        if g_dec is not None:
            flags |= F_GRAD_DEC
        if grad_rays:
            flags |= F_GRAD_RAYS
        if forward_only:
            flags |= F_FORWARD_ONLY
        if defer_loss:
            flags |= _lib.F_DEFER_LOSS
        a.voxel_size, a.step_size, a.truncation = float(voxel_size), float(step_size), float(truncation)
        a.max_distance, a.max_depth = float(max_distance), float(max_depth)
        a.w_rgb, a.w_depth, a.w_fs, a.w_sdf = [float(x) for x in weights]
        a.rays_o, a.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
        if target_rgb is not None:
            target_rgb = _lib.require_cuda(target_rgb.reshape(-1, 3), "target_rgb", torch.float32)
            target_depth = _lib.require_cuda(target_depth.reshape(-1), "target_depth", torch.float32)
            if target_rgb.shape[0] != R or target_depth.shape[0] != R:
                raise RuntimeError("targets must have one entry per ray")
        a.target_rgb = None if target_rgb is None else target_rgb.data_ptr()
        a.target_depth = None if target_depth is None else target_depth.data_ptr()
        a.centres, a.structure, a.vertex_idx, a.emb = (centres.data_ptr(), structure.data_ptr(), vidx.data_ptr(),
                                                     emb.data_ptr())
        a.dec = _decoder_struct(dec_params)
        a.dec_ws = self.dec_ws[width].data_ptr()
        # traversal cache of the octree walk: 128 B per octree row, built by the first sample() / step() on a map and reused
        # until the map changes (new tensors, or an in-place edit that torch's version counters see; a caller that edits the
        # map behind torch's back -- raw kernels -- calls invalidate_map())
        need = 128 * int(centres.shape[0])
        if self.node_cache is None or self.node_cache.numel() < need:
            self.node_cache = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._map_key = None
        a.node_cache, a.node_cache_bytes = self.node_cache.data_ptr(), self.node_cache.numel()
        key = (centres.data_ptr(), structure.data_ptr(), int(centres.shape[0]), centres._version, structure._version,
               map_states.get("generation") if hasattr(map_states, "get") else None)
        if key == self._map_key and self._map_built:
            flags |= _lib.F_NODE_CACHE_VALID
        else:
            self._map_key, self._map_built = key, False
        a.flags = flags
        # the workspace holds the wgrad operands (decoder gradients) and, for any backward, the forward's ReLU masks and the
        # feature rows of the stand-alone trilinear kernels (width 128: 3.3 kB per sample, width 256: 7.3 kB)
        if (g_dec is not None or g_emb is not None or grad_rays) and not forward_only:
            need_ws = int(self.lib.pslam_wgrad_ws_bytes_w(self.sample_cap, width))
            if self.wgrad_ws is None or self.wgrad_ws.numel() < need_ws:
                self.wgrad_ws = torch.empty(need_ws, dtype=torch.uint8, device=self.device)
            a.wgrad_ws, a.wgrad_ws_bytes = self.wgrad_ws.data_ptr(), self.wgrad_ws.numel()
        else:
            a.wgrad_ws, a.wgrad_ws_bytes = None, 0
        if noise is not None:
            _lib.require_cuda(noise, "uniform_noise", torch.float32)
            a.noise, a.noise_stride = noise.data_ptr(), int(noise.shape[-1])
        else:
            a.noise, a.noise_stride = None, 0
        a.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.seed_dev = None if seed_dev is None else seed_dev.data_ptr()   # int64 device scalar added to the seed
        for name in ("hit_idx", "hit_min", "hit_max", "hit_count", "hit_ray", "ray_rank", "samp_off", "samp_vox",
                     "samp_ray", "samp_z", "samp_dist", "samp_out", "samp_w", "samp_gout", "ray_out", "scratch_i",
                     "scratch_f", "counters", "loss", "loss_raw", "g_rays_o", "g_rays_d"):
            setattr(a, name, getattr(self, name).data_ptr())
        a.g_emb = None if g_emb is None else _lib.require_cuda(g_emb, "g_emb", torch.float32).data_ptr()
        if g_dec is not None:
            for t, p in zip(g_dec, dec_params):
                if t.shape != p.shape:
                    raise RuntimeError("decoder gradient shapes must match the parameters")
                _lib.require_cuda(t, "decoder gradient", torch.float32)
            a.g_dec = _decoder_struct(g_dec, DecoderGradT)
        else:
            a.g_dec = DecoderGradT()
        self._keep = (rays_o, rays_d, target_rgb, target_depth, centres, structure, vidx, emb, list(dec_params), noise,
                      g_emb, g_dec, seed_dev)
        self.R = R
        return self

    # ------------------------------------------------------------------ stages (no host sync)
    def _call(self, fn, what):
        _lib.check(fn(C.byref(self.args), _lib.stream_ptr(self.device)), what)

    def _map_cache_built(self):
        # the launch just enqueued built the traversal cache of the bound map: later launches on this stream reuse it
        if not self._map_built and self.args.node_cache:
            self._map_built = True
            self.args.flags |= _lib.F_NODE_CACHE_VALID

    def invalidate_map(self):
        """The bound map's octree arrays were edited in place outside torch: the next sample() rebuilds the traversal cache."""
        self._map_key, self._map_built = None, False
        self.args.flags &= ~_lib.F_NODE_CACHE_VALID

    def sample(self):
        self._call(self.lib.pslam_render_sample, "pslam_render_sample")
        self._map_cache_built()

    def forward(self):
        self._call(self.lib.pslam_render_forward, "pslam_render_forward")

    def backward(self):
        self._call(self.lib.pslam_render_backward, "pslam_render_backward")

    def backward_ext(self, g_color=None, g_depth=None, g_sdf=None, g_weight=None):
        """Backward from upstream gradients w.r.t. the rendered outputs (rank / CSR order)."""
        _lib.check(self.lib.pslam_render_backward_ext(C.byref(self.args), ptr(g_color), ptr(g_depth), ptr(g_sdf), ptr(g_weight),
                                                      _lib.stream_ptr(self.device)), "pslam_render_backward_ext")

    def finalize_loss(self, rows):
        """Multi-GPU: rows = all ranks' ``loss_raw`` ([nranks,16] float64 on this device)."""
        _lib.require_cuda(rows, "loss rows")
        _lib.check(self.lib.pslam_loss_finalize(C.byref(self.args), ptr(rows), int(rows.shape[0]),
                                                _lib.stream_ptr(self.device)), "pslam_loss_finalize")

    def stage(self, i):
        _lib.check(self.lib.pslam_render_stage(C.byref(self.args), int(i), _lib.stream_ptr(self.device)), f"stage {i}")
        if int(i) == 0:
            self._map_cache_built()

    def step(self):
        self._call(self.lib.pslam_render_step, "pslam_render_step")
        self._map_cache_built()

    # ------------------------------------------------------------------ results (synchronising)
    def counts(self):
        c = self.counters.tolist()
        if c[C_OVERFLOW] & 1:
            raise RuntimeError(f"sample capacity exceeded ({self.sample_cap}); build the pipeline with a larger samples_per_ray")
        if c[C_OVERFLOW] & 4:
            raise RuntimeError("a decoder weight, activation or gradient left the range of the 3xF16 tensor-core build "
                               "(|16 x value| >= 32752); select the 3xTF32 build: pslam_set_option(PSLAM_OPT_DECODER, 0)")
        if c[C_OVERFLOW] & 2:
            raise RuntimeError("octree traversal stack overflow (the reference asserts here, intersect_gpu.cu:235)")
        if c[C_OVERFLOW] & 16:
            raise RuntimeError("a peer rank did not arrive at a cross-GPU exchange (csrc/peer.cu)")
        return dict(R_h=c[C_RH], P=c[C_P], n_samples=c[C_NSAMP], S=c[C_S])

    # ------------------------------------------------------------------ flags of a whole loop of steps
    def _raise_for(self, bits, steps):
        if bits & 1:
            raise RuntimeError(f"sample capacity exceeded ({self.sample_cap}) in one of the last {steps} steps: the tail of the batch was "
                               "dropped; build the pipeline with a larger samples_per_ray")
        if bits & 2:
            raise RuntimeError("octree traversal stack overflow (the reference asserts here, intersect_gpu.cu:235)")
        if bits & 16:
            raise RuntimeError("a peer rank did not arrive at a cross-GPU exchange of one of the last steps (csrc/peer.cu gives up after "
                               "~2 s): the ranks no longer run the same sequence of steps")
        if bits & 8:
            raise AssertionError("no ray hits the map")          # the reference's assert, render_helpers.py:388
        if bits & 4:
            # an operand left the f16 window of the default decoder build: its gradients saturated (satfinite, never NaN).
            # Continue on the fp32-range 3xTF32 build instead of raising; the caller may redo the affected loop.
            import warnings
            _lib.check(self.lib.pslam_set_option(1, 0), "set_option")
            warnings.warn("a decoder weight, activation or gradient left the range of the 3xF16 tensor-core build "
                          "(|16 x value| >= 32752): switched this process to the 3xTF32 build (PSLAM_OPT_DECODER = 0)")
            return "range"
        return None

    def check(self):
        """Flags of every step since the last check (one host sync): raises for a dropped batch tail, a traversal stack
        overflow or a step without hit rays; returns "range" after switching to the 3xTF32 decoder when the 3xF16 window was
        left (see ``_raise_for``), else None.  The fused loops call this once per loop instead of syncing per step."""
        c = self.counters.tolist()
        bits, steps = c[_lib.C_STICKY] | c[C_OVERFLOW] | (8 if (c[_lib.C_STEPS] > 0 and c[C_RH] == 0) else 0), c[_lib.C_STEPS]
        self.counters[_lib.C_STICKY: _lib.C_STEPS + 1] = 0
        return self._raise_for(bits, steps)

    def check_async(self):
        """Non-blocking form: starts a copy of the counters into pinned memory and returns a callable that, once the copy has
        landed (it waits if necessary), does what ``check`` does.  ``GraphTracker`` evaluates it at the next frame."""
        host = torch.empty(_lib.C_COUNT, dtype=torch.int32, pin_memory=True)
        host.copy_(self.counters, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.counters[_lib.C_STICKY: _lib.C_STEPS + 1] = 0

        def finish():
            ev.synchronize()
            c = host.tolist()
            bits = c[_lib.C_STICKY] | c[C_OVERFLOW] | (8 if (c[_lib.C_STEPS] > 0 and c[C_RH] == 0) else 0)
            return self._raise_for(bits, c[_lib.C_STEPS])
        return finish

    def losses(self):
        l = self.loss.tolist()
        return dict(loss=l[L_TOTAL], color_loss=l[L_COLOR], depth_loss=l[L_DEPTH], fs_loss=l[L_FS], sdf_loss=l[L_SDF])

    def intersections(self, max_distance=None):
        """The reference's ``ray_intersect_vox`` result (voxel_helpers.py:558-595): (dict of [1,R,P]
        tensors sorted by entry depth and trimmed, hits [1,R])."""
        c = self.counts()
        R, P = self.R, max(c["P"], 0)
        md = float(self.args.max_distance if max_distance is None else max_distance)
        cnt = self.hit_count[:R].long()
        slot = torch.arange(P, device=self.device)
        valid = slot[None, :] < cnt[:, None]
        # buffers are slot-major with row length R (the bound batch size)
        idx = self.hit_idx[: self.n_max * R].view(self.n_max, R)[:P].t()
        tmin = self.hit_min[: self.n_max * R].view(self.n_max, R)[:P].t()
        tmax = self.hit_max[: self.n_max * R].view(self.n_max, R)[:P].t()
        out = {
            "min_depth": torch.where(valid, tmin, torch.full_like(tmin, md)).unsqueeze(0).contiguous(),
            "max_depth": torch.where(valid, tmax, torch.full_like(tmax, md)).unsqueeze(0).contiguous(),
            "intersected_voxel_idx": torch.where(valid, idx, torch.full_like(idx, -1)).unsqueeze(0).contiguous(),
        }
        return out, (cnt > 0).unsqueeze(0)

    def _padded(self, c):
        Rh, S = c["R_h"], c["S"]
        off = self.samp_off[: Rh + 1].long()
        cnt = (off[1:] - off[:-1])
        k = torch.arange(S, device=self.device)
        mask = k[None, :] < cnt[:, None]
        src = (off[:-1, None] + k[None, :]).clamp(max=max(self.sample_cap - 1, 0))
        return mask, src

    def samples(self):
        """The reference's ``ray_sample`` result on the hit rays (voxel_helpers.py:637-663)."""
        c = self.counts()
        mask, src = self._padded(c)
        z = torch.where(mask, self.samp_z[src], torch.full_like(src, MAX_DEPTH, dtype=torch.float32))
        dist = torch.where(mask, self.samp_dist[src], torch.zeros_like(z))
        vox = torch.where(mask, self.samp_vox[src], torch.full_like(self.samp_vox[src], -1))
        return {"sampled_point_depth": z, "sampled_point_distance": dist, "sampled_point_voxel_idx": vox}

    def outputs(self):
        """The reference's ``render_rays`` dict (render_helpers.py:547-556) from the last forward."""
        c = self.counts()
        Rh = c["R_h"]
        mask, src = self._padded(c)
        z = torch.where(mask, self.samp_z[src], torch.full_like(src, MAX_DEPTH, dtype=torch.float32))
        sdf = torch.where(mask, self.samp_out[src, 3], torch.ones_like(z))
        w = torch.where(mask, self.samp_w[src], torch.zeros_like(z))
        ro = self.ray_out[:Rh]
        return {
            "weights": w, "color": ro[:, :3].clone(), "depth": ro[:, 3].clone(), "z_vals": z, "sdf": sdf,
            "ray_mask": (self.hit_count[: self.R] > 0).view(1, -1), "raw": ro[:, 4:5].clone(),
            "sample_mask": mask,
        }
