"""Drop-in for the render-path part of ``src/variations/render_helpers.py``.

* ``get_features_vox`` (``:105-156``), differentiable w.r.t. the sample positions and the embedding
  table through the CUDA trilinear kernels;
* ``render_rays`` (``:351-556``): same arguments, same returned dict, differentiable w.r.t.
  ``rays_o``, ``rays_d``, ``map_states['voxel_vertex_emb']`` and the decoder parameters.  Forward =
  fused intersection + sampling + lookup + decoder + compositing; backward = one fused kernel
  chain fed with the caller's gradients (any Criterion works on the returned tensors);
* ``bundle_adjust_frames`` (``:559-676``) and ``track_frame`` (``:679-761``): the reference's loops with
  the whole iteration (render + Criterion + backward) as ONE device-side call and no host sync.

The reference's per-call debug dumps (``np.savetxt``, ``:403-405``) and the dead ``resnet`` argument
are accepted and ignored.
"""
import itertools
import weakref
from copy import deepcopy

import torch

from .. import _lib
from ..pipeline import RenderPipeline

_seed_counter = itertools.count(1)


def _next_seed():
    """Sampling-noise seed: follows torch.manual_seed, advances per call, needs no device sync."""
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter)) & 0xFFFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------
# pipelines are pooled by (device, capacity); one is pinned while an autograd graph refers to it
# ------------------------------------------------------------------------------------------
_pool = {}


def _acquire(num_rays, device):
    cap = 1 << max(10, (int(num_rays) - 1).bit_length())
    key = (str(device), cap)
    free = _pool.setdefault(key, [])
    return (free.pop() if free else RenderPipeline(cap, device)), key


def _release(pipe, key):
    _pool.setdefault(key, []).append(pipe)


class _Lease:
    """Returns the pipeline to the pool when the autograd node (or the caller) lets go of it."""

    def __init__(self, pipe, key):
        self.pipe, self.key = pipe, key
        self._fin = weakref.finalize(self, _release, pipe, key)


def decoder_params_of(sdf_network):
    """The 10 decoder tensors in C-ABI order from our Decoder or the reference's (nrgbd.py:106-113)."""
    if hasattr(sdf_network, "param_list"):
        return sdf_network.param_list()
    if isinstance(sdf_network, (list, tuple)):
        return list(sdf_network)
    m = sdf_network
    return [m.pts_linears[0].weight, m.pts_linears[0].bias, m.pts_linears[1].weight, m.pts_linears[1].bias,
            m.sdf_out.weight, m.sdf_out.bias, m.color_out[0].weight, m.color_out[0].bias,
            m.color_out[2].weight, m.color_out[2].bias]


def _device_states(map_states, device):
    """map_states tensors on the device with the dtypes the kernels take (the reference calls
    ``.cuda()`` on them at every use, render_helpers.py:108-110)."""
    out = {
        "voxel_center_xyz": map_states["voxel_center_xyz"].to(device, torch.float32).contiguous(),
        "voxel_vertex_idx": map_states["voxel_vertex_idx"].to(device, torch.int32).contiguous(),
        "voxel_vertex_emb": map_states["voxel_vertex_emb"].to(device, torch.float32),
    }
    if "voxel_structure" in map_states:      # (the meshing queries pass the three tables of the surface voxels only)
        out["voxel_structure"] = map_states["voxel_structure"].to(device, torch.int32).contiguous()
    return out


# ------------------------------------------------------------------------------------------
# get_features_vox
# ------------------------------------------------------------------------------------------
class _TrilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, emb, vox_idx, centres, vertex_idx, voxel_size):
        lib = _lib.lib()
        xyz_c, emb_c = xyz.contiguous().float(), emb.contiguous()
        n = xyz_c.shape[0]
        feat = torch.empty(n, 16, device=xyz.device, dtype=torch.float32)
        _lib.check(lib.pslam_trilinear_fwd(n, _lib.ptr(xyz_c), _lib.ptr(vox_idx), _lib.ptr(centres), _lib.ptr(vertex_idx),
                                           _lib.ptr(emb_c), float(voxel_size), _lib.ptr(feat), _lib.stream_ptr(xyz.device)),
                   "trilinear forward")
        ctx.save_for_backward(xyz_c, emb_c, vox_idx, centres, vertex_idx)
        ctx.voxel_size = float(voxel_size)
        return feat

    @staticmethod
    def backward(ctx, g_feat):
        lib = _lib.lib()
        xyz, emb, vox_idx, centres, vertex_idx = ctx.saved_tensors
        n = xyz.shape[0]
        g_xyz = torch.empty_like(xyz) if ctx.needs_input_grad[0] else None
        g_emb = torch.zeros_like(emb) if ctx.needs_input_grad[1] else None
        _lib.check(lib.pslam_trilinear_bwd(n, _lib.ptr(xyz), _lib.ptr(vox_idx), _lib.ptr(centres), _lib.ptr(vertex_idx), _lib.ptr(emb),
                                           ctx.voxel_size, _lib.ptr(g_feat.contiguous()), _lib.ptr(g_emb), _lib.ptr(g_xyz),
                                           _lib.stream_ptr(xyz.device)), "trilinear backward")
        return g_xyz, g_emb, None, None, None, None


def get_features_vox(samples, map_states, voxel_size):
    """render_helpers.py:105-156: samples {'sampled_point_xyz' [p,3], 'sampled_point_voxel_idx' [p],
    'sampled_point_distance' [p]} -> {'dists', 'emb' [p,16]}."""
    xyz = samples["sampled_point_xyz"]
    ms = _device_states(map_states, xyz.device)
    idx = samples["sampled_point_voxel_idx"].to(torch.int32).contiguous()
    feats = _TrilinearFn.apply(xyz, ms["voxel_vertex_emb"], idx, ms["voxel_center_xyz"], ms["voxel_vertex_idx"], voxel_size)
    return {"dists": samples["sampled_point_distance"], "emb": feats}


# ------------------------------------------------------------------------------------------
# meshing / evaluation queries (render_helpers.py:243-328): the same two kernels (trilinear lookup + decoder), no gradients
# ------------------------------------------------------------------------------------------
def _field_values(sdf_network, ms, xyz, idx, voxel_size, chunk=1 << 20):
    """(r, g, b, sdf) of arbitrary in-voxel points: ``pslam_trilinear_fwd`` + ``pslam_decoder_fwd`` over chunks of
    ``chunk`` points (the reference goes through get_features_vox + get_values in chunks of 32 voxels)."""
    lib = _lib.lib()
    dev = xyz.device
    dec = [p.detach().to(dev).contiguous() for p in decoder_params_of(sdf_network)]
    from ..pipeline import _decoder_struct, check_decoder_params
    import ctypes as C
    width = check_decoder_params(dec)
    ds = _decoder_struct(dec)
    ws = torch.empty(int(lib.pslam_decoder_ws_count(width)), device=dev)
    n = xyz.shape[0]
    out = torch.empty(n, 4, device=dev)
    centres, vidx, emb = ms["voxel_center_xyz"], ms["voxel_vertex_idx"], ms["voxel_vertex_emb"].detach().contiguous()
    st = _lib.stream_ptr(dev)
    feat = torch.empty(min(n, chunk), 16, device=dev)
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        x, i = xyz[a:b].contiguous(), idx[a:b].contiguous()
        _lib.check(lib.pslam_trilinear_fwd(b - a, _lib.ptr(x), _lib.ptr(i), _lib.ptr(centres), _lib.ptr(vidx), _lib.ptr(emb), float(voxel_size),
                                           _lib.ptr(feat), st), "trilinear forward")
        _lib.check(lib.pslam_decoder_fwd(b - a, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out[a:b]), st), "decoder forward")
    return out


@torch.no_grad()
def get_scores(sdf_network, map_states, voxel_size, bits=8):
    """render_helpers.py:243-296: (r, g, b, sdf) on a ``bits``^3 lattice inside every row of the map, as a CPU tensor
    ``[N, bits, bits, bits, 4]`` (what the mesher's marching cubes reads).  Rows must have their 8 vertex ids (surface voxels),
    as in the reference, where ``F.embedding`` faults on -1."""
    emb = map_states["voxel_vertex_emb"]
    dev = emb.device if emb.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ms = _device_states(map_states, dev)
    res = int(bits)
    n = ms["voxel_center_xyz"].shape[0]
    if n == 0:
        return torch.zeros(0, res, res, res, 4)
    if bool((ms["voxel_vertex_idx"] < 0).any()):
        raise RuntimeError("get_scores: every row needs its 8 vertex ids (pass the surface voxels, as the reference's mesher does)")
    lin = torch.linspace(-0.5, 0.5, res, device=dev)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3) * voxel_size
    xyz = (grid + ms["voxel_center_xyz"].unsqueeze(1)).reshape(-1, 3).float()
    idx = torch.arange(n, device=dev, dtype=torch.int32).repeat_interleave(res ** 3)
    vals = _field_values(sdf_network, ms, xyz, idx, voxel_size)
    return vals.view(n, res, res, res, 4).cpu()


@torch.no_grad()
def eval_points(sdf_network, map_states, sampled_xyz, sampled_idx, voxel_size):
    """render_helpers.py:299-328: colours of given points inside given voxels, CPU ``[p, 3]`` (``None`` for no points)."""
    emb = map_states["voxel_vertex_emb"]
    dev = emb.device if emb.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ms = _device_states(map_states, dev)
    xyz = sampled_xyz.reshape(-1, 3).to(dev).float()
    if xyz.shape[0] == 0:
        return None
    idx = sampled_idx.reshape(-1).to(dev).to(torch.int32)
    return _field_values(sdf_network, ms, xyz, idx, voxel_size)[:, :3].cpu()


class SharedMap:
    """Map hand-off between the mapping and the tracking loop without the reference's round trip through the host
    (``Mapping.update_share_data``: deepcopy -> CPU -> manager -> ``.cuda()``, src/mapping.py:236-247, src/tracking.py:116-125).

    Two device-resident slots; ``publish`` copies the map tensors and the decoder into the inactive slot on the publisher's
    stream (device-to-device, no sync) and then flips a version word; ``acquire`` returns the tensors of the last complete
    version after making the reader's stream wait for that copy.  The tensors are ordinary CUDA tensors, so they also travel
    through ``torch.multiprocessing`` queues as CUDA IPC handles when mapping and tracking are separate processes, as in
    the reference (src/voxslam.py:26)."""

    KEYS = ("voxel_vertex_idx", "voxel_center_xyz", "voxel_structure", "voxel_vertex_emb")

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.version = 0          # completed publishes; slot = (version - 1) & 1

    def publish(self, map_states, sdf_network):
        s = self.version & 1
        dec = decoder_params_of(sdf_network)
        cur = self.slots[s]
        fresh = (cur is None or any(cur["map"][k].shape != map_states[k].shape for k in self.KEYS)
                 or any(a.shape != b.shape for a, b in zip(cur["dec"], dec)))
        if fresh:
            cur = {"map": {k: torch.empty_like(map_states[k], device=self.device) for k in self.KEYS},
                   "dec": [torch.empty_like(p, device=self.device) for p in dec]}
            self.slots[s] = cur
        for k in self.KEYS:
            cur["map"][k].copy_(map_states[k].detach(), non_blocking=True)
        for d, p in zip(cur["dec"], dec):
            d.copy_(p.detach(), non_blocking=True)
        self.events[s].record(torch.cuda.current_stream(self.device))
        self.version += 1
        return self.version

    def acquire(self):
        """(map_states dict, decoder parameter list, version) of the latest complete publish; ``None`` before the first."""
        if self.version == 0:
            return None
        s = (self.version - 1) & 1
        torch.cuda.current_stream(self.device).wait_event(self.events[s])
        cur = self.slots[s]
        return cur["map"], cur["dec"], self.version


# ------------------------------------------------------------------------------------------
# render_rays
# ------------------------------------------------------------------------------------------
class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, emb, cfg, *dec_params):
        device = rays_o.device
        R = rays_o.reshape(-1, 3).shape[0]
        pipe, key = _acquire(R, device)
        lease = _Lease(pipe, key)
        ms = dict(cfg["map_states"])
        ms["voxel_vertex_emb"] = emb.detach().contiguous()
        dec = [p.detach().contiguous() for p in dec_params]
        pipe.bind(rays_o.detach().float().contiguous(), rays_d.detach().float().contiguous(), ms, dec,
                  voxel_size=cfg["voxel_size"], step_size=cfg["step_size"], truncation=cfg["truncation"],
                  max_distance=cfg["max_distance"], noise=cfg["noise"], seed=cfg["seed"], forward_only=True)
        pipe.sample()
        pipe.forward()
        c = pipe.counts()                         # host sync, as the reference has at render_helpers.py:388
        ctx.lease, ctx.R, ctx.counts = lease, R, c
        ctx.dec, ctx.ms, ctx.cfg = dec, ms, cfg
        if c["R_h"] == 0 or c["n_samples"] == 0:
            ctx.empty = True
            z = torch.zeros(0, 0, device=device)
            return z, torch.zeros(0, 3, device=device), torch.zeros(0, device=device), z, z, (pipe.hit_count[:R] > 0).view(1, -1), torch.zeros(0, 1, device=device)
        ctx.empty = False
        o = pipe.outputs()
        ctx.sample_mask = o["sample_mask"]
        for k in ("z_vals", "ray_mask", "raw"):
            ctx.mark_non_differentiable(o[k])
        return o["weights"], o["color"], o["depth"], o["sdf"], o["z_vals"], o["ray_mask"], o["raw"]

    @staticmethod
    def backward(ctx, g_weights, g_color, g_depth, g_sdf, *_):
        n_dec = len(ctx.dec)
        if ctx.empty:
            return (None,) * (4 + n_dec)
        pipe = ctx.lease.pipe
        need_o, need_d, need_e = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_p = any(ctx.needs_input_grad[4:])
        emb = ctx.ms["voxel_vertex_emb"]
        g_emb = torch.zeros_like(emb) if need_e else None
        g_dec = [torch.zeros_like(p) for p in ctx.dec] if need_p else None
        a = pipe.args
        # same intermediates, now with gradient targets
        flags = a.flags & ~(_lib.F_GRAD_EMB | _lib.F_GRAD_DEC | _lib.F_GRAD_RAYS | _lib.F_FORWARD_ONLY)
        if need_e:
            flags |= _lib.F_GRAD_EMB
            a.g_emb = g_emb.data_ptr()
        if need_p:
            from ..pipeline import DecoderGradT, _decoder_struct
            flags |= _lib.F_GRAD_DEC
            a.g_dec = _decoder_struct(g_dec, DecoderGradT)
            if int(ctx.dec[0].shape[0]) == 128:
                if pipe.wgrad_ws is None:
                    pipe.wgrad_ws = torch.empty(int(pipe.lib.pslam_wgrad_ws_bytes(pipe.sample_cap)), dtype=torch.uint8, device=pipe.device)
                a.wgrad_ws, a.wgrad_ws_bytes = pipe.wgrad_ws.data_ptr(), pipe.wgrad_ws.numel()
        if need_o or need_d:
            flags |= _lib.F_GRAD_RAYS
        a.flags = flags
        m = ctx.sample_mask
        g_sdf_csr = g_sdf[m].contiguous() if g_sdf is not None else None
        g_w_csr = g_weights[m].contiguous() if g_weights is not None else None
        pipe.backward_ext(None if g_color is None else g_color.contiguous(), None if g_depth is None else g_depth.contiguous(),
                          g_sdf_csr, g_w_csr)
        R = ctx.R
        g_o = pipe.g_rays_o[:R].clone().view(1, R, 3) if need_o else None
        g_d = pipe.g_rays_d[:R].clone().view(1, R, 3) if need_d else None
        return (g_o, g_d, g_emb, None, *(g_dec if need_p else [None] * n_dec))


def render_rays(rays_o, rays_d, map_states, sdf_network, resnet, step_size, voxel_size, truncation, max_voxel_hit,
                max_distance, chunk_size=10000, profiler=None, return_raw=False, noise=None, seed=None):
    """render_helpers.py:351-556.  rays_o / rays_d [1,R,3]; returns the reference's dict
    {weights [R_h,S], color [R_h,3], depth [R_h], z_vals [R_h,S], sdf [R_h,S] (pad 1), ray_mask [1,R],
    raw ([R_h,1] z_min, or None)} -- or ``(None, 0)`` when no sample exists (:433).  Raises
    AssertionError when no ray hits the map (:388).  ``noise`` ([>=R_h, M] uniform numbers) replays a
    recorded draw (used by the parity tests); by default the device-side counter noise is used."""
    if profiler is not None:
        profiler.tick("render_rays_fused")
    device = rays_o.device
    ms = _device_states(map_states, device)
    emb = map_states["voxel_vertex_emb"]
    if not emb.is_cuda:
        emb = emb.to(device)
    dec = decoder_params_of(sdf_network)
    cfg = dict(map_states=ms, voxel_size=float(voxel_size), step_size=float(step_size), truncation=float(truncation),
               max_distance=float(max_distance), noise=noise, seed=_next_seed() if seed is None else int(seed))
    weights, color, depth, sdf, z_vals, ray_mask, raw = _RenderFn.apply(rays_o, rays_d, emb, cfg, *dec)
    if profiler is not None:
        profiler.tok("render_rays_fused")
    assert ray_mask.sum() > 0, "no ray hits the map"   # render_helpers.py:388
    if z_vals.numel() == 0:
        return None, 0                                   # render_helpers.py:433
    return {"weights": weights, "color": color, "depth": depth, "z_vals": z_vals, "sdf": sdf, "ray_mask": ray_mask,
            "raw": raw if return_raw else None}


# ------------------------------------------------------------------------------------------
# fused iteration used by the SLAM loops
# ------------------------------------------------------------------------------------------
def _frame_on_device(frame, device):
    """Contiguous fp32 device copies of a frame's per-pixel tensors ([HW,3] camera rays, [HW,3] colour, [HW] depth), made
    once per frame object (the reference indexes the frame's own tensors and calls ``.cuda()`` on the result per iteration,
    render_helpers.py:621-640)."""
    cached = getattr(frame, "_pslam_dev", None)
    key = (frame.rays_d.data_ptr(), frame.rgb.data_ptr(), frame.depth.data_ptr(), str(device))
    if cached is None or cached[0] != key:
        t = (frame.rays_d.reshape(-1, 3).to(device, torch.float32).contiguous(), frame.rgb.reshape(-1, 3).to(device, torch.float32).contiguous(),
             frame.depth.reshape(-1).to(device, torch.float32).contiguous())
        cached = (key, t)
        try:
            frame._pslam_dev = cached
        except Exception:
            pass
    return cached[1]


def _criterion_cfg(loss_criteria):
    return dict(weights=(loss_criteria.rgb_weight, loss_criteria.depth_weight, loss_criteria.fs_weight, loss_criteria.sdf_weight),
                truncation=loss_criteria.truncation, max_depth=loss_criteria.max_dpeth)


class FusedIteration:
    """One optimisation iteration = ``pslam_render_step`` (render + Criterion + backward).  Gradients land
    in ``g_emb`` / ``g_dec`` / ``pipe.g_rays_o`` / ``pipe.g_rays_d``; the loss block stays on the device."""

    def __init__(self, max_rays, device, width):
        self.pipe = RenderPipeline(max_rays, device)
        self.g_emb = None
        self.g_dec = None
        self.width = width

    def run(self, rays_o, rays_d, rgb, depth, ms, dec, crit, *, voxel_size, step_size, max_distance, tracking, grad_emb,
            grad_dec, grad_rays, seed, zero_grads=True, noise=None):
        if grad_emb and (self.g_emb is None or self.g_emb.shape != ms["voxel_vertex_emb"].shape):
            self.g_emb = torch.zeros_like(ms["voxel_vertex_emb"])
        if grad_dec and self.g_dec is None:
            # one flat buffer: a single fill clears all ten gradients (and a caller may all-reduce it in one go)
            from ..parallel import FlatGrads
            self._flat = FlatGrads(torch.zeros(0, 16, device=dec[0].device), dec)
            self.g_dec = self._flat.g_dec
        if zero_grads:           # (the fused Adam of bundle_adjust_frames clears the gradients it consumed itself)
            if grad_emb:
                self.g_emb.zero_()
            if grad_dec:
                self._flat.flat.zero_()
        self.pipe.bind(rays_o, rays_d, ms, dec, voxel_size=voxel_size, step_size=step_size, truncation=crit["truncation"],
                       max_distance=max_distance, max_depth=crit["max_depth"], target_rgb=rgb, target_depth=depth, seed=seed, noise=noise,
                       weights=crit["weights"], tracking=tracking, g_emb=self.g_emb if grad_emb else None,
                       g_dec=self.g_dec if grad_dec else None, grad_rays=grad_rays)
        self.pipe.step()
        return self.pipe.loss


def bundle_adjust_frames(keyframe_graph, map_states, sdf_network, resnet, loss_criteria, voxel_size, step_size, N_rays=512,
                         num_iterations=10, truncation=0.1, max_voxel_hit=10, max_distance=10, learning_rate=[1e-2, 5e-3],
                         embed_optim=None, model_optim=None, resnet_optim=None, update_pose=True, device_sampling=True, noise_fn=None):
    """render_helpers.py:559-676.  Same loop and arguments; per iteration ONE fused call produces the loss and every gradient
    and the caller's optimizers are stepped.  Around it, for keyframes whose pose is the 6-vector of se3pose.py with a plain
    Adam: pixels are drawn and rays assembled on the device (``device_sampling``; uniform without replacement, like the
    frame's own ``sample_rays``, but not its random stream), dL/dpose and the pose Adam are two kernels; the embedding / decoder
    Adam is one launch on the optimizers' own state.  Any other frame or optimizer keeps the reference's torch route.
    ``noise_fn(iteration)`` (optional, deterministic replay): the uniform sampling noise [>= R_h, M] of that iteration, in place of
    the counter-based generator -- what the reference draws with ``torch.rand`` inside ``ray_sample`` (voxel_helpers.py:328)."""
    from .. import _lib
    lib = _lib.lib()
    emb = map_states["voxel_vertex_emb"]
    device = emb.device if emb.is_cuda else torch.device("cuda", torch.cuda.current_device())
    optimizers = [embed_optim] + ([model_optim] if model_optim is not None else [])
    # Keyframes whose pose is the 6-vector (t, w) of se3pose.py on this device with a plain Adam of their own take the
    # pose side of the iteration (rotation of the sampled rays, dL/dpose, Adam) as two kernels per frame (csrc/pose.cu)
    # instead of ~250 torch launches per frame and iteration; any other frame keeps the torch route.
    fused = {}
    for keyframe in keyframe_graph:
        pose_obj, opt = getattr(keyframe, "pose", None), getattr(keyframe, "optim", None)
        pdata = getattr(pose_obj, "data", None)
        ok = (torch.is_tensor(pdata) and tuple(pdata.shape) == (6,) and pdata.dtype == torch.float32 and pdata.is_cuda
              and pdata.device == device and isinstance(opt, torch.optim.Adam) and len(opt.param_groups) == 1
              and len(opt.param_groups[0]["params"]) == 1 and opt.param_groups[0]["params"][0] is pdata
              and not opt.param_groups[0].get("amsgrad", False) and opt.param_groups[0].get("weight_decay", 0) == 0
              and not opt.param_groups[0].get("maximize", False))
        if ok:
            fused[id(keyframe)] = True
        elif keyframe.stamp != 0 and update_pose:
            optimizers += [keyframe.optim]
    ms = _device_states(map_states, device)
    dec_params = decoder_params_of(sdf_network)
    crit = _criterion_cfg(loss_criteria)
    crit["truncation"] = truncation
    n_frames = max(len(keyframe_graph), 1)
    it = FusedIteration(N_rays * n_frames, device, int(dec_params[0].shape[0]))
    # Optimizer step of the embedding table and the decoder as one launch on the optimizers' own state (csrc/optim.cu) when
    # they are plain fp32 CUDA Adams; it reads the fused step's gradient buffers directly and clears them in the same pass.
    from ..optim import FusedAdam
    fused_adam = None
    if emb.is_cuda and embed_optim is not None:
        fa = FusedAdam([embed_optim, model_optim], row_tensors=[emb])
        if fa.fused and all(p.is_cuda for p in dec_params):
            fused_adam = fa
    torch_optimizers = [o for o in optimizers if fused_adam is None or (o is not embed_optim and o is not model_optim)]
    # Device-side ray selection and assembly for the fused-pose frames: n distinct uniform pixels per
    # frame and iteration from a keyed permutation (pslam_sample_pixels) instead of the frame's Gumbel top-k over all H*W pixels
    # (src/utils/sample_util.py:4-20), and one gather + rotate kernel per frame (pslam_track_assemble) writing straight into
    # the concatenated batch instead of three boolean-mask gathers and a torch.cat (render_helpers.py:620-646).
    dev_frames = {}
    if device_sampling:
        for frame in keyframe_graph:
            if id(frame) in fused and int(frame.rays_d.reshape(-1, 3).shape[0]) >= N_rays:
                dev_frames[id(frame)] = _frame_on_device(frame, device)
    R_all = N_rays * n_frames
    buf_o, buf_d = torch.empty(R_all, 3, device=device), torch.empty(R_all, 3, device=device)
    buf_rgb, buf_depth = torch.empty(R_all, 3, device=device), torch.empty(R_all, device=device)
    buf_idx = torch.empty(R_all, dtype=torch.int64, device=device)
    base_seed = _next_seed()
    for iteration in range(num_iterations):
        all_dev = len(dev_frames) == len(keyframe_graph) and len(keyframe_graph) > 0
        rays_o, rays_d, rgb_samples, depth_samples = [], [], [], []
        cam_dirs = []
        off = 0
        for k, frame in enumerate(keyframe_graph):
            if id(frame) in dev_frames:
                cam_all, rgb_all, depth_all = dev_frames[id(frame)]
                n = N_rays
                idx = buf_idx[off:off + n]
                _lib.check(lib.pslam_sample_pixels(n, cam_all.shape[0], (base_seed + 0x9E3779B97F4A7C15 * (iteration * n_frames + k + 1)) & 0xFFFFFFFFFFFFFFFF,
                                                   None, _lib.ptr(idx), _lib.stream_ptr(device)), "pslam_sample_pixels")
                o_k, d_k = buf_o[off:off + n], buf_d[off:off + n]
                _lib.check(lib.pslam_track_assemble(n, _lib.ptr(frame.pose.data), _lib.ptr(idx), _lib.ptr(cam_all), _lib.ptr(rgb_all),
                                                    _lib.ptr(depth_all), _lib.ptr(o_k), _lib.ptr(d_k), _lib.ptr(buf_rgb[off:off + n]),
                                                    _lib.ptr(buf_depth[off:off + n]), _lib.stream_ptr(device)), "pslam_track_assemble")
                cam_dirs += [(cam_all, idx)]
                rays_o += [o_k]; rays_d += [d_k]; rgb_samples += [buf_rgb[off:off + n]]; depth_samples += [buf_depth[off:off + n]]
                off += n
                continue
            frame.sample_rays(N_rays)
            sample_mask = frame.sample_mask.to(device)
            sampled_rays_d = frame.rays_d.to(device)[sample_mask]
            if id(frame) in fused:
                cam = sampled_rays_d.float().contiguous()
                n = cam.shape[0]
                idx = torch.arange(n, device=device)
                o_k, d_k = torch.empty_like(cam), torch.empty_like(cam)
                _lib.check(lib.pslam_track_assemble(n, _lib.ptr(frame.pose.data), _lib.ptr(idx), _lib.ptr(cam), None, None, _lib.ptr(o_k),
                                                    _lib.ptr(d_k), None, None, _lib.stream_ptr(device)), "pslam_track_assemble")
                cam_dirs += [(cam, idx)]
                rays_d += [d_k]
                rays_o += [o_k]
            else:
                pose = frame.get_pose().to(device)
                cam_dirs += [None]
                sampled_rays_d = sampled_rays_d @ pose[:3, :3].transpose(-1, -2)
                rays_d += [sampled_rays_d]
                rays_o += [pose[:3, 3].reshape(1, -1).expand_as(sampled_rays_d)]
            rgb_samples += [frame.rgb.to(device)[sample_mask]]
            depth_samples += [frame.depth.to(device)[sample_mask]]
            off += rays_d[-1].shape[0]
        counts = [t.shape[0] for t in rays_d]
        if all_dev:      # every frame wrote its slice of the batch buffers: nothing to concatenate
            n_tot = sum(counts)
            rays_o, rays_d, rgb_samples, depth_samples = buf_o[:n_tot], buf_d[:n_tot], buf_rgb[:n_tot], buf_depth[:n_tot]
        else:
            rays_d = torch.cat(rays_d, dim=0)
            rays_o = torch.cat(rays_o, dim=0)
            rgb_samples = torch.cat(rgb_samples, dim=0).float().contiguous()
            depth_samples = torch.cat(depth_samples, dim=0).float().contiguous()
        torch_pose = rays_d.requires_grad or rays_o.requires_grad
        need_pose = torch_pose or (update_pose and any(id(f) in fused and f.stamp != 0 for f in keyframe_graph))
        ms["voxel_vertex_emb"] = emb.detach() if emb.is_cuda else emb.detach().to(device)
        dec = [p.detach() for p in dec_params]
        it.run(rays_o.detach().float().contiguous(), rays_d.detach().float().contiguous(), rgb_samples, depth_samples, ms, dec, crit,
               voxel_size=voxel_size, step_size=step_size, max_distance=max_distance, tracking=False, grad_emb=True,
               grad_dec=model_optim is not None, grad_rays=need_pose, seed=_next_seed(), zero_grads=fused_adam is None or iteration == 0,
               noise=None if noise_fn is None else noise_fn(iteration))
        for optim in torch_optimizers:
            optim.zero_grad()
        if fused_adam is None:
            emb.grad = it.g_emb if emb.is_cuda else it.g_emb.to(emb.device)
            if model_optim is not None:
                for p, g in zip(dec_params, it.g_dec):
                    p.grad = g.clone()
        if torch_pose:
            R = rays_o.shape[0]
            torch.autograd.backward([rays_o, rays_d], [it.pipe.g_rays_o[:R], it.pipe.g_rays_d[:R]])
        for optim in torch_optimizers:
            optim.step()
        if fused_adam is not None:
            grads = {emb: it.g_emb}
            if model_optim is not None:
                grads.update({p: g for p, g in zip(dec_params, it.g_dec)})
            fused_adam.step(grads=grads, zero_grad=True)
        off = 0
        for frame, cd, n in zip(keyframe_graph, cam_dirs, counts):
            if cd is not None and frame.stamp != 0 and update_pose and n > 0:
                cam, idx = cd
                opt, pdata = frame.optim, frame.pose.data
                grp, st = opt.param_groups[0], opt.state[pdata]
                if len(st) == 0:                                   # what torch.optim.Adam creates on its first step
                    st["step"] = (torch.zeros((), dtype=torch.float32, device=device) if grp.get("capturable", False)
                                  else torch.tensor(0.0, dtype=torch.float32))
                    st["exp_avg"] = torch.zeros_like(pdata)
                    st["exp_avg_sq"] = torch.zeros_like(pdata)
                on_dev = st["step"].is_cuda
                if not on_dev:
                    st["step"] += 1                                # host-side count: passed by value
                _lib.check(lib.pslam_track_pose_step(n, _lib.ptr(pdata), _lib.ptr(idx), _lib.ptr(cam), _lib.ptr(it.pipe.g_rays_o[off:off + n]),
                                                     _lib.ptr(it.pipe.g_rays_d[off:off + n]), _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"]),
                                                     _lib.ptr(st["step"]) if on_dev else None, 0.0 if on_dev else float(st["step"]),
                                                     float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                                                     None, _lib.stream_ptr(device)), "pslam_track_pose_step")
            off += n
    it.pipe.check()      # capacity / stack / no-hit flags of every iteration above, one sync (the reference syncs >= 7 times per iteration)


def track_frame(frame_pose, curr_frame, map_states, sdf_network, resnet, loss_criteria, voxel_size, N_rays=512, step_size=0.05,
                num_iterations=10, truncation=0.1, learning_rate=1e-3, max_voxel_hit=10, max_distance=10, profiler=None,
                depth_variance=False, noise_fn=None):
    """render_helpers.py:679-761: optimise the 6-vector pose of ``curr_frame`` against the fixed map.
    Returns (pose, optim, hit_mask) like the reference.  ``noise_fn``: see ``bundle_adjust_frames``."""
    device = torch.device("cuda", torch.cuda.current_device())
    init_pose = deepcopy(frame_pose).to(device)
    init_pose.requires_grad_(True)
    # the pose side of the iteration (rotation of the sampled rays, dL/dpose, Adam) runs as two kernels of the library
    # (csrc/pose.cu) when the pose is the 6-vector (t, w) of se3pose.py; any other pose object keeps the torch route
    pdata = getattr(init_pose, "data", None)
    fused_pose = (torch.is_tensor(pdata) and tuple(pdata.shape) == (6,) and pdata.dtype == torch.float32 and pdata.is_cuda
                  and len(list(init_pose.parameters())) == 1)
    optim = torch.optim.Adam(init_pose.parameters(), lr=learning_rate, capturable=fused_pose)
    if fused_pose:
        from .. import _lib
        lib = _lib.lib()
        st = optim.state[init_pose.data]
        st["step"] = torch.zeros((), dtype=torch.float32, device=device)
        st["exp_avg"] = torch.zeros_like(init_pose.data)
        st["exp_avg_sq"] = torch.zeros_like(init_pose.data)
        grp = optim.param_groups[0]
        idx_all = torch.arange(N_rays, device=device)
    ms = _device_states(map_states, device)
    ms["voxel_vertex_emb"] = ms["voxel_vertex_emb"].detach().contiguous()
    dec = [p.detach().to(device).contiguous() for p in decoder_params_of(sdf_network)]
    crit = _criterion_cfg(loss_criteria)
    crit["truncation"] = truncation
    it = FusedIteration(N_rays, device, int(dec[0].shape[0]))
    hit_mask = None
    for iteration in range(num_iterations):
        curr_frame.sample_rays(N_rays)
        sample_mask = curr_frame.sample_mask.to(device)
        ray_dirs = curr_frame.rays_d.to(device)[sample_mask]
        rgb = curr_frame.rgb.to(device)[sample_mask].float().contiguous()
        depth = curr_frame.depth.to(device)[sample_mask].float().contiguous()
        if fused_pose:
            ray_dirs = ray_dirs.float().contiguous()
            R = ray_dirs.shape[0]
            idx = idx_all[:R] if R <= idx_all.numel() else torch.arange(R, device=device)
            ray_start_iter, ray_dirs_iter = torch.empty_like(ray_dirs), torch.empty_like(ray_dirs)
            _lib.check(lib.pslam_track_assemble(R, _lib.ptr(init_pose.data), _lib.ptr(idx), _lib.ptr(ray_dirs), None, None,
                                                _lib.ptr(ray_start_iter), _lib.ptr(ray_dirs_iter), None, None, _lib.stream_ptr(device)),
                       "pslam_track_assemble")
        else:
            ray_dirs_iter = ray_dirs @ init_pose.rotation().transpose(-1, -2)
            ray_start_iter = init_pose.translation().reshape(1, -1).expand_as(ray_dirs_iter)
        it.run(ray_start_iter.detach().float().contiguous(), ray_dirs_iter.detach().float().contiguous(), rgb, depth, ms, dec, crit,
               voxel_size=voxel_size, step_size=step_size, max_distance=max_distance, tracking=depth_variance, grad_emb=False,
               grad_dec=False, grad_rays=True, seed=_next_seed(), noise=None if noise_fn is None else noise_fn(iteration))
        R = ray_dirs_iter.shape[0]
        if fused_pose:
            _lib.check(lib.pslam_track_pose_step(R, _lib.ptr(init_pose.data), _lib.ptr(idx), _lib.ptr(ray_dirs), _lib.ptr(it.pipe.g_rays_o),
                                                 _lib.ptr(it.pipe.g_rays_d), _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"]),
                                                 _lib.ptr(st["step"]), 0.0, float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]),
                                                 float(grp["eps"]), None, _lib.stream_ptr(device)), "pslam_track_pose_step")
        else:
            optim.zero_grad()
            torch.autograd.backward([ray_start_iter, ray_dirs_iter], [it.pipe.g_rays_o[:R], it.pipe.g_rays_d[:R]])
            optim.step()
        hit_mask = it.pipe.hit_count[:R] > 0
    it.pipe.check()      # flags of all iterations (one sync per frame)
    return init_pose, optim, hit_mask


# ------------------------------------------------------------------------------------------
# CUDA-graph tracker: the whole tracking iteration (pixel sampling, ray assembly from the pose,
# fused render + loss + backward, pose gradient + Adam) captured once and replayed per iteration
# ------------------------------------------------------------------------------------------
class GraphTracker:
    """Per-frame pose optimisation of ``track_frame`` (render_helpers.py:679-761) without per-iteration
    host work.  The reference's 30 x 1024-ray loop costs ~2 ms of Python/launch overhead per iteration
    around ~0.3 ms of GPU work; here one iteration is a CUDA graph.  Differences from the plain
    ``track_frame`` above: pixels are drawn on the device (``pslam_sample_pixels``: distinct, uniform) instead of by the
    frame's own ``sample_rays``, and the frame's tensors are copied into static buffers once per frame."""

    def __init__(self, n_pixels, map_states, sdf_network, loss_criteria, voxel_size, N_rays=1024, step_size=0.02, truncation=0.1,
                 learning_rate=0.01, max_distance=10.0, depth_variance=True, device=None, fused_pose=True):
        from ..se3pose import OptimizablePose
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        self.N, self.HW = int(N_rays), int(n_pixels)
        self.ms = _device_states(map_states, d)
        self.ms["voxel_vertex_emb"] = self.ms["voxel_vertex_emb"].detach().contiguous()
        self.dec = [p.detach().to(d).contiguous() for p in decoder_params_of(sdf_network)]
        self.crit = _criterion_cfg(loss_criteria)
        self.crit["truncation"] = truncation
        self.cfg = dict(voxel_size=voxel_size, step_size=step_size, max_distance=max_distance, tracking=depth_variance)
        self.rays_d_all = torch.zeros(self.HW, 3, device=d)
        self.rgb_all = torch.zeros(self.HW, 3, device=d)
        self.depth_all = torch.zeros(self.HW, device=d)
        self.pose = OptimizablePose(torch.zeros(6)).to(d)
        self.optim = torch.optim.Adam(self.pose.parameters(), lr=learning_rate, capturable=True)
        self.it = FusedIteration(self.N, d, int(self.dec[0].shape[0]))
        self.counter = torch.zeros(1, dtype=torch.int64, device=d)
        self.base_seed = _next_seed()
        self.hit_mask = torch.zeros(self.N, dtype=torch.bool, device=d)
        self.graph = None
        # fused_pose: ray assembly from the pose and the pose's Adam step are two kernels of the library
        # (csrc/pose.cu) instead of ~250 torch launches (Rodrigues series + autograd + optimizer) per iteration;
        # the Adam state they update is the one torch.optim.Adam owns, so the returned optimizer stays consistent
        self.fused_pose = bool(fused_pose)
        if self.fused_pose:
            p = self.pose.data
            st = self.optim.state[p]
            st["step"] = torch.zeros((), dtype=torch.float32, device=d)
            st["exp_avg"] = torch.zeros_like(p)
            st["exp_avg_sq"] = torch.zeros_like(p)
            self.rays_o = torch.zeros(self.N, 3, device=d)
            self.rays_d = torch.zeros(self.N, 3, device=d)
            self.rgb = torch.zeros(self.N, 3, device=d)
            self.depth = torch.zeros(self.N, device=d)
            self.idx = torch.zeros(self.N, dtype=torch.int64, device=d)

    def _iteration_fused(self):
        from .. import _lib
        lib, d = _lib.lib(), self.device
        idx = self.idx          # distinct pixels, uniform, like the frame's own sample_rays (no replacement)
        pose = self.pose.data
        _lib.check(lib.pslam_track_sample_assemble(self.N, self.HW, self.base_seed ^ 0x5851F42D4C957F2D, _lib.ptr(self.counter), _lib.ptr(pose),
                                                   _lib.ptr(self.rays_d_all), _lib.ptr(self.rgb_all), _lib.ptr(self.depth_all), _lib.ptr(idx),
                                                   _lib.ptr(self.rays_o), _lib.ptr(self.rays_d), _lib.ptr(self.rgb), _lib.ptr(self.depth),
                                                   _lib.stream_ptr(d)), "pslam_track_sample_assemble")
        pipe = self.it.pipe
        pipe.bind(self.rays_o, self.rays_d, self.ms, self.dec, voxel_size=self.cfg["voxel_size"], step_size=self.cfg["step_size"],
                  truncation=self.crit["truncation"], max_distance=self.cfg["max_distance"], max_depth=self.crit["max_depth"],
                  target_rgb=self.rgb, target_depth=self.depth, seed=self.base_seed, seed_dev=self.counter,
                  weights=self.crit["weights"], tracking=self.cfg["tracking"], grad_rays=True)
        pipe.step()
        g = self.optim.param_groups[0]
        st = self.optim.state[pose]
        # the pose step is the iteration's last kernel: it also writes the hit mask and advances the iteration counter that seeds
        # the next iteration's pixel selection and sampling noise (three torch launches per iteration otherwise)
        _lib.check(lib.pslam_track_pose_step_iter(self.N, _lib.ptr(pose), _lib.ptr(idx), _lib.ptr(self.rays_d_all), _lib.ptr(pipe.g_rays_o),
                                                  _lib.ptr(pipe.g_rays_d), _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"]),
                                                  _lib.ptr(st["step"]), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                                  float(g["eps"]), _lib.ptr(self.counter), _lib.ptr(pipe.hit_count), _lib.ptr(self.hit_mask),
                                                  _lib.stream_ptr(d)), "pslam_track_pose_step_iter")

    def _iteration(self):
        if self.fused_pose:
            return self._iteration_fused()
        idx = torch.randint(0, self.HW, (self.N,), device=self.device)
        ray_dirs = self.rays_d_all[idx] @ self.pose.rotation().transpose(-1, -2)
        ray_start = self.pose.translation().reshape(1, -1).expand_as(ray_dirs)
        self.counter.add_(1)
        pipe = self.it.pipe
        pipe.bind(ray_start.detach().contiguous(), ray_dirs.detach().contiguous(), self.ms, self.dec, voxel_size=self.cfg["voxel_size"],
                  step_size=self.cfg["step_size"], truncation=self.crit["truncation"], max_distance=self.cfg["max_distance"],
                  max_depth=self.crit["max_depth"], target_rgb=self.rgb_all[idx].contiguous(), target_depth=self.depth_all[idx].contiguous(),
                  seed=self.base_seed, seed_dev=self.counter, weights=self.crit["weights"], tracking=self.cfg["tracking"], grad_rays=True)
        pipe.step()
        self.optim.zero_grad(set_to_none=False)
        torch.autograd.backward([ray_start, ray_dirs], [pipe.g_rays_o[: self.N], pipe.g_rays_d[: self.N]])
        self.optim.step()
        self.hit_mask.copy_(pipe.hit_count[: self.N] > 0)

    def _reset(self, frame, init_pose):
        self.rays_d_all.copy_(frame.rays_d.reshape(-1, 3))
        self.rgb_all.copy_(frame.rgb.reshape(-1, 3))
        self.depth_all.copy_(frame.depth.reshape(-1))
        with torch.no_grad():
            self.pose.data.copy_(init_pose.data.to(self.device))
        for st in self.optim.state.values():      # a fresh Adam per frame, like the reference (:700)
            for v in st.values():
                if torch.is_tensor(v):
                    v.zero_()

    def track(self, init_pose, frame, num_iterations=30):
        if getattr(self, "_pending_check", None) is not None:      # flags of the previous frame's iterations (copied without a sync)
            pending, self._pending_check = self._pending_check, None
            pending()
        self._reset(frame, init_pose)
        if self.graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):                 # warm-up: allocations, lazy kernel attributes, Adam state
                    self._iteration()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._reset(frame, init_pose)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._iteration()
            self._reset(frame, init_pose)
        for _ in range(num_iterations):
            self.graph.replay()
        self._pending_check = self.it.pipe.check_async()
        return self.pose, self.optim, self.hit_mask

    def finish(self):
        """Evaluates the flags of the last tracked frame (``track`` checks each frame's flags when the next one starts)."""
        if getattr(self, "_pending_check", None) is not None:
            pending, self._pending_check = self._pending_check, None
            pending()
