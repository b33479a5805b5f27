"""Drop-in for ``src/variations/nrgbd.py``: the SDF / colour decoder.

Same constructor, parameter names (``state_dict`` keys ``pts_linears.{0,1}``, ``sdf_out``,
``color_out.{0,2}``), ``get_values`` / ``get_sdf`` / ``forward`` as ``nrgbd.py:80-146``; the math runs
in the CUDA decoder kernels (tcgen05 for width 128, fp32 SIMT for width 256) through the C ABI,
with a hand-written backward (activations recomputed on chip, not stored).

Only the configuration the SLAM configs use is supported: ``depth=2, skips=[], embedder='none',
in_dim=16, sdf_dim=128`` and ``width`` 128 (Replica) or 256 (ScanNet / ARKit); anything else raises.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ..pipeline import DecoderGradT, _decoder_struct, check_decoder_params


class Same(nn.Module):
    """nrgbd.py:71-78 (embedder 'none')."""

    def __init__(self, in_dim):
        super().__init__()
        self.embedding_size = in_dim

    def forward(self, x):
        return x.squeeze(0)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, *params):
        lib = _lib.lib()
        feat = feat.contiguous()
        _lib.require_cuda(feat, "emb", torch.float32)
        params_c = [p.contiguous() for p in params]
        width = check_decoder_params(params_c)
        n = feat.shape[0]
        out = torch.empty(n, 4, device=feat.device, dtype=torch.float32)
        ws = torch.empty(int(lib.pslam_decoder_ws_count(width)), device=feat.device, dtype=torch.float32)
        ds = _decoder_struct(params_c)
        _lib.check(lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out),
                                         _lib.stream_ptr(feat.device)), "decoder forward")
        ctx.save_for_backward(feat, *params_c)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.lib()
        feat, *params = ctx.saved_tensors
        width = int(params[0].shape[0])
        n = feat.shape[0]
        g_out = g_out.contiguous()
        need_p = any(ctx.needs_input_grad[1:])
        g_feat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
        ws = torch.empty(int(lib.pslam_decoder_ws_count(width)), device=feat.device, dtype=torch.float32)
        ds = _decoder_struct(params)
        g_params, gs, wws = None, None, None
        if need_p:
            # one aligned block per gradient (vector reductions need 16-byte alignment)
            g_params = [torch.zeros_like(p) for p in params]
            gs = _decoder_struct(g_params, DecoderGradT)
            if width == 128:
                wws = torch.empty(int(lib.pslam_wgrad_ws_bytes(n)), dtype=torch.uint8, device=feat.device)
        _lib.check(lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(g_out), _lib.ptr(g_feat),
                                         C.byref(gs) if need_p else None, _lib.ptr(wws), 0 if wws is None else wws.numel(),
                                         _lib.stream_ptr(feat.device)), "decoder backward")
        return (g_feat, *(g_params if need_p else [None] * len(params)))


class Decoder(nn.Module):
    def __init__(self, depth=8, width=256, in_dim=3, sdf_dim=128, skips=[4], multires=6, embedder="nerf",
                 local_coord=False, **kwargs):
        super().__init__()
        if not (depth == 2 and list(skips) == [] and embedder == "none" and in_dim == 16 and sdf_dim == 128
                and width in (128, 256)):
            raise NotImplementedError(
                "the B200 render path implements the decoder the SLAM configs use (depth=2, skips=[], "
                "embedder='none', in_dim=16, sdf_dim=128, width 128 or 256; configs/replica/replica.yaml:17-23)")
        self.D, self.W, self.skips = depth, width, skips
        self.pe = Same(in_dim)
        self.pts_linears = nn.ModuleList([nn.Linear(in_dim, width), nn.Linear(width, width)])
        self.sdf_out = nn.Linear(width, 1 + sdf_dim)
        self.color_out = nn.Sequential(nn.Linear(sdf_dim + in_dim, width), nn.ReLU(), nn.Linear(width, 3), nn.Sigmoid())

    def param_list(self):
        """The 10 tensors in the order of the C ABI (W1 b1 W2 b2 W3 b3 W4 b4 W5 b5)."""
        return [self.pts_linears[0].weight, self.pts_linears[0].bias, self.pts_linears[1].weight, self.pts_linears[1].bias,
                self.sdf_out.weight, self.sdf_out.bias, self.color_out[0].weight, self.color_out[0].bias,
                self.color_out[2].weight, self.color_out[2].bias]

    def get_values(self, x):
        """[p,16] -> [p,4] = (r, g, b, sdf), nrgbd.py:116-135."""
        x = self.pe(x)
        return _DecoderFn.apply(x.reshape(-1, x.shape[-1]), *self.param_list())

    def get_sdf(self, inputs):
        return self.get_values(inputs["emb"])[:, 3]

    def forward(self, inputs):
        outputs = self.get_values(inputs["emb"])
        return {"color": outputs[:, :3], "sdf": outputs[:, 3]}
