"""Fused optimizer step of the mapping loop: ``torch.optim.Adam`` over the embedding table and the
decoder (reference ``src/mapping.py:81-82``, stepped at ``src/variations/render_helpers.py:667-676``) as ONE launch of the
library (``pslam_adam_step``, ``csrc/optim.cu``), in place on the optimizers' own state tensors, so the ``torch.optim.Adam``
objects the caller owns stay consistent (``state_dict`` / later torch steps keep working).

Embedding rows that never received a gradient are skipped (exactly what the dense step does to them: nothing), and the
gradients can be cleared in the same pass.
"""
import ctypes as C

import torch

from . import _lib


def _eligible(opt):
    if not isinstance(opt, torch.optim.Adam) or type(opt) is not torch.optim.Adam:
        return False
    for g in opt.param_groups:
        if g.get("amsgrad", False) or g.get("weight_decay", 0) != 0 or g.get("maximize", False) or g.get("differentiable", False):
            return False
        for p in g["params"]:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                return False
    return True


class FusedAdam:
    """Steps one or more ``torch.optim.Adam`` objects (same betas / eps; each param group keeps its own lr) with one kernel.

    ``grads``: optional mapping param -> gradient tensor to read instead of ``param.grad`` (the fused render step leaves its
    gradients in its own buffers; no ``.grad`` assignment or clone is needed).  Falls back to ``optimizer.step()`` for
    anything that is not a plain fp32 CUDA Adam."""

    MAX_TENSORS = 16

    def __init__(self, optimizers, row_tensors=()):
        self.opts = [o for o in optimizers if o is not None]
        self.fused = all(_eligible(o) for o in self.opts) and len(self.opts) > 0
        self.row_ids = {id(p) for p in row_tensors}
        self.row_active = {}
        self.lib = _lib.lib()

    def _state(self, opt, group, p):
        st = opt.state[p]
        if len(st) == 0:                       # what torch.optim.Adam creates on its first step
            st["step"] = (torch.zeros((), dtype=torch.float32, device=p.device) if group.get("capturable", False)
                          else torch.tensor(0.0, dtype=torch.float32))
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def step(self, grads=None, zero_grad=True):
        if not self.fused:
            for o in self.opts:
                o.step()
            return False
        entries, host_step, betas, eps = [], None, None, None
        for opt in self.opts:
            for group in opt.param_groups:
                for p in group["params"]:
                    g = grads.get(p) if grads is not None else None
                    if g is None:
                        g = p.grad
                    if g is None:
                        continue
                    if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape):
                        raise RuntimeError("FusedAdam: gradients must be contiguous fp32 CUDA tensors of the parameter's shape")
                    st = self._state(opt, group, p)
                    on_dev = st["step"].is_cuda
                    if not on_dev:
                        st["step"] += 1
                        hs = float(st["step"])
                        if host_step is not None and hs != host_step:
                            raise RuntimeError("FusedAdam: parameters with host-side step counts must share one count")
                        host_step = hs
                    b = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
                    if betas is not None and b != (betas[0], betas[1], eps):
                        raise RuntimeError("FusedAdam: all parameter groups must share betas and eps")
                    betas, eps = (b[0], b[1]), b[2]
                    t = _lib.AdamTensorT()
                    t.param, t.grad, t.exp_avg, t.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                    t.step = st["step"].data_ptr() if on_dev else None
                    t.n, t.lr = p.numel(), float(group["lr"])
                    if id(p) in self.row_ids and p.dim() == 2 and p.shape[1] == 16:
                        ra = self.row_active.get(id(p))
                        if ra is None or ra.numel() != p.shape[0]:
                            # rows that already carry state (an optimizer stepped by torch before) are live
                            ra = (st["exp_avg"].abs().amax(1) + st["exp_avg_sq"].amax(1) > 0).to(torch.uint8).contiguous()
                            self.row_active[id(p)] = ra
                        t.row_active, t.row = ra.data_ptr(), 16
                    else:
                        t.row_active, t.row = None, 0
                    entries.append(t)
        for i in range(0, len(entries), self.MAX_TENSORS):
            chunk = entries[i:i + self.MAX_TENSORS]
            arr = (_lib.AdamTensorT * len(chunk))(*chunk)
            dev = self.opts[0].param_groups[0]["params"][0].device
            _lib.check(self.lib.pslam_adam_step(arr, len(chunk), 0.0 if host_step is None else host_step, betas[0], betas[1], eps,
                                                1 if zero_grad else 0, _lib.stream_ptr(dev)), "pslam_adam_step")
        return True
