"""Pose parameter at the boundary of the render path: a 6-vector (t, w) with the rotation given by
Rodrigues' formula, the two scalar functions sin(x)/x and (1-cos x)/x^2 evaluated by their 11-term
Maclaurin series as the reference does (``src/se3pose.py:24-34, 62-91``).  Stays in torch: the
render kernels return dL/d(rays_o, rays_d) and autograd carries them through these few 3x3 ops."""
import math

import torch
import torch.nn as nn

_NTERMS = 11
# sin(x)/x = sum (-1)^i x^(2i)/(2i+1)!   ;   (1-cos x)/x^2 = sum (-1)^i x^(2i)/(2i+2)!
_COEF_A = [(-1.0) ** i / math.factorial(2 * i + 1) for i in range(_NTERMS)]
_COEF_B = [(-1.0) ** i / math.factorial(2 * i + 2) for i in range(_NTERMS)]


def _series(x, coef):
    x2 = x * x
    acc = torch.zeros_like(x)
    p = torch.ones_like(x)
    for c in coef:
        acc = acc + c * p
        p = p * x2
    return acc


def skew(w):
    w0, w1, w2 = w.unbind(dim=-1)
    z = torch.zeros_like(w0)
    return torch.stack([torch.stack([z, -w2, w1], -1), torch.stack([w2, z, -w0], -1), torch.stack([-w1, w0, z], -1)], -2)


class OptimizablePose(nn.Module):
    def __init__(self, init_pose):
        super().__init__()
        self.register_parameter("data", nn.Parameter(init_pose.detach().clone().float()))

    def copy_from(self, pose):
        self.data = nn.Parameter(pose.data.detach().clone())

    def rotation(self):
        w = self.data[3:]
        wx = skew(w)
        theta = w.norm(dim=-1)[..., None, None]
        eye = torch.eye(3, device=w.device, dtype=torch.float32)
        return eye + _series(theta, _COEF_A) * wx + _series(theta, _COEF_B) * (wx @ wx)

    def translation(self):
        return self.data[:3]

    def matrix(self):
        Rt = torch.eye(4, device=self.data.device)
        Rt[:3, :3] = self.rotation()
        Rt[:3, 3] = self.translation()
        return Rt

    @classmethod
    def log(cls, R, eps=1e-7):
        trace = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
        theta = ((trace - 1) / 2).clamp(-1 + eps, 1 - eps).acos()[..., None, None] % math.pi
        lnR = 1 / (2 * _series(theta, _COEF_A) + 1e-8) * (R - R.transpose(-2, -1))
        return torch.stack([lnR[..., 2, 1], lnR[..., 0, 2], lnR[..., 1, 0]], dim=-1)

    @classmethod
    def from_matrix(cls, Rt):
        return cls(torch.cat([Rt[:3, 3], cls.log(Rt[:3, :3])], dim=-1))
