"""Debug helper: find rays whose rendered depth differs from the oracle and print why."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import util
from oracle import render_oracle as ro
from proud_slam_b200 import scene as sc
from proud_slam_b200.pipeline import RenderPipeline

device = torch.device("cuda:0")
s, ms = util.build_scene("tiny")
dec = ro.decoder_params(width=128, seed=1)
rays_o, rays_d, rgb, depth = sc.sample_batch(s, [0, 1], 300, seed=5)
depth = depth * (1.0 + 0.01 * torch.randn(depth.shape, generator=torch.Generator().manual_seed(4)))
rays_o.requires_grad_(True); rays_d.requires_grad_(True)
inv = util.device_rcp(rays_d.detach().reshape(-1, 3), device)
out, loss, parts = util.oracle_step(rays_o, rays_d, rgb, depth, ms, dec, voxel_size=s.voxel_size, inv_dir=inv,
                                    generator=torch.Generator().manual_seed(11))
noise = out["_dbg"]["noise"]
msd = {k: v.detach().to(device) for k, v in ms.items()}
decd = [p.detach().to(device) for p in dec]
pipe = RenderPipeline(rays_o.shape[1], device, samples_per_ray=96)
pipe.bind(rays_o.detach().to(device), rays_d.detach().to(device), msd, decd, voxel_size=s.voxel_size,
          step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0, target_rgb=rgb.to(device),
          target_depth=depth.to(device), noise=noise.reshape(-1, noise.shape[-1]).to(device).contiguous(), forward_only=True)
pipe.step()
o = pipe.outputs()
d = (o["depth"].cpu() - out["depth"].detach()).abs()
bad = torch.nonzero(d > 1e-4).view(-1)
print("bad rays", bad.tolist(), "of", d.numel())
for q in bad[:3].tolist():
    print("ray", q, "gpu depth", float(o["depth"][q]), "cpu", float(out["depth"][q]), "zmin gpu", float(o["raw"][q]), "cpu", float(out["raw"][q]))
    sg = o["sdf"][q].cpu(); sc_ = out["sdf"][q].detach()
    print(" sdf gpu", sg[:40].tolist())
    print(" sdf cpu", sc_[:40].tolist())
    print(" z", out["z_vals"][q][:40].tolist())
    print(" w gpu", o["weights"][q].cpu()[:40].tolist())
    print(" w cpu", out["weights"][q].detach()[:40].tolist())
