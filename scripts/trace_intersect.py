"""Per-warp timeline of k_intersect_warp at the bench workload (pslam_debug_intersect_trace): walk vs sort vs write-out."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import RenderPipeline
from proud_slam_b200.parallel import FlatGrads
dev = torch.device("cuda:0"); lib = _lib.lib()
s, ms_cpu, batch, n_oct, n_vox = bench.build_workload(0)
ms = {k: v.to(dev).contiguous() for k, v in ms_cpu.items()}
dec = bench.decoder_params(128, dev)
fg = FlatGrads(ms["voxel_vertex_emb"], dec)
b = [t.to(dev) for t in batch]
pipe = RenderPipeline(b[0].shape[0], dev, samples_per_ray=64)
pipe.bind(b[0], b[1], ms, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0, max_depth=10.0,
          target_rgb=b[2], target_depth=b[3], noise=None, seed=1, weights=bench.CRIT_W, g_emb=fg.g_emb, g_dec=fg.g_dec, grad_rays=True)
for _ in range(3): pipe.step()
torch.cuda.synchronize()
rpw = int(os.environ.get('PSLAM_INTERSECT_RPW', 0)) or (1 if b[0].shape[0] <= 64 * 148 else (2 if b[0].shape[0] <= 256 * 148 else 4))   # intersect.cu: rays per warp
nb = (b[0].shape[0] + 8 * rpw - 1) // (8 * rpw)
buf = torch.zeros(nb * 8 * 8, dtype=torch.int64, device=dev)
lib.pslam_debug_intersect_trace(_lib.ptr(buf)); pipe.stage(0); torch.cuda.synchronize(); lib.pslam_debug_intersect_trace(None)
t = buf.cpu().view(nb * 8, 8)
t = t[t[:, 0] > 0]
g0 = int(t[:, 0].min())
print("warps", t.shape[0], "kernel span (globaltimer, ns): first entry -> last exit", int(t[:, 5].max()) - g0, "last entry at", int(t[:, 0].max()) - g0)
d = lambda a, c: (t[:, a] - t[:, c]).float()
for name, v in (("walk", d(2, 1)), ("rank + write", d(3, 2)), ("block tail", d(4, 3)), ("total clk", d(4, 1))):
    print(f"  {name:14s} clocks: mean {v.mean():8.0f}  median {v.median():8.0f}  max {v.max():8.0f}")
trips = t[:, 6].float()
print("  trips per ray: mean", trips.mean().item(), "max", int(trips.max()), "| clocks per expansion:", (d(2, 1) / trips.clamp(min=1)).mean().item())
print("  hits per ray: mean", t[:, 7].float().mean().item(), "max", int(t[:, 7].max()))
