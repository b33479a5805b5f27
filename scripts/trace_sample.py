"""Per-warp timeline of k_sample_warp at the bench workload (pslam_debug_sample_trace): where the 40 us go."""
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
R = b[0].shape[0]
rpb = int(os.environ.get("PSLAM_SAMPLE_RPB", 0)) or (8 if (R + 15) // 16 < 148 else (16 if (R + 15) // 16 <= 8 * 148 else 32))   # sample.cu: sample_rays_per_block
nb = (R + rpb - 1) // rpb
buf = torch.zeros(nb * 8 * 8, dtype=torch.int64, device=dev)
pipe.stage(0); torch.cuda.synchronize()
lib.pslam_debug_sample_trace(_lib.ptr(buf)); pipe.stage(1); torch.cuda.synchronize(); lib.pslam_debug_sample_trace(None)
t = buf.cpu().view(nb * 8, 8)
t = t[t[:, 0] > 0]
g0 = int(t[:, 0].min())
print("warps", t.shape[0], "kernel span (globaltimer, ns): first entry -> last exit", int(t[:, 6].max()) - g0, "last entry at", int(t[:, 0].max()) - g0)
d = lambda a, c: (t[:, a] - t[:, c]).float()
for name, v in (("stage hits", d(2, 1)), ("sample rays", d(3, 2)), ("scan + look-back", d(4, 3)), ("copy-out", d(5, 4)), ("total clk", d(5, 1))):
    print(f"  {name:18s} clocks: mean {v.mean():8.0f}  median {v.median():8.0f}  max {v.max():8.0f}")
print("  largest sample count per warp: mean", t[:, 7].float().mean().item(), "max", int(t[:, 7].max()))
print("  per-warp wall (ns): mean", (t[:, 6] - t[:, 0]).float().mean().item(), "max", int((t[:, 6] - t[:, 0]).max()))
