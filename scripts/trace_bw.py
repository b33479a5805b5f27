"""Timeline of CTA 0 of the fused backward kernel (csrc/field_bw.cu, pslam_debug_bw_trace) on the bench workload.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from proud_slam_b200 import _lib
from proud_slam_b200.parallel import FlatGrads
from proud_slam_b200.pipeline import RenderPipeline

dev = torch.device("cuda:0")
lib = _lib.lib()
s, ms_cpu, batch, n_oct, n_vox = bench.build_workload(0)
ms = {k: v.to(dev).contiguous() for k, v in ms_cpu.items()}
dec = bench.decoder_params(128, dev)
fg = FlatGrads(ms["voxel_vertex_emb"], dec)
inp = [t.to(dev) for t in batch]
pipe = RenderPipeline(inp[0].shape[0], dev, samples_per_ray=64)
pipe.bind(inp[0], inp[1], ms, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
          target_rgb=inp[2], target_depth=inp[3], seed=1, weights=bench.CRIT_W, g_emb=fg.g_emb, g_dec=fg.g_dec, grad_rays=True)
for _ in range(3):
    pipe.step()
torch.cuda.synchronize()
print(pipe.counts())
buf = torch.zeros(4 * 2 * 16, dtype=torch.int64, device=dev)
for k in range(5):
    pipe.stage(k)
torch.cuda.synchronize()
lib.pslam_debug_bw_trace(_lib.ptr(buf))
pipe.stage(9)
torch.cuda.synchronize()
lib.pslam_debug_bw_trace(None)
t = buf.cpu().view(4, 2, 16)
for it in (0, 1, 2):
    t0 = int(t[it, 0, 0])
    w = [int(x) - t0 for x in t[it, 0, :11]]
    i = [int(x) - t0 for x in t[it, 1, :15]]
    print(f"tile {it}: worker: start 0, g_hc published {w[1]}")
    for ph in range(5):
        print(f"   phase {ph}: issuer saw operand {i[3 * ph]}, chain issued {i[3 * ph + 1]} | worker saw accumulators {w[2 + 2 * ph]}"
              + (f", epilogue published {w[3 + 2 * ph]}" if ph < 4 else ""))
    print(f"   issuer done with the tile {i[14]}; next tile starts {int(t[it + 1, 0, 0]) - t0}")
e = [int(x) - int(t[3, 0, 11]) for x in t[3, 0, 11:16]]
print(f"CTA 0: entry 0, worker loop starts {e[1]}, all MMAs done {e[2]}, drain done {e[3]}, cluster released {e[4]} clocks")
print("tile starts of CTA 0 (clocks after entry):", [int(t[i, 0, 0]) - int(t[3, 0, 11]) for i in range(4)])
