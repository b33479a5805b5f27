"""Timeline of CTA 0 of the saving forward (kFwdSave) and the saved backward (kBwdSaved) inside the fused pipeline at the
bench workload: per layer, when the accumulators were seen, when the staging buffer was free, when the epilogue was done."""
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
buf = torch.zeros(640, dtype=torch.int64, device=dev)
for stage, name, layers in ((2, "kFwdSave", range(0, 5)), (6, "kBwdSaved", range(5, 10))):
    for k in range(stage if stage == 2 else 5): pipe.stage(k)
    torch.cuda.synchronize(); buf.zero_()
    lib.pslam_debug_bf_trace(_lib.ptr(buf)); pipe.stage(stage); torch.cuda.synchronize(); lib.pslam_debug_bf_trace(None)
    t = buf.cpu()[:320].view(4, 10, 8)
    for tile in (1, 3):
        t0 = int(t[tile, layers[0], 0])
        print(name, "tile", tile, "(t0 = MMA warp waits for the tile's first A)")
        for l in layers:
            print("   layer", l, "mma_wait", int(t[tile, l, 0]) - t0, "issued", int(t[tile, l, 2]) - t0, "D_seen", int(t[tile, l, 3]) - t0,
                  "stage_free", int(t[tile, l, 7]) - t0, "epi_done", int(t[tile, l, 4]) - t0)
    print("   tile period", int(t[2, layers[0], 0]) - int(t[1, layers[0], 0]), int(t[3, layers[0], 0]) - int(t[2, layers[0], 0]))
