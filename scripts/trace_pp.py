"""Timeline of CTA 0 of the two-tiles-in-flight decoder kernels (csrc/field_pp.cu, pslam_debug_pp_trace) on the bench
workload: clock64 stamps of the two worker groups and of the MMA-issuing thread, per iteration.  GPU only."""
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
buf = torch.zeros(1024, dtype=torch.int64, device=dev)


def show(stage, nph, title):
    for k in range(min(stage, 5)):
        pipe.stage(k)
    torch.cuda.synchronize()
    buf.zero_()
    lib.pslam_debug_pp_trace(_lib.ptr(buf))
    pipe.stage(stage)
    torch.cuda.synchronize()
    lib.pslam_debug_pp_trace(None)
    span = buf.cpu()[256:256 + 296].view(148, 2)
    t = buf.cpu()[:256].view(4, 4, 16)
    print("==", title)
    print("  iteration starts:", [int(t[i, 0, 0]) - int(t[0, 0, 0]) for i in range(4)])
    for it in (1, 2):
        t0 = int(t[it, 0, 0])
        for g in (0, 1):
            w = [int(x) - t0 for x in t[it, g, : 2 + 2 * nph]]
            i = [int(x) - t0 for x in t[it, 2 + g, : 2 * nph]]
            print(f"  it {it} group {g}: start {w[0]} published {w[1]} | " +
                  " | ".join(f"ph{k}: A_seen {i[2 * k]} issued {i[2 * k + 1]} D_seen {w[2 + 2 * k]} epi_done {w[3 + 2 * k] if 3 + 2 * k < len(w) else '-'}"
                             for k in range(nph)))
        print("  next iteration starts", int(t[it + 1, 0, 0]) - t0)
    clk, ns = int(t[3, 3, 13]) - int(t[3, 3, 12]), int(t[3, 3, 15]) - int(t[3, 3, 14])
    print(f"  CTA 0 worker span: {clk} clocks in {ns} ns = {clk / max(ns, 1):.3f} GHz")
    t00 = int(span[:, 0].min())
    st, en = (span[:, 0] - t00).tolist(), (span[:, 1] - t00).tolist()
    print("  CTA entry ns: min %d max %d | exit ns: min %d median %d max %d | CTA 0: %d..%d, worker loop starts at %d"
          % (min(st), max(st), min(en), sorted(en)[74], max(en), st[0], en[0], int(t[3, 3, 14]) - t00))


def timed_stage(stage, reps=5):
    ts = []
    for _ in range(reps):
        for k in range(min(stage, 5)):
            pipe.stage(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.stage(stage); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


show(8, 4, "forward (kFwdSave)")
show(9, 5, "backward (kBwdSaved)")
print("kernel ms: fwd", timed_stage(8), "bwd", timed_stage(9))
# the same without decoder gradients: ReLU masks only, nothing spilled
pipe.bind(inp[0], inp[1], ms, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
          target_rgb=inp[2], target_depth=inp[3], seed=1, weights=bench.CRIT_W, g_emb=fg.g_emb, g_dec=None, grad_rays=True)
for _ in range(2):
    pipe.step()
torch.cuda.synchronize()
show(8, 4, "forward, no spill")
show(9, 5, "backward, no spill")
print("kernel ms (no spill): fwd", timed_stage(8), "bwd", timed_stage(9))
