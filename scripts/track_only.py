"""The tracking leg of bench.py alone (for ncu --graph-profiling node launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
s, ms_cpu, batch, n_oct, n_vox = bench.build_workload(0)
ms = {k: v.to(dev).contiguous() for k, v in ms_cpu.items()}
dec = bench.decoder_params(128, dev)
print(bench.tracking_bench(s, ms, dec, dev, frames=int(sys.argv[1]) if len(sys.argv) > 1 else 5))
