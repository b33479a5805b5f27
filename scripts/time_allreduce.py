"""Times pslam_peer_allreduce alone (one process per GPU under torchrun): back-to-back launches, events, max over ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from proud_slam_b200.parallel import PeerExchange
n = int(sys.argv[1]) if len(sys.argv) > 1 else 385576
n = (n + 3) // 4 * 4
px = PeerExchange(n, dev)
px.flat.fill_(1.0)
dist.barrier(); torch.cuda.synchronize()
for _ in range(20): px.allreduce()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): px.allreduce()
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 200 * 1e3], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
x = torch.ones(n, device=dev)
dist.barrier(); torch.cuda.synchronize()
for _ in range(20): dist.all_reduce(x)
torch.cuda.synchronize()
a.record()
for _ in range(200): dist.all_reduce(x)
b.record(); torch.cuda.synchronize()
t2 = torch.tensor([a.elapsed_time(b) / 200 * 1e3], device=dev)
dist.all_reduce(t2, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"n={n} floats ({n * 4 / 1e6:.2f} MB) world={world}: peer all-reduce {t.item():.1f} us, NCCL all_reduce {t2.item():.1f} us, fail={int(px.fail.item())}")
dist.barrier(); dist.destroy_process_group()
