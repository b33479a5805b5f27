"""Pure-write / pure-read / copy HBM bandwidth on this GPU (torch fill, sum, copy over 2 GiB), GPU only."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29   # floats = 2 GiB
x = torch.empty(n, device=dev); y = torch.empty(n, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best * 1e-3
print("write (fill)  GB/s", 4 * n / t(lambda: x.zero_()) / 1e9)
print("read (sum)    GB/s", 4 * n / t(lambda: x.sum()) / 1e9)
print("copy (r+w)    GB/s", 8 * n / t(lambda: y.copy_(x)) / 1e9)
