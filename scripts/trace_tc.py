import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import _decoder_struct
dev = torch.device("cuda:0"); lib = _lib.lib()
n = 148 * 128 * 3
dec = [p.detach().to(dev) for p in ro.decoder_params(width=128, seed=2)]
feat = torch.randn(n, 16, device=dev) * 0.05
ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev)
out = torch.zeros(n, 4, device=dev)
ds = _decoder_struct(dec)
buf = torch.zeros(640, dtype=torch.int64, device=dev)
for rep in range(2):
    lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(dev))
torch.cuda.synchronize()
lib.pslam_debug_tc_trace(_lib.ptr(buf))
lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(dev))
torch.cuda.synchronize()
lib.pslam_debug_tc_trace(None)
t = buf.cpu()[:320].view(4, 10, 8)
t0 = int(t[0, 0, 6])
names = ["mma_wait_A", "mma_A_seen", "mma_committed", "wrk_D_seen", "-", "wrk_A_produced", "gather_start"]
for tile in range(3):
    print("tile", tile)
    for l in range(5):
        row = {names[s]: int(t[tile, l, s]) - t0 for s in (0, 1, 2, 3, 5, 6) if int(t[tile, l, s])}
        print("  L%d" % (l + 1), row)
