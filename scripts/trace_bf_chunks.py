import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import _decoder_struct
dev = torch.device("cuda:0"); lib = _lib.lib()
n = 148 * 128 * 10
dec = [p.detach().to(dev) for p in ro.decoder_params(width=128, seed=2)]
feat = torch.randn(n, 16, device=dev) * 0.05
ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev)
out = torch.empty(n, 4, device=dev)
ds = _decoder_struct(dec)
buf = torch.zeros(640, dtype=torch.int64, device=dev)
fwd = lambda: lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(dev))
fwd(); fwd(); torch.cuda.synchronize()
lib.pslam_debug_bf_trace(_lib.ptr(buf)); fwd(); torch.cuda.synchronize(); lib.pslam_debug_bf_trace(None)
t = buf.cpu()[320:].view(40, 8); t0 = int(t[0, 0])
for c in range(4):
    print("chunk", c, "wait_start", int(t[c,0])-t0, "wait_end", int(t[c,1])-t0, "issued6", int(t[c,2])-t0, "committed", int(t[c,3])-t0, "synced", int(t[c,4])-t0)
