import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import _decoder_struct, DecoderGradT
dev = torch.device("cuda:0"); lib = _lib.lib()
n = 148 * 128 * 3
dec = [p.detach().to(dev) for p in ro.decoder_params(width=128, seed=2)]
feat = torch.randn(n, 16, device=dev) * 0.05
g_out = torch.randn(n, 4, device=dev)
ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev)
g_feat = torch.zeros(n, 16, device=dev)
ds = _decoder_struct(dec)
gd = [torch.zeros_like(p) for p in dec]; gs = _decoder_struct(gd, DecoderGradT)
wws = torch.empty(int(lib.pslam_wgrad_ws_bytes(n)), dtype=torch.uint8, device=dev)
buf = torch.zeros(640, dtype=torch.int64, device=dev)
def run(with_grad):
    return lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(g_out), _lib.ptr(g_feat),
                                 C.byref(gs) if with_grad else None, _lib.ptr(wws) if with_grad else None, wws.numel() if with_grad else 0, _lib.stream_ptr(dev))
for with_grad in (False, True):
    for rep in range(2): run(with_grad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(with_grad); b.record(); torch.cuda.synchronize()
    print("with_grad", with_grad, "ms", a.elapsed_time(b), "for", n, "samples")
    lib.pslam_debug_tc_trace(_lib.ptr(buf)); run(with_grad); torch.cuda.synchronize(); lib.pslam_debug_tc_trace(None)
    t = buf.cpu()[:320].view(4, 10, 8); t0 = int(t[1, 0, 6])
    for l in range(10):
        print("  layer", l, "A_seen", int(t[1, l, 1]) - t0, "committed", int(t[1, l, 2]) - t0, "D_seen", int(t[1, l, 3]) - t0, "A_next_produced", int(t[1, l + 1, 5]) - t0 if l < 9 else "-")
    print("  next tile gather", int(t[2, 0, 6]) - t0)
