"""Timeline of CTA 0 of the 3xBF16 decoder kernels (clock64 stamps per role, pslam_debug_bf_trace) and
stand-alone timings of forward / backward without and with the wgrad spill.  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import _decoder_struct, DecoderGradT
dev = torch.device("cuda:0"); lib = _lib.lib()
tiles_per_cta = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = 148 * 128 * tiles_per_cta
dec = [p.detach().to(dev) for p in ro.decoder_params(width=128, seed=2)]
feat = torch.randn(n, 16, device=dev) * 0.05
g_out = torch.randn(n, 4, device=dev)
ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev)
out = torch.empty(n, 4, device=dev)
g_feat = torch.zeros(n, 16, device=dev)
ds = _decoder_struct(dec)
gd = [torch.zeros_like(p) for p in dec]; gs = _decoder_struct(gd, DecoderGradT)
wws = torch.empty(int(lib.pslam_wgrad_ws_bytes(n)), dtype=torch.uint8, device=dev)
buf = torch.zeros(640, dtype=torch.int64, device=dev)
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
def fwd():
    return lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(dev))
def bwd(with_grad):
    return lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(g_out), _lib.ptr(g_feat),
                                 C.byref(gs) if with_grad else None, _lib.ptr(wws) if with_grad else None, wws.numel() if with_grad else 0, _lib.stream_ptr(dev))
def timed(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
def trace(fn, layers):
    """slots: 0 = MMA thread starts waiting for the layer's A, 2 = its MMAs are issued and committed, 3 = workers see the
    accumulators, 7 = staging decided, 4 = epilogue done, 5 = last quarter of the next A published, 6 = tile starts"""
    lib.pslam_debug_bf_trace(_lib.ptr(buf)); fn(); torch.cuda.synchronize(); lib.pslam_debug_bf_trace(None)
    t = buf.cpu()[:320].view(4, 10, 8); t0 = int(t[1, 0, 6])
    for l in range(layers):
        print("  layer", l, "mma_wait", int(t[1, l, 0]) - t0, "issued", int(t[1, l, 2]) - t0, "D_seen", int(t[1, l, 3]) - t0,
              "epi_done", int(t[1, l, 4]) - t0, "A_next_produced", int(t[1, l + 1, 5]) - t0 if l < layers - 1 else "-")
    print("  next tile starts", int(t[2, 0, 6]) - t0)
print("samples", n, "= tiles/CTA", tiles_per_cta)
print("fwd ms", timed(fwd)); trace(fwd, 5)
for with_grad in (False, True):
    print("bwd with_grad", with_grad, "ms (dgrad + wgrad kernels)", timed(lambda: bwd(with_grad))); trace(lambda: bwd(with_grad), 10)
if True:
    lib.pslam_debug_bf_trace(_lib.ptr(buf)); bwd(True); torch.cuda.synchronize(); lib.pslam_debug_bf_trace(None)
    t = buf.cpu()[320:].view(40, 8)
    t0 = int(t[0, 0])
    print("wgrad steps (producer issue, mma full seen, mma issued):")
    for g in range(0, 24):
        print("  step", g, int(t[g, 0]) - t0, int(t[g, 4]) - t0, int(t[g, 5]) - t0)
