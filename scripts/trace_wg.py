import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import _decoder_struct, DecoderGradT
dev = torch.device("cuda:0"); lib = _lib.lib()
n = 148 * 128 * 3
dec = [p.detach().to(dev) for p in ro.decoder_params(width=128, seed=2)]
feat = torch.randn(n, 16, device=dev) * 0.05; g_out = torch.randn(n, 4, device=dev)
ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev); g_feat = torch.zeros(n, 16, device=dev)
ds = _decoder_struct(dec); gd = [torch.zeros_like(p) for p in dec]; gs = _decoder_struct(gd, DecoderGradT)
wws = torch.empty(int(lib.pslam_wgrad_ws_bytes(n)), dtype=torch.uint8, device=dev)
buf = torch.zeros(640, dtype=torch.int64, device=dev)
run = lambda: lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(feat), _lib.ptr(ws), _lib.ptr(g_out), _lib.ptr(g_feat), C.byref(gs), _lib.ptr(wws), wws.numel(), _lib.stream_ptr(dev))
run(); run(); torch.cuda.synchronize()
lib.pslam_debug_tc_trace(_lib.ptr(buf)); run(); torch.cuda.synchronize(); lib.pslam_debug_tc_trace(None)
t = buf.cpu()[320:].view(40, 8)
# k_field_tc<1> wrote into the same buffer first (tile/layer stamps), then k_wgrad_tc overwrote g<40 slots 0..5
t0 = int(t[12, 0])
for g in range(12, 30):
    print(g, "prod_issue", int(t[g,0])-t0, "raw_seen", int(t[g,1])-t0, "opfree_seen", int(t[g,2])-t0, "xform_done", int(t[g,3])-t0, "mma_seen", int(t[g,4])-t0, "mma_commit", int(t[g,5])-t0)
