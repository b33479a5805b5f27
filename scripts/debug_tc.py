import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from proud_slam_b200.pipeline import DecoderGradT, _decoder_struct
dev = torch.device("cuda:0"); lib = _lib.lib()
torch.set_printoptions(precision=4, linewidth=200)
for mode in (4, 2):
    for (N, K) in ((16, 8), (16, 32)):
        if mode == 2 and K % 32: continue
        g = torch.Generator().manual_seed(1)
        if mode == 4:
            A = torch.randn(128, K, generator=g); B = torch.randn(N, K, generator=g); ref = A.double() @ B.double().t()
        else:
            A = torch.randn(K, 128, generator=g); B = torch.randn(K, N, generator=g); ref = A.double().t() @ B.double()
        D = torch.full((128, N), -7.0, device=dev)
        rc = lib.pslam_debug_umma_gemm(_lib.ptr(A.to(dev)), _lib.ptr(B.to(dev)), _lib.ptr(D), N, K, mode, _lib.stream_ptr(dev))
        torch.cuda.synchronize()
        Dc = D.cpu().double()
        ok = ((Dc - ref).abs() < 1e-3 * ref.abs().max())
        print("mode", mode, N, K, "rc", rc, "frac ok", float(ok.float().mean()), "ok rows", torch.nonzero(ok.all(1)).view(-1).tolist()[:20],
              "ok cols", torch.nonzero(ok.all(0)).view(-1).tolist())
        print(" D[0:3,:6]", Dc[0:3, :6].tolist()); print(" R[0:3,:6]", ref[0:3, :6].tolist())
        if mode == 4:
            # hypotheses: which A rows does the result row 0 use?
            for r in range(0, 16):
                cand = (A[r].double() @ B.double().t())
                if (cand - Dc[0]).abs().max() < 1e-2: print("  D row 0 == A row", r)
# fwd / bwd on a process whose TMEM holds unrelated data
for n in (64, 300, 1000):
    dec = ro.decoder_params(width=128, seed=2)
    g = torch.Generator().manual_seed(n)
    feat = (torch.randn(n, 16, generator=g) * 0.05).requires_grad_(True)
    rgb, sdf = ro.decoder_forward(dec, feat)
    g_out = torch.randn(n, 4, generator=g)
    (torch.cat([rgb, sdf[:, None]], 1) * g_out).sum().backward()
    decd = [p.detach().to(dev) for p in dec]
    ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=dev)
    ds = _decoder_struct(decd)
    out = torch.zeros(n, 4, device=dev)
    lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(feat.detach().to(dev)), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    e = (out.cpu() - torch.cat([rgb, sdf[:, None]], 1).detach()).abs().max(1).values
    print("n", n, "FWD bad rows", torch.nonzero(e > 1e-4).view(-1).tolist()[:10])
    # pollute TMEM with another problem
    lib.pslam_debug_umma_gemm(_lib.ptr(torch.randn(128, 128, device=dev)), _lib.ptr(torch.randn(128, 128, device=dev)), _lib.ptr(torch.zeros(128, 128, device=dev)), 128, 128, 1, _lib.stream_ptr(dev))
    for rep in range(3):
        g_feat = torch.zeros(n, 16, device=dev)
        rc = lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(feat.detach().to(dev)), _lib.ptr(ws), _lib.ptr(g_out.to(dev)), _lib.ptr(g_feat),
                                   None, None, 0, _lib.stream_ptr(dev))
        torch.cuda.synchronize()
        err = (g_feat.cpu() - feat.grad).abs().max(1).values / feat.grad.abs().max()
        bad = torch.nonzero(err > 1e-4).view(-1)
        print("n", n, "BWD rep", rep, "max row err", float(err.max()), "bad rows", bad.numel(), bad[:8].tolist(), bad[-3:].tolist())
        if bad.numel():
            r = int(bad[0]); print("   row", r, "got", g_feat[r, :6].cpu().tolist(), "want", feat.grad[r, :6].tolist())
