"""Kernels 3 and 4 in isolation through the C ABI: trilinear lookup and the decoder MLP,
forward and backward, against the oracle's torch restatement (fp32, 1e-4 relative)."""
import pytest
import torch

from oracle import render_oracle as ro
from proud_slam_b200 import _lib
from tests import util
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _dec_struct(params):
    from proud_slam_b200.pipeline import _decoder_struct
    return _decoder_struct(params)


@pytest.fixture
def decoder_build(request):
    """PSLAM_OPT_DECODER for the duration of one test: 0 = tcgen05 3xTF32, 1 = SIMT fp32, 2 = tcgen05 3xF16."""
    with util.decoder_build(request.param) as mode:
        yield mode


# ReLU'(0) margin: pre-activations closer to 0 than the build's own rounding error may flip a mask
NEAR = {0: 2e-6, 1: 2e-6, 2: 4e-6}


@pytest.mark.parametrize("decoder_build", [0, 2], indirect=True)
@pytest.mark.parametrize("use_ws", [True, False])   # True: tcgen05 wgrad (width 128); False: SIMT wgrad
@pytest.mark.parametrize("width,n", [(128, 1), (128, 64), (128, 1000), (256, 333), (128, 20000)])
def test_decoder_forward_backward(width, n, use_ws, decoder_build, device):
    from proud_slam_b200.pipeline import DecoderGradT, _decoder_struct
    import ctypes as C
    lib = _lib.lib()
    near_eps = NEAR[decoder_build]
    dec = ro.decoder_params(width=width, seed=2)
    g = torch.Generator().manual_seed(n)
    feat = (torch.randn(n, 16, generator=g) * 0.05).requires_grad_(True)
    rgb, sdf = ro.decoder_forward(dec, feat)
    g_out = torch.randn(n, 4, generator=g)
    # ReLU'(0) is a hard decision: a pre-activation within fp32 rounding of 0 flips one unit's mask
    # between two correct implementations and changes that sample's gradient by O(1).  Samples
    # with such a unit get zero upstream gradient on both sides (they then contribute nothing).
    with torch.no_grad():
        W1, b1, W2, b2, W3, b3, W4, b4, W5, b5 = dec
        a1 = feat @ W1.t() + b1
        a2 = torch.relu(a1) @ W2.t() + b2
        t = (torch.relu(a2) @ W3.t() + b3)[:, 1:]
        a4 = torch.cat([t, feat], 1) @ W4.t() + b4
        near = (a1.abs().min(1).values < near_eps) | (a2.abs().min(1).values < near_eps) | (a4.abs().min(1).values < near_eps)
        g_out[near] = 0.0
    (torch.cat([rgb, sdf[:, None]], 1) * g_out).sum().backward()

    decd = [p.detach().to(device) for p in dec]
    featd = feat.detach().to(device)
    ws = torch.empty(int(lib.pslam_decoder_ws_count(width)), device=device)
    out = torch.empty(n, 4, device=device)
    ds = _decoder_struct(decd)
    _lib.check(lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(featd), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(device)), "fwd")
    assert rel_err(out[:, :3], rgb.detach()) < TOL
    assert rel_err(out[:, 3], sdf.detach()) < TOL

    gd = [torch.zeros_like(p) for p in decd]
    gs = _decoder_struct(gd, DecoderGradT)
    g_feat = torch.empty(n, 16, device=device)
    g_outd = g_out.to(device)
    wws = torch.empty(int(lib.pslam_wgrad_ws_bytes(n)), dtype=torch.uint8, device=device) if use_ws else None
    _lib.check(lib.pslam_decoder_bwd(n, C.byref(ds), _lib.ptr(featd), _lib.ptr(ws), _lib.ptr(g_outd), _lib.ptr(g_feat),
                                     C.byref(gs), _lib.ptr(wws), 0 if wws is None else wws.numel(), _lib.stream_ptr(device)), "bwd")
    torch.cuda.synchronize()
    assert rel_err(g_feat, feat.grad) < TOL
    for i in range(10):
        assert rel_err(gd[i], dec[i].grad) < TOL, f"param {i}"


def test_decoder_rejects_bad_width(device):
    import ctypes as C
    from proud_slam_b200.pipeline import DecoderT
    ds = DecoderT()
    ds.width = 64
    rc = _lib.lib().pslam_decoder_fwd(4, C.byref(ds), None, None, None, None)
    assert rc < 0
    assert b"width" in _lib.lib().pslam_last_error()


@pytest.mark.parametrize("n", [1, 777, 50000])
def test_trilinear_forward_backward(n, device):
    lib = _lib.lib()
    s, ms = util.build_scene("tiny", emb_scale=0.5)
    leaf = torch.nonzero((ms["voxel_vertex_idx"] >= 0).all(-1)).view(-1)
    g = torch.Generator().manual_seed(n)
    vox = leaf[torch.randint(0, leaf.numel(), (n,), generator=g)]
    xyz = (ms["voxel_center_xyz"].detach()[vox] + (torch.rand(n, 3, generator=g) - 0.5) * s.voxel_size).requires_grad_(True)
    f = ro.get_features_vox(xyz, vox, ms, s.voxel_size)
    g_f = torch.randn(n, 16, generator=g)
    (f * g_f).sum().backward()

    d = lambda t: t.detach().to(device).contiguous()
    feat = torch.empty(n, 16, device=device)
    args = (n, _lib.ptr(d(xyz)), _lib.ptr(d(vox.int())), _lib.ptr(d(ms["voxel_center_xyz"])), _lib.ptr(d(ms["voxel_vertex_idx"])),
            _lib.ptr(d(ms["voxel_vertex_emb"])), float(s.voxel_size))
    keep = [d(xyz), d(vox.int()), d(ms["voxel_center_xyz"]), d(ms["voxel_vertex_idx"]), d(ms["voxel_vertex_emb"])]
    args = (n,) + tuple(_lib.ptr(t) for t in keep) + (float(s.voxel_size),)
    _lib.check(lib.pslam_trilinear_fwd(*args, _lib.ptr(feat), _lib.stream_ptr(device)), "tri fwd")
    assert rel_err(feat, f.detach()) < TOL
    g_emb = torch.zeros_like(keep[4])
    g_xyz = torch.empty(n, 3, device=device)
    g_fd = g_f.to(device)
    _lib.check(lib.pslam_trilinear_bwd(*args, _lib.ptr(g_fd), _lib.ptr(g_emb), _lib.ptr(g_xyz), _lib.stream_ptr(device)), "tri bwd")
    torch.cuda.synchronize()
    assert rel_err(g_emb, ms["voxel_vertex_emb"].grad) < TOL
    assert rel_err(g_xyz, xyz.grad) < TOL


@pytest.mark.parametrize("N,K", [(16, 8), (128, 16), (128, 128), (144, 128), (128, 144), (16, 128)])
@pytest.mark.parametrize("split3", [0, 1])
def test_umma_gemm_primitives(N, K, split3, device):  # modes 0/1 of the debug kernel
    """tcgen05.mma (A from tensor memory, B through a shared-memory descriptor), TMEM ld/st and the
    operand layouts of csrc/umma.cuh, against an fp64 matmul.  1xTF32 ~1e-3, 3xTF32 ~fp32."""
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = (A.double() @ B.double().t())
    Ad, Bd = A.to(device), B.to(device)
    D = torch.zeros(128, N, device=device)
    _lib.check(_lib.lib().pslam_debug_umma_gemm(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, split3, _lib.stream_ptr(device)), "umma")
    torch.cuda.synchronize()
    err = rel_err(D, ref)
    assert err < (2e-6 if split3 else 3e-3), err
    if not split3:
        assert err > 1e-6    # really went through TF32 tensor cores


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_decoder_forward_both_builds(mode, device):
    """tcgen05 (3xTF32 / 3xF16) and SIMT fp32 builds of the width-128 decoder agree with the oracle."""
    import ctypes as C
    from proud_slam_b200.pipeline import _decoder_struct
    lib = _lib.lib()
    with util.decoder_build(mode):
        dec = ro.decoder_params(width=128, seed=4)
        n = 5000
        feat = torch.randn(n, 16, generator=torch.Generator().manual_seed(1)) * 0.05
        rgb, sdf = ro.decoder_forward([p.detach() for p in dec], feat)
        decd = [p.detach().to(device) for p in dec]
        ws = torch.empty(int(lib.pslam_decoder_ws_count(128)), device=device)
        out = torch.empty(n, 4, device=device)
        ds = _decoder_struct(decd)
        featd = feat.to(device)
        _lib.check(lib.pslam_decoder_fwd(n, C.byref(ds), _lib.ptr(featd), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr(device)), "fwd")
        torch.cuda.synchronize()
        tol = 2e-5 if mode == 2 else 1e-5
        assert rel_err(out[:, :3], rgb) < tol
        assert rel_err(out[:, 3], sdf) < tol


@pytest.mark.parametrize("N,K", [(16, 8), (16, 32), (128, 32), (144, 64)])
def test_umma_gemm_both_operands_from_smem(N, K, device):
    """The wgrad form: A and B both through shared-memory descriptors (K-major), 3xTF32."""
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    Ad, Bd = A.to(device), B.to(device)
    D = torch.zeros(128, N, device=device)
    _lib.check(_lib.lib().pslam_debug_umma_gemm(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, 4, _lib.stream_ptr(device)), "umma ss")
    torch.cuda.synchronize()
    assert rel_err(D, ref) < 2e-6


@pytest.mark.parametrize("N,K", [(16, 16), (128, 16), (128, 128), (144, 128), (128, 144), (16, 128)])
def test_umma_f16_gemm_a_from_tmem(N, K, device):
    """kind::f16 MMAs of the 3xF16 build: A packed two-per-column in tensor memory, B K-major in shared memory."""
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    Ad, Bd = A.to(device), B.to(device)
    D = torch.zeros(128, N, device=device)
    _lib.check(_lib.lib().pslam_debug_umma_gemm_bf(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, 0, _lib.stream_ptr(device)), "umma bf ts")
    torch.cuda.synchronize()
    err = rel_err(D, ref)
    assert err < 2e-6, err



@pytest.mark.parametrize("N,K", [(16, 16), (16, 64), (128, 16), (128, 64), (144, 32)])
def test_umma_f16_gemm_mn_major_smem(N, K, device):
    """The 3xF16 wgrad form: reduction over samples, both operands MN-major (sample-major) in shared memory."""
    g = torch.Generator().manual_seed(N + K)
    At = torch.randn(K, 128, generator=g)
    Bt = torch.randn(K, N, generator=g)
    ref = At.double().t() @ Bt.double()
    Ad, Bd = At.to(device), Bt.to(device)
    D = torch.zeros(128, N, device=device)
    _lib.check(_lib.lib().pslam_debug_umma_gemm_bf(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), N, K, 1, _lib.stream_ptr(device)), "umma bf ss")
    torch.cuda.synchronize()
    err = rel_err(D, ref)
    assert err < 2e-6, err
