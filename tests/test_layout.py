"""Repository contract checks that need no GPU: the product never touches oracle/, the C-ABI
library loads and exports every symbol the header declares, and the Python mirror matches."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "proud_slam_b200")


def _py_files(d):
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                yield os.path.join(base, f)


def test_product_never_imports_oracle_or_reference():
    bad = []
    for path in _py_files(PKG):
        src = open(path).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "/root/reference" in src or "liboracle" in src:
            bad.append(path)
    assert not bad, bad


def test_gpu_entry_points_do_not_read_reference_at_runtime():
    for name in ("bench.py", "__graft_entry__.py"):
        src = open(os.path.join(ROOT, name)).read()
        assert "ref_import" not in src, name


def test_no_forbidden_layers():
    """No Triton / torch.compile / CPU fallback in the product."""
    for path in _py_files(PKG):
        src = open(path).read()
        assert "import triton" not in src and "torch.compile" not in src, path


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "proud_slam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pslam_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from proud_slam_b200 import _build
    path = _build.build_library()
    h = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/proud_slam_b200.h but not exported"
    assert h.pslam_abi_version() == 2


def test_process_wide_options_are_validated():
    """pslam_set_option (no device work): the documented keys / values are accepted, anything else is an error with a message."""
    from proud_slam_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "proud_slam_b200.h")).read()
    keys = {name: int(val) for name, val in re.findall(r"#define (PSLAM_OPT_[A-Z_]+) (\d+)", hdr)}
    assert keys == {"PSLAM_OPT_DECODER": 1, "PSLAM_OPT_SAVE_ACT": 2, "PSLAM_OPT_PDL": 3, "PSLAM_OPT_TILES": 4, "PSLAM_OPT_FUSED_WGRAD": 5,
                    "PSLAM_OPT_FUSED_SCATTER": 6, "PSLAM_OPT_WALK": 7}
    try:
        for key, values in ((1, (0, 1, 2)), (2, (0, 1)), (3, (0, 1)), (4, (1, 2)), (5, (0, 1))):
            for v in values:
                assert lib.pslam_set_option(key, v) == 0
            assert lib.pslam_set_option(key, 7) != 0
            assert b"unknown option" in lib.pslam_last_error()
        assert lib.pslam_set_option(99, 0) != 0
    finally:
        for key, v in ((1, 2), (2, 1), (3, 1), (4, 2), (5, 1)):   # the defaults
            lib.pslam_set_option(key, v)


def test_python_mirror_matches_header_and_struct():
    from proud_slam_b200 import _lib
    lib = _lib.lib()   # checks sizeof / offsetof of pslam_render_t against the ctypes mirror
    declared = set(_header_functions())
    assert set(_lib.exported_symbols()) <= declared
    assert lib.pslam_decoder_ws_count(128) == 128 * 128 + 288 * 128 + 229376   # SIMT pack + tcgen05 weight stream
    assert lib.pslam_decoder_ws_count(256) == 256 * 256 + 288 * 256 + 278528 + 16   # SIMT pack + the width-256 tcgen05 stream (field_w256.cu)
    assert lib.pslam_decoder_ws_count(7) == -1
    assert lib.pslam_render_scratch_i_count(8192) > 0


def test_argument_errors_do_not_need_a_gpu():
    """Bad arguments are rejected before any CUDA call, with a message (never abort/exit)."""
    from proud_slam_b200 import _lib
    lib = _lib.lib()
    rc = lib.pslam_svo_intersect(0, 1, 1, 0.2, 50, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"positive" in lib.pslam_last_error()
    rc = lib.pslam_inverse_cdf_sampling(1, 1, 1, 1, -1.0, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.pslam_last_error()
    with pytest.raises(RuntimeError, match="failed"):
        _lib.check(rc, "inverse_cdf_sampling")


def test_every_cu_cites_the_reference():
    for f in os.listdir(os.path.join(PKG, "csrc")):
        if f.endswith((".cu", ".cpp")):
            src = open(os.path.join(PKG, "csrc", f)).read()
            assert re.search(r"\.(cu|cpp|py|h):\d+", src), f"{f} cites no reference file:line"
