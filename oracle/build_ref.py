"""oracle/build_ref.py -- TEST INFRASTRUCTURE.

Builds the reference's OWN CUDA extension ``grid`` (third_party/sparse_voxels: ray/octree
intersection and inverse-CDF sampling kernels) for sm_100 from the sources where they lie under
/root/reference, into ``oracle/_ref/grid.so`` (git-ignored, but it travels to the GPU box with the
repository snapshot).  No reference source is copied and the reference's build system is not run:
this is one nvcc command over its six source files.

On the GPU box ``tests/test_gpu_ref_grid.py`` loads it and compares our kernels with the
reference's kernels bit for bit -- that is what pins parity for kernels 1-2 (the reference ships no
golden vectors and has no CPU implementation of them).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/third_party/sparse_voxels"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "grid.so")
SOURCES = ["src/binding.cpp", "src/intersect.cpp", "src/intersect_gpu.cu", "src/sample.cpp", "src/sample_gpu.cu", "src/octree.cpp"]


def available():
    return os.path.isdir(os.path.join(REF, "src"))


def build(force=False):
    if not available():
        raise RuntimeError("/root/reference is not mounted")
    if os.path.exists(OUT) and not force:
        return OUT
    from torch.utils import cpp_extension as ce
    import torch
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [os.path.join(REF, "include")] + ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-shared", "--compiler-options", "-fPIC",
            "-gencode", "arch=compute_100,code=sm_100", "-DTORCH_EXTENSION_NAME=grid", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), "--expt-relaxed-constexpr"]
           + ["-I" + i for i in inc] + [os.path.join(REF, s) for s in SOURCES]
           + ["-L" + libdir, "-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python",
              "-Xlinker", "-rpath," + libdir, "-o", OUT])
    subprocess.check_call(cmd)
    return OUT


def build_if_possible():
    if available():
        return build()
    return OUT if os.path.exists(OUT) else None


def load():
    """Imports oracle/_ref/grid.so as a module (needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(OUT):
        return None
    spec = importlib.util.spec_from_file_location("grid", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
