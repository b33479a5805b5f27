"""Builds ``libproud_b200.so`` in-tree with nvcc for sm_100a (no JIT cache: the
built library travels with the repository snapshot).  Every source is compiled to
its own object (in parallel, only when it or a header changed) and linked."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libproud_b200.so")
SOURCES = ["api.cu", "intersect.cu", "sample.cu", "field.cu", "field_tc.cu", "field_bf.cu", "field_pp.cu", "field_bw.cu", "field_w256.cu",
           "composite.cu", "peer.cu", "pose.cu", "optim.cu", "octree_dev.cu", "octree_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC"]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "proud_slam_b200.h"))
    return deps


def _obj_of(src):
    return os.path.join(OBJ, os.path.basename(src) + ".o")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in sources() + _headers())


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())

    def compile_one(src):
        obj = _obj_of(src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        subprocess.check_call(cmd, cwd=CSRC)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(compile_one, sources()))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [_obj_of(s) for s in sources()], cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
