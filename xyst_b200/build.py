"""Build recipes for the native libraries (nvcc for sm_100a, g++ for the host mirror).

    python -m xyst_b200.build            # build everything in-tree

The built .so files live next to this file; they are git-ignored but travel to the
GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fopenmp",
              "-Xcompiler", "-fPIC", "-shared", "-Xlinker", "-soname=libxyst_b200.so",
              "-I" + os.path.join(ROOT, "include")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_device(force=False, verbose=False, out=None, defines=()):
    """libxyst_b200.so: CUDA kernels + the C ABI of include/xyst_b200.h."""
    src = [os.path.join(HERE, "csrc", "xyst_b200.cu")]
    cdir = os.path.join(HERE, "csrc")
    dep = src + [os.path.join(cdir, f) for f in os.listdir(cdir) if f.endswith(".cuh")] + \
        [os.path.join(ROOT, "include", "xyst_b200.h")]
    out = out or os.path.join(HERE, "libxyst_b200.so")
    if force or _stale(out, dep):
        cmd = [NVCC] + CUDA_FLAGS + ["-D" + d for d in defines] + \
            (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + src + ["-ldl"]
        subprocess.run(cmd, check=True)
    return out


def build_host(force=False):
    """libxyst_host.so: C++ host mirror of the reference's solver-side interface."""
    hdir = os.path.join(HERE, "host")
    src = sorted(os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".cpp")) \
        if os.path.isdir(hdir) else []
    if not src:
        return None
    dep = src + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".hpp")] + \
        [os.path.join(ROOT, "include", "xyst_b200.h"), os.path.join(ROOT, "include", "xyst_host.h")]
    out = os.path.join(HERE, "libxyst_host.so")
    if force or _stale(out, [d for d in dep if os.path.exists(d)]):
        cmd = ["g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-fopenmp", "-Wall",
               "-I" + os.path.join(ROOT, "include"), "-o", out] + src + \
              ["-L" + HERE, "-lxyst_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return out


def build_all(force=False):
    return [build_device(force), build_host(force)]


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
