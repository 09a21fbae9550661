"""In-tree build of libwassgpu.so (nvcc, sm_100a only) and of the C++ host executable.

    python -m wass_b200.build          # incremental
    python -m wass_b200.build --force

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwassgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", CXX]
# fp32/fp64 parity kernels: reproduce the reference's operation order literally (no FMA contraction)
PER_FILE = {"geom_kernels.cu": ["-fmad=false"], "capi_geom.cu": ["-fmad=false"], "rectify.cu": ["-fmad=false"], "resize_kernels.cu": ["-fmad=false"]}


def _newer(srcs, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cu + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in cu:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer([s] + [d for d in deps if not d.endswith(".cu")], o):
            cmd = [NVCC] + ARCH + FLAGS + PER_FILE.get(os.path.basename(s), []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _newer(objs, LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", CXX, "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        subprocess.check_call(cmd)
    build_host(force or bool(procs))
    return LIB


def build_host(force=False):
    """The drop-in `wass_stereo` executable (C++ host over the C ABI): wass_b200/bin/wass_stereo."""
    host = os.path.join(CSRC, "host")
    srcs = sorted(glob.glob(os.path.join(host, "*.cpp")))
    exe = os.path.join(HERE, "bin", "wass_stereo")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    deps = srcs + glob.glob(os.path.join(host, "*.hpp")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [LIB]
    if force or _newer(deps, exe):
        cmd = [CXX, "-O2", "-std=c++17", "-Wall", "-pthread", "-o", exe] + srcs + ["-L" + HERE, "-lwassgpu", "-lz", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
