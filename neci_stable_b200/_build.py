"""Build recipes for the native libraries (run by __graft_entry__.build()).

  libneci_gpu.so   CUDA engine + C ABI (include/neci_gpu.h), sm_100a only
  libneci_host.so  CPU host-side mirror of the Fortran host's setup code
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
GPU_LIB = os.path.join(HERE, "libneci_gpu.so")
HOST_LIB = os.path.join(HERE, "libneci_host.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_gpu(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("neci_gpu.cu", "kernels.cuh", "spawn_kernel.cuh", "device_system.cuh", "device_common.cuh")]
    srcs.append(os.path.join(ROOT, "include", "neci_gpu.h"))
    if not force and not _newer(GPU_LIB, srcs):
        return GPU_LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           # IEEE fp64 without contraction: the parity tests compare with a CPU restatement bit for bit
           "-fmad=false", "-shared", "-Xcompiler", "-fPIC", "-o", GPU_LIB, os.path.join(CSRC, "neci_gpu.cu"), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return GPU_LIB


def build_gpu_variant(tag, defines, verbose=False):
    """A tuning variant of the engine beside the shipped one: libneci_gpu_<tag>.so compiled with extra -D macros
    (K1_BLOCK, K1_CTAS_PER_SM, K1_SPT ...).  Select it at run time with NECI_GPU_LIB=<path> (capi.GPU_LIB).
    Used by profiles/tools/round2_first_call.sh to compare launch configurations in one GPU call."""
    out = os.path.join(HERE, "libneci_gpu_%s.so" % tag)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
           "-shared", "-Xcompiler", "-fPIC"] + ["-D%s" % d for d in defines] + ["-o", out, os.path.join(CSRC, "neci_gpu.cu"), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return out


def build_host(force=False):
    srcs = [os.path.join(CSRC, "host", f) for f in ("neci_host.cpp", "core_space.cpp")]
    if not force and not _newer(HOST_LIB, srcs):
        return HOST_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.check_call([cxx, "-O3", "-mpopcnt", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-shared", "-o", HOST_LIB] + srcs)
    return HOST_LIB


def build_all(force=False):
    build_host(force)
    build_gpu(force)
