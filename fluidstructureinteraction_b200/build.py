"""Build recipe for the native parts (run by __graft_entry__.build()).

Everything is compiled in-tree so the shared objects travel to the GPU box with the snapshot:
  fluidstructureinteraction_b200/libgpupcg_cuda.so   nvcc, sm_100a only (the product)
  fluidstructureinteraction_b200/host/libfoamlike.so  C++ host mirror of the foam-extend solver API
  oracle/liboracle.so                                 gcc, the CPU restatement (test infrastructure)
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fluidstructureinteraction_b200")
CSRC = os.path.join(PKG, "csrc")
CUDA_LIB = os.path.join(PKG, "libgpupcg_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",              # belt and braces: the kernels already use __d*_rn intrinsics
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def cuda_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "gpupcg.h"))
    return srcs, hdrs


def build_cuda(force=False, verbose=False):
    srcs, hdrs = cuda_sources()
    if not force and _newer(CUDA_LIB, srcs + hdrs):
        return CUDA_LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", CUDA_LIB] + srcs + ["-ldl"]
    out = _run(cmd, cwd=CSRC)
    if verbose:
        print(out)
    return CUDA_LIB


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    _run(["make", "-C", odir] + (["-B"] if force else []))
    return os.path.join(odir, "liboracle.so")


def build_host(force=False):
    hdir = os.path.join(PKG, "host")
    if os.path.exists(os.path.join(hdir, "Makefile")):
        _run(["make", "-C", hdir] + (["-B"] if force else []))


def build_all(force=False):
    build_cuda(force)
    build_host(force)
    build_oracle(force)
