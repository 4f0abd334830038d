"""fluidstructureinteraction_b200 -- B200-native D-equation hot path (gpuPCG lduMatrix::solver + GPU assembly).

The product is libgpupcg_cuda.so (hand-written sm_100a CUDA behind the C ABI in include/gpupcg.h) and the
C++ foam-extend adaptor under foam/.  This Python package is the test / benchmark harness around the C ABI.
"""
from .capi import GpuMesh, GpuPCG, GpuPcgError, load_library, plan_export, device_report, comm_unique_id, LIB_PATH  # noqa: F401
from . import meshes, cases, geometry  # noqa: F401

__all__ = ["GpuPCG", "GpuPcgError", "load_library", "plan_export", "device_report", "comm_unique_id",
           "meshes", "cases", "LIB_PATH"]
