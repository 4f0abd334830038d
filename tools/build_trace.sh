#!/bin/sh
# trace build of the CUDA library (extra per-row / per-entry time stamps in the forward sweep) for tools/hop_trace.py
cd "$(dirname "$0")/../fluidstructureinteraction_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -DGPUPCG_DEEPTRACE -Xcompiler -fPIC -shared -o ../libgpupcg_cuda_trace.so gpupcg.cu assembly.cu plan.cpp comm.cpp -ldl
