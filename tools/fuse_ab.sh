#!/bin/bash
# bench step (1 M cells) with / without the x/r update riding in the forward pencil sweep, over ring sizes
run() { echo -n "$* : "; env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --no-64m --no-8m --no-evolve --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), 'ms/step', '%.3e' % d['value'])"; }
run GPUPCG_FUSE_XR=0
run GPUPCG_FUSE_XR=1
run GPUPCG_FUSE_XR=1 GPUPCG_PENCIL_SMEM_KB=150
run GPUPCG_FUSE_XR=1 GPUPCG_PENCIL_SMEM_KB=200
run GPUPCG_FUSE_XR=1 GPUPCG_PENCIL_SPLIT=0 GPUPCG_PENCIL_CTAS_PER_SM=1 GPUPCG_PENCIL_SMEM_KB=200
run GPUPCG_FUSE_XR=0 GPUPCG_PENCIL_SMEM_KB=200
