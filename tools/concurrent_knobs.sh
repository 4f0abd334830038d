#!/bin/bash
# knob matrix of the concurrent-component experiment (tools/concurrent_components.py): ring budget / CTAs per SM of the sweeps
OUT=${1:-gpurun_out/concurrent_knobs.log}; : > $OUT
for cfg in "100 2" "64 1" "48 1" "32 1" "64 2" "48 2" "24 1"; do
  set -- $cfg
  echo "=== GPUPCG_PENCIL_SMEM_KB=$1 GPUPCG_PENCIL_CTAS_PER_SM=$2" >> $OUT
  GPUPCG_SPIN_TIMEOUT_MS=3000 GPUPCG_PENCIL_SMEM_KB=$1 GPUPCG_PENCIL_CTAS_PER_SM=$2 timeout 120 python tools/concurrent_components.py 2>&1 | grep "batched\|serial   \|concurrent (stagger   0" >> $OUT
done
cat $OUT
