#!/bin/bash
# DIC precondition timings (scalar / batched-3) at 1 M and 8 M cells over the pencil-sweep knobs
OUT=$1; : > $OUT
for split in 1 0; do for cps in 1 2 3; do
  echo "== SPLIT=$split CTAS_PER_SM=$cps" >> $OUT
  GPUPCG_PENCIL_SPLIT=$split GPUPCG_PENCIL_CTAS_PER_SM=$cps python tools/pencil_probe.py big 2>&1 | grep shape | cut -c1-150 >> $OUT
done; done
