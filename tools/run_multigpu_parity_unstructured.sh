#!/bin/bash
# usage: tools/run_multigpu_parity_unstructured.sh N OUT -- unstructured sub-domains (decomposed jittered-tet box, shipped
# 3dTube solid mesh cut by `method simple`): ordered sums (bit-identical bar) and production reductions at N ranks
N=$1; OUT=$2
run() { echo "=== $*" >> $OUT; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_parity.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM_THREADS\|^$" >> $OUT; echo "rc=$?" >> $OUT; }
: > $OUT
run MESH=tet ORDERED=1
run MESH=tet ORDERED=0
run MESH=tet ORDERED=1 CELLS=120,40,48
run MESH=tube ORDERED=1
run MESH=tube ORDERED=0
