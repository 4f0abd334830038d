#!/bin/bash
# 8 ranks: ordered sums over the mailboxes (bit-identical), production reductions, a larger case
N=$1; OUT=$2
run() { echo "=== $*" >> $OUT; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_parity.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM_THREADS\|^$" >> $OUT; echo "rc=$?" >> $OUT; }
: > $OUT
run ORDERED=1
run ORDERED=0
run ORDERED=1 CELLS=64,32,32
run ORDERED=0 CELLS=64,32,32
