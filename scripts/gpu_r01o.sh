#!/bin/bash
# r01o (2 GPUs): banded mode with graph-captured steps, eager for comparison, and the 8192 x 16384 grid
N=${1:-2}
OUT=gpurun_out/r01o; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 200 $TR scripts/banded_check.py 512 40 > $OUT/banded_512_graph.txt 2>&1
KAMINO_BANDED_GRAPH=0 timeout 200 $TR scripts/banded_check.py 512 40 > $OUT/banded_512_eager.txt 2>&1
timeout 200 $TR scripts/banded_check.py 2048 20 > $OUT/banded_2048_graph.txt 2>&1
timeout 500 $TR scripts/banded_check.py 8192 6 > $OUT/banded_8192_graph.txt 2>&1
for f in $OUT/banded_*.txt; do echo "== $f"; grep -E "banded|Error|error|Traceback" $f | tail -6; done
