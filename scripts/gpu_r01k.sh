#!/bin/bash
# r01k: branch-free interior backtraces: parity, then register-budget A/B at C2 / C3
OUT=gpurun_out/r01k; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -3 $OUT/pytest_gpu.txt
bash scripts/gpu_ab.sh r01k "KAMINO_ADVECT=5 KAMINO_ADVECT=4 KAMINO_ADVECT=6" "c2 c3"
