#!/bin/bash
# r01n: compact solve tables + new GPU tests: parity, then A/B
OUT=gpurun_out/r01n; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -15 $OUT/pytest_gpu.txt
bash scripts/gpu_ab.sh r01n "KAMINO_TRI_COMPACT=0 KAMINO_TRI_COMPACT=1 KAMINO_ADVECT=7" "c2 c3"
bash scripts/gpu_ab.sh r01n "KAMINO_TRI_COMPACT=0 KAMINO_TRI_COMPACT=1" "c4"
