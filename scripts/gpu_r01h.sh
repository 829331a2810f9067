#!/bin/bash
# r01h: advect cell-path diet + interleaved tile/particle blocks + 512-thread geometric: parity, then A/B
OUT=gpurun_out/r01h; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -3 $OUT/pytest_gpu.txt
bash scripts/gpu_ab.sh r01h "KAMINO_ADVECT_MIX=0,KAMINO_GEO_THREADS=256 KAMINO_ADVECT_MIX=1,KAMINO_GEO_THREADS=512 KAMINO_ADVECT_MIX=1,KAMINO_ADVECT=4 KAMINO_ADVECT_MIX=1,KAMINO_ADVECT=6" "c2"
bash scripts/gpu_ab.sh r01h "KAMINO_ADVECT=5 KAMINO_ADVECT=4" "c3 c1"
