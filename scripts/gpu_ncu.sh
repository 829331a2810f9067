#!/bin/bash
# ncu full capture of the five kernels of one step.  gpurun -- 'bash scripts/gpu_ncu.sh tag [workload]'
TAG=${1:-ncu}; W=${2:-c2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_$W \
    python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$W.log 2>&1
ls -la $OUT
