#!/bin/bash
# ncu full capture (with source) of ONE kernel by name regex.  gpurun -- 'bash scripts/gpu_ncu_kernel.sh tag regex [workload] [count]'
TAG=${1:-ncuk}; RE=${2:-advect}; W=${3:-c2}; C=${4:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s 8 -c $C -o $OUT/prof_${RE}_$W -f \
    python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_${RE}_$W.log 2>&1
ls -la $OUT
