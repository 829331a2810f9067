#!/bin/bash
# last GPU seconds of the round: one full ncu capture of the five kernels of a C2 step with the final code
OUT=gpurun_out/r01p; mkdir -p $OUT
timeout 110 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_c2 \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
