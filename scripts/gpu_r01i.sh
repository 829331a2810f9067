#!/bin/bash
# r01i: particles as their own kernel on a parallel graph branch (3 rotating velocity buffers): parity, A/B, attribution
OUT=gpurun_out/r01i; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -3 $OUT/pytest_gpu.txt
KAMINO_FORK_PARTICLES=0 timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_nofork.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_nofork.txt
tail -2 $OUT/pytest_gpu_nofork.txt
bash scripts/gpu_ab.sh r01i "KAMINO_FORK_PARTICLES=0 KAMINO_FORK_PARTICLES=1 KAMINO_FORK_SEQUENTIAL=1 KAMINO_PARTICLE_BLOCKS=4" "c2 c1"
timeout 300 python scripts/step_mask_timing.py c2 1000 2>&1 | tee $OUT/mask_c2.txt
KAMINO_FORK_PARTICLES=1 timeout 200 python bench.py --workload c3 --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
python -c "
import json; r=json.loads(open('$OUT/bench_c3.json').read().strip().splitlines()[-1]); print('c3', r['value'], r['kernel_us'])"
