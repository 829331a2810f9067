#!/bin/bash
# r01g: particle lattice tiling A/B + in-graph attribution of the step time
OUT=gpurun_out/r01g; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -3 $OUT/pytest_gpu.txt
bash scripts/gpu_ab.sh r01g "KAMINO_PARTICLE_TILES=0 KAMINO_PARTICLE_TILES=1" "c2 c1"
timeout 300 python scripts/step_mask_timing.py c2 1000 2>&1 | tee $OUT/mask_c2.txt
timeout 300 python scripts/step_mask_timing.py c3 100 2>&1 | tee $OUT/mask_c3.txt
