#!/bin/bash
# First GPU session of the next round.
#  (1) parity, incl. the tests written after the r01 GPU budget ran out (checkpoint restart, SPIKE over
#      virtual ranks): -rxX prints which of the non-strict xfails passed;
#  (2) every opt-in experiment switch prepared in r01 must first pass the parity tests that touch its
#      kernel, then it is measured at C2 / C3:
#        KAMINO_PDL_TAIL=1        last-wave blocks release the programmatic dependents at entry
#        KAMINO_TILE_STRIDE=64    advection tiles padded to 64 floats per row (bank conflicts)
#        KAMINO_GEO_PREFETCH=1    geometric: software-pipelined input loads
#        KAMINO_GEO_COLS=64       geometric 8 x 64 tiles (balance at C2)
#        KAMINO_TRI_L=8|32        theta-solve chunk length at 512 rows; 16|64 at 2048 rows
#        KAMINO_TRI_W=2|4         slot-group width of the theta solve
#        KAMINO_FFT_MINBLOCKS=3   80-register FFT kernels at C3 (three 256-thread blocks per SM)
OUT=gpurun_out/r02a; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -rxX > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -12 $OUT/pytest_gpu.txt
SUBSET="reference_dump or graph_steps or one_step_at_c2 or live_reference or properties_at_full_size or banded_step"
for v in KAMINO_PDL_TAIL=1 KAMINO_TILE_STRIDE=64 KAMINO_GEO_PREFETCH=1 KAMINO_GEO_COLS=64 KAMINO_TRI_L=8 KAMINO_TRI_L=32 KAMINO_TRI_W=2 KAMINO_FFT_MINBLOCKS=3; do
  env $v timeout 600 python -m pytest tests -q -m gpu -k "$SUBSET" > $OUT/pytest_$v.txt 2>&1
  echo "$v parity: $(tail -1 $OUT/pytest_$v.txt)"
done
bash scripts/gpu_ab.sh r02a "KAMINO_GEO_COLS=128 KAMINO_PDL_TAIL=1 KAMINO_TILE_STRIDE=64 KAMINO_GEO_PREFETCH=1 KAMINO_GEO_COLS=64 KAMINO_TRI_L=8 KAMINO_TRI_L=32 KAMINO_TRI_W=2" "c2"
bash scripts/gpu_ab.sh r02a "KAMINO_TRI_L=32 KAMINO_PDL_TAIL=1 KAMINO_TILE_STRIDE=64 KAMINO_GEO_PREFETCH=1 KAMINO_TRI_L=16 KAMINO_TRI_L=64 KAMINO_TRI_W=4 KAMINO_FFT_MINBLOCKS=3" "c3"
