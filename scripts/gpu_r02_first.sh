#!/bin/bash
# First GPU session of the next round: (1) parity incl. the tests that were written after the r01 GPU budget
# ran out (checkpoint restart, SPIKE over virtual ranks), (2) the opt-in experiment switches prepared in r01:
#   KAMINO_GEO_COLS=64      geometric 8 x 64 tiles (balance at C2)
#   KAMINO_TRI_L=8|32       theta-solve chunk length at 512 rows; 16|64 at 2048 rows
#   KAMINO_TRI_W=2          narrower slot groups (more blocks) at C2
#   KAMINO_FFT_MINBLOCKS=3  80-register FFT kernels at C3 (three 256-thread blocks per SM)
#   KAMINO_GEO_PREFETCH=1   geometric: software-pipelined input loads
#   KAMINO_PDL_TAIL=1       last-wave blocks release the programmatic dependents at entry
OUT=gpurun_out/r02a; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -rxX > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -12 $OUT/pytest_gpu.txt
bash scripts/gpu_ab.sh r02a "KAMINO_GEO_COLS=128 KAMINO_GEO_PREFETCH=1 KAMINO_PDL_TAIL=1 KAMINO_GEO_COLS=64 KAMINO_TRI_L=8 KAMINO_TRI_L=32 KAMINO_TRI_W=2" "c2"
bash scripts/gpu_ab.sh r02a "KAMINO_TRI_L=32 KAMINO_GEO_PREFETCH=1 KAMINO_PDL_TAIL=1 KAMINO_TRI_L=16 KAMINO_TRI_L=64 KAMINO_TRI_W=4 KAMINO_FFT_MINBLOCKS=3" "c3"
# bit-identity of the tail-trigger variant (it must not change a result)
KAMINO_PDL_TAIL=1 timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu_pdltail.txt 2>&1; tail -3 $OUT/pytest_gpu_pdltail.txt
