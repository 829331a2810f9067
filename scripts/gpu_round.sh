#!/bin/bash
# One GPU-box session: parity tests, both bench arms, ncu launch list + full capture.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python scripts/gpu_quickcheck.py > $OUT/quickcheck.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 1000 --warmup 20 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 python bench.py --steps 1000 --warmup 20 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench_reference.json $OUT/bench.json
for w in c1 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 600 python bench.py --impl reference --workload c3 --steps 50 --warmup 5 > $OUT/bench_reference_c3.json 2>> $OUT/bench_reference.err
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
# one full capture of each of the five kernels (step 4 after warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_c2 \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_c3 \
    python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c3.log 2>&1
ls -la $OUT
