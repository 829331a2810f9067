#!/bin/bash
# Trimmed GPU-box session: parity tests, both bench arms at C2, C3/C4 bench lines, ncu launch list + full captures.
# Usage (from the repo root): gpurun --timeout 1200 -- 'bash scripts/gpu_round_lite.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
cp /root/repo/MEASURED_PEAKS.json $OUT/ 2>/dev/null
timeout 600 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
timeout 300 python bench.py --impl reference --steps 1000 --warmup 20 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 300 python bench.py --steps 1000 --warmup 20 > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench_reference.json $OUT/bench.json
for w in c3 c4; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_c2 \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 30 -c 5 -o $OUT/prof_c3 \
    python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_c3.log 2>&1
ls -la $OUT
