#!/bin/bash
# Last single-GPU session of a round: parity suite, the driver's own bench lines, ncu launch list + full captures.
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
timeout 60 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference_driver.json 2> $OUT/bench_reference_driver.err
timeout 90 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err
python scripts/show_bench.py $OUT/bench_reference_driver.json $OUT/bench_driver.json
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 50 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --reps 1 --no-ensemble --no-banded --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -s 32 -c 5 -o $OUT/prof_c2 -f \
    python bench.py --steps 20 --warmup 3 --reps 1 --no-ensemble --no-banded --no-cpu-baseline > $OUT/ncu_full.log 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -s 32 -c 5 -o $OUT/prof_c3 -f \
    python bench.py --workload c3 --steps 20 --warmup 3 --reps 1 --no-cpu-baseline > $OUT/ncu_full_c3.log 2>&1
ls $OUT
