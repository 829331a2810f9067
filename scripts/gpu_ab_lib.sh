#!/bin/bash
# A/B of alternative builds of the library (kaminogpu_b200/build/variants/libkamino_<v>.so; "default" = the product build)
# on the C2 / C3 bench lines, interleaved twice.   gpurun -- 'bash scripts/gpu_ab_lib.sh tag "default old check" "c2 c3"'
TAG=${1:-ablib}; VARIANTS=${2:-"default"}; WL=${3:-"c2"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for round in 1 2; do
  for v in $VARIANTS; do
    for w in $WL; do
      steps=1000; [ $w = c3 ] && steps=100
      lib=""; [ $v != default ] && lib=$PWD/kaminogpu_b200/build/variants/libkamino_$v.so
      KAMINO_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps $steps --warmup 10 --reps 10 --no-cpu-baseline --no-ensemble --no-banded > $OUT/bench_${w}_${v}_$round.json 2> $OUT/bench_${w}_${v}_$round.err
      python - <<PY
import json
try:
    r = json.loads(open("$OUT/bench_${w}_${v}_$round.json").read().strip().splitlines()[-1])
    print("$v $w round $round: steps/s %.0f ms/step %.4f (min %.4f)" % (r["value"], r["ms_per_step"], r["reps"]["ms_per_step_min"]), {k: round(x, 2) for k, x in r["kernel_us"].items()})
except Exception as e:
    print("$v $w FAILED", e)
PY
    done
  done
done
