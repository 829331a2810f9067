#!/bin/bash
# A/B of environment switches (VAR=a, or VAR1=a,VAR2=b for combinations) on the C2 / C3 bench lines.  gpurun -- 'bash scripts/gpu_ab.sh tag "VAR=a VAR=b ..." [workloads]'
TAG=${1:-ab}; VARIANTS=${2:-"KAMINO_NONE=0"}; WL=${3:-"c2 c3"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARIANTS; do
  for w in $WL; do
    steps=1000; [ $w = c3 ] && steps=100; [ $w = c4 ] && steps=100
    env $(echo $v | tr "," " ") timeout 600 python bench.py --workload $w --steps $steps --warmup 10 --no-cpu-baseline > $OUT/bench_${w}_$v.json 2> $OUT/bench_${w}_$v.err
    python - <<PY
import json
try:
    r = json.loads(open("$OUT/bench_${w}_$v.json").read().strip().splitlines()[-1])
    print("$v $w steps/s %.0f ms/step %.4f e2e %.0f cold %.4f" % (r["value"], r["ms_per_step"], r["e2e"]["value"], r["cold"]["ms_per_step"]), {k: round(x, 1) for k, x in r["kernel_us"].items()})
except Exception as e:
    print("$v $w FAILED", e)
PY
  done
done
