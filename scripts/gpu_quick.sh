#!/bin/bash
# Short GPU session: parity tests + C2/C3 bench lines (no ncu).  gpurun -- 'bash scripts/gpu_quick.sh tag'
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
timeout 600 python bench.py --workload c3 --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
python - <<PY
import json
for f in ("$OUT/bench.json", "$OUT/bench_c3.json"):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "steps/s %.0f  ms/step %.4f  e2e %.0f" % (r["value"], r["ms_per_step"], r["e2e"]["value"]))
        print("   kernel_us", {k: round(v, 2) for k, v in r["kernel_us"].items()})
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-2000:])
PY
