#!/bin/bash
# One GPU-box session: parity tests, the driver-style bench lines of both arms, long bench lines, launch list.
# Usage (from the repo root): gpurun --timeout 1200 -- 'bash scripts/gpu_session.sh <tag> [ncu]'
TAG=${1:-r02}; NCU=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
# what the driver runs at round end
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference_driver.json 2> $OUT/bench_reference_driver.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err
# long lines
timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --workload c3 --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
timeout 300 python bench.py --impl cli --steps 1000 > $OUT/bench_cli.json 2> $OUT/bench_cli.err
python - <<PY
import json
for f in ("bench_reference_driver", "bench_driver", "bench", "bench_c3", "bench_cli"):
    try:
        r = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "steps/s %.0f  ms/step %.4f" % (r["value"], r["ms_per_step"]), "e2e %.0f" % r["e2e"]["value"] if "e2e" in r else "",
              r.get("reps", ""), r.get("ensemble_c4", {}).get("value", ""), r.get("warning", ""))
        if "kernel_us" in r:
            print("   kernel_us", {k: round(v, 2) for k, v in r["kernel_us"].items()}, "l2_copy_gbs %.0f" % r["roofline"]["l2_copy_gbs"],
                  "pcie d2h %.1f" % r["e2e"]["pcie_d2h_gbs"], "cold %.4f" % r["cold"]["ms_per_step"], r["clocks"])
    except Exception as e:
        print(f, "FAILED", e); print(open("$OUT/%s.err" % f).read()[-2000:])
PY
if [ -n "$NCU" ]; then
  # launch list (cold-cache, serialised: shares only) and one full capture of the five kernels of a step
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 50 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 20 --warmup 3 --reps 1 --no-ensemble --no-banded --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -s 32 -c 5 -o $OUT/prof_c2 -f \
      python bench.py --steps 20 --warmup 3 --reps 1 --no-ensemble --no-banded --no-cpu-baseline > $OUT/ncu_full.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -s 32 -c 5 -o $OUT/prof_c3 -f \
      python bench.py --workload c3 --steps 20 --warmup 3 --reps 1 --no-cpu-baseline > $OUT/ncu_full_c3.log 2>&1
fi
ls $OUT | head -40
