#!/bin/bash
# N-GPU session (gpurun --gpus N): ensemble scaling lines + theta-band check.  gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh tag 2'
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu-baseline > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err
timeout 300 $TR bench.py --gpus $N --workload c4 --steps 200 --warmup 10 --no-cpu-baseline > $OUT/bench_c4_n$N.json 2> $OUT/bench_c4_n$N.err
timeout 300 $TR scripts/banded_check.py 512 22 > $OUT/banded_512_n$N.txt 2>&1
timeout 300 $TR scripts/banded_check.py 2048 12 > $OUT/banded_2048_n$N.txt 2>&1
tail -4 $OUT/banded_512_n$N.txt $OUT/banded_2048_n$N.txt
python - <<PY
import json
for w in ("c2", "c4"):
    try:
        r = json.loads(open("$OUT/bench_%s_n$N.json" % w).read().strip().splitlines()[-1])
        print(w, "n_gpus", r["n_gpus"], "value %.0f" % r["value"], "ms/step %.4f" % r["ms_per_step"], "e2e %.0f" % r["e2e"]["value"])
    except Exception as e:
        print(w, "FAILED", e); print(open("$OUT/bench_%s_n$N.err" % w).read()[-1500:])
PY
