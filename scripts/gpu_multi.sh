#!/bin/bash
# N-GPU session (gpurun --gpus N): theta-band check over NCCL against the single-GPU step, then the driver-style bench
# line (ensemble weak scaling + ensemble_c4 + banded keys) of both arms.   gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh tag 2'
TAG=${1:-multi}; N=${2:-2}; SIZES=${3:-"512:20 2048:10 8192:10"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for sz in $SIZES; do
  nT=${sz%%:*}; st=${sz##*:}; dt=0.005; [ $nT = 8192 ] && dt=0.0025
  NCCL_DEBUG=WARN timeout 400 $TR scripts/dist_check.py $nT $st $dt 1.0 10 > $OUT/dist_${nT}_n$N.txt 2>&1
  echo "dist_check $nT exit $?"; grep "^rank" $OUT/dist_${nT}_n$N.txt | sort | head -8
done
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/bench_reference_n$N.json 2> $OUT/bench_reference_n$N.err
python - <<PY
import json
for f in ("bench_n$N", "bench_reference_n$N"):
    try:
        r = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "n_gpus", r["n_gpus"], "value %.0f" % r["value"], "ms/step %.4f" % r["ms_per_step"], "e2e %.0f" % r["e2e"]["value"], r.get("reps", ""))
        for k in ("ensemble_c4", "banded"):
            if k in r: print("   ", k, json.dumps(r[k]))
    except Exception as e:
        print(f, "FAILED", e); print(open("$OUT/%s.err" % f).read()[-2500:])
PY
