#!/bin/bash
# N-GPU session: theta-band check only (both transports timed).   gpurun --gpus 2 -- 'bash scripts/gpu_dist.sh tag 2 "2048:10 8192:10"'
TAG=${1:-dist}; N=${2:-2}; SIZES=${3:-"2048:10 8192:10"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for sz in $SIZES; do
  nT=${sz%%:*}; st=${sz##*:}; dt=0.005; [ $nT = 8192 ] && dt=0.0025
  NCCL_DEBUG=WARN timeout 300 $TR scripts/dist_check.py $nT $st $dt 1.0 10 > $OUT/dist_${nT}_n$N.txt 2>&1
  echo "dist_check $nT exit $?"; grep "^rank" $OUT/dist_${nT}_n$N.txt | sort | head -8; grep -i "error\|Traceback" $OUT/dist_${nT}_n$N.txt | head -5
done
