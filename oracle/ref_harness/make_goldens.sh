#!/bin/bash
# TEST INFRASTRUCTURE ONLY (oracle/). Runs the reference CUDA build (oracle/_ref/kamino_ref,
# built by oracle/ref_harness/Makefile from the sources under /root/reference) on the GPU box
# and leaves raw state dumps under gpurun_out/ref_dumps/<case>/ . The small cases are copied
# into tests/golden/ by tests/golden/import_ref_dumps.py and committed.
# Usage (from the repo root, under gpurun): bash oracle/ref_harness/make_goldens.sh
set -e
REF=oracle/_ref/kamino_ref
OUT=gpurun_out/ref_dumps
mkdir -p $OUT
run_case () {  # name nTheta particleDensity dt radius nSteps phaseSteps
  mkdir -p $OUT/$1
  $REF dump $2 $3 $4 $5 $6 $OUT/$1 - $7 > $OUT/$1/stdout.txt 2>&1 || { echo "case $1 failed"; tail -5 $OUT/$1/stdout.txt; }
}
run_case t16   16  4 0.005 5.0 3 3
run_case t32   32  4 0.005 5.0 12 2
run_case t64   64  1 0.005 5.0 3 1
run_case t128 128  1 0.005 5.0 100 1
ls -la $OUT/*
