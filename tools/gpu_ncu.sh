#!/bin/bash
# ncu --set full capture of selected tick kernels in steady state: tools/gpu_ncu.sh <tag> <kernel regex> [launch count]
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=$1; KRE=$2; CNT=${3:-2}
mkdir -p gpurun_out
SMALL="--streams 16384 --steps 4 --warmup 2 --settle 300 --no-cpu-baseline --no-other-configs --check-legs 0"
python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 310 -c $CNT -o gpurun_out/prof_$TAG python bench.py $SMALL > gpurun_out/ncu_f_$TAG.log 2>&1
tail -3 gpurun_out/ncu_f_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
