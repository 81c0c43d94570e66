#!/bin/bash
# One GPU visit of round 2: parity tests, driver-shaped bench, steady-state bench, optional ncu.
# usage: tools/gpu_r02.sh <tag> [ncu]
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
free -g | head -2 >> gpurun_out/smi_$TAG.txt; nproc >> gpurun_out/smi_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.txt
tail -5 gpurun_out/pytest_gpu_$TAG.txt
timeout 900 python bench.py --warmup 5 --steps 20 > gpurun_out/bench_drv_$TAG.json 2> gpurun_out/bench_drv_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_drv_$TAG.json
timeout 900 python bench.py --no-other-configs --no-cpu-baseline --check-legs 0 > gpurun_out/bench_long_$TAG.json 2> gpurun_out/bench_long_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_long_$TAG.json
timeout 600 python bench.py --impl reference --steps 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_ref_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
if [ "$2" = "ncu" ]; then
  SMALL="--streams 16384 --steps 4 --warmup 2 --settle 300 --no-cpu-baseline --no-other-configs --check-legs 0"
  python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL > gpurun_out/ncu_l_$TAG.log 2>&1
  python bench.py $SMALL > gpurun_out/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"k_front|k_delay|k_echo" -s 915 -c 6 -o gpurun_out/prof_$TAG python bench.py $SMALL > gpurun_out/ncu_f_$TAG.log 2>&1
  ls -la gpurun_out/prof_$TAG.ncu-rep   # (the profiled library is the snapshot's own build: not copied back, gpurun_out is capped at 64 MiB)
fi
ls -la gpurun_out | tail -20
