#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list + one full capture of k_tick.
# usage: tools/gpu_round.sh <tag> [bench args...]
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-r01}; shift
BARGS="$@"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.txt
tail -5 gpurun_out/pytest_gpu_$TAG.txt
python bench.py $BARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 1 $BARGS > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
SMALL="--steps 6 --warmup 3 --no-cpu-baseline $BARGS"
python bench.py $SMALL > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $SMALL > gpurun_out/ncu_l_$TAG.log 2>&1
python bench.py $SMALL > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_front|k_delay|k_echo" -s 9 -c 6 -o gpurun_out/prof_$TAG python bench.py $SMALL > gpurun_out/ncu_f_$TAG.log 2>&1
cp webrtc-audio-processing_b200/libwap_b200.so gpurun_out/libwap_b200_$TAG.so
ls -la gpurun_out
