#!/bin/bash
# GPU visit for the multi-channel path: its parity tests, then the config-4 bench line (device + e2e + spot check).
# usage: tools/gpu_mc.sh <tag> [all]
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-mc}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
free -g | head -2 >> gpurun_out/smi_$TAG.txt; nproc >> gpurun_out/smi_$TAG.txt
if [ "$2" = "all" ]; then
  timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.txt
else
  timeout 900 python -m pytest tests/test_multichannel.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.txt
fi
tail -5 gpurun_out/pytest_gpu_$TAG.txt
timeout 900 python bench.py --mc 1 --rate 48000 --ns 0 --streams 8192 --steps 50 --warmup 5 --no-other-configs --check-legs 4 > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_cfg4_$TAG.json; tail -3 gpurun_out/bench_cfg4_$TAG.err
timeout 600 python bench.py --mc 1 --rate 16000 --ns 0 --streams 16384 --steps 50 --warmup 5 --no-other-configs --no-cpu-baseline --check-legs 4 > gpurun_out/bench_mc16k_$TAG.json 2> gpurun_out/bench_mc16k_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_mc16k_$TAG.json
ls -la gpurun_out | tail -12
