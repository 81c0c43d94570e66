#!/bin/bash
# First GPU look: tests + a quick NS-only throughput probe.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python tools/ns_probe.py 2>&1 | tee gpurun_out/ns_probe.txt
