#!/bin/bash
# GPU tests on the current library, then A/B of library builds on the headline workload
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02f.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02f.txt
tail -4 gpurun_out/pytest_gpu_r02f.txt
bash tools/gpu_variants.sh r02f variants/libwap_b200_base.so variants/libwap_b200_fft.so variants/libwap_b200_fftfront.so -- --steps 100 --warmup 5 --check-legs 4
