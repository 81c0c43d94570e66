#!/bin/bash
# config 4 with the noise suppressor on both channels (device + e2e + spot check)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --mc 1 --rate 48000 --ns 1 --ns-level 1 --streams 8192 --steps 50 --warmup 5 --no-other-configs --check-legs 4 > gpurun_out/bench_cfg4_ns.json 2> gpurun_out/bench_cfg4_ns.err; echo "bench rc=$?"
cat gpurun_out/bench_cfg4_ns.json | cut -c1-400
