#!/bin/bash
# ncu launch list + one --set full capture of the multi-channel tick kernels (config 4, 2048 legs, after the switch).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SMALL="--mc 1 --rate 48000 --ns 0 --streams 2048 --steps 4 --warmup 2 --settle 520 --no-cpu-baseline --no-other-configs --check-legs 0"
python bench.py $SMALL > gpurun_out/plain_mc.log 2>&1 || exit 1
# 4 launches per tick (front, delay, echo, post); skip the settle phase
ncu --metrics gpu__time_duration.sum --clock-control none -s 2090 -c 64 --csv --log-file gpurun_out/launches_mc.csv python bench.py $SMALL > gpurun_out/ncu_l_mc.log 2>&1
python bench.py $SMALL > gpurun_out/plain2_mc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mc_front|k_mc_echo|k_mc_post" -s 1570 -c 6 -o gpurun_out/prof_mc python bench.py $SMALL > gpurun_out/ncu_f_mc.log 2>&1
ls -la gpurun_out | tail
