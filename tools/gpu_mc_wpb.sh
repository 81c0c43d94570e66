#!/bin/bash
# k_mc_echo variants on the config-4 bench (device numbers only).  usage: tools/gpu_mc_wpb.sh "<wpb list>"
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for w in ${1:-4}; do
  WAP_MC_ECHO_WPB=$w timeout 600 python bench.py --mc 1 --rate 48000 --ns 0 --streams 8192 --steps 50 --warmup 5 --no-other-configs --no-cpu-baseline --check-legs 2 > gpurun_out/bench_cfg4_wpb$w.json 2> gpurun_out/bench_cfg4_wpb$w.err
  echo "wpb=$w rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg4_wpb$w.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], [ (k['name'], round(k['ms_per_launch'],3)) for k in d['roofline']['kernels']], d['parity_spot_check']['pass'])
PY
done
