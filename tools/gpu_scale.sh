#!/bin/bash
# Multi-GPU lines on one 8-GPU box: BASELINE config 3 (NS-only kHigh 48 kHz) at 1/2/4/8 GPUs and the headline config at 8.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {  # N tag args...
  N=$1; TAG=$2; shift; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/scale_${TAG}_${N}gpu.json 2> gpurun_out/scale_${TAG}_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/scale_${TAG}_${N}gpu.json 2> gpurun_out/scale_${TAG}_${N}gpu.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${TAG}_${N}gpu.json"))
    print("${TAG}", d["n_gpus"], "GPUs:", round(d["value"]), "legs device;", round(d["e2e"]["value"]), "e2e;", round(d["ms_per_step"],3), "ms/tick")
except Exception as e:
    print("${TAG} ${N} failed", e)
PY
}
CFG3="--rate 48000 --aec 0 --ns-level 2 --streams 16384 --steps 100 --warmup 10 --settle 300 --no-cpu-baseline --no-other-configs --check-legs 4"
for N in 1 2 4 8; do run $N cfg3 $CFG3; done
run 8 cfg2 --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs --check-legs 4
