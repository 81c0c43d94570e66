#!/bin/bash
# k_echo occupancy experiment: WAP_ECHO_SMEM_PAD_KB (extra shared memory per CTA => fewer legs in flight per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for PAD in 0 20 40; do
  WAP_ECHO_SMEM_PAD_KB=$PAD python bench.py --no-cpu-baseline --no-other-configs --steps 60 --warmup 5 --check-legs 0 > gpurun_out/pad_$PAD.json 2> gpurun_out/pad_$PAD.err
  python - <<PY
import json
d=json.load(open("gpurun_out/pad_$PAD.json"))
print("pad $PAD KB:", round(d["value"]), "legs;", round(d["ms_per_step"],3), "ms/tick", [(k["name"], round(k["ms_per_launch"],3)) for k in d["roofline"]["kernels"]])
PY
done
