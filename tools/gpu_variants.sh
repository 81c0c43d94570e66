#!/bin/bash
# Bench several builds of the library back to back (same box): tools/gpu_variants.sh <tag> <lib1> <lib2> ... -- [bench args]
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=$1; shift
LIBS=()
while [ "$1" != "--" ] && [ -n "$1" ]; do LIBS+=("$1"); shift; done
shift
mkdir -p gpurun_out
for L in "${LIBS[@]}"; do
  N=$(basename $L .so)
  WAP_B200_LIB=$PWD/$L python bench.py --no-cpu-baseline --no-other-configs "$@" > gpurun_out/var_${TAG}_$N.json 2> gpurun_out/var_${TAG}_$N.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_${TAG}_$N.json"))
print("$N", round(d["value"]), "legs;", round(d["ms_per_step"],3), "ms/tick; e2e", round(d["e2e"]["value"]), [(k["name"], round(k["ms_per_launch"],3)) for k in d["roofline"]["kernels"]], d.get("parity_spot_check",{}).get("max_abs_diff_lsb"))
PY
done
