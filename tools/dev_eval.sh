#!/bin/bash
# Development aid (GPU box): parity subset + short bench for each engine build given on the command line.
#   tools/dev_eval.sh spacegym_b200/libsg_dev_v1.so ...   -> gpurun_out/dev_<name>.json / .log
# Builds made with -DSG_DEV_ONLY_GOAL3P only carry the GoalContinuous3P-v0 / f64 instantiation.
# SG_DEV_TESTS=0 skips the parity subset (for builds whose semantics predate the oracle's).
mkdir -p gpurun_out
for spec in "$@"; do
  # spec = path/to/lib.so[:VAR=value[,VAR=value...]]  (engine environment switches for this run)
  lib="${spec%%:*}"
  vars=""
  if [ "$spec" != "$lib" ]; then vars="${spec#*:}"; fi
  name=$(basename "$lib" .so)
  if [ -n "$vars" ]; then name="${name}_$(echo "$vars" | tr ',=' '__')"; fi
  unset SG_PDL SG_TAIL SG_TAIL_CTAS SG_SUBBATCHES SG_GRAPHS
  for kv in $(echo "$vars" | tr ',' ' '); do export "$kv"; done
  export SPACEGYM_B200_LIB="$PWD/$lib"
  if [ "${SG_DEV_TESTS:-1}" != "0" ]; then
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
      -k "(golden and goal3p-) or (lockstep and goal3p-) or fused_substeps_and_synthetic_actions and goal3p or full_size" \
      > gpurun_out/dev_${name}.log 2>&1
    echo "== $name tests: $(tail -1 gpurun_out/dev_${name}.log)"
  fi
  python bench.py --steps 3000 --warmup 20 --no-cpu-baseline --no-f32 > gpurun_out/dev_${name}.json 2>> gpurun_out/dev_${name}.log
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-f32 > gpurun_out/dev_${name}_s20.json 2>> gpurun_out/dev_${name}.log
  python - <<PY
import json
d=json.load(open("gpurun_out/dev_${name}.json"))
s=json.load(open("gpurun_out/dev_${name}_s20.json"))
k=d.get("kernel_times") or {}
sh=d.get("single_handle") or {}
print("== ${name} value %.4g (20 steps: %.4g) single %.4g e2e %.4g hot %.1f cont %.1f tail %.1f us clocks %s" % (d["value"], s["value"], sh.get("value",0), d["e2e"]["value"], k.get("hot_us",0), k.get("cont_us",0), k.get("tail_us",0), d["clocks"]["sm_mhz"]))
PY
done
