#!/bin/bash
# A/B harness: runs bench.py (device-timed part only) under several environment settings on ONE box, interleaved twice.
# usage: tools/ab.sh "NAME1:ENV1=..,ENV2=.." "NAME2:..."   -> gpurun_out/ab_<NAME>_<rep>.json
for rep in 1 2; do
  for spec in "$@"; do
    name="${spec%%:*}"; envs="${spec#*:}"
    ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
      timeout 300 python bench.py --no-extras --no-cpu-baseline --no-latency --steps 15 > gpurun_out/ab_${name}_${rep}.json 2> gpurun_out/ab_${name}_${rep}.err )
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${name}_${rep}.json").read().strip().splitlines()[-1])
    print("${name} rep${rep}: %.1f img/s  %.3f ms  e2e %.1f  gemm %.0f TF/s  clk %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("${name} rep${rep}: ERR", e)
PY
  done
done
