#!/bin/bash
# A/B over environment settings.  Usage: gpu_env_ab.sh TAG WORKLOAD "ENV1=.. ENV2=.." "..." ...   (use "-" for no env)
set -u
TAG=$1; W=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
i=0
for spec in "$@"; do
  i=$((i+1))
  F=$OUT/${TAG}_${W}_$i.json
  if [ "$spec" = "-" ]; then timeout 300 python bench.py --workload $W --steps 100 --no-cpu-baseline --no-extras > $F 2> $OUT/${TAG}_${W}_$i.err
  else timeout 300 env $spec python bench.py --workload $W --steps 100 --no-cpu-baseline --no-extras > $F 2> $OUT/${TAG}_${W}_$i.err; fi
  python - <<PY
import json
try:
    d=json.load(open("$F")); print("$spec", "ms/launch %.4f frac %.3f" % (d["ms_per_step"], d["roofline"]["frac"]), d["clocks"]["reasons"])
except Exception as e: print("$spec", "failed", e)
PY
done
