#!/bin/bash
# A/B of alternative builds of the same ABI (VFK_LIB).  Usage: gpu_variants.sh TAG "lib:workload:extra-args" ...
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for spec in "$@"; do
  LIB=${spec%%:*}; rest=${spec#*:}; W=${rest%%:*}; ARGS=${rest#*:}; [ "$ARGS" = "$rest" ] && ARGS=""
  if [ "$LIB" = "default" ]; then unset VFK_LIB; else export VFK_LIB=$PWD/vfclik_b200/libvfk_$LIB.so; fi
  F=$OUT/${TAG}_${LIB}_${W}.json
  timeout 400 python bench.py --workload $W --no-cpu-baseline $ARGS > $F 2> $OUT/${TAG}_${LIB}_${W}.err
  echo "$LIB $W exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$F"))
    k=d["extras"].get("k_fused",{})
    print("  ms/launch %.4f frac %.3f  kfused %s ms %s | e2e %s ms | clocks %s %s" % (d["ms_per_step"], d["roofline"]["frac"], k.get("value"), k.get("ms_per_launch"), (d.get("e2e") or {}).get("ms_per_step"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    if "fp64_config2" in d["extras"]: print("  fp64_config2", d["extras"]["fp64_config2"])
    if (d.get("e2e") or {}).get("roofline"): print("  e2e roofline", d["e2e"]["roofline"], d["e2e"].get("k100"), d["e2e"].get("numa"))
except Exception as e: print("  parse failed", e)
PY
done
unset VFK_LIB
