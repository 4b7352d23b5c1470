#!/bin/bash
# Kernel tuning round: default build tests + variants (bench short form) + FP32 accuracy report.
set -u
TAG=${1:-t01}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
B="python bench.py --steps 300 --warmup 5 --no-cpu-baseline"
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 $B > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value %.3e"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "k100 %.3e"%d["extras"]["k_fused"]["value"], "fp64 %.3e"%d["extras"].get("fp64_config2",{}).get("value",0))
except Exception as e: print("$name failed", e)
PY
}
run default A=1
run ne_only VFK_LIB=$PWD/build/libvfk_ne.so
run fk_only VFK_LIB=$PWD/build/libvfk_fk.so
for l in ne fk; do VFK_LIB=$PWD/build/libvfk_$l.so timeout 300 python scripts/fp32_error.py 1048576 2>&1 | grep -E "twist|qdot_vf"; done
