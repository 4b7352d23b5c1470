#!/bin/bash
# Lane-split kernel: parity test, then A/B timings (split vs one thread per instance) on configs 5 and 2.
set -u
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
for W in config5 config2; do
  for NS in 0 1; do
    if [ $NS = 1 ]; then unset VFK_SPLIT; else export VFK_SPLIT=1; fi
    timeout 300 python bench.py --workload $W --steps 100 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_${W}_nosplit${NS}.json 2> $OUT/${TAG}_${W}_nosplit${NS}.err
    echo "$W nosplit=$NS exit $?"
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_${W}_nosplit${NS}.json"))
    print("$W nosplit=$NS", "ms/launch %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "kfused", d["extras"].get("k_fused",{}).get("value"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("parse failed", e)
PY
  done
done
unset VFK_SPLIT
