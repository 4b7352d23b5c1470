#!/bin/bash
# N-GPU record of the final build: the full bench line and the CPU reference arm.  Usage: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_n_final.sh TAG N'
set -u
TAG=${1:-n8}
NG=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG"
timeout 600 $RUN --steps 300 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
timeout 300 $RUN --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference exit $?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("value %.4g frac %.3f | e2e %.4g ms %.3f floor %.3f frac %.3f" % (d["value"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e["roofline"]["floor_ms_per_step"], e["roofline"]["frac"]))
for k,v in d["extras"].items(): print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"))
print("   sustained", d["roofline"].get("sustained"))
print(open("$OUT/${TAG}_bench_reference.json").read()[:300])
PY
