#!/bin/bash
# 8-GPU run: e2e with and without NUMA binding (short), then the full bench line.  Usage: gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_n8.sh TAG'
set -u
TAG=${1:-n8}
NG=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $OUT/${TAG}_lscpu.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG"
VFK_NO_NUMA_BIND=1 timeout 300 $RUN --steps 50 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_nobind.json 2> $OUT/${TAG}_nobind.err
echo "nobind exit $?"
timeout 300 $RUN --steps 50 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_bind.json 2> $OUT/${TAG}_bind.err
echo "bind exit $?"
timeout 600 $RUN --steps 300 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
python - <<PY
import json
for name in ("nobind","bind","bench"):
    try:
        d=json.loads(open("$OUT/${TAG}_%s.json"%name).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(name, "value %.4g frac %.3f | e2e %.4g ms %.3f floor %.3f frac %.3f numa %s" % (d["value"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e["roofline"]["floor_ms_per_step"], e["roofline"]["frac"], e.get("numa")))
        if name=="bench":
            for k,v in d["extras"].items(): print("  ", k, v.get("value"), (v.get("roofline") or {}).get("frac"))
            print("   sustained", d["roofline"].get("sustained"))
    except Exception as ex: print(name, "parse failed", ex)
PY
