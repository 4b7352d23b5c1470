#!/bin/bash
# One gpurun call: GPU parity tests + smoke (+ optional bench).  Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_tests.sh TAG [bench]'
set -u
TAG=${1:-t}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -25 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
if [ "${2:-}" = "bench" ]; then
  timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench exit $?"
  tail -c 2500 $OUT/${TAG}_bench.json
fi
