#!/bin/bash
# One gpurun call: GPU parity tests, the default bench, then the ncu launch list and one full capture.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r01'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 300 $SHORT > $OUT/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_cycle_kernel -s 3 -c 2 -f -o $OUT/${TAG}_prof $SHORT > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
