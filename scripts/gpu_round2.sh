#!/bin/bash
# One gpurun call: GPU parity tests, smoke, default bench, then ncu launch metrics (time, instructions, DRAM bytes) of a short bench.
# Usage (from the build container):  gpurun --timeout 1800 -- 'bash scripts/gpu_round2.sh TAG'
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
tail -c 3000 $OUT/${TAG}_bench.json
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct \
   --clock-control none -k regex:vfk_cycle_kernel -c 24 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
ls -la $OUT | tail -12
