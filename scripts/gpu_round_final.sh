#!/bin/bash
# Round record: GPU tests, smoke, default bench, reference arm, ncu launch list + full captures (configs 3, 2, 5).
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round_final.sh TAG'
set -u
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
echo "smoke exit $?"; cat $OUT/${TAG}_smoke.log | tail -2
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
echo "reference arm exit $?"
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread \
   --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_list.log 2>&1
echo "ncu list exit $?"
for W in config3 config2 config5; do
  CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
  timeout 300 $CMD > $OUT/${TAG}_plain_$W.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_cycle_kernel -s 3 -c 2 -f -o $OUT/${TAG}_${W}_prof $CMD > $OUT/${TAG}_ncu_$W.log 2>&1
  echo "ncu $W exit $?"
done
# the K-fused launch of the headline shape
CMD="python bench.py --steps 2 --warmup 3 --kcycles 100 --no-cpu-baseline --no-extras"
timeout 300 $CMD > $OUT/${TAG}_plain_k100.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_cycle_kernel -s 3 -c 1 -f -o $OUT/${TAG}_k100_prof $CMD > $OUT/${TAG}_ncu_k100.log 2>&1
echo "ncu k100 exit $?"
ls -la $OUT | grep $TAG
