#!/bin/bash
# One ncu --set full capture of a bench workload's cycle kernel.  Usage: gpu_ncu_one.sh TAG WORKLOAD [extra bench args]
set -u
TAG=$1; W=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-extras $*"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_ -s 3 -c 2 -f -o $OUT/${TAG}_${W} $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
tail -2 $OUT/${TAG}_plain.log | cut -c1-400
