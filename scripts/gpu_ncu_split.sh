#!/bin/bash
# ncu captures of the config 5 launch: lane-split kernel and the one-thread-per-instance kernel.
set -u
TAG=${1:-ncu}
W=${2:-config5}
OUT=gpurun_out
mkdir -p $OUT
export VFK_SPLIT=1
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $CMD > $OUT/${TAG}_plain_split.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_ -s 3 -c 2 -f -o $OUT/${TAG}_${W}_split $CMD > $OUT/${TAG}_ncu_split.log 2>&1
echo "ncu split exit $?"
unset VFK_SPLIT
timeout 300 $CMD > $OUT/${TAG}_plain_solo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_ -s 3 -c 2 -f -o $OUT/${TAG}_${W}_solo $CMD > $OUT/${TAG}_ncu_solo.log 2>&1
echo "ncu solo exit $?"
ls -la $OUT | tail -6
