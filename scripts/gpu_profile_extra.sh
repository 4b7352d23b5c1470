#!/bin/bash
# ncu --set full captures of the other bounding regimes: K-fused (FP32 issue / pipe bound) and FP64 config 2.
# One capture per gpurun call would be the rule for long kernels; these two are short (2 launches each).
TAG=${1:-r01l}
OUT=gpurun_out
mkdir -p $OUT
K100="python bench.py --kcycles 100 --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 300 $K100 > $OUT/${TAG}_k100_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vfk_cycle_kernel -s 1 -c 1 -f -o $OUT/${TAG}_k100_prof $K100 > $OUT/${TAG}_k100_ncu.log 2>&1
echo "k100 ncu exit $?"
