#!/bin/bash
# Session A/B: GPU tests on the default build, then default vs a variant build on the listed workloads, twice each.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_s3_ab.sh TAG VARIANT [workloads...]'
set -u
TAG=$1; VAR=$2; shift 2
WL=${@:-config3}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
for W in $WL; do
 for rep in 1 2; do
  for lib in libvfk.so libvfk_$VAR.so; do
    VFK_LIB=$PWD/vfclik_b200/$lib timeout 300 python bench.py --workload $W --warmup 5 --no-cpu-baseline 2>$OUT/${TAG}_${W}_${lib}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d.get('extras',{}).get('k_fused',{}); print('$W $lib', '%.5f'%d['ms_per_step'], '%.4f'%d['roofline']['frac'], 'sust %.4f'%d['roofline'].get('sustained',{}).get('frac',0), 'K100 %.4g'%k.get('value',0), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
 done
done
