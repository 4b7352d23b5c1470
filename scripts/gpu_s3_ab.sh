#!/bin/bash
# Session A/B: the default build against variant builds (vfclik_b200/libvfk_<name>.so) on the listed workloads, twice each.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_s3_ab.sh TAG "var1 var2" "config3 config2" [pytest]'
set -u
TAG=$1; VARS=$2; WL=${3:-config3}
OUT=gpurun_out; mkdir -p $OUT
if [ "${4:-}" = "pytest" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
fi
for W in $WL; do
 for rep in 1 2; do
  for v in default $VARS; do
    if [ $v = default ]; then lib=libvfk.so; else lib=libvfk_$v.so; fi
    VFK_LIB=$PWD/vfclik_b200/$lib timeout 300 python bench.py --workload $W --warmup 5 --no-cpu-baseline 2>$OUT/${TAG}_${W}_${v}.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d.get('extras',{}).get('k_fused',{}); f2=d.get('extras',{}).get('fp64_config2',{}); print('$W %-10s' % '$v', '%.5f'%d['ms_per_step'], '%.4f'%d['roofline']['frac'], 'sust %.4f'%d['roofline'].get('sustained',{}).get('frac',0), 'K100 %.4g'%k.get('value',0), 'c2 %.5f'%f2.get('ms_per_launch',0), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
 done
done
for v in $VARS; do VFK_LIB=$PWD/vfclik_b200/libvfk_$v.so timeout 300 python scripts/fp32_error.py 2>&1 | tail -1 | sed "s/^/$v fp32_error: /"; done
timeout 300 python scripts/fp32_error.py 2>&1 | tail -1 | sed "s/^/default fp32_error: /"
