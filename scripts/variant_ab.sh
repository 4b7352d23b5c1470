#!/bin/bash
# A/B of a build variant against the default library: scripts/variant_ab.sh <variant .so name>
for lib in libvfk.so $1 libvfk.so $1; do
  VFK_LIB=$PWD/vfclik_b200/$lib timeout 250 python bench.py --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['ms_per_step'], '%.4f'%d['roofline']['frac'], 'K100 %.4g'%d['extras']['k_fused']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
VFK_LIB=$PWD/vfclik_b200/$1 python scripts/fp32_error.py 2>&1 | tail -1
python scripts/fp32_error.py 2>&1 | tail -1
