#!/bin/bash
# A/B of a sin / cos table build (libvfk_<NAME>.so) against the default library + parity tests on the variant.
# Usage: gpurun -- 'bash scripts/gpu_tab_ab.sh TAG NAME'
set -u
TAG=$1; NAME=$2
OUT=gpurun_out; mkdir -p $OUT
VFK_LIB=$PWD/vfclik_b200/libvfk_$NAME.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest (variant) exit $?"; tail -3 $OUT/${TAG}_pytest.log
bash scripts/gpu_variants.sh $TAG default:config3 $NAME:config3 default:config5 $NAME:config5 default:config3 $NAME:config3 default:config5 $NAME:config5 | grep -v "e2e roofline\|fp64_config2"
VFK_LIB=$PWD/vfclik_b200/libvfk_$NAME.so python scripts/fp32_error.py 2>&1 | tail -1
python scripts/fp32_error.py 2>&1 | tail -1
