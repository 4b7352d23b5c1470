#!/bin/bash
# A/B of the table-form sin / cos (-DVFK_SINCOS_TABLE build, libvfk_sctab.so) against the default library; parity tests on the variant.
set -u
TAG=${1:-sctab}
OUT=gpurun_out; mkdir -p $OUT
VFK_LIB=$PWD/vfclik_b200/libvfk_sctab.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest (variant) exit $?"; tail -3 $OUT/${TAG}_pytest.log
bash scripts/gpu_variants.sh $TAG default:config3 sctab:config3 default:config5 sctab:config5 default:config3 sctab:config3 default:config5 sctab:config5
VFK_LIB=$PWD/vfclik_b200/libvfk_sctab.so python scripts/fp32_error.py 2>&1 | tail -2
python scripts/fp32_error.py 2>&1 | tail -2
