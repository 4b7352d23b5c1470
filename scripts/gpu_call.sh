#!/bin/bash
# tests then variants
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -6 $OUT/${TAG}_pytest.log
bash scripts/gpu_variants.sh $TAG "$@"
