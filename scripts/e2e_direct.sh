#!/bin/bash
# Host-buffer session: direct host I/O (zero-copy kernel, VFK_SESSION_DIRECT=1, default) against the chunked copy
# pipeline (=0), same run; then the PCIe request-size probe that explains the result.
for d in 1 0; do
  VFK_SESSION_DIRECT=$d timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('direct $d', 'e2e %.3e'%d['e2e']['value'], 'ms %.3f'%d['e2e']['ms_per_step'], 'launches', d['e2e']['gpu_launches'])"; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_probe scripts/pcie_probe.cu && timeout 120 /tmp/pcie_probe
