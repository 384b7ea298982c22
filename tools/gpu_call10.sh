#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02g.log
for t in 1 0; do echo "=== TMA=$t"; SEPFWI_TMA=$t timeout 600 python tools/quick_perf.py 0 401 c3,c3x8,c5s 2>&1 | grep -E "grad:|per-kernel"; done | tee gpurun_out/qp10.log
