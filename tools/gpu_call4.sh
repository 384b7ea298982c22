#!/bin/bash
mkdir -p gpurun_out
D=$PWD/sep-2023_b200
run() { echo "=== $1 pad=$2"; SEPFWI_LIB=$D/$1 SEPFWI_SMEM_PAD=$2 timeout 600 python tools/quick_perf.py 0 201 c3x8,c5s 2>&1 | grep -E "per-kernel|lib:|grad:"; }
(run libsepfwi.so 0; run libsepfwi.so 36864; run libsepfwi_cg.so 0; run libsepfwi_cg.so 36864; run libsepfwi_cg_minb3.so 0) 2>&1 | tee gpurun_out/qp4.log
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r02d.log
