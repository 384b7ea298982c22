#!/bin/bash
# round 2, GPU call 1: full -m gpu suite, variant sweep, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi1.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02a.log
tail -5 gpurun_out/pytest_gpu_r02a.log
D=$PWD/sep-2023_b200
for v in "default 1" "default 0" "minb3 1" "minb3 0"; do
  set -- $v
  lib=$D/libsepfwi.so; [ "$1" = "minb3" ] && lib=$D/libsepfwi_minb3.so
  SEPFWI_LIB=$lib SEPFWI_MERGE_BWD=$2 timeout 600 python tools/quick_perf.py 0 401 c3,c3x8,ref,c5s > gpurun_out/qp_$1_m$2.log 2>&1
  cat gpurun_out/qp_$1_m$2.log
done
for lib in libsepfwi.so libsepfwi_relaxed.so; do SEPFWI_LIB=$D/$lib timeout 300 python tools/quick_perf.py 0 2001 c2,c2x8 2>&1 | tee gpurun_out/qp_c2_$lib.log; done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r02a.json
