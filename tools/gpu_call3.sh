#!/bin/bash
mkdir -p gpurun_out
D=$PWD/sep-2023_b200
timeout 600 python tools/debug_c4.py 301 2>&1 | tee gpurun_out/debug_c4.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r02c.log
for lib in libsepfwi.so libsepfwi_u1.so libsepfwi_u6.so; do
  for m in 1 0; do
    SEPFWI_LIB=$D/$lib SEPFWI_MERGE_BWD=$m timeout 600 python tools/quick_perf.py 0 401 c3,c3x8,ref,c5s 2>&1 | tee gpurun_out/qp3_${lib}_m$m.log
  done
done
ncu --set full --import-source on --clock-control none -k k_stream_bwd -s 8 -c 1 -o gpurun_out/r02_k_stream_bwd_c5_v2 python tools/profile_step.py c5s 24 > gpurun_out/prof_full_bwd2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k_stream_bwd_c5_v2.ncu-rep | tee gpurun_out/r02_k_stream_bwd_c5_v2_ncu_summary.txt
