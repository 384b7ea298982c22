#!/bin/bash
# round 2, GPU call 2: one-pass reverse-time kernel -- parity, timing, ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02b.log
tail -8 gpurun_out/pytest_gpu_r02b.log
for m in 1 0; do
  SEPFWI_MERGE_BWD=$m timeout 600 python tools/quick_perf.py 0 401 c3,c3x8,ref,c5s 2>&1 | tee gpurun_out/qp2_m$m.log
done
for lz in 26 62 122; do
  SEPFWI_LZ=$lz timeout 300 python tools/quick_perf.py 0 201 c3x8,c5s 2>&1 | tee gpurun_out/qp2_m1_lz$lz.log
done
ncu --set full --import-source on --clock-control none -k k_stream_bwd -s 8 -c 1 -o gpurun_out/r02_k_stream_bwd_c5 python tools/profile_step.py c5s 24 > gpurun_out/prof_full_bwd.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k_stream_bwd_c5.ncu-rep | tee gpurun_out/r02_k_stream_bwd_c5_ncu_summary.txt
