#!/bin/bash
mkdir -p gpurun_out
D=$PWD/sep-2023_b200
SEPFWI_LIB=$D/libsepfwi_cg_minb3.so ncu --set full --import-source on --clock-control none -k k_stream_fwd -s 30 -c 1 -o gpurun_out/x_fwd_cg_minb3 python tools/profile_step.py c5s 24 > gpurun_out/x1.log 2>&1
SEPFWI_LIB=$D/libsepfwi_minb3.so ncu --set full --import-source on --clock-control none -k k_stream_fwd -s 30 -c 1 -o gpurun_out/x_fwd_minb3 python tools/profile_step.py c5s 24 > gpurun_out/x2.log 2>&1
ncu --set full --import-source on --clock-control none -k k_stream_fwd -s 30 -c 1 -o gpurun_out/r02_k_stream_fwd_c5 python tools/profile_step.py c5s 24 > gpurun_out/x3.log 2>&1
for f in x_fwd_cg_minb3 x_fwd_minb3 r02_k_stream_fwd_c5; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_summary.txt; ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv; done
cat gpurun_out/x_fwd_cg_minb3_summary.txt gpurun_out/r02_k_stream_fwd_c5_summary.txt
