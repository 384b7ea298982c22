#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k k_stream_adj -s 12 -c 1 -o gpurun_out/r02_k_stream_adj_c3x8 python tools/profile_step.py c3 30 8 > gpurun_out/x12.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_k_stream_adj_c3x8.ncu-rep | tee gpurun_out/r02_k_stream_adj_c3x8_ncu_summary.txt
ncu -i gpurun_out/r02_k_stream_adj_c3x8.ncu-rep --page source --csv --print-source cuda > gpurun_out/r02_k_stream_adj_c3x8_source.csv 2>/dev/null
wc -l gpurun_out/r02_k_stream_adj_c3x8_source.csv
