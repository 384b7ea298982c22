#!/bin/bash
# Captures the ncu evidence committed under profiles/ (run on the GPU box through gpurun):
#   launch lists (gpu__time_duration per launch) of a short forward + gradient run on the C5-size and C2/C3 grids,
#   and one --set full capture per streaming kernel on the HBM-bound grid.
set -x
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_stream_launches_c5_2080x8064.csv python tools/profile_step.py c5s 30 > gpurun_out/prof_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_stream_launches_c3_416x1764.csv python tools/profile_step.py c3 60 > gpurun_out/prof_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_stream_launches_c2_480x1064.csv python tools/profile_step.py c2 100 > gpurun_out/prof_c2.log 2>&1
for k in k_stream_fwd k_stream_recon k_stream_adj; do
  ncu --set full --import-source on --clock-control none -k $k -s 8 -c 2 -o gpurun_out/${R}_${k}_c5 python tools/profile_step.py c5s 24 > gpurun_out/prof_full_$k.log 2>&1
done
ncu --set full --import-source on --clock-control none -k k_stream_fwd -s 30 -c 2 -o gpurun_out/${R}_k_stream_fwd_c2 python tools/profile_step.py c2 60 > gpurun_out/prof_full_c2.log 2>&1
ls -la gpurun_out
