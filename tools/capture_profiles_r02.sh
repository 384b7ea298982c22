#!/bin/bash
# Round-2 evidence (run on the GPU box through gpurun): bench lines, ncu launch lists, ncu --set full summaries of the time-stepping
# kernels on the headline workload (C3 single-shot gradient) and on the HBM-bound grid.
set -x
R=r02
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_reference_arm.json 2> gpurun_out/${R}_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches_c3_416x1764.csv python tools/profile_step.py c3 100 > gpurun_out/prof_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c5_2080x8064.csv python tools/profile_step.py c5s 30 > gpurun_out/prof_c5.log 2>&1
# default path: the reverse-time step of these two workloads is ONE launch (k_stream_bwd)
for k in k_stream_fwd k_stream_bwd; do
  ncu --set full --import-source on --clock-control none -k $k -s 60 -c 1 -o gpurun_out/${R}_${k}_c3 python tools/profile_step.py c3 80 > gpurun_out/prof_full_c3_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/${R}_${k}_c3.ncu-rep > gpurun_out/${R}_${k}_c3_ncu_summary.txt
done
ncu --set full --import-source on --clock-control none -k k_stream_bwd -s 8 -c 1 -o gpurun_out/${R}_k_stream_bwd_c5 python tools/profile_step.py c5s 24 > gpurun_out/prof_full_c5_bwd.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_k_stream_bwd_c5.ncu-rep > gpurun_out/${R}_k_stream_bwd_c5_ncu_summary.txt
# the two-launch form (what batches of many waves run): SEPFWI_MERGE=0
for k in k_stream_recon k_stream_adj; do
  SEPFWI_MERGE=0 ncu --set full --import-source on --clock-control none -k $k -s 60 -c 1 -o gpurun_out/${R}_${k}_c3 python tools/profile_step.py c3 80 > gpurun_out/prof_full_c3_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/${R}_${k}_c3.ncu-rep > gpurun_out/${R}_${k}_c3_ncu_summary.txt
  SEPFWI_MERGE=0 ncu --set full --import-source on --clock-control none -k $k -s 8 -c 1 -o gpurun_out/${R}_${k}_c5 python tools/profile_step.py c5s 24 > gpurun_out/prof_full_c5_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/${R}_${k}_c5.ncu-rep > gpurun_out/${R}_${k}_c5_ncu_summary.txt
done
ls -la gpurun_out | tail -30
