#!/bin/bash
# ncu evidence for the shared-memory-resident forward loop on the C2 workload (run on the GPU box through gpurun):
#   launch list of the bench command and one --set full capture of the cooperative launch at the full nt = 4001.
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_resident_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_resident -s 1 -c 1 -o gpurun_out/${R}_k_resident_fwd_c2 python tools/res_probe.py c2 4001 pr,vx,vz,ett > gpurun_out/ncu_res_full.log 2>&1
tail -2 gpurun_out/ncu_res_full.log
