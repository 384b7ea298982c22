#!/bin/bash
mkdir -p gpurun_out
D=$PWD/sep-2023_b200
run() { echo "=== $1 force=$2"; SEPFWI_LIB=$D/$1 SEPFWI_FORCE=$2 timeout 600 python tools/quick_perf.py 0 201 $3 2>&1 | grep -E "per-kernel|grad:"; }
(run libsepfwi.so 0 c3,c3x8,ref,c5s; run libsepfwi.so 1 c3,c3x8,ref,c5s; run libsepfwi_io2.so 1 c3x8,c5s; run libsepfwi_io3.so 1 c3,c3x8,c5s; run libsepfwi_io3cg.so 1 c3x8,c5s) 2>&1 | tee gpurun_out/qp8.log
