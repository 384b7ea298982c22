#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/quick_perf.py 0 201 c3x8,c3x16,c3x32,c3x64 2>&1 | grep -E "grad:|per-kernel" | tee gpurun_out/qp11.log
