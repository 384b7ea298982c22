#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k scratch 2>&1 | tail -5
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/san_probe.py 30 > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
