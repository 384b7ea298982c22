#!/bin/bash
# Chunk-height sweep used to calibrate stream_plan's cost model: per-kernel us per step for forced (Lz, Le) on the BASELINE grids.
#   bash tools/sweep_plan.sh > gpurun_out/sweep_plan.log      (SEPFWI_LZ / SEPFWI_LZE are read at handle creation)
for wl in c3 c3x8 ref; do
  echo "== $wl default: $(timeout 200 python tools/quick_perf.py 0 201 $wl 2>&1 | grep -E 'per-kernel|ref ')"
  for lze in 1 2 3 4 6; do for lz in 4 8 12 20 32 48; do
    echo "$wl LZ=$lz LZE=$lze: $(SEPFWI_LZ=$lz SEPFWI_LZE=$lze timeout 200 python tools/quick_perf.py 0 201 $wl 2>&1 | grep -E 'per-kernel|ref ' | tr -s ' ')"
  done; done
done
