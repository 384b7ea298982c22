#!/bin/bash
for wl in c3x64 c5s; do
  echo "== $wl default: $(timeout 300 python tools/quick_perf.py 0 121 $wl 2>&1 | grep -E 'per-kernel')"
  for cfg in "8 4" "12 4" "20 4" "20 8" "32 6" "32 12" "50 8" "50 16" "80 12"; do set -- $cfg
    echo "$wl LZ=$1 LZE=$2: $(SEPFWI_LZ=$1 SEPFWI_LZE=$2 timeout 300 python tools/quick_perf.py 0 121 $wl 2>&1 | grep -E 'per-kernel')"
  done
done
