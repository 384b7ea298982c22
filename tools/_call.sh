#!/bin/bash
timeout 600 python tools/quick_perf.py 0 401 c2x8,c3x8,c3x16,c3x64,c5s 2>&1 | grep -E 'fwd |per-kernel' | tr -s ' '
