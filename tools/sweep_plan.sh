SEPFWI_PLAN_DEBUG=1 python tools/grad_probe.py c3 61 8 2>&1 | grep -E "stream_plan|c3" | sort | uniq | tail -5
for lze in 4 8 12 16 24 32; do for lz in 14 26 44 62; do SEPFWI_LZ=$lz SEPFWI_LZE=$lze python tools/grad_probe.py c3 61 8 2>&1 | tail -1; done; done
