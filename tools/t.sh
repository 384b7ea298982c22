SEPFWI_PLAN_DEBUG=1 python tools/quick_perf.py 0 401 2>&1 | grep -v "^stream_plan" ; SEPFWI_PLAN_DEBUG=1 python tools/quick_perf.py 0 101 2>&1 | grep "^stream_plan" | sort | uniq
