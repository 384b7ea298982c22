#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "sponge or elastic_solver or c1_gpu" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02f.log
timeout 600 python bench.py --steps 2 --warmup 3 --extras c2_forward --no-ncu 2>gpurun_out/b7.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['c2_forward'], indent=1)); print(d['value'], d['e2e'], d['loop_ms'])"
