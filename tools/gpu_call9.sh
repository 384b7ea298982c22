#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q -k "multi_gpu or one_process_per_gpu" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r02_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','shot_gradients_per_s','e2e','fwi','c4_strong','c5_full'):
    print(k, json.dumps(d.get(k))[:700])
P
