#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r02_bench_n8.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','shot_gradients_per_s','e2e','fwi','c4_strong','c5_full'):
    print(k, json.dumps(d.get(k))[:600])
P
