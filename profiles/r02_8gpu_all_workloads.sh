#!/bin/bash
mkdir -p gpurun_out
for w in C5 C1 C2 C3 C4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --workload $w --steps 3 --warmup 3 > gpurun_out/r02_bench_${w}_8gpu.json 2> gpurun_out/r02_bench_${w}_8gpu.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${w}_8gpu.json')); print('$w', d['n_gpus'], round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), round(d['ms_per_step'],2), 'ms')
except Exception as e: print('$w failed', e)"
done
