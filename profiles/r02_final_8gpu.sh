#!/bin/bash
# round 2, final code: C5 at 8 GPUs and at 1 GPU with the per-call host trace of the e2e loop (PIMC_E2E_TRACE)
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "split_phase" 2>&1 | tail -2
PIMC_E2E_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02x_8gpu_trace.err | tail -1 > gpurun_out/r02f_bench_C5_8gpu.json
grep "e2e trace rank 0" gpurun_out/r02x_8gpu_trace.err
CUDA_VISIBLE_DEVICES=0 PIMC_E2E_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02x_1gpu_trace.err | tail -1 > gpurun_out/r02f_bench_C5_1gpu_nocpu.json
grep "e2e trace" gpurun_out/r02x_1gpu_trace.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r02f_bench_C5_8gpu.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/r02f_bench_C5_1gpu_nocpu.json').read().strip().splitlines()[-1])
print('8gpu', a['value']/1e6, a['e2e']['value']/1e6, '1gpu', b['value']/1e6, b['e2e']['value']/1e6, 'eff', a['value']/8/b['value'], a['e2e']['value']/8/b['e2e']['value'])
PY
