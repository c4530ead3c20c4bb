#!/bin/bash
# round 2, final code: C5 at 8 GPUs (torchrun, one rank per GPU) -- the e2e loop with split-phase transfers
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02f_bench_C5_8gpu.err | tail -1 > gpurun_out/r02f_bench_C5_8gpu.json
cat gpurun_out/r02f_bench_C5_8gpu.json | cut -c1-400
tail -3 gpurun_out/r02f_bench_C5_8gpu.err
