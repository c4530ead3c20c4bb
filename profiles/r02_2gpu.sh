#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_2gpu_driver_nccl.txt
timeout 900 python -m pytest tests/test_gpu_statistics.py -q -m gpu -s -k "two_ranks" >> gpurun_out/r02_2gpu_driver_nccl.txt 2>&1
tail -5 gpurun_out/r02_2gpu_driver_nccl.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_C5_2gpu.json 2> gpurun_out/r02_bench_C5_2gpu.err
tail -c 600 gpurun_out/r02_bench_C5_2gpu.json
