#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "split_phase or batched_state" > gpurun_out/r02v_tests.txt 2>&1
tail -12 gpurun_out/r02v_tests.txt
timeout 400 python bench.py --workload C5 --steps 6 > gpurun_out/r02v_bench_C5.json 2> gpurun_out/r02v_bench_C5.err; tail -3 gpurun_out/r02v_bench_C5.err
python -c "
import json
d=json.load(open('gpurun_out/r02v_bench_C5.json')); print('C5', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3)); print(d.get('cpu_baseline'))"
