#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trajectory or geometry or full_size or permuted or odd_rot" > gpurun_out/r02f_tests.txt 2>&1
tail -12 gpurun_out/r02f_tests.txt
timeout 300 python bench.py --workload C5 --no-cpu --steps 4 > gpurun_out/r02f_C5.json 2> gpurun_out/r02f_C5.err; tail -2 gpurun_out/r02f_C5.err
python -c "
import json
d=json.load(open('gpurun_out/r02f_C5.json')); print('C5', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), d['config']['geometry'])"
timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -3
