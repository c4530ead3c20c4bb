#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/r02r_tests.txt 2>&1
tail -4 gpurun_out/r02r_tests.txt
for w in C5 C1 C2 C3 C4; do
  timeout 400 python bench.py --workload $w --no-cpu --steps 3 > gpurun_out/r02r_bench_$w.json 2> gpurun_out/r02r_bench_$w.err
  python -c "
import json
d=json.load(open('gpurun_out/r02r_bench_$w.json')); print('$w', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), d['config']['geometry'])"
done
