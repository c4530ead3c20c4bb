#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_bars.py -q -m gpu -x > gpurun_out/r02d_tests.txt 2>&1
tail -15 gpurun_out/r02d_tests.txt
python bench.py --workload C5 --no-cpu --steps 5 > gpurun_out/r02d_bench_C5.json 2> gpurun_out/r02d_bench_C5.err
python bench.py --workload C1 --no-cpu --steps 5 > gpurun_out/r02d_bench_C1.json 2> gpurun_out/r02d_bench_C1.err
PIMC_NO_GEO=1 python bench.py --workload C5 --no-cpu --steps 3 > gpurun_out/r02d_bench_C5_nogeo.json 2>/dev/null
PIMC_NO_POLY1D=1 python bench.py --workload C5 --no-cpu --steps 3 > gpurun_out/r02d_bench_C5_nopoly.json 2>/dev/null
for f in gpurun_out/r02d_bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms frac', round(d['roofline']['frac'],3), d['config']['geometry'])"; done
python profiles/stage_times.py C5 8 2>&1 | tail -8
