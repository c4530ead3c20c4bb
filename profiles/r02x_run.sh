#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_worm.py -q -m gpu -x 2>&1 | tail -2
for w in C1 C2 C3 C4; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/r02g_bench_${w}_1gpu.json; python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench_${w}_1gpu.json').read().strip().splitlines()[-1]); print('$w', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), round(d['roofline']['frac'],3))"; done
