#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trajectory or geometry or full_size" > gpurun_out/r02e_tests.txt 2>&1
tail -4 gpurun_out/r02e_tests.txt
run() { tag=$1; shift; python bench.py --workload C5 --no-cpu --steps 4 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$tag', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), d['config']['geometry'])"; }
run default
run cpc32_t256 --cpc 32 --threads 256
run cpc32_t256_team32 --cpc 32 --threads 256 --team 32
run cpc32_t512 --cpc 32 --threads 512
run cpc16_t256 --cpc 16 --threads 256
python profiles/stage_times.py C5 8 2>&1 | tail -3
python profiles/stage_times.py C5 8 32 256 64 2>&1 | tail -3
