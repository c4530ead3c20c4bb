#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload C5 --no-cpu --steps 5 > gpurun_out/r02m_C5.json 2> gpurun_out/r02m_C5.err; tail -2 gpurun_out/r02m_C5.err
PIMC_NO_BIS_PIPE=1 timeout 300 python bench.py --workload C5 --no-cpu --steps 3 > gpurun_out/r02m_C5_nobispipe.json 2>/dev/null
PIMC_NO_ROT_RUN=1 timeout 300 python bench.py --workload C5 --no-cpu --steps 3 > gpurun_out/r02m_C5_norotrun.json 2>/dev/null
for f in gpurun_out/r02m_C5*.json; do python -c "
import json
d=json.load(open('$f')); print('$f', round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3))"; done
timeout 900 python -m pytest tests/test_gpu_statistics.py -q -m gpu -s -k "top_in_helium or tip4p or linear_dopant" 2>&1 | grep -E "sigma|passed|failed|Error" | head -60
