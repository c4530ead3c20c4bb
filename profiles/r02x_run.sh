#!/bin/bash
PIMCGPU_LIB=moribs-pimc_b200/csrc/libpimcgpu_pre.so timeout 200 python profiles/dbg_bitident.py /tmp/pre.npz 2>&1 | tail -1
timeout 200 python profiles/dbg_bitident.py /tmp/new.npz 2>&1 | tail -1
python - <<'PY'
import numpy as np
a=np.load('/tmp/pre.npz'); b=np.load('/tmp/new.npz')
bad=[k for k in a.files if not np.array_equal(a[k], b[k])]
print("keys", len(a.files), "differing:", bad)
PY
for w in C1 C2 C3; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/r02h_bench_${w}_1gpu.json; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_${w}_1gpu.json').read().strip().splitlines()[-1]); print('$w', round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), round(d['roofline']['frac'],3))"; done
