#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_bars.py -q -m gpu -s -k "composed or own_table_plane or C3" > gpurun_out/r02c_tests.txt 2>&1
tail -5 gpurun_out/r02c_tests.txt
# C5: rot-only launch (60 steps after the sweep at time 1152) and one bisection step, with per-line shared-memory / L2 columns
for part in "rot 1153 60" "bis 1152 1"; do
  set -- $part
  ncu --set full --clock-control none --import-source on -k regex:pimc_steps -s 1 -c 1 -f -o gpurun_out/r02c_C5_$1 python profiles/prof_run.py C5 8 $2 $3 > gpurun_out/r02c_C5_$1.log 2>&1
  python profiles/regions.py gpurun_out/r02c_C5_$1.ncu-rep 40 > gpurun_out/r02c_C5_$1_regions.txt 2>&1
  python profiles/hotlines.py gpurun_out/r02c_C5_$1.ncu-rep 12 >> gpurun_out/r02c_C5_$1_regions.txt 2>&1
  rm -f gpurun_out/r02c_C5_$1.ncu-rep
done
# team width experiments (C1, C4)
for t in 0 8 16 32; do python bench.py --workload C1 --team $t --no-cpu --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C1 team', $t, d['value']/1e6, d['config']['geometry'])"; done > gpurun_out/r02c_team.txt 2>&1
for t in 0 2 4 8; do python bench.py --workload C4 --team $t --no-cpu --steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 team', $t, d['value']/1e6, d['config']['geometry'])"; done >> gpurun_out/r02c_team.txt 2>&1
cat gpurun_out/r02c_team.txt
