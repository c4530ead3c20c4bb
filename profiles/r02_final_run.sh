#!/bin/bash
# round 2, final code: bench lines for every BASELINE configuration on one GPU (C5 with its cpu_baseline leg), the worm
# variants, the 9-chain C5 side line, then the ncu --set full capture of the C5 move kernel and the launch list of the bench command
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 2> gpurun_out/r02f_bench_C5.err | tail -1 > gpurun_out/r02f_bench_C5_1gpu.json
for w in C1 C2 C3 C4; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02f_bench_$w.err | tail -1 > gpurun_out/r02f_bench_${w}_1gpu.json; done
for w in C2 C3; do python bench.py --workload $w --worm --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02f_bench_${w}worm.err | tail -1 > gpurun_out/r02f_bench_${w}worm_1gpu.json; done
python bench.py --chains 9 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r02f_bench_C5x9.err | tail -1 > gpurun_out/r02f_bench_C5_9chains_1gpu.json
profiles/capture.sh r02f C5 8 1024 1024 > /dev/null 2>&1
python profiles/lines_by_number.py gpurun_out/r02f_C5.ncu-rep 0.4 > gpurun_out/r02f_C5_lines.txt 2>&1; rm -f gpurun_out/r02f_C5.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02f_launches_bench_c5.log 2>&1
for f in gpurun_out/r02f_bench_*_1gpu.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], round(d['value']/1e6,1), 'M/s', round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('value'))
PY
done
