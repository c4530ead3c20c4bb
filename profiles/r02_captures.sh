#!/bin/bash
# round 2: ncu --set full capture of the move kernel for every BASELINE configuration + launch list of the bench command
mkdir -p gpurun_out
profiles/capture.sh r02 C5 8 1024 1024
python profiles/lines_by_number.py gpurun_out/r02_C5.ncu-rep 0.4 > gpurun_out/r02_C5_lines.txt 2>&1; rm -f gpurun_out/r02_C5.ncu-rep
profiles/capture.sh r02 C1 148 512 512
python profiles/lines_by_number.py gpurun_out/r02_C1.ncu-rep 0.4 > gpurun_out/r02_C1_lines.txt 2>&1; rm -f gpurun_out/r02_C1.ncu-rep
profiles/capture.sh r02 C2 148 512 512; rm -f gpurun_out/r02_C2.ncu-rep
profiles/capture.sh r02 C3 148 1024 1024; rm -f gpurun_out/r02_C3.ncu-rep
profiles/capture.sh r02 C4 148 128 128; rm -f gpurun_out/r02_C4.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench_c5.log 2>&1
ls -la gpurun_out | head -40
