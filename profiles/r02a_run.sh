#!/bin/bash
mkdir -p gpurun_out
profiles/capture.sh r02a C1 148 512 512
profiles/capture.sh r02a C2 148 512 512; rm -f gpurun_out/r02a_C2.ncu-rep
profiles/capture.sh r02a C3 148 1024 1024; rm -f gpurun_out/r02a_C3.ncu-rep
profiles/capture.sh r02a C4 148 128 128; rm -f gpurun_out/r02a_C4.ncu-rep
python profiles/all_workloads.py > gpurun_out/r02a_all_workloads.txt 2>&1
ls -la gpurun_out/; du -sh gpurun_out
