#!/bin/bash
# round 2, final code: ncu --set full of the free-running top variant (pimc_steps_kernel<10>) on C1 and C3
mkdir -p gpurun_out
profiles/capture.sh r02f C1 148 512 512 > /dev/null 2>&1; rm -f gpurun_out/r02f_C1.ncu-rep
profiles/capture.sh r02f C3 148 1024 1024 > /dev/null 2>&1; rm -f gpurun_out/r02f_C3.ncu-rep
head -30 gpurun_out/r02f_C1_steps_kernel.txt
