#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_worm.py -q -m gpu -x 2>&1 | tail -3
PROF_WORM=1 TL_WORKLOAD=C2 TL_CHAINS=148 PIMCGPU_LIB=moribs-pimc_b200/csrc/libpimcgpu_tl.so timeout 200 python profiles/timeline.py 0 0 0 1029 2 400 > gpurun_out/r02x_worm_timeline.txt 2>&1
PROF_WORM=1 timeout 200 python profiles/prof_run.py C2 148 512 512 2>&1 | tail -1
PROF_WORM=1 timeout 200 python profiles/prof_run.py C3 148 1024 1024 2>&1 | tail -1
