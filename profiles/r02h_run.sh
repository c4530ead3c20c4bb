#!/bin/bash
mkdir -p gpurun_out
PIMCGPU_LIB=moribs-pimc_b200/csrc/libpimcgpu_tl.so timeout 120 python profiles/timeline.py 0 0 0 1029 6 50 > gpurun_out/r02h_timeline.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trajectory or geometry or full_size or permuted or odd_rot" > gpurun_out/r02h_tests.txt 2>&1
tail -5 gpurun_out/r02h_tests.txt
timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2
