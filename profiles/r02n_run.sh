#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_statistics.py -q -m gpu -s -k "boundary or reference_main or top_in_helium or linear_dopant or tip4p" > gpurun_out/r02n_tests.txt 2>&1
grep -E "sigma|passed|failed|Error|assert" gpurun_out/r02n_tests.txt | head -60
