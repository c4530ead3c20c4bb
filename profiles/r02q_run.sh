#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_statistics.py tests/test_gpu_boundary.py tests/test_gpu_parity_bars.py -q -m gpu -s -k "top_in_helium or linear_dopant or tip4p or rotden_composed or reference_main or worm_deck or writes_reference" > gpurun_out/r02q_tests.txt 2>&1
grep -E "sigma \(bar|passed|failed|^FAILED|Error" gpurun_out/r02q_tests.txt | head -70
