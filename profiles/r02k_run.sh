#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trajectory or geometry or full_size or permuted or odd_rot" > gpurun_out/r02k_tests.txt 2>&1
tail -3 gpurun_out/r02k_tests.txt
timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2
PIMC_NO_ROT_RUN=1 timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2
bash profiles/r02j_run.sh 2>&1 | grep -E "dram__bytes_read|lts__t_sector_hit|gpu__time|smsp__inst_executed"
