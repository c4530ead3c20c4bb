#!/bin/bash
for gh in 0 1 2; do for ch in 1 2; do echo "GEO_HINT=$gh CELL_HINT=$ch"; PIMC_GEO_HINT=$gh PIMC_CELL_HINT=$ch timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2 | head -1; done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "trajectory or geometry or full_size" 2>&1 | tail -3
