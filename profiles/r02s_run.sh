#!/bin/bash
for v in "0 1" "1 1" "2 1" "1 2" "0 2" "1 0"; do set -- $v; echo "GEO_HINT=$1 CELL_HINT=$2"; PIMC_GEO_HINT=$1 PIMC_CELL_HINT=$2 timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2; done
