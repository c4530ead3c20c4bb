#!/bin/bash
for v in "0 0" "1 0" "1 1"; do set -- $v; echo "CELL4=$1 HINT=$2"; PIMC_CELL4=$1 PIMC_CELL_HINT=$2 timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2; done
PIMC_CELL4=0 bash profiles/r02j_run.sh 2>&1 | grep -E "dram__bytes_read|lts__t_sector_hit|gpu__time|smsp__inst_executed"
