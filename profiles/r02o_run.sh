#!/bin/bash
for h in 0 1; do echo "hint(no_allocate)=$h"; PIMC_CELL_HINT=$h timeout 300 python profiles/stage_times.py C5 8 2>&1 | tail -2; done
