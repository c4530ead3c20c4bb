#!/bin/bash
for ch in 8 9; do timeout 300 python profiles/stage_times.py C5 $ch 2>&1 | tail -2; done
python bench.py --workload C5 --chains 9 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1
