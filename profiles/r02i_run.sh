#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0 0" "32 256 64" "32 256 32"; do
  set -- $cfg
  timeout 300 python profiles/stage_times.py C5 8 $1 $2 $3 2>&1 | tail -3
done
