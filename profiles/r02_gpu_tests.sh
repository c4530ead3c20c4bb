#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu > gpurun_out/r02p_tests.txt 2>&1
tail -15 gpurun_out/r02p_tests.txt
