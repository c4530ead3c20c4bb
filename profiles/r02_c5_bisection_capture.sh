#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pimc_steps -s 1 -c 1 -f -o gpurun_out/r02g_C5_bis python profiles/prof_run.py C5 8 1152 1 > gpurun_out/r02g_C5_bis.log 2>&1
python profiles/lines_by_number.py gpurun_out/r02g_C5_bis.ncu-rep 0.2 > gpurun_out/r02g_C5_bis_lines.txt 2>&1
python profiles/regions.py gpurun_out/r02g_C5_bis.ncu-rep 10 > gpurun_out/r02g_C5_bis_regions.txt 2>&1
ncu -i gpurun_out/r02g_C5_bis.ncu-rep --page raw --csv 2>/dev/null | python profiles/raw_metrics.py >> gpurun_out/r02g_C5_bis_regions.txt
rm -f gpurun_out/r02g_C5_bis.ncu-rep
