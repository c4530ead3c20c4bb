#!/bin/bash
# ncu --set full capture of the move kernel for one workload + text summary (metrics, stalls, hottest source lines).
# usage: profiles/capture.sh <tag> <workload> <chains> <warm_steps> <nsteps> [cpc threads team]   (run on the GPU box, under gpurun)
set -u
tag=$1; w=$2; shift 2
rep=gpurun_out/${tag}_${w}
ncu --set full --clock-control none --import-source on -k regex:pimc_steps -s 1 -c 1 -f -o $rep python profiles/prof_run.py $w "$@" > $rep.log 2>&1
out=gpurun_out/${tag}_${w}_steps_kernel.txt
{
  echo "# ncu --set full --clock-control none --import-source on -k regex:pimc_steps -s 1 -c 1 python profiles/prof_run.py $w $*"
  grep "us/step" $rep.log
  ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | python profiles/raw_metrics.py
  echo "# stall reasons (all samples) and hottest source lines: python profiles/hotlines.py <report>"
  python profiles/hotlines.py $rep.ncu-rep 28
} > $out 2>&1
tail -n 60 $out
