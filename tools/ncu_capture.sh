#!/bin/bash
# One ncu --set full capture per named kernel case (tools/kernel_cases.py); run under gpurun, ONE GPU.  The reports are large
# (3-6 MB each), so the raw and source pages are exported to (gzipped) CSV on the box and the .ncu-rep is dropped.
# usage: tools/ncu_capture.sh case:regex:skip [case:regex:skip ...]   (skip = launches of that kernel to skip: warm-up calls)
set -u
mkdir -p gpurun_out/ncu
for spec in "$@"; do
  case=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest#*:}
  python tools/kernel_cases.py $case 3 > gpurun_out/ncu/plain_$case.log 2>&1 || { echo "plain run of $case failed"; tail -3 gpurun_out/ncu/plain_$case.log; continue; }
  tail -1 gpurun_out/ncu/plain_$case.log
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/ncu_$case python tools/kernel_cases.py $case 3 > gpurun_out/ncu/ncu_$case.log 2>&1
  echo "ncu $case rc $?"
  ncu -i /tmp/ncu_$case.ncu-rep --page raw --csv > gpurun_out/ncu/$case.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$case.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu/$case.source.csv.gz
  rm -f /tmp/ncu_$case.ncu-rep
done
