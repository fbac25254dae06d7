#!/bin/bash
# Everything the committed evidence of a round is made from, in ONE gpurun call (one GPU):
#   1. bench.py (C2) and the other BASELINE configs, the reference arm            -> gpurun_out/<tag>_bench_*.json
#   2. the launch list of steady-state steps (ncu gpu__time_duration, eager mode) -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full captures of the kernel cases given as arguments             -> gpurun_out/ncu/
# usage: bash tools/profile_round.sh <tag> [case:regex:skip ...]      (numbers printed under ncu are never bench values)
set -u
tag=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/${tag}_bench_C2.json 2> gpurun_out/${tag}_bench_C2.err; echo "bench C2 rc $?"; python tools/kt.py gpurun_out/${tag}_bench_C2.json | head -3
for c in C1 C3 C4 C5; do
  timeout 300 python bench.py --config $c --no-cpu-baseline --no-kernel-roofline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc $?"
  python tools/kt.py gpurun_out/${tag}_bench_$c.json | head -1
done
timeout 400 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "reference arm rc $?"; tail -c 300 gpurun_out/${tag}_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2700 --launch-count 1700 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-roofline > gpurun_out/${tag}_launches_run.log 2>&1; echo "launch list rc $?"
python tools/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1; head -12 gpurun_out/${tag}_launches_summary.txt
if [ $# -gt 0 ]; then timeout 900 bash tools/ncu_capture.sh "$@"; fi
