#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --skip-train --skip-side-legs --skip-cpu-baseline --skip-sweep --skip-cold"
$CMD > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
CMD2="python tools/kbench.py --profile"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo -o gpurun_out/r2_halo $CMD2 > gpurun_out/ncu2.log 2>&1
echo "set full rc=$?"; ls -la gpurun_out/r2_halo.ncu-rep; tail -3 gpurun_out/ncu2.log
python tools/kbench.py --markdown > gpurun_out/r2_kbench.md 2> gpurun_out/r2_kbench.err; tail -45 gpurun_out/r2_kbench.md
