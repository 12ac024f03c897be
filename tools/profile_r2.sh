#!/bin/bash
# round-2 profile evidence (run under gpurun, 1 GPU): ncu launch list of the bench command + one --set full capture of the conv kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r2.csv
ncu --set full --clock-control none --import-source on -k regex:conv_ws -s 34 -c 8 -f -o gpurun_out/ncu_ws_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/ncu_ws_r2.ncu-rep
