#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ws -c 1 -f -o gpurun_out/ws_light ./build/conv_probe 0 4 296 0 21 > gpurun_out/ncu_ws_light.log 2>&1
tail -3 gpurun_out/ncu_ws_light.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -c 1 -f -o gpurun_out/tc2_light ./build/conv_probe 0 2 296 0 21 > gpurun_out/ncu_tc2_light.log 2>&1
tail -3 gpurun_out/ncu_tc2_light.log
ls -la gpurun_out/*.ncu-rep
