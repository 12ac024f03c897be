#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ws -c 1 -f -o gpurun_out/ws3_final ./build/conv_probe 0 4 296 0 32 > gpurun_out/ncu_ws3_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ws -c 1 -f -o gpurun_out/ws3_light ./build/conv_probe 0 4 296 0 21 > gpurun_out/ncu_ws3_light.log 2>&1
ls -la gpurun_out/ws3_*.ncu-rep
