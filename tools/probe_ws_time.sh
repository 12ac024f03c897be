#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
{
for strip in 21 17 32 0; do
  for v in 2 3 4; do
    echo "== timing variant $v strip $strip"; timeout 300 ./build/conv_probe 0 $v 2368 20 $strip 2>&1 | grep -E "TIMING|RESULT|error|failed"
  done
done
echo "== stem"; for v in 2 4; do timeout 300 ./build/conv_probe 1 $v 2368 20 1 2>&1 | grep -E "TIMING|RESULT|error|failed"; done
} > gpurun_out/probe_ws_time.log 2>&1
cat gpurun_out/probe_ws_time.log
