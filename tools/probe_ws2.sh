#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
{
timeout 120 ./build/mma_rate_probe 4000
for strip in 21 17 0; do
  echo "== trace variant 4 strip $strip"; JG_TRACE=1 timeout 300 ./build/conv_probe 0 4 296 10 $strip 2>&1 | grep -E "TIMING|RESULT|error|failed|ws CTA0|per sub-tile"
done
} > gpurun_out/probe_ws2.log 2>&1
cat gpurun_out/probe_ws2.log
