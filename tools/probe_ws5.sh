#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
{
for strip in 21 0; do
  echo "== trace variant 4 strip $strip"; JG_TRACE=1 timeout 300 ./build/conv_probe 0 4 592 10 $strip 2>&1 | grep -E "TIMING|RESULT|error|failed|ws CTA0|MMA warp"
done
} > gpurun_out/probe_ws5.log 2>&1
cat gpurun_out/probe_ws5.log
