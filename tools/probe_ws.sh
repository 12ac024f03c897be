#!/bin/bash
# correctness then timing of the weights-stationary conv kernel against the CTA-pair kernel (run under gpurun)
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
{
for cfg in "0 4 2 0 21" "0 4 2 0 17" "1 4 2 0 1" "0 4 2 0 32" "0 4 2 0 0" "0 4 2 0 8" "0 4 37 0 0"; do
  echo "== conv_probe $cfg"; timeout 120 ./build/conv_probe $cfg 2>&1 | tail -12
done
if [ "$1" != "quick" ]; then
for strip in 21 17 32 0; do
  for v in 2 3 4; do
    echo "== timing variant $v strip $strip"; timeout 300 ./build/conv_probe 0 $v 2368 20 $strip 2>&1 | grep -E "TIMING|RESULT|error|failed"
  done
done
echo "== stem"; for v in 2 4; do timeout 300 ./build/conv_probe 1 $v 2368 20 1 2>&1 | grep -E "TIMING|RESULT|error|failed"; done
fi
} > gpurun_out/probe_ws.log 2>&1
tail -80 gpurun_out/probe_ws.log
