#!/bin/bash
export JG_PROBE_FOLDED=1
mkdir -p gpurun_out
{
for b in base DJG_WS_VSLOTS4 DJG_WS_HELPER_VALIDITY_FIRST DJG_WS_FINAL_UPFRONT; do
  for rep in 1 2; do
  echo "== $b strip 32 rep $rep"; timeout 120 ./build/conv_probe_$b 0 4 592 5 32 2>&1 | grep -E "TIMING|RESULT|error|failed"
  done
done
echo "== sanitizer"; timeout 300 compute-sanitizer --tool memcheck ./build/conv_probe_base 0 4 8 0 32 2>&1 | tail -15
} > gpurun_out/probe_ws7.log 2>&1
cat gpurun_out/probe_ws7.log
