#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --config 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_ws.json 2> gpurun_out/bench_c2_ws.err; echo "bench ws rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2_ws.json'))
print('ws:', round(d['value'],1),'Mbp/s e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),d['roofline']['kernel'],d['clocks'])
PY
JG_CONV_IMPL=3 python bench.py --config 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r1k.json 2> gpurun_out/bench_c2_r1k.err; echo "bench r1 kernels rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2_r1k.json'))
print('r1 kernels:', round(d['value'],1),'Mbp/s e2e',round(d['e2e']['value'],1),'roofline',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),d['roofline']['kernel'],d['clocks'])
PY
python tools/layer_profile.py > gpurun_out/layer_profile_ws.log 2>&1; tail -25 gpurun_out/layer_profile_ws.log
