#!/bin/bash
# round-2 check of the window-resident kernel: new GPU tests, then bench --config 3 with and without it (same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident or 500bp or segmentation" 2>&1 | tail -5
python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_resident.json 2> gpurun_out/bench_c3_resident.err; tail -c 2500 gpurun_out/bench_c3_resident.json
JG_RESIDENT=0 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_perlayer.json 2> gpurun_out/bench_c3_perlayer.err
python - <<'PY'
import json
for f in ("bench_c3_resident", "bench_c3_perlayer"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["windows_per_s"]), "windows/s device", round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), d["roofline"]["kernel"], "frac", round(d["roofline"]["frac"], 3), d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
