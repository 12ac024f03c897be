#!/bin/bash
# run under: gpurun --gpus 8 -- bash tools/multigpu_check_r2_n8.sh   (N = 1 and N = 8 bench lines on one box, then the driver checks)
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n1_samebox.err; echo "bench n1 rc=$?"
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 4 --warmup 3 --scaling $sc > gpurun_out/bench_n8_$sc.json 2> gpurun_out/bench_n8_$sc.err; echo "bench $sc rc=$?"
done
python - <<'PY'
import json
for f in ("bench_n1_samebox", "bench_n8_weak", "bench_n8_strong"):
    for line in open(f"gpurun_out/{f}.json"):
        if line.startswith('{'):
            d = json.loads(line); print(f, d['n_gpus'], round(d['value'], 1), 'Mbp/s e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 1), d['clocks'])
PY
timeout 600 python tools/check_multigpu_driver.py 8 > gpurun_out/mg_driver_n8.log 2>&1; echo "driver check rc=$?"; tail -3 gpurun_out/mg_driver_n8.log
timeout 600 python tools/stream_check.py 1 8 > gpurun_out/mg_stream_n8.log 2>&1; echo "stream check rc=$?"; grep -A 12 '"runs"' gpurun_out/stream_check.json | head -16
grep -i "nranks\|NCCL WARN" gpurun_out/bench_n8_weak.err | head -5
