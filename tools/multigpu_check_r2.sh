#!/bin/bash
# run under: gpurun --gpus N -- bash tools/multigpu_check_r2.sh N
N=${1:-2}
mkdir -p gpurun_out
python tools/check_multigpu_driver.py $N > gpurun_out/mg_driver.log 2>&1; echo "driver check rc=$?"; tail -3 gpurun_out/mg_driver.log
python tools/stream_check.py 1 $N > gpurun_out/mg_stream.log 2>&1; echo "stream check rc=$?"; grep '^{' gpurun_out/mg_stream.log | tail -2
for sc in weak strong; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 4 --warmup 3 --scaling $sc > gpurun_out/bench_n${N}_$sc.json 2> gpurun_out/bench_n${N}_$sc.err; echo "bench $sc rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_n${N}_$sc.json'):
    if line.startswith('{'):
        d=json.loads(line); print('$sc', d['n_gpus'], round(d['value'],1), 'Mbp/s e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1), 'share', round(d['roofline']['kernel_share_of_step'],3))
PY
done
tail -2 gpurun_out/bench_n${N}_weak.err
