#!/bin/bash
# run under: gpurun --gpus N -- bash tools/bench_n.sh N [weak|strong]   -- one torchrun bench line on N GPUs
N=${1:-2}; SC=${2:-weak}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 4 --warmup 3 --scaling $SC --no-cpu-baseline > gpurun_out/bench_n${N}_$SC.json 2> gpurun_out/bench_n${N}_$SC.err
echo "rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_n${N}_$SC.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print('$SC', d['n_gpus'], round(d['value'], 1), 'Mbp/s e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 1), 'rank ms', d.get('rank_ms_per_step'), 'share', round(d['roofline']['kernel_share_of_step'], 3))
PY
tail -4 gpurun_out/bench_n${N}_$SC.err
