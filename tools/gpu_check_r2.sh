#!/bin/bash
# round-2 GPU check: parity tests, smoke, a short bench of every config (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for c in 2 3 4; do
  python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err; echo "bench config $c rc=$?"
  tail -c 1500 gpurun_out/bench_c$c.json; tail -3 gpurun_out/bench_c$c.err
done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
