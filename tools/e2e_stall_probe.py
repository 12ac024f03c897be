"""Where does the host block during an end-to-end step? (development aid)"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from jaeger_b200 import B200Engine, parse_project, standin_1p4m_config
from bench import synth_batch, FSIZE, STRIDE
eng = B200Engine(spec=parse_project(standin_1p4m_config()), workspace_gb=24)
batches = [synth_batch(i + 1, int(64e6)) for i in range(6)]
pinned = [torch.from_numpy(s).pin_memory() for s, _ in batches]
import jaeger_b200.engine as E
orig_h2d = E.B200Engine._h2d
log = []
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); dt = time.perf_counter() - t
        if dt > 0.02: log.append((name, round(dt * 1e3, 1)))
        return r
    return w
for name in ("_h2d", "pack", "encode", "forward", "aggregate", "_empty"):
    setattr(E.B200Engine, name, timed(name, getattr(E.B200Engine, name)))
with torch.cuda.stream(eng._stream()):
    for rep in range(3):
        for i in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter(); log.clear()
            x = pinned[i].to(eng.tdev, non_blocking=True)
            t1 = time.perf_counter()
            agg, w, c = eng.classify_long(x, batches[i][1], FSIZE, STRIDE)
            t2 = time.perf_counter()
            host = {k: agg[k].cpu() for k in ("pred_sum", "pred_var", "consensus", "per_class_counts", "entropy", "energy", "rel_pos")}
            torch.cuda.synchronize(); t3 = time.perf_counter()
            print(f"step {rep}.{i}: total {1e3*(t3-t0):.0f} ms  h2d-call {1e3*(t1-t0):.1f}  classify-call {1e3*(t2-t1):.1f}  d2h+sync {1e3*(t3-t2):.0f}  slow calls {log}", flush=True)
