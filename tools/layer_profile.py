"""Per-conv-launch CUDA-event times of one forward chunk (development aid)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from jaeger_b200 import B200Engine, parse_project, standin_1p4m_config
from bench import synth_batch
FSIZE, STRIDE = 2000, 1500
cfg = standin_1p4m_config()
if os.environ.get("JG_DYT"):            # the MaskedDYT variant of the architecture
    sys.path.insert(0, "tests")
    from helpers import to_dyt
    cfg = to_dyt(cfg)
eng = B200Engine(spec=parse_project(cfg), workspace_gb=24)
seq, lens = synth_batch(1, int(16e6))
with torch.cuda.stream(eng._stream()):
    x = torch.from_numpy(seq).to(eng.tdev)
    for _ in range(3):
        eng.classify_long(x, lens, FSIZE, STRIDE)
    eng.set_profiling(True)
    for _ in range(3):
        agg, w, c = eng.classify_long(x, lens, FSIZE, STRIDE)
    prof = eng.get_profile()
lc = 665
for i, (c, (ms, n, win)) in enumerate(zip(eng.plan.launches, prof)):
    k, cin, cout = c.kernel.shape
    fl = 2.0 * 6 * (lc - c.cum_shrink_in - c.shrink) * k * cin * cout * win
    kind = "final" if c.scale2 is not None else ("conv2" if c.sc_buf >= 0 else ("stem" if i == 0 else "conv1"))
    print(f"layer {i:2d} {kind:6s} {ms/n:7.3f} ms/launch  {fl/ms/1e9:7.1f} TFLOP/s  windows/launch {win/n:.0f}")
print("impl", os.environ.get("JG_CONV_IMPL", "auto"))
