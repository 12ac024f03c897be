"""First end-to-end run of the device pipeline against the oracle (development aid)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from jaeger_b200 import B200Engine, WindowSource, parse_project, standin_1p4m_config, init_random
from oracle import seqwin, encode, forward as ofw
from tests.helpers import random_contigs

spec = parse_project(standin_1p4m_config())
w = init_random(spec, 0)
recs = random_contigs(1, [2000, 3500, 5000, 9000, 2100, 12000, 800, 1500])
src = WindowSource(records=recs, fsize=2000, stride=1500, min_len=500)
eng = B200Engine(spec=spec, weights=w)
t0 = time.time(); y = eng.predict(src); print("predict s", time.time() - t0)
wins = list(seqwin.fragment_windows(recs, 2000, 1500, min_len=2000)) + \
       list(seqwin.fragment_windows(recs, 2000, 1500, min_len=500, max_len=1999))
print("windows", len(wins), y["prediction"].shape)
ok = True
for i, wd in enumerate(wins):
    got = (y["meta_0"][i].decode(), int(y["meta_1"][i]), int(y["meta_2"][i]), int(y["meta_3"][i]), int(y["meta_4"][i]),
           int(y["meta_5"][i]), int(y["meta_6"][i]), int(y["meta_7"][i]), int(y["meta_8"][i]), y["meta_9"][i].decode())
    exp = (wd.header, wd.index, wd.is_last, wd.ordinal, wd.seqlen, wd.g, wd.c, wd.a, wd.t, wd.gc_skew)
    if got != exp:
        ok = False; print("META MISMATCH", i, got, exp)
print("meta ok", ok)
long_w = [x for x in wins if len(x.seq) == 2000]
tok = encode.encode_windows([x.seq for x in long_w], 2000)
ref = ofw.forward(spec, w, tok)
n = len(long_w)
for k in ref:
    d = np.abs(ref[k].reshape(n, -1) - y[k][:n].reshape(n, -1))
    print(k, "max|ref|", np.abs(ref[k]).max(), "max|diff|", d.max(), "mean|diff|", d.mean())
print("argmax agree", (ref["prediction"].argmax(1) == y["prediction"][:n].argmax(1)).mean())
eng2 = B200Engine(spec=spec, weights=w, use_ref_kernels=True)
y2 = eng2.predict(src)
for k in ref:
    d = np.abs(y2[k].reshape(len(wins), -1) - y[k].reshape(len(wins), -1))
    print("tc vs cuda-core", k, d.max())
# short pass vs oracle (padded batch semantics)
short_w = [x for x in wins if len(x.seq) < 2000]
if short_w:
    tok_s = encode.encode_windows([x.seq for x in short_w], 2000)
    ref_s = ofw.forward(spec, w, tok_s)
    for k in ref_s:
        d = np.abs(ref_s[k].reshape(len(short_w), -1) - y[k][n:].reshape(len(short_w), -1))
        print("short", k, "max|diff|", d.max())
