"""Per-contig label agreement of the device path with the CPU oracle AT SCALE (north star: >= 99.9 %).
  (A) the reference's real `default` weights (config 1 graph) on a synthetic FASTA of N_A contigs: device vs oracle/legacy.py;
  (B) the stand-in 1.4 M architecture with its classifier kernel scaled so that |logit| ~ 5-10, N_B contigs: device vs
      oracle/forward.py (fp32, all host threads).
Writes profiles/label_agreement_r2.json (bench.py copies it into its JSON line).  Run on the GPU box:
    python tools/label_agreement.py [N_A] [N_B]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from jaeger_b200 import B200Engine, WindowSource, init_random, parse_project, standin_1p4m_config
from jaeger_b200 import codon_tables as ct
from oracle import encode as oenc, forward as ofw, legacy as oleg, seqwin
from tests.helpers import agreement_contigs, agreement_stats
from tests.test_gpu_parity import _legacy_fixture

n_a = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_b = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
torch.set_num_threads(torch.get_num_threads())
out = {}

w, _ = _legacy_fixture()
recs = agreement_contigs(101, n_a)
eng = B200Engine(legacy_weights=w)
y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, outputs=("prediction",), lazy_meta=True))
wins = list(seqwin.fragment_windows(recs, 2000, 1500))
table = dict(zip(oenc.CODONS, ct.LEGACY_AA_ID))
t0 = time.time()
tok = np.stack([oenc.encode_window_legacy(x.seq, 2000, table) for x in wins]).astype(np.uint8)
ref = np.concatenate([oleg.forward(w, tok[b:b + 256])["output"] for b in range(0, len(tok), 256)])
st = agreement_stats(y["prediction"], ref, np.array([x.is_last for x in wins]))
st["oracle_seconds"] = round(time.time() - t0, 1)
st["model"] = "reference `default` weights (legacy graph, config 1), synthetic contigs 2-6.6 kbp, GC 0.25-0.75"
out["legacy_real_weights"] = st
print(json.dumps(st), flush=True)
eng.close()

if n_b <= 0:      # keep the stand-in result of an earlier run (its oracle takes minutes)
    old = ROOT / "profiles" / "label_agreement_r2.json"
    if old.exists():
        out["standin_scaled_classifier"] = json.loads(old.read_text()).get("standin_scaled_classifier")
    for d in ("profiles", "gpurun_out"):
        (ROOT / d / "label_agreement_r2.json").write_text(json.dumps(out, indent=1))
    sys.exit(0)
spec = parse_project(standin_1p4m_config())
wt = init_random(spec, 0)
scale = 25.0
wt["classifier"][0]["kernel"] = wt["classifier"][0]["kernel"] * scale
recs = agreement_contigs(202, n_b)
eng = B200Engine(spec=spec, weights=wt)
y = eng.predict(WindowSource(records=recs, fsize=2000, stride=1500, outputs=("prediction",), lazy_meta=True))
wins = list(seqwin.fragment_windows(recs, 2000, 1500))
t0 = time.time()
tok = oenc.encode_windows([x.seq for x in wins], 2000)
ref = np.concatenate([ofw.forward(spec, wt, tok[b:b + 96])["prediction"] for b in range(0, len(tok), 96)])
st = agreement_stats(y["prediction"], ref, np.array([x.is_last for x in wins]))
st["oracle_seconds"] = round(time.time() - t0, 1)
st["model"] = f"stand-in 1.4M architecture, random init seed 0, classifier kernel x {scale:g}"
out["standin_scaled_classifier"] = st
print(json.dumps(st), flush=True)
eng.close()
(ROOT / "profiles").mkdir(exist_ok=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
for d in ("profiles", "gpurun_out"):
    (ROOT / d / "label_agreement_r2.json").write_text(json.dumps(out, indent=1))
