"""`fragment_generator(dustmask=True)` of the REFERENCE (seqops/io.py:74-147) with pydustmasker (>= 1.0.3, not installable
here) replaced by a stub whose `DustMasker(seq, window_size, score_threshold).mask()` returns the oracle's SDUST restatement
(oracle/dust.py).  Pins what the reference does around the masker: masking runs on the stripped, upper-cased record, windows
are cut from the soft-masked string, the A / C / G / T counts are case-sensitive (masked bases do not count, so N% and G+C
move), gc_skew formatting, two-pass short contigs.  SDUST itself stays unpinned.
Writes tests/golden/fragments_dustmask.json (window rows with the sequence as crc32 + length, like fragments_synthetic.json).

usage:  python tests/golden/make_dust_flow_goldens.py
"""
from __future__ import annotations

import json
import sys
import tempfile
import types
import zlib
from pathlib import Path

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(OUT.parent.parent))

from oracle import dust as odust          # noqa: E402
from tests.helpers import dust_contigs   # noqa: E402


class _Fasta:
    def __init__(self, path, build_index=False):
        self.path = path

    def __iter__(self):
        name, chunks = None, []
        with open(self.path) as fh:
            for line in fh:
                if line.startswith(">"):
                    if name is not None:
                        yield name, "".join(chunks)
                    name, chunks = line[1:].split()[0], []
                else:
                    chunks.append(line.strip())
        if name is not None:
            yield name, "".join(chunks)


class _DustMasker:
    def __init__(self, sequence, window_size=64, score_threshold=20):
        self.seq, self.w, self.t = sequence, window_size, score_threshold

    def mask(self):
        return odust.mask(self.seq, self.w, self.t)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


_stub("pyfastx", Fasta=_Fasta)
_stub("pydustmasker", DustMasker=_DustMasker)


def main():
    from jaeger.seqops import io as rio
    recs = dust_contigs()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fa = Path(tmp) / "d.fasta"
        fa.write_text("".join(f">{n} desc\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n" for n, s in recs))
        for key, kw in {"2000_1500": dict(fragsize=2000, stride=1500, min_len=None, max_len=None),
                        "2000_1500_short": dict(fragsize=2000, stride=1500, min_len=500, max_len=1999),
                        "500_500": dict(fragsize=500, stride=500, min_len=None, max_len=None)}.items():
            rows = []
            for s in rio.fragment_generator(str(fa), dustmask=True, **kw):
                f = s.split(",")
                rows.append([zlib.crc32(f[0].encode()), len(f[0])] + f[1:])
            out[key] = rows
            print(key, len(rows), "windows;", sum(1 for r in rows if int(r[7]) + int(r[8]) + int(r[9]) + int(r[10]) < r[1]), "with masked / unknown bases")
    (OUT / "fragments_dustmask.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
