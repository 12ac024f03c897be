"""Six-frame tokens from the reference's numba encoder (jaeger.dataops.convert._process_batch_numba) at the crops where
it agrees with the predict path's TF encoder by construction (nt = 3 * codons + 5, seqops/crop.py:26-37): 2048 -> 681 and
500 -> 165 codons per frame (2000 -> 665 is tests/golden/tokens_2000.npz from make_goldens.py).  Full-length windows only:
for shorter sequences the two reference encoders use different arithmetic.  Writes tests/golden/tokens_more_crops.npz.

usage:  python tests/golden/make_token_goldens_more_crops.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference/src")


def main():
    from jaeger.dataops import convert as rconv
    codon_lut, ascii_lut, comp_lut = rconv._build_numba_lookups()
    rng = np.random.default_rng(2048)
    out = {}
    for crop, lc, n in ((2048, 681, 10), (500, 165, 24)):
        arr = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, crop)).copy()
        arr[1, 100:140] = ord("N")                    # unknown run
        arr[2, [7, 250, 499]] = [ord("R"), ord("Y"), ord("K")]
        arr[3, -9:] = ord("N")                        # unknown codons at the window end (start of the reverse frames)
        arr[4, :4] = ord("N")
        tok = rconv._process_batch_numba(arr, np.full(n, crop, np.int64), crop, lc, codon_lut, comp_lut, ascii_lut)
        out[f"seqs_{crop}"] = np.array([a.tobytes().decode() for a in arr])
        out[f"tokens_{crop}"] = tok.astype(np.uint8)
        print(crop, tok.shape, int((tok == 0).sum()), "unknown tokens")
    np.savez_compressed(OUT / "tokens_more_crops.npz", **out)


if __name__ == "__main__":
    main()
