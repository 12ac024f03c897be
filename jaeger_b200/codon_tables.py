"""Codon tables of the reference encoder as data (seqops/maps.py:3-68, 137-202, 408-473,
475-540; preprocess/v1/maps.py TRIMER_INT) and the 64-entry device LUT built from them.

The device packs bases as A=0 C=1 T=2 G=3 (complement = code ^ 2); LUT index =
b0*16 + b1*4 + b2.  LUT value = token the model consumes: id + 1 for the v2 encoder
(0 = unknown codon), the amino-acid id for the legacy encoder (0 = unknown).
"""
from __future__ import annotations

import numpy as np

_ORDER = "TCAG"
# standard-table order: second base outermost, then first base, then third (TTT TTC TTA TTG CTT ...)
CODONS = [b + a + c for a in _ORDER for b in _ORDER for c in _ORDER]
CODON_ID = list(range(64))
AA_ID = [1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 9, 9,
         9, 9, 10, 10, 0, 0, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 0, 18, 19, 19,
         19, 19, 6, 6, 19, 19, 20, 20, 20, 20]
MURPHY10_ID = [1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 3, 3, 3, 3,
               5, 5, 5, 5, 1, 1, 0, 0, 6, 6, 7, 7, 7, 7, 8, 8, 7, 7, 7, 7, 9, 9, 0, 1, 8, 8, 8, 8,
               3, 3, 8, 8, 10, 10, 10, 10]
PC5_ID = [1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 2, 2, 2, 4, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4,
          4, 4, 1, 1, 0, 0, 1, 1, 3, 3, 3, 3, 5, 5, 5, 5, 5, 5, 4, 4, 0, 1, 5, 5, 5, 5, 4, 4, 5, 5,
          4, 4, 4, 4]
# legacy `default` model: codon -> amino-acid id 1..21 (preprocess/v1/maps.py TRIMER_INT)
LEGACY_AA_ID = [1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8,
                9, 9, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18,
                11, 19, 20, 20, 20, 20, 6, 6, 20, 20, 21, 21, 21, 21]
TABLES = {"CODON_ID": CODON_ID, "AA_ID": AA_ID, "MURPHY10_ID": MURPHY10_ID, "PC5_ID": PC5_ID}

_CODE = {"A": 0, "C": 1, "T": 2, "G": 3}


def device_lut(ids, plus_one: bool = True) -> np.ndarray:
    lut = np.zeros(64, dtype=np.uint8)
    for codon, i in zip(CODONS, ids):
        idx = _CODE[codon[0]] * 16 + _CODE[codon[1]] * 4 + _CODE[codon[2]]
        lut[idx] = i + 1 if plus_one else i
    return lut
