"""Oracle (test infrastructure): window string -> six-frame codon tokens.

Restates the TensorFlow string-op encoders with plain Python/NumPy:
  * `process_string_inference`   seqops/encode.py:203-318  (tables :20-41, seqops/maps.py)
  * legacy `process_string`      preprocess/v1/convert.py:56-125 (tables preprocess/v1/maps.py)
Pinned against the reference's own TF-free numba encoder
(`jaeger.dataops.convert._process_batch_numba`, dataops/convert.py:664-743) through
tests/golden/make_goldens.py, and against tests/unit/test_crop.py / test_inference_crop.py
frame-length known answers (2000 -> 665, 1500 -> 498, 500 -> 165).
"""
from __future__ import annotations

import numpy as np

# seqops/maps.py:3-68
_BASES = "TCAG"
# second base outermost, then first base, then third (TTT TTC TTA TTG CTT ...)
CODONS = [b + a + c for a in _BASES for b in _BASES for c in _BASES]
CODON_ID = list(range(64))
# seqops/maps.py:137-202 (AA_ID), :408-473 (MURPHY10_ID), :475-540 (PC5_ID); values copied as data
AA_ID = [1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 9, 9,
         9, 9, 10, 10, 0, 0, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 0, 18, 19, 19,
         19, 19, 6, 6, 19, 19, 20, 20, 20, 20]
MURPHY10_ID = [1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 3, 3, 3, 3,
               5, 5, 5, 5, 1, 1, 0, 0, 6, 6, 7, 7, 7, 7, 8, 8, 7, 7, 7, 7, 9, 9, 0, 1, 8, 8, 8, 8,
               3, 3, 8, 8, 10, 10, 10, 10]
PC5_ID = [1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 2, 2, 2, 4, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4,
          4, 4, 1, 1, 0, 0, 1, 1, 3, 3, 3, 3, 5, 5, 5, 5, 5, 5, 4, 4, 0, 1, 5, 5, 5, 5, 4, 4, 5, 5,
          4, 4, 4, 4]
CODON_MAPS = {"CODON_ID": CODON_ID, "AA_ID": AA_ID, "MURPHY10_ID": MURPHY10_ID, "PC5_ID": PC5_ID}

# encode.py:27-33: complement keeps case, anything else -> "N"
_COMPLEMENT = {"A": "T", "T": "A", "G": "C", "C": "G", "a": "t", "t": "a", "g": "c", "c": "g"}


def frame_offset(crop_size: int) -> int:
    """encode.py:232-236: offset_lut[crop_size % 3] with offset_lut = (-2, -1, 0)."""
    return (-2, -1, 0)[crop_size % 3]


def codons_per_frame(n_bases: int, crop_size: int) -> int:
    """Length of every frame produced for a window of n_bases (<= crop_size) bases."""
    off = frame_offset(crop_size)
    n_tri = max(0, n_bases - 2)
    return len(range(0, n_tri)[0:-3 + off:3])


def encode_window(seq: str, crop_size: int, codon_id=CODON_ID, masking: bool = False,
                  codons=CODONS) -> np.ndarray:
    """encode.py:229-302 for one window: int32 [6, Lc] of codon ids, -1 for unknown codons.
    The model input is  id + 1  (token, 0 = unknown) or one_hot(id) (all-zero row = unknown)."""
    table = {c: i for c, i in zip(codons, codon_id)}
    fwd = list(seq[:crop_size])
    rev = [_COMPLEMENT.get(b, "N") for b in fwd[::-1]]
    if not masking:
        fwd = [b.upper() for b in fwd]
        rev = [b.upper() for b in rev]
    off = frame_offset(crop_size)
    tri_f = ["".join(fwd[i:i + 3]) for i in range(len(fwd) - 2)]
    tri_r = ["".join(rev[i:i + 3]) for i in range(len(rev) - 2)]
    frames = [tri_f[0:-3 + off:3], tri_f[1:-2 + off:3], tri_f[2:-1 + off:3],
              tri_r[0:-3 + off:3], tri_r[1:-2 + off:3], tri_r[2:-1 + off:3]]
    lc = min(len(f) for f in frames)
    assert all(len(f) == lc for f in frames), [len(f) for f in frames]
    return np.array([[table.get(t, -1) for t in f] for f in frames], dtype=np.int32).reshape(6, lc)


def tokens_from_ids(ids: np.ndarray) -> np.ndarray:
    """encode.py:302: float(id + 1); returned as uint8 (0 = unknown / masked)."""
    return (ids + 1).astype(np.uint8)


def encode_windows(seqs, crop_size: int, codon_id=CODON_ID, masking: bool = False) -> np.ndarray:
    """Batch of windows -> uint8 tokens [n, 6, Lc_max], right-padded with 0 exactly like
    `padded_batch(padding_values=0.0)` (commands/predict.py:159-183)."""
    enc = [tokens_from_ids(encode_window(s, crop_size, codon_id, masking)) for s in seqs]
    lc = max((e.shape[1] for e in enc), default=0)
    out = np.zeros((len(enc), 6, lc), dtype=np.uint8)
    for i, e in enumerate(enc):
        out[i, :, :e.shape[1]] = e
    return out


# ---- legacy (model "default") encoder ---------------------------------------------------
# preprocess/v1/maps.py: TRIMERS x TRIMER_INT map every codon to an amino-acid id 1..21
# (stop codons = 11), unknown -> 0 (hash-table default, convert.py:21).  The standard genetic
# code in the reference's amino-acid numbering, indexed like CODONS above:
LEGACY_AA = "FFLLLLLLIIIMVVVVSSSSPPPPTTTTAAAAYY**HHQQNNKKDDEECC*WRRRRSSRRGGGG"


def legacy_table(trimers, trimer_int):
    return {t: v for t, v in zip(trimers, trimer_int)}


def encode_window_legacy(seq: str, crop_size: int, table: dict) -> np.ndarray:
    """preprocess/v1/convert.py:75-99: no crop, NO upper-casing (soft-masked bases fall out
    of the table -> 0), default 0.  Returns int32 [6, Lc]."""
    off = frame_offset(crop_size)
    fwd = list(seq)
    rev = [_COMPLEMENT.get(b, "N") for b in fwd[::-1]]
    tri_f = ["".join(fwd[i:i + 3]) for i in range(len(fwd) - 2)]
    tri_r = ["".join(rev[i:i + 3]) for i in range(len(rev) - 2)]
    frames = [tri_f[0:-3 + off:3], tri_f[1:-2 + off:3], tri_f[2:-1 + off:3],
              tri_r[0:-3 + off:3], tri_r[1:-2 + off:3], tri_r[2:-1 + off:3]]
    lc = min(len(f) for f in frames)
    return np.array([[table.get(t, 0) for t in f[:lc]] for f in frames], dtype=np.int32)
