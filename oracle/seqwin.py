"""Oracle (test infrastructure): FASTA records -> fragment windows + per-window metadata.

Restates, without TensorFlow / pyfastx:
  * `_window_indices`                       seqops/io.py:38-71
  * `fragment_generator`                    seqops/io.py:74-147
  * `safe_divide`, `signal_l`               utils/misc.py:117-147
  * `validate_fasta_entries`                utils/fs.py:99-115
The low-complexity soft-mask (pydustmasker, seqops/io.py:105-108) is a "next" row: this
oracle takes an optional per-base boolean soft-mask instead of computing one.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


def read_fasta(path: str):
    """(name, sequence) pairs with pyfastx semantics: name = header up to the first
    whitespace, sequence = all lines concatenated without whitespace (call sites:
    seqops/io.py:98-104, utils/fs.py:106)."""
    name, chunks = None, []
    with open(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                fields = line[1:].split()
                name = fields[0] if fields else ""
                chunks = []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        yield name, "".join(chunks)


def validate_fasta_entries(path: str, min_len: int = 2048) -> int:
    """utils/fs.py:99-115: number of records; raises when none reaches min_len."""
    num = gt = 0
    for _, seq in read_fasta(path):
        num += 1
        gt += len(seq) >= min_len
    if gt == 0:
        raise ValueError(f"all records in {path} are < {min_len}bp")
    return num


def window_indices(seqlen: int, fragsize: int, stride: int | None, dynamic_stride: bool = False,
                   dynamic_stride_threshold: float = 10.0) -> list[int]:
    """seqops/io.py:38-71."""
    if not dynamic_stride or seqlen >= dynamic_stride_threshold * fragsize:
        step = fragsize if stride is None else stride
        return list(range(0, seqlen - (fragsize - 1), step))
    n_windows = max(1, math.ceil(seqlen / fragsize))
    if n_windows == 1:
        return [0]
    raw_stride = (seqlen - fragsize) / (n_windows - 1)
    indices = [int(round(i * raw_stride)) for i in range(n_windows)]  # banker's rounding
    indices[-1] = seqlen - fragsize
    seen, unique = set(), []
    for idx in indices:
        if idx not in seen:
            seen.add(idx)
            unique.append(idx)
    return unique


def safe_divide(numerator, denominator):
    """utils/misc.py:117-123."""
    try:
        return round(numerator / denominator, 2)
    except ZeroDivisionError:
        return 0


@dataclass
class Window:
    seq: str          # window bases (upper-case; soft-masked bases lower-case)
    header: str       # meta_0
    index: int        # meta_1 window start
    is_last: int      # meta_2
    ordinal: int      # meta_3
    seqlen: int       # meta_4
    g: int            # meta_5
    c: int            # meta_6
    a: int            # meta_7
    t: int            # meta_8
    gc_skew: str      # meta_9, formatted "{: .3f}"

    def csv(self) -> str:
        return (f"{self.seq},{self.header},{self.index},{self.is_last},{self.ordinal},"
                f"{self.seqlen},{self.g},{self.c},{self.a},{self.t},{self.gc_skew}")


def apply_softmask(seq: str, softmask) -> str:
    if softmask is None:
        return seq
    return "".join(ch.lower() if m else ch for ch, m in zip(seq, softmask))


def fragment_windows(records, fragsize: int, stride: int | None, *, softmasks=None,
                     dynamic_stride: bool = False, dynamic_stride_threshold: float = 10.0,
                     min_len: int | None = None, max_len: int | None = None):
    """seqops/io.py:74-147 as a generator of Window objects.

    records: iterable of (name, sequence); softmasks: optional dict name -> bool sequence
    (stands in for the DustMasker call at io.py:105-108)."""
    if min_len is None:
        min_len = fragsize
    for name, raw in records:
        seqlen = len(raw)
        sequence = raw.strip().upper()
        if softmasks is not None and name in softmasks:
            sequence = apply_softmask(sequence, softmasks[name])
        header = name.strip().replace(",", "___")
        if max_len is not None and seqlen > max_len:
            continue
        if seqlen >= fragsize:
            indices = window_indices(seqlen, fragsize, stride, dynamic_stride,
                                     dynamic_stride_threshold)
            for i, index in enumerate(indices):
                w = sequence[index:index + fragsize]
                g, c, a, t = w.count("G"), w.count("C"), w.count("A"), w.count("T")
                skew = safe_divide(g - c, g + c)
                yield Window(w, header, index, int(i == len(indices) - 1), i, seqlen, g, c, a, t,
                             f"{skew: .3f}")
        elif seqlen >= min_len:
            g, c, a, t = (sequence.count(x) for x in "GCAT")
            skew = safe_divide(g - c, g + c)
            yield Window(sequence, header, 0, 1, 0, seqlen, g, c, a, t, f"{skew: .3f}")
