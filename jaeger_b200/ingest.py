"""Chunked FASTA ingest: a file of any size through two fixed pinned host buffers.

The reference iterates the records of the input with pyfastx and holds every window's outputs in host memory
(seqops/io.py:97-104, nnlib/inference.py:341-373).  For the 100 Gbp configuration that cannot work: `FastaChunks` hands out
runs of WHOLE records of at most `chunk_bases` bases, parsed by the native reader (jg_fasta_open / jg_fasta_next) straight
into one of two pinned buffers, so that chunk k + 1 is parsed (on a helper thread; the C call releases the GIL) and copied to
the device while chunk k is being classified.  A byte range gives a rank its own slice of a plain-text file
(`rank_byte_range`): N ranks read N disjoint slices, nothing is parsed twice.
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

import numpy as np
import torch

from ._cabi import check, lib


def rank_byte_range(path: str | Path, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of rank's slice of an uncompressed file; a record belongs to the slice that holds its '>'."""
    size = os.path.getsize(path)
    return size * rank // world, (size * (rank + 1) // world if rank + 1 < world else -1)


def is_gzip(path: str | Path) -> bool:
    with open(path, "rb") as fh:
        return fh.read(2) == b"\x1f\x8b"


class FastaChunks:
    """Iterator over (names, bases [pinned uint8 tensor view], offsets [n + 1] int64) chunks of whole records."""

    def __init__(self, path: str | Path, chunk_bases: int = 256_000_000, byte_range: tuple[int, int] = (0, -1),
                 max_records: int = 4_000_000, prefetch: bool = True, pin: bool | None = None):
        self.path, self.chunk_bases, self.max_records = str(path), int(chunk_bases), int(max_records)
        self.pin = torch.cuda.is_available() if pin is None else pin
        h = ctypes.c_void_p()
        check(lib.jg_fasta_open(self.path.encode(), int(byte_range[0]), int(byte_range[1]), ctypes.byref(h)))
        self.reader = h
        self.bufs = [self._alloc(self.chunk_bases), self._alloc(self.chunk_bases)]
        self.name_cap = 64 * 1024 * 1024
        self.names = [ctypes.create_string_buffer(self.name_cap), ctypes.create_string_buffer(self.name_cap)]
        self.prefetch = prefetch
        self._k = 0
        self._next = None            # (thread, result holder) of the chunk being parsed ahead
        self.total_records = self.total_bases = 0

    def _alloc(self, n: int) -> torch.Tensor:
        t = torch.empty(max(int(n), 1), dtype=torch.uint8)
        return t.pin_memory() if self.pin else t

    def _read(self, slot: int):
        """Parse the next chunk into buffer `slot` (grown when one record is longer than the buffer)."""
        n, nb, nn, need = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        offsets = np.zeros(self.max_records + 1, dtype=np.int64)
        while True:
            buf = self.bufs[slot]
            rc = lib.jg_fasta_next(self.reader, self.chunk_bases, buf.numel(), self.max_records, self.name_cap,
                                   ctypes.c_void_p(buf.data_ptr()), offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                   self.names[slot], ctypes.byref(n), ctypes.byref(nb), ctypes.byref(nn), ctypes.byref(need))
            if rc == 3 and n.value == 0:          # one record longer than the buffer: grow it and read the record again
                self.bufs[slot] = self._alloc(need.value)
                continue
            if rc not in (0, 3):
                check(rc)
            break
        if n.value == 0:
            return None
        names = [x.decode() for x in self.names[slot].raw[:nn.value].split(b"\0")[:n.value]]
        return names, self.bufs[slot][:nb.value], offsets[:n.value + 1].copy()

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is not None:
            th, holder = self._next
            th.join()
            self._next = None
            if "error" in holder:
                raise holder["error"]
            cur = holder["chunk"]
        else:
            cur = self._read(self._k & 1)
        if cur is None:
            self.close()
            raise StopIteration
        self._k += 1
        if self.prefetch:                        # parse the following chunk into the other buffer meanwhile
            holder: dict = {}
            slot = self._k & 1

            def work():
                try:
                    holder["chunk"] = self._read(slot)
                except Exception as e:           # surfaced on the consumer's thread
                    holder["error"] = e
            th = threading.Thread(target=work, daemon=True)
            th.start()
            self._next = (th, holder)
        self.total_records += len(cur[0])
        self.total_bases += int(cur[2][-1])
        return cur

    def close(self):
        if self._next is not None:
            self._next[0].join()
            self._next = None
        if getattr(self, "reader", None):
            lib.jg_fasta_close(self.reader)
            self.reader = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
