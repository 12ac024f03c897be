"""Contig sharding over the GPUs of one node and the final gather of per-contig records.

The reference has no multi-process inference path (commands/predict.py:602 exposes one GPU).
Contigs are independent units of the hot path: every rank classifies its own
shard with no collective on the data path, and one gather of fixed-width per-contig records at
the end rebuilds the input order on rank 0.  Works over NCCL (GPU tensors) and gloo (CPU tensors).
"""
from __future__ import annotations

import heapq

import numpy as np
import torch
import torch.distributed as dist


def window_counts(lens: np.ndarray, fsize: int, stride: int) -> np.ndarray:
    """Windows per contig for the fixed-stride long pass: len(range(0, L - (fsize-1), stride))."""
    lens = np.asarray(lens, dtype=np.int64)
    n = (lens - fsize) // stride + 1
    return np.where(lens >= fsize, n, 0)


def shard_contigs(lens: np.ndarray, world: int, fsize: int, stride: int) -> list[np.ndarray]:
    """Longest-processing-time greedy bin packing on window counts (the FLOP proxy).
    Returns, per rank, the sorted indices of its contigs; contigs without windows go to the
    lightest rank so every contig has an owner."""
    w = window_counts(lens, fsize, stride)
    order = np.argsort(-w, kind="stable")
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    owner = np.empty(len(w), dtype=np.int64)
    for c in order:
        load, r = heapq.heappop(heap)
        owner[c] = r
        heapq.heappush(heap, (load + int(w[c]), r))
    return [np.flatnonzero(owner == r) for r in range(world)]


def gather_contig_records(records: torch.Tensor, shards: list[np.ndarray], n_total: int, dst: int = 0,
                          group=None) -> torch.Tensor | None:
    """records [len(shards[rank]), width] (any float/int dtype): this rank's per-contig records, row k belonging to
    global contig shards[rank][k].  `shards` is the partition every rank computed for itself with `shard_contigs`
    (it is deterministic), so all sizes and destinations are known up front: ONE pre-sized gather, no size
    exchange, no device-to-host read, nothing that blocks the host.  Returns on `dst` the [n_total, width] table
    in global contig order (rows nobody owns stay zero), None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(shards) != world or records.shape[0] != len(shards[rank]):
        raise ValueError(f"rank {rank}: {records.shape[0]} records for a shard of {len(shards[rank])} contigs ({len(shards)} shards, world {world})")
    dev = records.device
    m = max(len(s) for s in shards)
    width = records.shape[1]
    pad = torch.zeros((m, width), dtype=records.dtype, device=dev)
    pad[:records.shape[0]] = records
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst, group=group)
    if rank != dst:
        return None
    table = torch.zeros((n_total, width), dtype=records.dtype, device=dev)
    for r_, ids in zip(out, shards):
        if len(ids):
            table[torch.as_tensor(np.asarray(ids), dtype=torch.int64, device=dev)] = r_[:len(ids)]
    return table


# ---- driver-level sharding (python -m torch.distributed.run ... -m jaeger_b200.predict) ----------
def dist_env() -> tuple[int, int, int]:
    """(world, rank, local_rank) from the torchrun environment; (1, 0, 0) when not launched by it."""
    import os
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_loaded(loaded, mine: np.ndarray):
    """Sub-select records `mine` (sorted global indices) of a loaded FASTA (names, bases, offsets):
    the shard's bases are copied back to back into a fresh (pinned when CUDA is present) buffer."""
    names, host, offsets = loaded
    lens = np.diff(offsets)[mine]
    sub_off = np.zeros(len(mine) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    buf = torch.empty(max(int(sub_off[-1]), 1), dtype=torch.uint8)
    if torch.cuda.is_available():
        buf = buf.pin_memory()
    dst, srcv = buf.numpy(), host.numpy()
    for k, c in enumerate(mine):
        dst[sub_off[k]:sub_off[k + 1]] = srcv[offsets[c]:offsets[c + 1]]
    return [names[c] for c in mine], buf[:int(sub_off[-1])], sub_off


def normalise_joined_columns(df):
    """Per-rank / per-chunk summary frames are concatenated.  The left-joined terminal-repeat columns (collect.py:527-532)
    must come out as ONE whole-file join gives them: `repeat_length` float64 with NaN where no repeat was found (the TSV then
    shows "15.000", as the reference's does) -- also when a part had a repeat on every contig (an int column) or on none."""
    if "repeat_length" in df.columns:
        import pandas as pd
        df["repeat_length"] = pd.to_numeric(df["repeat_length"], errors="coerce").astype("float64")
    return df


def exchange_errors(err: BaseException | None, world: int, rank: int) -> None:
    """Called by every rank right before the result gather: a rank that failed in its own work (engine.predict,
    contig_table ...) tells the others instead of leaving them blocked in the collective.  Re-raises the local error
    on the rank that failed and a RuntimeError naming the failed ranks everywhere else."""
    if world <= 1:
        if err is not None:
            raise err
        return
    import torch.distributed as dist
    msgs: list = [None] * world
    dist.all_gather_object(msgs, None if err is None else f"{type(err).__name__}: {err}")
    if err is not None:
        raise err
    bad = [f"rank {r}: {m}" for r, m in enumerate(msgs) if m]
    if bad:
        raise RuntimeError("another rank failed before the gather -- " + "; ".join(bad))


def merge_rank_frames(frames, keep_order_columns: bool = False):
    """Per-rank summary tables (each with helper columns `_pass`, `_gid` = pass of the contig's windows
    and its index in the FASTA) -> one table in the single-process row order: long-pass contigs in
    FASTA order, then short-pass contigs in FASTA order."""
    import pandas as pd
    frames = [f for f in frames if f is not None and len(f)]
    if not frames:
        return None                      # no rank produced a row: the callers write no table, like the single-process run
    df = normalise_joined_columns(pd.concat(frames, ignore_index=True))
    df = df.sort_values(["_pass", "_gid"], kind="stable").reset_index(drop=True)
    return df if keep_order_columns else df.drop(columns=["_pass", "_gid"])
