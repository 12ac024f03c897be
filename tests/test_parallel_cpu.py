"""CPU tests of the multi-GPU host logic with the gloo backend (world_size 2)."""
import os
from pathlib import Path
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaeger_b200.parallel import gather_contig_records, shard_contigs, window_counts


def test_window_counts_match_planner_formula():
    lens = np.array([1999, 2000, 3499, 3500, 50000, 0, 2001])
    assert window_counts(lens, 2000, 1500).tolist() == [len(range(0, int(L) - 1999, 1500)) for L in lens]


def test_shard_contigs_balances_and_partitions():
    rng = np.random.default_rng(0)
    lens = rng.integers(2000, 50001, size=20000)
    lens[:5] = 5_000_000                                     # a few genomes
    for world in (1, 2, 4, 8):
        shards = shard_contigs(lens, world, 2000, 1500)
        allc = np.sort(np.concatenate(shards))
        assert np.array_equal(allc, np.arange(len(lens)))    # a partition
        loads = np.array([window_counts(lens[s], 2000, 1500).sum() for s in shards])
        assert loads.max() / loads.mean() < 1.01             # balanced to < 1 %


def _worker(rank, world, port, n_total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    lens = rng.integers(2000, 50001, size=n_total)
    shards = shard_contigs(lens, world, 2000, 1500)
    mine = shards[rank]
    # a stand-in for the per-contig records every rank computes on its own GPU
    rec = torch.tensor(np.stack([mine * 3.0 + 1.0, lens[mine].astype(np.float64)], axis=1))
    table = gather_contig_records(rec, shards, n_total, dst=0)
    if rank == 0:
        want = np.stack([np.arange(n_total) * 3.0 + 1.0, lens.astype(np.float64)], axis=1)
        assert np.array_equal(table.numpy(), want)
    else:
        assert table is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_contig_records_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 501), nprocs=2, join=True)


def test_native_fasta_loader_matches_line_reader(tmp_path):
    """jg_fasta_scan / jg_fasta_load (host-only entry points): names, bases and offsets equal the
    record iteration the reference gets from pyfastx (seqops/io.py:98-104)."""
    from jaeger_b200.engine import load_fasta, read_fasta
    p = tmp_path / "x.fa"
    p.write_bytes(b">a desc here\nACGT\nNNAC\r\n\n>b\tx\nGG\n>c\n>d\nTTTT")
    for path in (p, Path(__file__).parent / "golden" / "synthetic_contigs.fasta"):
        names, host, off = load_fasta(path)
        ref = list(read_fasta(path))
        assert names == [r[0] for r in ref]
        assert bytes(host.numpy()) == b"".join(r[1] for r in ref)
        assert np.diff(off).tolist() == [len(r[1]) for r in ref]
    import gzip
    gz = tmp_path / "x.fa.gz"
    gz.write_bytes(gzip.compress(p.read_bytes()))
    names, host, off = load_fasta(gz)
    assert names == ["a", "b", "c", "d"] and bytes(host.numpy()) == b"ACGTNNACGGTTTT" and off.tolist() == [0, 8, 10, 10, 14]
    empty = tmp_path / "e.fa"
    empty.write_bytes(b"")
    names, host, off = load_fasta(empty)
    assert names == [] and host.numel() == 0 and off.tolist() == [0]


def test_driver_sharding_helpers_round_trip():
    """shard_loaded gives each rank exactly its records; merge_rank_frames restores the single-process
    row order (long-pass contigs in FASTA order, then short-pass contigs)."""
    import pandas as pd
    from jaeger_b200.engine import WindowSource
    from jaeger_b200.parallel import merge_rank_frames, shard_contigs, shard_loaded
    rng = np.random.default_rng(3)
    lens = [2500, 700, 9000, 2000, 1200, 30000, 4100, 650]
    recs = [(f"c{i}", "".join(rng.choice(list("ACGT"), n))) for i, n in enumerate(lens)]
    loaded = WindowSource(records=recs).load()
    eff = np.where(np.array(lens) < 2000, 2000, np.array(lens))
    shards = shard_contigs(eff, 3, 2000, 1500)
    assert sorted(np.concatenate(shards).tolist()) == list(range(len(lens)))
    frames = []
    for mine in shards:
        names, buf, off = shard_loaded(loaded, mine)
        assert names == [recs[c][0] for c in mine]
        for k, c in enumerate(mine):
            assert bytes(buf.numpy()[off[k]:off[k + 1]]) == recs[c][1].encode()
        # a rank's table: its long-pass contigs first, then its short ones (the engine's window order)
        order = [c for c in mine if lens[c] >= 2000] + [c for c in mine if lens[c] < 2000]
        frames.append(pd.DataFrame({"contig_id": [f"c{c}" for c in order], "_pass": [int(lens[c] < 2000) for c in order],
                                    "_gid": order}))
    merged = merge_rank_frames(frames + [None])
    want = [f"c{c}" for c in range(len(lens)) if lens[c] >= 2000] + [f"c{c}" for c in range(len(lens)) if lens[c] < 2000]
    assert merged["contig_id"].tolist() == want and list(merged.columns) == ["contig_id"]


def test_chunked_fasta_reader_equals_whole_file_loader_and_partitions_by_byte_range(tmp_path):
    """jg_fasta_open / jg_fasta_next (ingest.FastaChunks): chunks of whole records -- any chunk size, records longer than a chunk,
    empty records, CRLF and blank lines, gzip -- concatenate to exactly what the one-pass loader returns, and the byte ranges of
    `rank_byte_range` give every record to exactly one rank, in file order (a record belongs to the slice holding its '>')."""
    import gzip
    from jaeger_b200.engine import load_fasta
    from jaeger_b200.ingest import FastaChunks, rank_byte_range
    rng = np.random.default_rng(0)
    p = tmp_path / "x.fa"
    with open(p, "w", newline="") as fh:
        for i in range(400):
            n = int(rng.integers(0, 4000)) if i % 50 else 60_000
            s = "".join(rng.choice(list("ACGTNacgt"), n))
            fh.write(f">rec{i} description {i}\n")
            for k in range(0, n, 70):
                fh.write(s[k:k + 70] + ("\r\n" if i % 7 == 0 else "\n"))
            if i % 11 == 0:
                fh.write("\n")
    names, host, off = load_fasta(p)
    whole = bytes(host.numpy())
    for chunk in (1000, 25_000, 10 ** 9):
        got_names, got, lens = [], b"", []
        for nm, b, o in FastaChunks(p, chunk_bases=chunk, pin=False):
            assert int(o[-1]) <= max(chunk, 60_000) and int(o[-1]) == b.numel()
            got_names += nm; got += bytes(b.numpy()); lens += np.diff(o).tolist()
        assert got_names == names and got == whole and lens == np.diff(off).tolist(), chunk
    for world in (2, 3, 8):
        seen = []
        for r in range(world):
            for nm, _, _ in FastaChunks(p, chunk_bases=50_000, byte_range=rank_byte_range(p, r, world), pin=False, prefetch=False):
                seen += nm
        assert seen == names, world
    gz = tmp_path / "x.fa.gz"
    gz.write_bytes(gzip.compress(p.read_bytes()))
    assert [n for nm, _, _ in FastaChunks(gz, chunk_bases=30_000, pin=False) for n in nm] == names
    import pytest
    from jaeger_b200._cabi import JaegerB200Error
    with pytest.raises(JaegerB200Error, match="uncompressed"):
        FastaChunks(gz, byte_range=(100, -1), pin=False)


def _error_exchange_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    from jaeger_b200.parallel import exchange_errors
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        exchange_errors(None, world, rank)                                   # nobody failed: returns on every rank
        try:
            exchange_errors(ValueError("bad shard") if rank == 1 else None, world, rank)
            q.put((rank, "no error"))
        except ValueError as e:
            q.put((rank, f"own:{e}"))
        except RuntimeError as e:
            q.put((rank, f"peer:{e}"))
    finally:
        dist.destroy_process_group()


def test_rank_failure_is_exchanged_before_the_gather():
    """A rank that fails in its own work must not leave the others blocked in the result gather (world size 2, gloo)."""
    import torch.multiprocessing as mp
    from jaeger_b200.parallel import merge_rank_frames
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 37
    procs = [ctx.Process(target=_error_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1] == "own:bad shard"
    assert got[0].startswith("peer:") and "rank 1: ValueError: bad shard" in got[0]
    assert merge_rank_frames([None, None]) is None                           # no rows anywhere: no table, no KeyError
