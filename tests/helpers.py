"""Shared synthetic inputs for the parity tests (seeded, no file I/O)."""
from __future__ import annotations

import numpy as np


def random_contigs(seed: int, lengths, n_run_every: int = 3, lower_every: int = 4):
    """Contigs of i.i.d. bases; every n_run_every-th gets a run of N, every lower_every-th
    gets a lower-case stretch and a few IUPAC codes."""
    rng = np.random.default_rng(seed)
    recs = []
    for i, n in enumerate(lengths):
        s = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).copy()
        if n_run_every and i % n_run_every == 1 and n > 400:
            a = int(rng.integers(0, n - 300))
            s[a:a + int(rng.integers(20, 300))] = ord("N")
        if lower_every and i % lower_every == 2 and n > 200:
            a = int(rng.integers(0, n - 150))
            s[a:a + 120] |= 0x20
            s[int(rng.integers(0, n))] = ord("R")
            s[int(rng.integers(0, n))] = ord("y")
        recs.append((f"c{i}", s.tobytes().decode()))
    return recs


def to_dyt(cfg: dict) -> dict:
    """The same architecture with MaskedDYT in place of every MaskedBatchNorm (the reference's
    train_config/nn_config_1500bp_nmd_merge_6_class_zeus.yaml family)."""
    for layer in cfg["model"]["representation_learner"]["hidden_layers"]:
        if layer["name"] == "masked_batchnorm":
            layer["name"] = "masked_dyt"
            layer["config"] = {}
        elif layer["name"] == "residual_block":
            layer["config"]["norm_type"] = "masked_dyt"
    return cfg
