"""jaeger_b200 -- B200-native (sm_100a) implementation of the `jaeger predict` hot path.

Everything that computes loads libjaeger_b200.so (`jaeger_b200._cabi`) and there is no CPU fallback: asking the
package for `B200Engine`, `Context`, `lib` ... raises if the library is missing.  The pure host-side description of
a model (`jaeger_b200.modelspec`: project.yaml parsing, random initialisation, the declared stand-in) has no native
part and can be imported on its own -- `bench.py --impl reference` and the oracle use it without loading the library.
"""
from __future__ import annotations

import importlib

from .modelspec import ModelSpec, init_random, load_project, parse_project, standin_1p4m_config  # noqa: F401

_NATIVE = {"Context": "_cabi", "JaegerB200Error": "_cabi", "lib": "_cabi",
           "B200Engine": "engine", "WindowSource": "engine", "read_fasta": "engine"}
_SUBMODULES = {"_cabi", "build", "codon_tables", "engine", "legacy", "modelspec", "parallel", "plan", "postprocess",
               "predict", "prophage", "refine", "termini", "weights", "ingest"}

__all__ = ["B200Engine", "WindowSource", "Context", "JaegerB200Error", "ModelSpec", "init_random",
           "load_project", "parse_project", "standin_1p4m_config", "read_fasta"]


def __getattr__(name: str):
    if name in _NATIVE:
        return getattr(importlib.import_module(f".{_NATIVE[name]}", __name__), name)
    if name in _SUBMODULES:
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
