"""jaeger_b200 -- B200-native (sm_100a) implementation of the `jaeger predict` hot path.

Importing this package loads libjaeger_b200.so; there is no CPU fallback.
"""
from ._cabi import Context, JaegerB200Error, lib  # noqa: F401
from .engine import B200Engine, WindowSource, read_fasta  # noqa: F401
from .modelspec import ModelSpec, init_random, load_project, parse_project, standin_1p4m_config  # noqa: F401

__all__ = ["B200Engine", "WindowSource", "Context", "JaegerB200Error", "ModelSpec", "init_random",
           "load_project", "parse_project", "standin_1p4m_config", "read_fasta"]
