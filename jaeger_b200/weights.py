"""Weight loading without TensorFlow / h5py.

* `read_tf_bundle(dir)`: pure-Python reader of a SavedModel `variables/` bundle
  (`variables.index` is a LevelDB-style SSTable of BundleEntryProto records pointing into
  `variables.data-00000-of-00001`), which is how the reference stores a model next to its
  `*_project.yaml` (nnlib/builder.py:1495-1529).
* `save_npz_weights` / `load_npz_weights`: the nested weights dict of `modelspec.py` flattened
  into an `.npz` (`<name>.weights.npz` next to the project file).  `tools/export_weights_npz.py`
  writes that file from a Keras model inside a reference environment.
* `load_saved_model_weights(path_dict, spec)`: what `B200Engine(path_dict)` calls.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Any

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16, 10: np.bool_}


def _varint(buf: bytes, pos: int) -> tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _parse_proto(buf: bytes) -> dict[int, list]:
    """Minimal protobuf wire parser: field number -> list of raw values."""
    out: dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _read_block(data: bytes, offset: int, size: int) -> list[tuple[bytes, bytes]]:
    """Entries of one uncompressed SSTable block (prefix-compressed keys + restart array)."""
    block = data[offset:offset + size]
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_tf_bundle(variables_dir: str | Path) -> dict[str, np.ndarray]:
    """All tensors of a TF checkpoint bundle, keyed by their object-graph path."""
    d = Path(variables_dir)
    index = (d / "variables.index").read_bytes()
    blob = (d / "variables.data-00000-of-00001").read_bytes()
    footer = index[-48:]
    pos = 0
    _, pos = _varint(footer, pos)            # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)      # index block handle
    idx_size, pos = _varint(footer, pos)
    tensors: dict[str, np.ndarray] = {}
    for _, handle in _read_block(index, idx_off, idx_size):
        p = 0
        off, p = _varint(handle, p)
        size, p = _varint(handle, p)
        for key, val in _read_block(index, off, size):
            if not key:
                continue                      # bundle header
            e = _parse_proto(val)
            dtype = _DTYPES.get(e.get(1, [0])[0])
            if dtype is None:
                continue
            shape = []
            if 2 in e:
                for dim in _parse_proto(e[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            o, n = e.get(4, [0])[0], e.get(5, [0])[0]
            tensors[key.decode()] = np.frombuffer(blob, dtype=dtype, count=n // np.dtype(dtype).itemsize, offset=o).reshape(shape).copy()
    return tensors


def _flatten(w: Any, prefix: str, out: dict[str, np.ndarray]) -> None:
    if isinstance(w, dict):
        for k, v in w.items():
            _flatten(v, f"{prefix}/{k}" if prefix else k, out)
    elif isinstance(w, (list, tuple)):
        for i, v in enumerate(w):
            _flatten(v, f"{prefix}/{i}", out)
    elif w is not None:
        out[prefix] = np.asarray(w)


def save_npz_weights(path: str | Path, weights: dict[str, Any]) -> None:
    flat: dict[str, np.ndarray] = {}
    _flatten(weights, "", flat)
    np.savez(path, **flat)


def load_npz_weights(path: str | Path) -> dict[str, Any]:
    z = np.load(path)
    root: dict[str, Any] = {}
    for key in z.files:
        parts = key.split("/")
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = z[key]

    def fix(n):
        if isinstance(n, dict):
            if n and all(k.isdigit() for k in n):        # a list; parameter-free members were not stored
                return [fix(n[str(i)]) if str(i) in n else {} for i in range(max(int(k) for k in n) + 1)]
            return {k: fix(v) for k, v in n.items()}
        return n

    w = fix(root)
    w.setdefault("embedding", None)
    # layers without parameters are dropped by the flattening: rebuild them as {}
    return w


def load_saved_model_weights(path_dict: dict[str, Any], spec) -> dict[str, Any]:
    project = Path(path_dict["project"])
    npz = project.with_name(project.name.replace("_project.yaml", ".weights.npz"))
    if npz.exists():
        w = load_npz_weights(npz)
        layers = list(w.get("layers", []))
        layers += [{} for _ in range(len(spec.layers) - len(layers))]     # trailing parameter-free layers
        w["layers"] = layers
        return w
    raise NotImplementedError(
        f"{npz.name} not found. The SavedModel / .weights.h5 key layout of layer-list models cannot be validated "
        "offline (no such model is vendored by the reference); export the weights once inside a reference "
        "environment with tools/export_weights_npz.py.")
