"""Test infrastructure: writes a TensorFlow checkpoint bundle (`variables.index` + `variables.data-00000-of-00001`)
the way a SavedModel's `variables/` directory holds it -- a LevelDB-style SSTable of BundleEntryProto records -- so
that the product's pure-Python reader and its bundle -> weights mapping can be exercised without TensorFlow.
Block checksums are written as zeros (the reader does not verify them)."""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9, np.dtype(np.float16): 19}
ATTR = "/.ATTRIBUTES/VARIABLE_VALUE"


def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _field(num: int, wt: int, payload: bytes) -> bytes:
    return _varint((num << 3) | wt) + payload


def _entry_proto(arr: np.ndarray, offset: int) -> bytes:
    shape = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(int(n))) for n in arr.shape))
    return (_field(1, 0, _varint(_DT[arr.dtype])) + _field(2, 2, _varint(len(shape)) + shape) + _field(4, 0, _varint(offset))
            + _field(5, 0, _varint(arr.nbytes)) + _field(6, 5, b"\x00\x00\x00\x00"))


def _block(entries: list[tuple[bytes, bytes]]) -> bytes:
    body = b"".join(_varint(0) + _varint(len(k)) + _varint(len(v)) + k + v for k, v in entries)
    return body + struct.pack("<II", 0, 1)                   # one restart point at offset 0


def write_bundle(variables_dir: str | Path, tensors: dict[str, np.ndarray], per_block: int = 7) -> None:
    d = Path(variables_dir)
    d.mkdir(parents=True, exist_ok=True)
    data = bytearray()
    entries = [(b"", _field(1, 0, _varint(1)))]               # BundleHeaderProto{num_shards: 1}
    for key in sorted(tensors, key=lambda k: k.encode()):
        arr = np.ascontiguousarray(tensors[key])
        entries.append((key.encode(), _entry_proto(arr, len(data))))
        data += arr.tobytes()
    index = bytearray()
    handles = []
    for a in range(0, len(entries), per_block):               # several data blocks, like a real index file
        blk = _block(entries[a:a + per_block])
        handles.append((entries[min(a + per_block, len(entries)) - 1][0], len(index), len(blk)))
        index += blk + b"\x00" + b"\x00\x00\x00\x00"         # compression type + crc
    meta_off, meta = len(index), _block([])
    index += meta + b"\x00" + b"\x00\x00\x00\x00"
    idx_off, idx = len(index), _block([(k, _varint(o) + _varint(n)) for k, o, n in handles])
    index += idx + b"\x00" + b"\x00\x00\x00\x00"
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(idx_off) + _varint(len(idx))
    index += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    (d / "variables.index").write_bytes(bytes(index))
    (d / "variables.data-00000-of-00001").write_bytes(bytes(data))


def _norm_attrs(nw: dict):
    """(bundle attribute, weights key) pairs of a norm layer: MaskedDYT, MaskedLayerNormalization (gamma / beta only), BatchNorm."""
    if "alpha" in nw:
        return (("alpha", "alpha"), ("gamma", "gamma"), ("beta", "beta"))
    if "mean" not in nw:
        return (("gamma", "gamma"), ("beta", "beta"))
    return (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "mean"), ("moving_variance", "var"))


def keras3_export_names(spec, weights: dict) -> dict[str, np.ndarray]:
    """The nested weights of `modelspec.init_random` keyed the way a Keras 3 `model.export()` keys them:
    `_operations/<i>/<attribute>` with built-in layers using `_kernel` / `_embeddings`, the reference's custom layers
    their own attribute names, residual stacks `blocks/<j>/conv1/kernel` ... (see weights.group_bundle)."""
    out: dict[str, np.ndarray] = {}
    op = 1                                                     # 0 is the InputLayer
    def put(path, arr):
        out[f"_operations/{path}{ATTR}"] = np.asarray(arr)
    if weights.get("embedding") is not None:
        put(f"{op}/_embeddings" if spec.uses_token_input else f"{op}/_kernel", weights["embedding"])
        op += 2                                                # a parameter-free op (Masking / Lambda) in between
    for layer, lw in zip(spec.layers, weights["layers"]):
        if layer.kind == "conv":
            put(f"{op}/kernel", lw["kernel"])
            if layer.cfg["use_bias"]:
                put(f"{op}/bias", lw["bias"])
        elif layer.kind == "norm":
            for a, b in _norm_attrs(lw):
                put(f"{op}/{a}", lw[b])
        elif layer.kind == "nmd":
            put(f"{op}/moving_mean", lw["moving_mean"])
        elif layer.kind == "resblock":
            for j, blk in enumerate(lw["blocks"]):
                for part in ("conv1", "conv2", "conv3"):            # conv3 / bn3: the bypass of a strided or use_1x1conv block
                    if part not in blk:
                        continue
                    put(f"{op}/blocks/{j}/{part}/kernel", blk[part]["kernel"])
                    if layer.cfg["use_bias"]:
                        put(f"{op}/blocks/{j}/{part}/bias", blk[part]["bias"])
                for part in ("bn1", "bn2", "bn3"):
                    if part not in blk:
                        continue
                    nw = blk[part]
                    for a, b in _norm_attrs(nw):
                        put(f"{op}/blocks/{j}/{part}/{a}", nw[b])
        op += 1
    op += 2                                                    # pooling, dropout
    put(f"{op}/seed_generator/state", np.array([1, 2], dtype=np.int64))
    for dlist in (weights.get("reliability") or [], weights["classifier"]):     # the heads in either order
        for dw in dlist:
            op += 2
            put(f"{op}/_kernel", dw["kernel"])
            put(f"{op}/bias", dw["bias"])
    out[f"optimizer/_iterations{ATTR}"] = np.array(7, dtype=np.int64)
    return out
