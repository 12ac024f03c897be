"""Oracle (test infrastructure): executes the reference's own serialized TensorFlow graph without TensorFlow.

The reference ships one SavedModel, `src/jaeger/data/models/test/jaeger_fragment_graph` (the legacy `default`
graph exported by Keras 3; its variables are byte-identical to `data/models/default/WRes_1024.h5`).
`saved_model.pb` holds the `serving_default` FunctionDef -- 6 324 nodes of 28 standard TF ops -- and the object
graph that binds its 79 resource arguments to checkpoint variables.  This module parses both (a 30-line protobuf
wire reader, no generated code) and interprets the function with NumPy / torch in float64, so the logits it
returns are what the reference's `InferModel` / legacy `predict` computes for the same tokens up to float32
rounding -- a pin for `oracle/legacy.py` (and through it for the CUDA path) that does not depend on anybody's
reading of nnlib/v1/layers.py.  Op semantics follow the TensorFlow op definitions (NHWC, VALID / SAME padding,
SpaceToBatchND / BatchToSpaceND as tf.nn.convolution lowers dilated convolutions, exact erfc GELU).

Used only by tests/golden/make_legacy_graph_goldens.py (the SavedModel lives under /root/reference, which exists
only in the build container); the goldens it writes travel with the repo.
"""
from __future__ import annotations

import collections
import struct
from pathlib import Path

import numpy as np
import torch

_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 19: np.float16}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _proto(buf) -> dict[int, list]:
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
            raise ValueError(f"wire type {wt}")
        out.setdefault(field, []).append(v)
    return out


def _sint(v: int, bits: int = 64) -> int:
    return v - (1 << bits) if v >= 1 << (bits - 1) else v


def _packed_ints(vals) -> list[int]:
    out = []
    for v in vals:
        if isinstance(v, (bytes, bytearray)):
            p = 0
            while p < len(v):
                x, p = _varint(v, p)
                out.append(_sint(x))
        else:
            out.append(_sint(v))
    return out


def _tensor(buf) -> np.ndarray:
    """TensorProto -> ndarray."""
    t = _proto(buf)
    dtype = _DT[t[1][0]]
    shape = [_sint(_proto(d).get(1, [0])[0]) for d in _proto(t[2][0]).get(2, [])] if 2 in t else []
    n = int(np.prod(shape)) if shape else 1
    if 4 in t and len(t[4][0]):
        return np.frombuffer(t[4][0], dtype=dtype).reshape(shape).copy()
    if dtype == np.float32 and 5 in t:
        vals = []
        for v in t[5]:
            vals += list(struct.unpack(f"<{len(v) // 4}f", v)) if len(v) != 4 else [struct.unpack("<f", v)[0]]
    elif dtype == np.float64 and 6 in t:
        vals = []
        for v in t[6]:
            vals += list(struct.unpack(f"<{len(v) // 8}d", v))
    elif dtype == np.int32 and 7 in t:
        vals = [_sint(x) for x in _packed_ints(t[7])]
    elif dtype == np.int64 and 10 in t:
        vals = _packed_ints(t[10])
    elif dtype == np.bool_ and 11 in t:
        vals = [bool(x) for x in _packed_ints(t[11])]
    else:
        vals = [0]
    arr = np.array(vals, dtype=dtype)
    if arr.size == 1 and n != 1:
        arr = np.full(n, arr[0], dtype=dtype)        # splat encoding
    return arr.reshape(shape)


def _attr(buf):
    """AttrValue -> python value."""
    a = _proto(buf)
    if 2 in a:
        return a[2][0].decode()
    if 3 in a:
        return _sint(a[3][0])
    if 4 in a:
        return struct.unpack("<f", a[4][0])[0]
    if 5 in a:
        return bool(a[5][0])
    if 6 in a:
        return ("type", a[6][0])
    if 8 in a:
        return _tensor(a[8][0])
    if 1 in a:
        lv = _proto(a[1][0])
        if 3 in lv:
            return _packed_ints(lv[3])
        if 2 in lv:
            return [s.decode() for s in lv[2]]
        return []
    return None


class SavedFunction:
    """`serving_default` of a SavedModel directory: `run(list of input arrays) -> {output name: ndarray}`."""

    def __init__(self, model_dir: str | Path, variables: dict[str, np.ndarray], prefix: str = "__inference_serving_default"):
        pb = (Path(model_dir) / "saved_model.pb").read_bytes()
        mg = _proto(_proto(pb)[2][0])                      # SavedModel.meta_graphs[0]
        lib = _proto(_proto(mg[2][0])[2][0])               # MetaGraphDef.graph_def.library
        fdef = None
        for f in lib[1]:
            fd = _proto(f)
            sig = _proto(fd[1][0])
            if sig[1][0].decode().startswith(prefix):
                fdef, self.name = fd, sig[1][0].decode()
        if fdef is None:
            raise ValueError(f"no function named {prefix}* in {model_dir}")
        sig = _proto(fdef[1][0])
        self.args = [(_proto(a)[1][0].decode(), _proto(a)[3][0]) for a in sig.get(2, [])]          # (name, dtype enum)
        self.rets = {}
        for r in fdef.get(4, []):
            e = _proto(r)
            self.rets[e[1][0].decode()] = e[2][0].decode()
        self.nodes = {}
        for n in fdef[3]:
            nd = _proto(n)
            attrs = {}
            for a in nd.get(5, []):
                e = _proto(a)
                attrs[e[1][0].decode()] = e[2][0]
            self.nodes[nd[1][0].decode()] = (nd[2][0].decode(), [i.decode() for i in nd.get(3, [])], attrs)
        # resource arguments -> checkpoint variables: SavedObjectGraph.concrete_functions[name].bound_inputs are object-graph
        # node ids; a variable's checkpoint key is its breadth-first path from the root (how TF names checkpoint entries)
        og = _proto(mg[7][0])
        objs = [_proto(n) for n in og[1]]
        path = {0: ""}
        queue = collections.deque([0])
        while queue:
            i = queue.popleft()
            for c in objs[i].get(1, []):
                ref = _proto(c)
                cid, name = ref.get(1, [0])[0], ref[2][0].decode()
                if cid not in path:
                    path[cid] = (path[i] + "/" if path[i] else "") + name
                    queue.append(cid)
        bound = None
        for e in og[2]:
            m = _proto(e)
            if m[1][0].decode() == self.name:
                bound = _packed_ints(_proto(m[2][0]).get(2, []))
        resources = [a for a, dt in self.args if dt == 20]
        if bound is None or len(bound) != len(resources):
            raise ValueError("bound_inputs do not match the function's resource arguments")
        self.resource = {}
        for arg, node_id in zip(resources, bound):
            key = path[node_id] + "/.ATTRIBUTES/VARIABLE_VALUE"
            self.resource[arg] = np.asarray(variables[key])
        self.inputs = [a for a, dt in self.args if dt != 20]

    # ---- interpreter ---------------------------------------------------------------------------------------------
    def run(self, inputs: list[np.ndarray], dtype=np.float64) -> dict[str, np.ndarray]:
        env: dict[str, list] = {}
        feed = {name: np.asarray(x) for name, x in zip(self.inputs, inputs)}
        fdt = dtype

        def value(ref: str):
            if ref in feed:
                return feed[ref]
            if ref in self.resource:
                return ("resource", ref)
            name, _, idx = ref.split(":") if ref.count(":") == 2 else (ref, None, "0")
            return evaluate(name)[int(idx)]

        def evaluate(name: str):
            stack = [name]
            while stack:
                cur = stack[-1]
                if cur in env:
                    stack.pop()
                    continue
                op, ins, _ = self.nodes[cur]
                pending = []
                for i in ins:
                    if i.startswith("^") or i in feed or i in self.resource:
                        continue
                    dep = i.split(":")[0]
                    if dep not in env:
                        pending.append(dep)
                if pending:
                    stack.extend(pending)
                    continue
                env[cur] = self._exec(cur, [value(i) for i in ins if not i.startswith("^")], fdt)
                stack.pop()
            return env[name]

        return {k: np.asarray(value(v)) for k, v in self.rets.items()}

    def _exec(self, name, x, fdt):
        op, _, raw = self.nodes[name]
        at = {k: _attr(v) for k, v in raw.items() if not k.startswith("_")}
        f = lambda a: np.asarray(a, dtype=fdt) if np.asarray(a).dtype.kind == "f" else np.asarray(a)
        if op == "Const":
            return [f(at["value"])]
        if op == "ReadVariableOp":
            return [f(self.resource[x[0][1]])]
        if op in ("Identity", "NoOp"):
            return [x[0] if x else None]
        if op == "Cast":
            dst = _DT[at["DstT"][1]]
            return [np.asarray(x[0]).astype(fdt if np.dtype(dst).kind == "f" else dst)]      # float -> int truncates like TF
        if op == "Less":
            return [np.less(x[0], x[1])]
        if op == "NotEqual":
            return [np.not_equal(x[0], x[1])]
        if op == "AddV2":
            return [np.add(x[0], x[1])]
        if op == "Sub":
            return [np.subtract(x[0], x[1])]
        if op == "Mul":
            return [np.multiply(x[0], x[1])]
        if op == "Neg":
            return [np.negative(x[0])]
        if op == "Rsqrt":
            return [1.0 / np.sqrt(x[0])]
        if op == "Erfc":
            return [torch.special.erfc(torch.as_tensor(x[0])).numpy()]
        if op == "FloorMod":
            return [np.mod(x[0], x[1])]
        if op == "SelectV2":
            return [np.where(x[0], x[1], x[2])]
        if op == "GatherV2":
            return [np.take(x[0], x[1], axis=int(x[2]))]
        if op == "ExpandDims":
            return [np.expand_dims(x[0], int(x[1]))]
        if op == "Squeeze":
            dims = at.get("squeeze_dims") or None
            return [np.squeeze(x[0], axis=tuple(dims) if dims else None)]
        if op == "Reshape":
            return [np.reshape(x[0], [int(v) for v in x[1]])]
        if op == "Shape":
            return [np.array(np.shape(x[0]), dtype=np.int32)]
        if op == "Pack":
            return [np.stack(x, axis=at.get("axis", 0) or 0)]
        if op == "StridedSlice":
            return [self._strided_slice(x, at)]
        if op == "BiasAdd":
            return [x[0] + x[1]]
        if op == "MatMul":
            a, b = x
            return [(a.T if at.get("transpose_a") else a) @ (b.T if at.get("transpose_b") else b)]
        if op == "Max":
            axes = tuple(int(v) for v in np.atleast_1d(x[1]))
            return [np.max(x[0], axis=axes, keepdims=bool(at.get("keep_dims", False)))]
        if op == "Conv2D":
            return [self._conv2d(x[0], x[1], at)]
        if op == "MaxPool":
            return [self._maxpool(x[0], at)]
        if op == "SpaceToBatchND":
            return [self._space_to_batch(x[0], x[1], x[2])]
        if op == "BatchToSpaceND":
            return [self._batch_to_space(x[0], x[1], x[2])]
        raise NotImplementedError(f"TF op {op} ({name})")

    @staticmethod
    def _strided_slice(x, at):
        data, begin, end, strides = x
        if at.get("ellipsis_mask") or at.get("new_axis_mask"):
            raise NotImplementedError("StridedSlice with ellipsis / new-axis masks")
        bm, em, sm = at.get("begin_mask", 0) or 0, at.get("end_mask", 0) or 0, at.get("shrink_axis_mask", 0) or 0
        idx = []
        for d in range(len(begin)):
            if sm >> d & 1:
                idx.append(int(begin[d]))
            else:
                idx.append(slice(None if bm >> d & 1 else int(begin[d]), None if em >> d & 1 else int(end[d]), int(strides[d])))
        return np.asarray(data)[tuple(idx)]

    @staticmethod
    def _same_pad(n, k, s):
        out = -(-n // s)
        total = max((out - 1) * s + k - n, 0)
        return total // 2, total - total // 2

    def _conv2d(self, x, w, at):
        """NHWC input, HWIO filter, dilations 1 (dilated convolutions arrive through SpaceToBatchND)."""
        if at.get("dilations") not in (None, [1, 1, 1, 1]) or at.get("data_format") not in (None, "NHWC"):
            raise NotImplementedError("Conv2D variant")
        s = at["strides"]
        xt = torch.as_tensor(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        wt = torch.as_tensor(np.ascontiguousarray(w)).permute(3, 2, 0, 1)
        if at["padding"] == "SAME":
            ph, pw = self._same_pad(x.shape[1], w.shape[0], s[1]), self._same_pad(x.shape[2], w.shape[1], s[2])
            xt = torch.nn.functional.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
        elif at["padding"] != "VALID":
            raise NotImplementedError(at["padding"])
        y = torch.nn.functional.conv2d(xt, wt, stride=(s[1], s[2]))
        return y.permute(0, 2, 3, 1).numpy()

    def _maxpool(self, x, at):
        k, s = at["ksize"], at["strides"]
        if at["padding"] != "VALID":
            raise NotImplementedError("MaxPool padding")
        xt = torch.as_tensor(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        y = torch.nn.functional.max_pool2d(xt, kernel_size=(k[1], k[2]), stride=(s[1], s[2]))
        return y.permute(0, 2, 3, 1).numpy()

    @staticmethod
    def _space_to_batch(x, block_shape, paddings):
        """tf.space_to_batch_nd: zero-pad the spatial dims, split each into (size / block, block) and move the block
        factors in front of the batch dimension."""
        block = [int(b) for b in block_shape]
        m = len(block)
        pads = [(0, 0)] + [(int(p[0]), int(p[1])) for p in paddings] + [(0, 0)] * (x.ndim - 1 - m)
        x = np.pad(x, pads)
        n = x.shape[0]
        shape = [n]
        for d in range(m):
            shape += [x.shape[1 + d] // block[d], block[d]]
        shape += list(x.shape[1 + m:])
        x = x.reshape(shape)
        perm = [2 + 2 * d for d in range(m)] + [0] + [1 + 2 * d for d in range(m)] + list(range(1 + 2 * m, x.ndim))
        x = x.transpose(perm)
        return x.reshape([n * int(np.prod(block))] + [shape[1 + 2 * d] for d in range(m)] + list(x.shape[1 + 2 * m:]))

    @staticmethod
    def _batch_to_space(x, block_shape, crops):
        block = [int(b) for b in block_shape]
        m = len(block)
        nb = int(np.prod(block))
        n = x.shape[0] // nb
        x = x.reshape(block + [n] + list(x.shape[1:]))
        # [b_0..b_{m-1}, n, s_0..s_{m-1}, rest] -> [n, s_0, b_0, s_1, b_1, ..., rest]
        perm = [m]
        for d in range(m):
            perm += [m + 1 + d, d]
        perm += list(range(2 * m + 1, x.ndim))
        x = x.transpose(perm)
        shape = [n] + [x.shape[1 + 2 * d] * x.shape[2 + 2 * d] for d in range(m)] + list(x.shape[1 + 2 * m:])
        x = x.reshape(shape)
        idx = [slice(None)] + [slice(int(c[0]), x.shape[1 + d] - int(c[1])) for d, c in enumerate(crops)]
        return x[tuple(idx)]
