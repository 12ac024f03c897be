"""Test infrastructure: a NumPy stand-in for the ~30 TensorFlow / Keras symbols that the `call` bodies of the reference's
fragment-model layers use (nnlib/v2/layers.py: GeLU, MaskedConv1D, MaskedBatchNorm, MaskedDYT, MaskedGlobalMax/AvgPooling;
nnlib/v2/nmd.py: NMDLayer), so that those bodies can be EXECUTED FROM THE REFERENCE'S OWN SOURCE in a container without
TensorFlow.  Everything computes in float64.  It does not emulate Keras' mask propagation between layers: layers are built
and called one at a time with explicit masks (`layer.build(shape)`, `layer.call(x, mask=...)`).
Install with `install()` before importing `jaeger.nnlib.v2.layers`."""
from __future__ import annotations

import sys
import types

import numpy as np
import torch


class Shape(tuple):
    @property
    def rank(self):
        return len(self)


class T(np.ndarray):
    """ndarray whose `.shape` has a `.rank`, like a tf.Tensor's TensorShape."""

    @property
    def shape(self):
        return Shape(np.ndarray.shape.__get__(self))

    def assign(self, value):
        self[...] = value


def t(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    return a.view(T) if a.ndim else a.reshape(()).view(T)


class _Dtype:
    def __init__(self, name, np_dtype):
        self.name, self.np = name, np_dtype


FLOAT32, FLOAT16, INT32, BOOL = _Dtype("float32", np.float64), _Dtype("float16", np.float64), _Dtype("int32", np.int32), _Dtype("bool", np.bool_)


def _np_dtype(d):
    if isinstance(d, _Dtype):
        return d.np
    if isinstance(d, str):
        return {"float32": np.float64, "float16": np.float64, "bfloat16": np.float64, "int32": np.int32, "bool": np.bool_}[d]
    if isinstance(d, np.dtype) or isinstance(d, type):
        return np.float64 if np.dtype(d).kind == "f" else d
    raise TypeError(d)


def cast(x, dtype):
    return t(np.asarray(x).astype(_np_dtype(dtype)))


def _same_pad(n, k, s, d):
    out = -(-n // s)
    total = max((out - 1) * s + d * (k - 1) + 1 - n, 0)
    return total // 2, total - total // 2


def conv1d(input, filters, stride, padding, dilations=None, data_format="NWC"):
    """tf.nn.conv1d, NWC input [N, W, Cin], filters [k, Cin, Cout]; SAME pads floor(total/2) left (TF)."""
    assert data_format == "NWC"
    s = int(stride if np.ndim(stride) == 0 else stride[0])
    d = int(1 if dilations is None else (dilations if np.ndim(dilations) == 0 else dilations[0]))
    x = torch.as_tensor(np.asarray(input, dtype=np.float64)).transpose(1, 2)
    w = torch.as_tensor(np.asarray(filters, dtype=np.float64)).permute(2, 1, 0).contiguous()
    if padding.upper() == "SAME":
        left, right = _same_pad(x.shape[2], w.shape[2], s, d)
        x = torch.nn.functional.pad(x, (left, right))
    elif padding.upper() != "VALID":
        raise ValueError(padding)
    return t(torch.nn.functional.conv1d(x, w, stride=s, dilation=d).transpose(1, 2).numpy())


def gelu(x, approximate=False):
    x = np.asarray(x, dtype=np.float64)
    if approximate:
        return t(0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3))))
    return t(0.5 * x * (1.0 + torch.erf(torch.as_tensor(x / np.sqrt(2.0))).numpy()))


def _axes(axis):
    return None if axis is None else (tuple(int(a) for a in axis) if np.ndim(axis) else int(axis))


WEIGHT_PROVIDER = None      # callable(layer name, weight name, shape) -> array; set by a golden generator
WEIGHT_LOG: list = []


def get_keras_mask(x):
    if isinstance(x, (list, tuple)):
        return [get_keras_mask(v) for v in x]
    return getattr(x, "_keras_mask", None)


class Layer:
    """keras.layers.Layer as far as the layers' __init__ / build / call need it, plus Keras 3's `__call__` mask rules
    (keras/src/layers/layer.py, restated from the library's documented behaviour -- Keras itself is not installable here):
      1. a `mask` argument of `call` that was not passed explicitly is filled from the `_keras_mask` of the first argument;
      2. after `call`, if the layer `supports_masking`, every output that does not ALREADY carry a mask (one set by an inner
         layer is kept: `_set_mask_metadata` returns early when all outputs have one) gets
         `compute_mask(first argument, mask of the first argument)`, outputs and masks paired in order and an existing mask
         never overwritten; the default `compute_mask` passes the incoming mask through;
      3. a layer that does not support masking leaves its outputs without a mask;
      4. `supports_masking` defaults to "the class overrides compute_mask" (`not utils.is_default(self.compute_mask)` in
         Layer.__init__), so the reference's MaskedAdd -- which overrides compute_mask without setting the flag -- propagates
         the mask of its first input.
    Rules 2 and 4 decide which mask leaves a ResidualBlock (conv2's, attached by the inner layers, rather than the block
    input's); they are written down from the Keras 3 sources as remembered, not executed from them."""

    def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
        self.name, self.trainable, self.built = name, trainable, False
        self.supports_masking = type(self).compute_mask is not Layer.compute_mask
        self.compute_dtype = self.variable_dtype = "float32"

    def add_weight(self, name=None, shape=(), initializer="zeros", trainable=True, dtype=None, **kw):
        shape = tuple(int(s) for s in shape)
        if WEIGHT_PROVIDER is not None:                      # seeded values, logged in creation order
            arr = t(np.asarray(WEIGHT_PROVIDER(self.name, name, shape), dtype=np.float64))
            WEIGHT_LOG.append((self.name, name, arr))
            return arr
        if callable(initializer) and not isinstance(initializer, str):
            return t(np.asarray(initializer(shape), dtype=np.float64))
        return t(np.ones(shape) if str(initializer) == "ones" else np.zeros(shape))

    def build(self, input_shape):
        self.built = True

    def compute_mask(self, inputs, previous_mask=None):
        return previous_mask

    def get_config(self):
        return {}

    def __call__(self, inputs, *args, **kwargs):
        import inspect
        if not self.built:
            first = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
            self.build(tuple(np.ndarray.shape.__get__(np.asarray(first))))
            self.built = True
        params = inspect.signature(self.call).parameters
        previous_mask = get_keras_mask(inputs)
        if "mask" in params and kwargs.get("mask") is None:
            kwargs["mask"] = previous_mask
        if "training" in params and "training" not in kwargs:
            kwargs["training"] = False
        kwargs = {k: v for k, v in kwargs.items() if k in params or any(p.kind == p.VAR_KEYWORD for p in params.values())}
        outputs = self.call(inputs, *args, **kwargs)
        if self.supports_masking:                       # Layer._set_mask_metadata
            flat = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
            if not all(get_keras_mask(o) is not None for o in flat):
                m = self.compute_mask(inputs, previous_mask)
                if m is not None:
                    masks = list(m) if isinstance(m, (list, tuple)) else [m]
                    for o, mk in zip(flat, masks):           # pairs outputs with masks in order; never overwrites
                        if get_keras_mask(o) is None and mk is not None:
                            o._keras_mask = mk
        return outputs


class Add(Layer):
    """keras.layers.Add (used by ResidualBlock when use_masking is off: no masks are around)."""

    def call(self, inputs):
        return t(sum(np.asarray(v) for v in inputs))


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, inputs):
        return t(np.concatenate([np.asarray(v) for v in inputs], axis=self.axis))


class Activation(Layer):
    """keras.layers.Activation: supports masking, applies keras.activations.get(name) (gelu = the tanh approximation)."""

    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.supports_masking = True
        self.activation = activation

    def call(self, inputs):
        x = np.asarray(inputs)
        if self.activation == "gelu":
            return gelu(x, approximate=True)
        if self.activation == "relu":
            return t(np.maximum(x, 0.0))
        if self.activation in (None, "linear"):
            return t(x)
        raise NotImplementedError(self.activation)


def _passthrough(*a, **k):
    return a[0] if a else None


class _Anything:
    """Import-time filler for symbols the target call bodies never touch (other layers, annotations, serializers)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __mro_entries__(self, bases):
        return (Layer,)

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __getitem__(self, item):
        return self


class _NS(types.SimpleNamespace):
    """Namespace whose unknown attributes are import-time fillers."""

    def __getattr__(self, name):
        return _Anything()


def install():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float16, tf.int32, tf.bool, tf.Tensor = FLOAT32, FLOAT16, INT32, BOOL, T
    tf.cast = cast
    tf.constant = lambda v, dtype=None: t(np.asarray(v) if dtype is None else np.asarray(v).astype(_np_dtype(dtype)))
    tf.shape = lambda x: t(np.array(np.shape(x), dtype=np.int32))
    tf.reshape = lambda x, shape: t(np.reshape(np.asarray(x), [int(s) for s in np.asarray(shape).reshape(-1)]))
    tf.expand_dims = lambda x, axis: t(np.expand_dims(np.asarray(x), int(axis)))
    tf.squeeze = lambda x, axis=None: t(np.squeeze(np.asarray(x), axis=_axes(axis)))
    tf.ones = lambda shape, dtype=FLOAT32: t(np.ones([int(s) for s in np.asarray(shape).reshape(-1)], dtype=_np_dtype(dtype)))
    tf.zeros_like = lambda x, dtype=None: t(np.zeros_like(np.asarray(x)))
    tf.concat = lambda vals, axis: t(np.concatenate([np.atleast_1d(np.asarray(v)) for v in vals], axis=int(axis)))
    tf.equal = lambda a, b: t(np.equal(np.asarray(a), np.asarray(b)))
    tf.where = lambda c, a, b: t(np.where(np.asarray(c), np.asarray(a), np.asarray(b)))
    tf.maximum = lambda a, b: t(np.maximum(np.asarray(a), np.asarray(b)))
    tf.square = lambda x: t(np.square(np.asarray(x)))
    tf.sqrt = lambda x: t(np.sqrt(np.asarray(x)))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: t(np.sum(np.asarray(x), axis=_axes(axis), keepdims=keepdims))
    tf.reduce_mean = lambda x, axis=None, keepdims=False: t(np.mean(np.asarray(x), axis=_axes(axis), keepdims=keepdims))
    tf.reduce_max = lambda x, axis=None, keepdims=False: t(np.max(np.asarray(x), axis=_axes(axis), keepdims=keepdims))
    tf.stop_gradient = _passthrough
    tf.reduce_logsumexp = lambda x, axis=None, keepdims=False: t(torch.logsumexp(torch.as_tensor(np.asarray(x, dtype=np.float64)), dim=_axes(axis), keepdim=keepdims).numpy())
    tf.norm = lambda x, ord="euclidean", axis=None, keepdims=False: t(np.sqrt(np.sum(np.square(np.asarray(x)), axis=_axes(axis), keepdims=keepdims)))
    tf.math = _NS(add_n=lambda xs: t(sum(np.asarray(v) for v in xs)), log=lambda x: t(np.log(np.asarray(x))), rsqrt=lambda x: t(1.0 / np.sqrt(np.asarray(x))), tanh=lambda x: t(np.tanh(np.asarray(x))),
                                    divide_no_nan=lambda a, b: t(np.where(np.asarray(b) == 0, 0.0, np.asarray(a) / np.where(np.asarray(b) == 0, 1.0, np.asarray(b)))))
    def _top_k(x, k=1):
        v = -np.sort(-np.asarray(x), axis=-1)[..., :k]
        return _NS(values=t(v))

    def _softmax(x, axis=-1):
        return t(torch.softmax(torch.as_tensor(np.asarray(x, dtype=np.float64)), dim=int(axis)).numpy())
    tf.nn = _NS(softmax=_softmax, top_k=_top_k, conv1d=conv1d, gelu=gelu, bias_add=lambda x, b, data_format=None: t(np.asarray(x) + np.asarray(b)),
                                  moments=lambda x, axes, keepdims=False: (t(np.mean(np.asarray(x), axis=_axes(axes), keepdims=keepdims)),
                                                                           t(np.var(np.asarray(x), axis=_axes(axes), keepdims=keepdims))))
    ker = types.ModuleType("tensorflow.keras")
    ker.layers = _NS(Layer=Layer, Add=Add, Activation=Activation, Concatenate=Concatenate)
    ker.Model = Layer
    ker.activations = _NS(get=lambda a: None if a in (None, "linear") else (lambda x: gelu(x, approximate=True)) if a == "gelu" else _Anything(),
                                            serialize=lambda a: a)
    ker.initializers = _NS(get=lambda x: x, serialize=lambda x: x, Constant=lambda v: (lambda shape: np.full(shape, v)))
    ker.regularizers = _NS(get=lambda x: x, serialize=lambda x: x)
    ker.backend = _NS(epsilon=lambda: 1e-7)
    filler = _Anything()
    for ns in (ker, ker.layers):
        orig = ns.__dict__

    def _mod_getattr(mod):
        def g(name):
            return filler
        return g
    tf.__getattr__ = _mod_getattr(tf)
    ker.__getattr__ = _mod_getattr(ker)
    tf.keras = ker
    sys.modules["tensorflow"], sys.modules["tensorflow.keras"], sys.modules["keras"] = tf, ker, ker
    return tf
