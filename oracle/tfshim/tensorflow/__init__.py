"""A stand-in for the ``tensorflow`` package, just large enough to EXECUTE the reference's own model files.

TEST INFRASTRUCTURE ONLY (the rule of ``oracle/``: only ``tests/`` and the golden-vector generators import it; nothing in
``ultrasound_modeling_b200/`` does).  Purpose: TensorFlow cannot be installed in this image, so the reference's
``TBI_ResNest.py`` / ``ResNest.py`` / ``Decoder.py`` / ``VisionTransformer.py`` could not run and the oracles were "parity
unpinned".  With ``oracle/tfshim`` first on ``sys.path`` those files import UNMODIFIED from ``/root/reference`` and their own
code -- the functional-API graph wiring, layer creation order and auto-naming, ``step`` / ``train_step``, ``my_loss_cat``,
``compute_loss``, the GradientTape / clip / Adam sequence -- runs end to end; ``tests/golden/make_golden_ref.py`` records what
they produce and ``tests/test_oracle_pinned.py`` holds the oracles to it.

What this pins and what it does not: the REFERENCE'S CODE is real; the PRIMITIVES underneath (conv, norm, pooling, Adam ...)
are restated here from TensorFlow/Keras' documented definitions, in float64 on torch CPU tensors with autograd as the tape.
They are written independently of the oracles' formulations on purpose: convolutions here are explicit tap sums over
TF's SAME padding rule (``pad_total = max((ceil(n/s)-1)*s + (k-1)*d + 1 - n, 0)``, ``pad_before = pad_total // 2``) and
the transposed convolution is a scatter over the same rule, where the oracles call ``torch.nn.functional.conv2d`` /
``conv_transpose2d`` with hand-derived paddings and crops.

Keras semantics implemented (the defaults the reference relies on):
  Conv2D / Conv2DTranspose / Dense: glorot_uniform kernels unless ``kernel_initializer`` is given, zero bias, kernels HWIO /
  HWOI / [in,out]; ``padding`` case-insensitive; BatchNormalization: eps 1e-3, called without ``training=True`` => moving
  statistics; LayerNormalization: last axis, eps 1e-3 unless given, biased variance; ELU alpha 1; LeakyReLU slope 0.3;
  AveragePooling2D pool 2 stride 2 valid; gelu exact (erf); ``tf.nn.dropout(x, rate)`` keeps with probability 1-rate and scales
  by 1/(1-rate); CategoricalCrossentropy on probabilities: smooth labels, renormalise, clip to [1e-7, 1-1e-7];
  ``tf.nn.compute_average_loss``: sum / global_batch_size; ``tf.clip_by_global_norm``; Adam (optimizer_v2 form):
  ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t); var -= lr_t*m/(sqrt(v)+1e-7)``; ``tape.gradient`` of a non-scalar target differentiates
  its sum.  Layer auto-names follow Keras (snake-cased class name + ``_<n>`` from the second instance on).
Tensors are immutable like TF's: ``a += b`` rebinds, it never writes into ``a`` (ResNest.py:177 depends on that).
"""
from __future__ import annotations

import math as _pymath
import re
import sys
import types
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as _np
import torch as _torch

__version__ = "2.shim"
_DT = _torch.float64            # every float dtype of the reference maps onto the working precision


class DType:
    def __init__(self, name, is_float):
        self.name, self.is_floating = name, is_float

    def __repr__(self):
        return "tf." + self.name


float32, float64, float16 = DType("float32", True), DType("float64", True), DType("float16", True)
int32, int64, bool_ = DType("int32", False), DType("int64", False), DType("bool", False)


def _torch_dtype(dt):
    if dt is None:
        return None
    if isinstance(dt, DType):
        return _DT if dt.is_floating else (_torch.bool if dt.name == "bool" else _torch.int64)
    if dt in (_np.float32, _np.float64, float):
        return _DT
    return _torch.int64


# ------------------------------------------------------------------------------------------------------------------
# tensors: eager values or nodes of a functional-API graph
# ------------------------------------------------------------------------------------------------------------------
class TensorShape(list):
    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return len(self)


class _Node:
    __slots__ = ("fn", "args", "kwargs", "layer")

    def __init__(self, fn, args, kwargs, layer=None):
        self.fn, self.args, self.kwargs, self.layer = fn, args, kwargs, layer


class Tensor:
    """Immutable value.  ``_node`` is None for an eager tensor; for a symbolic one (functional API) it records how to
    recompute the value and ``_v`` holds the value for the example input (shape inference)."""
    __array_priority__ = 1000

    def __init__(self, value, node=None, index=None):
        self._v = value
        self._node = node
        self._index = index

    # ---- introspection
    @property
    def shape(self):
        return TensorShape(self._v.shape)

    @property
    def dtype(self):
        return float32 if self._v.is_floating_point() else (bool_ if self._v.dtype == _torch.bool else int64)

    def numpy(self):
        return self._v.detach().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self._v.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __int__(self):
        return int(self._v)

    def __index__(self):
        return int(self._v)

    def __float__(self):
        return float(self._v)

    def __bool__(self):
        return bool(self._v)

    def __len__(self):
        return self._v.shape[0]

    def __iter__(self):
        return (self[i] for i in range(self._v.shape[0]))

    def __repr__(self):
        return f"<shim tf.Tensor shape={tuple(self._v.shape)}{' symbolic' if self._node is not None else ''}>"

    __hash__ = object.__hash__

    # ---- arithmetic (no in-place forms: ``+=`` falls back to ``__add__`` and rebinds)
    def __add__(self, o): return _apply(lambda a, b: a + b, self, o)
    def __radd__(self, o): return _apply(lambda a, b: b + a, self, o)
    def __sub__(self, o): return _apply(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return _apply(lambda a, b: b - a, self, o)
    def __mul__(self, o): return _apply(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return _apply(lambda a, b: b * a, self, o)
    def __truediv__(self, o): return _apply(lambda a, b: a / b, self, o)
    def __rtruediv__(self, o): return _apply(lambda a, b: b / a, self, o)
    def __neg__(self): return _apply(lambda a: -a, self)
    def __pow__(self, o): return _apply(lambda a, b: a ** b, self, o)
    def __eq__(self, o): return _apply(lambda a, b: a == b, self, o)
    def __ne__(self, o): return _apply(lambda a, b: a != b, self, o)
    def __getitem__(self, idx): return _apply(lambda a: a[idx], self)


def _as_torch(x):
    """eager python / numpy / Variable / Tensor value -> torch"""
    if isinstance(x, Tensor):
        return x._v
    if isinstance(x, Variable):
        return x._t
    if isinstance(x, _torch.Tensor):
        return x
    if isinstance(x, _np.ndarray):
        t = _torch.from_numpy(_np.ascontiguousarray(x))
        return t.to(_DT) if t.is_floating_point() else t
    if isinstance(x, (list, tuple)) and any(isinstance(e, (Tensor, Variable, _torch.Tensor, _np.ndarray, list, tuple)) for e in x):
        return type(x)(_as_torch(e) for e in x)
    return x


def _is_sym(x):
    if isinstance(x, Tensor):
        return x._node is not None
    if isinstance(x, (list, tuple)):
        return any(_is_sym(e) for e in x)
    return False


def _wrap(out, node):
    if isinstance(out, (list, tuple)):
        return type(out)(Tensor(o, node, i) if isinstance(o, _torch.Tensor) else o for i, o in enumerate(out))
    if isinstance(out, _torch.Tensor):
        return Tensor(out, node)
    return out


def _apply(fn: Callable, *args, _layer=None, **kwargs):
    """Run ``fn`` on torch values; if any argument is symbolic the result is symbolic too (a node is recorded)."""
    sym = _is_sym(args) or _is_sym(tuple(kwargs.values()))
    out = fn(*[_as_torch(a) for a in args], **{k: _as_torch(v) for k, v in kwargs.items()})
    return _wrap(out, _Node(fn, args, kwargs, _layer) if sym else None)


class Variable:
    def __init__(self, initial_value, trainable=True, name=None, dtype=None):
        t = _as_torch(initial_value)
        if not isinstance(t, _torch.Tensor):
            t = _torch.as_tensor(t, dtype=_DT)
        self._t = t.detach().clone().to(_DT).requires_grad_(bool(trainable))
        self.trainable = trainable
        self.name = (name or "Variable") + ":0"

    @property
    def shape(self):
        return TensorShape(self._t.shape)

    def numpy(self):
        return self._t.detach().numpy().copy()

    def assign(self, value):
        v = _as_torch(value)
        v = _torch.as_tensor(v, dtype=_DT) if not isinstance(v, _torch.Tensor) else v.to(_DT)
        assert tuple(v.shape) == tuple(self._t.shape), (self.name, tuple(v.shape), tuple(self._t.shape))
        with _torch.no_grad():
            self._t.copy_(v)
        return self

    def value(self):
        return Tensor(self._t)

    def __repr__(self):
        return f"<shim tf.Variable {self.name} shape={tuple(self._t.shape)}>"


# ------------------------------------------------------------------------------------------------------------------
# primitives (torch in, torch out), TensorFlow's documented definitions
# ------------------------------------------------------------------------------------------------------------------
def _same_pad(n, k, s, d):
    out = -(-n // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - n, 0)
    return out, total // 2, total - total // 2


def _conv2d(x, kernel, stride=(1, 1), padding="SAME", dilation=(1, 1)):
    """tf.nn.conv2d: x NHWC, kernel HWIO; explicit sum over the kh*kw taps."""
    n, h, w, cin = x.shape
    kh, kw, kcin, cout = kernel.shape
    assert kcin == cin, (kcin, cin)
    sh, sw = stride
    dh, dw = dilation
    if padding.upper() == "SAME":
        oh, pt, pb = _same_pad(h, kh, sh, dh)
        ow, pl, pr = _same_pad(w, kw, sw, dw)
    else:
        oh, pt, pb = (h - (kh - 1) * dh - 1) // sh + 1, 0, 0
        ow, pl, pr = (w - (kw - 1) * dw - 1) // sw + 1, 0, 0
    xp = _torch.zeros(n, h + pt + pb, w + pl + pr, cin, dtype=x.dtype)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    y = None
    for i in range(kh):
        for j in range(kw):
            win = xp[:, i * dh: i * dh + (oh - 1) * sh + 1: sh, j * dw: j * dw + (ow - 1) * sw + 1: sw, :]
            t = win @ kernel[i, j]
            y = t if y is None else y + t
    return y


def _conv2d_transpose(x, kernel, stride=(2, 2), padding="SAME"):
    """tf.nn.conv2d_transpose: the gradient of conv2d w.r.t. its input.  x NHWC [n,h,w,cin]; kernel HWOI [kh,kw,cout,cin];
    output [n, h*s, w*s, cout] for SAME.  Scatter form: every input pixel adds x @ k[i,j]^T at (s*y + i - pad_before)."""
    n, h, w, cin = x.shape
    kh, kw, cout, kcin = kernel.shape
    assert kcin == cin, (kcin, cin)
    sh, sw = stride
    if padding.upper() == "SAME":
        oh, ow = h * sh, w * sw
        _, pt, _ = _same_pad(oh, kh, sh, 1)
        _, pl, _ = _same_pad(ow, kw, sw, 1)
    else:
        oh, ow, pt, pl = (h - 1) * sh + kh, (w - 1) * sw + kw, 0, 0
    fh, fw = (h - 1) * sh + kh, (w - 1) * sw + kw
    full = _torch.zeros(n, max(fh, oh + pt), max(fw, ow + pl), cout, dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            contrib = x @ kernel[i, j].transpose(0, 1)
            canvas = _torch.zeros_like(full)
            canvas[:, i: i + (h - 1) * sh + 1: sh, j: j + (w - 1) * sw + 1: sw, :] = contrib
            full = full + canvas
    return full[:, pt: pt + oh, pl: pl + ow, :]


def _avg_pool(x, pool, stride):
    n, h, w, c = x.shape
    oh, ow = (h - pool[0]) // stride[0] + 1, (w - pool[1]) // stride[1] + 1
    y = None
    for i in range(pool[0]):
        for j in range(pool[1]):
            t = x[:, i: i + (oh - 1) * stride[0] + 1: stride[0], j: j + (ow - 1) * stride[1] + 1: stride[1], :]
            y = t if y is None else y + t
    return y / float(pool[0] * pool[1])


def _softmax(x, axis=-1):
    e = _torch.exp(x - x.max(dim=axis, keepdim=True).values.detach())
    return e / e.sum(dim=axis, keepdim=True)


def _sigmoid(x):
    return 1.0 / (1.0 + _torch.exp(-x))


def _gelu(x, approximate=False):
    if approximate:
        return 0.5 * x * (1.0 + _torch.tanh(_pymath.sqrt(2.0 / _pymath.pi) * (x + 0.044715 * x ** 3)))
    return 0.5 * x * (1.0 + _torch.erf(x / _pymath.sqrt(2.0)))


def _elu(x, alpha=1.0):
    return _torch.where(x > 0, x, alpha * (_torch.exp(_torch.clamp(x, max=0.0)) - 1.0))


def _leaky(x, alpha=0.3):
    return _torch.where(x >= 0, x, alpha * x)


_ACTIVATIONS = {None: None, "linear": None, "softmax": _softmax, "sigmoid": _sigmoid, "relu": lambda x: _torch.clamp(x, min=0.0),
                "gelu": _gelu, "elu": _elu}

# dropout masks: a test may queue keep-masks (consumed in call order); otherwise a seeded generator draws them
_dropout_queue: List[Any] = []
_dropout_log: List[_torch.Tensor] = []
_rng = _torch.Generator().manual_seed(0)


def shim_queue_dropout_masks(masks: Sequence[Any]):
    _dropout_queue[:] = list(masks)


def shim_seed(seed: int):
    _rng.manual_seed(seed)


def _dropout(x, rate):
    if rate == 0:
        return x
    if _dropout_queue:
        keep = _as_torch(_dropout_queue.pop(0))
        keep = _torch.as_tensor(keep).to(_torch.bool)
        assert tuple(keep.shape) == tuple(x.shape), (tuple(keep.shape), tuple(x.shape))
    else:
        keep = _torch.rand(x.shape, generator=_rng, dtype=_DT) >= rate
    _dropout_log.append(keep)
    return x * keep.to(x.dtype) * (1.0 / (1.0 - rate))


# ------------------------------------------------------------------------------------------------------------------
# tf.* functions
# ------------------------------------------------------------------------------------------------------------------
def _shape_list(shape):
    if isinstance(shape, (Tensor, _torch.Tensor)):
        shape = _as_torch(shape).tolist()
    return [int(d) for d in shape]


def convert_to_tensor(value, dtype=None, name=None):
    if isinstance(value, Tensor):
        return value if dtype is None else cast(value, dtype)
    if isinstance(value, Variable):
        return value.value()
    if isinstance(value, (list, tuple)) and any(isinstance(e, (Tensor, list, tuple)) for e in value):
        value = _np.array(_np.asarray([_np.asarray(e) for e in value]))
    t = _torch.as_tensor(_np.asarray(value))
    td = _torch_dtype(dtype)
    if td is not None:
        t = t.to(td)
    elif t.is_floating_point():
        t = t.to(_DT)
    return Tensor(t)


constant = convert_to_tensor


def cast(x, dtype, name=None):
    return _apply(lambda a: (a if isinstance(a, _torch.Tensor) else _torch.as_tensor(a)).to(_torch_dtype(dtype)), x)


def _binary(f):
    def op(x=None, y=None, name=None):
        return _apply(lambda a, b: f(_t(a), _t(b)), x, y)
    return op


def _t(a):
    return a if isinstance(a, _torch.Tensor) else _torch.as_tensor(a, dtype=_DT if isinstance(a, float) else None)


add = _binary(lambda a, b: a + b)
subtract = _binary(lambda a, b: a - b)
multiply = _binary(lambda a, b: a * b)
divide = _binary(lambda a, b: a / b)
equal = _binary(lambda a, b: a == b)


def pow(x, y, name=None):  # noqa: A001
    if isinstance(x, int) and isinstance(y, int):
        return Tensor(_torch.tensor(x ** y))
    return _apply(lambda a, b: _t(a) ** b, x, y)


def concat(values, axis, name=None):
    return _apply(lambda vs: _torch.cat(list(vs), dim=axis), list(values))


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None):
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return _apply(lambda a: a.sum() if ax is None else a.sum(dim=ax, keepdim=keepdims), input_tensor)


def reduce_mean(input_tensor, axis=None, keepdims=False, name=None):
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return _apply(lambda a: a.mean() if ax is None else a.mean(dim=ax, keepdim=keepdims), input_tensor)


def expand_dims(input, axis, name=None):  # noqa: A002
    return _apply(lambda a: a.unsqueeze(axis), input)


def reshape(tensor, shape, name=None):
    shp = _shape_list(shape)
    return _apply(lambda a: a.reshape(shp), tensor)


def transpose(a, perm=None, name=None):
    return _apply(lambda t: t.permute(*perm) if perm is not None else t.t(), a)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    def f(x, y):
        if transpose_a:
            x = x.transpose(-1, -2)
        if transpose_b:
            y = y.transpose(-1, -2)
        return x @ y
    return _apply(f, a, b)


def shape(input, name=None):  # noqa: A002
    return Tensor(_torch.tensor(list(_as_torch(input).shape), dtype=_torch.int64))


def zeros(shape, dtype=None, name=None):  # noqa: A002
    return Tensor(_torch.zeros(_shape_list(shape), dtype=_torch_dtype(dtype) or _DT))


def ones(shape, dtype=None, name=None):  # noqa: A002
    return Tensor(_torch.ones(_shape_list(shape), dtype=_torch_dtype(dtype) or _DT))


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return _apply(lambda a: _torch.minimum(_torch.maximum(a, _t(float(clip_value_min))), _t(float(clip_value_max))), t)


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
    ts = [_as_torch(t) for t in t_list]
    norm = _torch.sqrt(sum((t * t).sum() for t in ts if t is not None))
    scale = clip_norm * _torch.minimum(1.0 / norm, _torch.tensor(1.0 / clip_norm, dtype=norm.dtype))
    return [None if t is None else Tensor(t * scale) for t in ts], Tensor(norm)


def function(func=None, **kwargs):
    """tf.function / tf.function(jit_compile=True): run the python eagerly."""
    if func is None:
        return lambda f: f
    return func


def print(*args, **kwargs):  # noqa: A001
    pass


class name_scope:
    def __init__(self, name=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class Module:
    def __init__(self, name=None):
        self._name = name

    @property
    def name(self):
        return self._name


# ---- tf.math / tf.nn / tf.random / tf.summary / tf.train / tf.distribute ------------------------------------------
math_ns = types.SimpleNamespace(
    multiply=multiply, add=add, subtract=subtract, divide=divide, reduce_mean=reduce_mean, reduce_sum=reduce_sum, pow=pow,
    log=lambda x, name=None: _apply(lambda a: _torch.log(a), x),
    exp=lambda x, name=None: _apply(lambda a: _torch.exp(a), x),
    sqrt=lambda x, name=None: _apply(lambda a: _torch.sqrt(_t(a).to(_DT)), x),
    argmax=lambda input, axis=None, name=None, output_type=None: _apply(lambda a: a.argmax(dim=axis), input),  # noqa: A002
    equal=equal,
)


def _nn_dropout(x, rate, noise_shape=None, seed=None, name=None):
    return _apply(lambda a: _dropout(a, rate), x)


def compute_average_loss(per_example_loss, sample_weight=None, global_batch_size=None):
    return _apply(lambda a: a.sum() / float(global_batch_size), per_example_loss)


nn = types.SimpleNamespace(
    relu=lambda x, name=None: _apply(lambda a: _torch.clamp(a, min=0.0), x),
    elu=lambda x, name=None: _apply(_elu, x),
    gelu=lambda x, approximate=False, name=None: _apply(lambda a: _gelu(a, approximate), x),
    softmax=lambda x, axis=-1, name=None: _apply(lambda a: _softmax(a, axis), x),
    sigmoid=lambda x, name=None: _apply(_sigmoid, x),
    dropout=_nn_dropout,
    compute_average_loss=compute_average_loss,
    conv2d=lambda input, filters, strides=1, padding="SAME", dilations=1, name=None: _apply(  # noqa: A002
        lambda a, k: _conv2d(a, k, _pair(strides), padding, _pair(dilations)), input, filters),
)


def _pair(v):
    if isinstance(v, (list, tuple)):
        return (int(v[0]), int(v[1])) if len(v) == 2 else (int(v[1]), int(v[2]))
    return (int(v), int(v))


random = types.SimpleNamespace(
    normal=lambda shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None:
        Tensor(_torch.randn(_shape_list(shape), generator=_rng, dtype=_DT) * stddev + mean),
    set_seed=lambda s: shim_seed(int(s)),
)


class _NullWriter:
    def as_default(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def flush(self):
        pass


summary = types.SimpleNamespace(create_file_writer=lambda *a, **k: _NullWriter(), scalar=lambda *a, **k: None,
                                image=lambda *a, **k: None)


class _Checkpoint:
    def __init__(self, **kw):
        self.items = kw

    def restore(self, path):
        return types.SimpleNamespace(expect_partial=lambda: None)

    def save(self, *a, **k):
        return None


train = types.SimpleNamespace(Checkpoint=_Checkpoint, latest_checkpoint=lambda d: None,
                              CheckpointManager=lambda ckpt, directory=None, max_to_keep=None: types.SimpleNamespace(save=lambda: None))
compat = types.SimpleNamespace(v1=types.SimpleNamespace(enable_eager_execution=lambda: None))
config = types.SimpleNamespace(list_physical_devices=lambda kind=None: [],
                               experimental=types.SimpleNamespace(set_memory_growth=lambda *a, **k: None))


# ---- GradientTape ----------------------------------------------------------------------------------------------
shim_last_gradients: Dict[str, _torch.Tensor] = {}


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, tensors):
        pass

    def gradient(self, target, sources, **kw):
        """d(sum(target))/d(sources) -- TF sums a non-scalar target."""
        src = list(sources)
        gs = _torch.autograd.grad(_as_torch(target).sum(), [s._t for s in src], allow_unused=True)
        shim_last_gradients.clear()
        for s, g in zip(src, gs):
            if g is not None:
                shim_last_gradients[s.name] = g.detach().clone()
        return [None if g is None else Tensor(g.detach()) for g in gs]


# ------------------------------------------------------------------------------------------------------------------
# tf.keras
# ------------------------------------------------------------------------------------------------------------------
_name_uids: Dict[str, int] = {}


def shim_reset_names():
    """keras.backend.clear_session(): auto-name counters restart."""
    _name_uids.clear()


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _unique_name(base):
    n = _name_uids.get(base, 0)
    _name_uids[base] = n + 1
    return base if n == 0 else f"{base}_{n}"


class _GlorotUniform:
    def __call__(self, shape, dtype=None):
        fan_in, fan_out = _fans(shape)
        limit = _pymath.sqrt(6.0 / (fan_in + fan_out))
        return (_torch.rand(tuple(shape), generator=_rng, dtype=_DT) * 2 - 1) * limit


class HeNormal:
    def __init__(self, seed=None):
        pass

    def __call__(self, shape, dtype=None):
        fan_in, _ = _fans(shape)
        sigma = _pymath.sqrt(2.0 / fan_in) / 0.87962566103423978
        t = _torch.empty(tuple(shape), dtype=_DT)
        _torch.nn.init.trunc_normal_(t, 0.0, sigma, -2 * sigma, 2 * sigma, generator=_rng)
        return t


def _fans(shape):
    shape = tuple(shape)
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for d in shape[:-2]:
        rf *= d
    return shape[-2] * rf, shape[-1] * rf


def _get_initializer(x):
    if x is None or (isinstance(x, str) and x.lower() == "glorot_uniform"):
        return _GlorotUniform()
    if isinstance(x, str) and x.lower().replace("_", "") == "henormal":
        return HeNormal()
    if isinstance(x, str) and x.lower() == "zeros":
        return lambda shape, dtype=None: _torch.zeros(tuple(shape), dtype=_DT)
    if isinstance(x, str) and x.lower() == "ones":
        return lambda shape, dtype=None: _torch.ones(tuple(shape), dtype=_DT)
    return x


class Layer(Module):
    _created = 0

    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        self._name = name if name is not None else _unique_name(_snake(type(self).__name__))
        self.trainable = trainable
        self.built = False
        self._weights: List[Variable] = []
        Layer._created += 1
        self._creation_index = Layer._created
        self.dtype = "float32"

    @property
    def name(self):
        return self._name

    def add_weight(self, name=None, shape=None, initializer=None, regularizer=None, trainable=True, dtype=None, **kw):
        v = Variable(_get_initializer(initializer)(shape), trainable=trainable, name=f"{self._name}/{name}")
        self._weights.append(v)
        return v

    @property
    def variables(self):
        out = list(self._weights)
        for sub in getattr(self, "_sublayers", lambda: [])():
            out += sub.variables
        return out

    weights = variables

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable]

    def build(self, input_shape):
        pass

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            first = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
            self.build(TensorShape(_as_torch(first).shape))
            self.built = True
        kwargs.pop("training", None)
        if hasattr(self, "_fn"):
            return _apply(self._fn, inputs, _layer=self)
        return self.call(inputs, *args, **kwargs)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", dilation_rate=(1, 1), activation=None,
                 use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros", kernel_regularizer=None,
                 name=None, **kw):
        super().__init__(name=name)
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.dilation_rate, self.use_bias = padding.upper(), _pair(dilation_rate), use_bias
        self.activation = _ACTIVATIONS[activation] if (activation is None or isinstance(activation, str)) else activation
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", self.kernel_size + (int(input_shape[-1]), self.filters), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.filters,), self.bias_initializer) if self.use_bias else None

    def _fn(self, x):
        y = _conv2d(x, self.kernel._t, self.strides, self.padding, self.dilation_rate)
        if self.bias is not None:
            y = y + self.bias._t
        return self.activation(y) if self.activation is not None else y


class Conv2DTranspose(Conv2D):
    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", self.kernel_size + (self.filters, int(input_shape[-1])), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.filters,), self.bias_initializer) if self.use_bias else None

    def _fn(self, x):
        y = _conv2d_transpose(x, self.kernel._t, self.strides, self.padding)
        if self.bias is not None:
            y = y + self.bias._t
        return self.activation(y) if self.activation is not None else y


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, name=None, **kw):
        super().__init__(name=name)
        self.units, self.use_bias = int(units), use_bias
        self.activation = _ACTIVATIONS[activation] if (activation is None or isinstance(activation, str)) else activation
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.units,), self.bias_initializer) if self.use_bias else None

    def _fn(self, x):
        y = x @ self.kernel._t
        if self.bias is not None:
            y = y + self.bias._t
        return self.activation(y) if self.activation is not None else y


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None, **kw):
        super().__init__(name=name)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = (int(input_shape[-1]),)
        self.gamma = self.add_weight("gamma", c, "ones")
        self.beta = self.add_weight("beta", c, "zeros")
        self.moving_mean = self.add_weight("moving_mean", c, "zeros", trainable=False)
        self.moving_variance = self.add_weight("moving_variance", c, "ones", trainable=False)

    def _fn(self, x):
        """inference form (no ``training=True`` anywhere on the reference's path): tf.nn.batch_normalization"""
        inv = self.gamma._t / _torch.sqrt(self.moving_variance._t + self.epsilon)
        return x * inv + (self.beta._t - self.moving_mean._t * inv)


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, name=None, **kw):
        super().__init__(name=name)
        assert axis == -1
        self.epsilon = epsilon

    def build(self, input_shape):
        c = (int(input_shape[-1]),)
        self.gamma = self.add_weight("gamma", c, "ones")
        self.beta = self.add_weight("beta", c, "zeros")

    def _fn(self, x):
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return (x - mean) * _torch.rsqrt(var + self.epsilon) * self.gamma._t + self.beta._t


class _Stateless(Layer):
    pass


class ELU(_Stateless):
    def __init__(self, alpha=1.0, name=None, **kw):
        super().__init__(name=name)
        self.alpha = alpha

    def _fn(self, x):
        return _elu(x, self.alpha)


class LeakyReLU(_Stateless):
    def __init__(self, alpha=0.3, name=None, **kw):
        super().__init__(name=name)
        self.alpha = alpha

    def _fn(self, x):
        return _leaky(x, self.alpha)


class ReLU(_Stateless):
    def __init__(self, name=None, **kw):
        super().__init__(name=name)

    def _fn(self, x):
        return _torch.clamp(x, min=0.0)


class Softmax(_Stateless):
    def __init__(self, axis=-1, name=None, **kw):
        super().__init__(name=name)
        self.axis = axis

    def _fn(self, x):
        return _softmax(x, self.axis)


class Dropout(_Stateless):
    """Keras Dropout is the identity unless called with training=True; the reference never does (and uses rate 0)."""
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name)
        self.rate = rate

    def _fn(self, x):
        return x


class AveragePooling2D(_Stateless):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kw):
        super().__init__(name=name)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size
        assert padding.lower() == "valid"

    def _fn(self, x):
        return _avg_pool(x, self.pool_size, self.strides)


class UpSampling2D(_Stateless):
    def __init__(self, size=(2, 2), name=None, **kw):
        super().__init__(name=name)
        self.size = _pair(size)

    def _fn(self, x):
        return x.repeat_interleave(self.size[0], dim=1).repeat_interleave(self.size[1], dim=2)


def Input(shape=None, batch_size=None, name=None, dtype=None, **kw):  # noqa: N802
    ex = _torch.zeros([int(batch_size) if batch_size else 1] + [int(d) for d in shape], dtype=_DT)
    return Tensor(ex, _Node("input", (), {}))


class Model:
    """tf.keras.Model(inputs, outputs) of the functional API: re-evaluates the recorded nodes on a new input."""

    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs = inputs, outputs
        self._layers: List[Layer] = []
        seen = set()

        def walk(t):
            if isinstance(t, (list, tuple)):
                for e in t:
                    walk(e)
                return
            if not isinstance(t, Tensor) or t._node is None or id(t._node) in seen:
                return
            seen.add(id(t._node))
            if t._node.fn == "input":
                return
            walk(t._node.args)
            walk(tuple(t._node.kwargs.values()))
            if t._node.layer is not None and t._node.layer not in self._layers:
                self._layers.append(t._node.layer)
        walk(outputs)
        # listed in layer CREATION order.  (Real Keras sorts a functional model's layers by graph depth; nothing on the
        # reference's path depends on the order -- gradients and variables are zipped from the same list.)
        self._layers.sort(key=lambda l: l._creation_index)

    @property
    def layers(self):
        return list(self._layers)

    @property
    def variables(self):
        return [v for l in self._layers for v in l._weights]

    weights = variables

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable]

    @property
    def losses(self):
        return []

    def compile(self, *a, **k):
        pass

    def summary(self, *a, **k):
        pass

    def save(self, *a, **k):
        pass

    def __call__(self, x, training=False, **kw):
        ins = self.inputs if isinstance(self.inputs, (list, tuple)) else [self.inputs]
        xs = x if isinstance(x, (list, tuple)) else [x]
        memo = {id(i._node): (_as_torch(convert_to_tensor(v)),) for i, v in zip(ins, xs)}

        def ev(t):
            if isinstance(t, (list, tuple)):
                return type(t)(ev(e) for e in t)
            if not isinstance(t, Tensor) or t._node is None:
                return _as_torch(t)
            nid = id(t._node)
            if nid not in memo:
                nd = t._node
                out = nd.fn(*[ev(a) for a in nd.args], **{k: ev(v) for k, v in nd.kwargs.items()})
                memo[nid] = out if isinstance(out, (list, tuple)) else (out,)
            return memo[nid][t._index or 0]

        def wrap(t):
            if isinstance(t, (list, tuple)):
                return type(t)(wrap(e) for e in t)
            return Tensor(ev(t)) if isinstance(t, Tensor) else t
        return wrap(self.outputs)

    predict = __call__


class _CategoricalCrossentropy:
    def __init__(self, from_logits=False, label_smoothing=0.0, axis=-1, reduction="auto", name=None):
        self.from_logits, self.label_smoothing, self.reduction = from_logits, label_smoothing, reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        ls = self.label_smoothing

        def f(yt, yp):
            yt = yt.to(_DT)
            if ls:
                yt = yt * (1.0 - ls) + ls / yt.shape[-1]
            if self.from_logits:
                logp = yp - _torch.logsumexp(yp, dim=-1, keepdim=True)
            else:
                yp = yp / yp.sum(dim=-1, keepdim=True)
                yp = _torch.clamp(yp, 1e-7, 1.0 - 1e-7)
                logp = _torch.log(yp)
            per = -(yt * logp).sum(dim=-1)
            if self.reduction in ("none", None):
                return per
            return per.mean()
        return _apply(f, y_true, y_pred)


class _Metric:
    def __init__(self, *a, **k):
        pass

    def update_state(self, *a, **k):
        pass

    def result(self):
        return Tensor(_torch.tensor(0.0, dtype=_DT))

    def reset_states(self):
        pass

    reset_state = reset_states


class _Adam:
    """Keras optimizer_v2.Adam, dense update, amsgrad off."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kw):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._m: Dict[int, _torch.Tensor] = {}
        self._v: Dict[int, _torch.Tensor] = {}

    lr = property(lambda self: self.learning_rate)

    def apply_gradients(self, grads_and_vars, **kw):
        self.iterations += 1
        t = self.iterations
        lr = float(self.learning_rate() if callable(self.learning_rate) else self.learning_rate)
        lr_t = lr * _pymath.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        with _torch.no_grad():
            for g, var in grads_and_vars:
                if g is None:
                    continue
                g = _as_torch(g)
                m = self._m.setdefault(id(var), _torch.zeros_like(var._t))
                v = self._v.setdefault(id(var), _torch.zeros_like(var._t))
                m.mul_(self.beta_1).add_(g * (1.0 - self.beta_1))
                v.mul_(self.beta_2).add_(g * g * (1.0 - self.beta_2))
                var._t.sub_(lr_t * m / (_torch.sqrt(v) + self.epsilon))


optimizers = types.SimpleNamespace(Adam=_Adam, Adamax=_Adam)

keras = types.SimpleNamespace(
    layers=types.SimpleNamespace(
        Layer=Layer, Input=Input, Conv2D=Conv2D, Conv2DTranspose=Conv2DTranspose, Dense=Dense, ELU=ELU, LeakyReLU=LeakyReLU,
        ReLU=ReLU, Softmax=Softmax, Dropout=Dropout, BatchNormalization=BatchNormalization,
        LayerNormalization=LayerNormalization, AveragePooling2D=AveragePooling2D, UpSampling2D=UpSampling2D,
        experimental=types.SimpleNamespace(SyncBatchNormalization=BatchNormalization)),
    activations=types.SimpleNamespace(
        softmax=lambda x, axis=-1: _apply(lambda a: _softmax(a, axis), x),
        sigmoid=lambda x: _apply(_sigmoid, x),
        gelu=lambda x, approximate=False: _apply(lambda a: _gelu(a, approximate), x),
        relu=lambda x: _apply(lambda a: _torch.clamp(a, min=0.0), x)),
    initializers=types.SimpleNamespace(HeNormal=HeNormal, GlorotUniform=_GlorotUniform, get=_get_initializer),
    regularizers=types.SimpleNamespace(get=lambda x: x, l2=lambda *a, **k: None),
    losses=types.SimpleNamespace(CategoricalCrossentropy=_CategoricalCrossentropy,
                                 Reduction=types.SimpleNamespace(NONE="none", SUM="sum", AUTO="auto")),
    metrics=types.SimpleNamespace(Precision=_Metric, Recall=_Metric, Mean=_Metric, CategoricalAccuracy=_Metric),
    optimizers=types.SimpleNamespace(Adam=_Adam, Adamax=_Adam,
                                     schedules=types.SimpleNamespace(PiecewiseConstantDecay=lambda b, v: (lambda: v[0]))),
    utils=types.SimpleNamespace(plot_model=lambda *a, **k: None),
    backend=types.SimpleNamespace(clear_session=shim_reset_names),
    Model=Model, Input=Input,
)

math = math_ns            # tf.math (the python module is _pymath in this file)
sys.modules[__name__ + ".keras"] = keras  # type: ignore[assignment]


def __getattr__(name):
    if name == "bool":
        return bool_
    raise AttributeError(name)


# ``from tensorflow.python.keras.utils import conv_utils`` (Decoder.py:3)
def _normalize_tuple(value, n, name):
    return tuple(value) if isinstance(value, (list, tuple)) else (int(value),) * n


for _modname, _attrs in (
        ("tensorflow.python", {}), ("tensorflow.python.keras", {}), ("tensorflow.python.keras.utils", {}),
        ("tensorflow.python.keras.utils.conv_utils", {"normalize_tuple": _normalize_tuple})):
    _m = types.ModuleType(_modname)
    _m.__dict__.update(_attrs)
    sys.modules.setdefault(_modname, _m)
sys.modules["tensorflow.python.keras.utils"].conv_utils = sys.modules["tensorflow.python.keras.utils.conv_utils"]
sys.modules["tensorflow.python.keras"].utils = sys.modules["tensorflow.python.keras.utils"]
sys.modules["tensorflow.python"].keras = sys.modules["tensorflow.python.keras"]
