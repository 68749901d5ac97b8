"""
TEST INFRASTRUCTURE -- a numpy-backed stand-in for the ``tensorflow`` module, just large
enough to EXECUTE THE UNMODIFIED REFERENCE (N3PDF/vegasflow v1.4.0, /root/reference) on a
machine without TensorFlow.  It exists to pin the oracle and the CUDA path to the reference's
own code: ``tests/golden/make_golden_from_reference.py`` puts ``/root/reference/src`` and
``/root/reference/examples`` on ``sys.path`` together with this directory and dumps what the
reference computes; ``tests/test_reference_pinned.py`` runs the reference's own unit test on it.

NEVER importable from ``vegasflow_b200`` (the product package has no tensorflow dependency and
no CPU path); only ``tests/`` put this directory on ``sys.path``.

What a ``tf.*`` symbol means here
---------------------------------
Every symbol the reference touches on the hot path (src/vegasflow/{configflow,utils,vflow,
vflowplus,plain,monte_carlo}.py, examples/{simgauss,drellyan_lo,singletop_lo}_tf.py) is mapped
onto the numpy operation with the same IEEE semantics: elementwise ops are one correctly
rounded fp64 operation each (numpy never contracts to FMA), casts truncate, gathers index.
Where TensorFlow/Eigen leaves the result build-dependent, the convention of SURVEY.md 8(c) is
adopted and is the ONLY interpretive choice made here:

  * ``reduce_sum`` / ``reduce_prod`` / ``segment_sum`` accumulate sequentially, left to right
    along the reduced axis (Eigen vectorises inner reductions in a packet-width dependent tree);
  * transcendental functions are numpy's (glibc / numpy SIMD), like the C oracle's;
  * ``tf.random.uniform`` draws from a seedable numpy generator with TensorFlow's affine map
    ``u * (maxval - minval) + minval``, or -- the parity seam -- returns arrays queued with
    ``tf.random.feed(array)`` verbatim ("the reference's own uniform draws").

Python scalars follow TensorFlow's conversion rules where it matters: a bare python float
becomes float32 and a bare int int32 when no tensor operand fixes the dtype (this is what
makes ``tf.math.pow(neval_eff / 2, 1 / n_dim)`` a float32 computation, vflowplus.py:113-123).
"""
import builtins as _b
import contextlib
import logging as _logging
import types as _types

import numpy as _np

__version__ = "2.16.1-numpy-shim"

# ---------------------------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------------------------
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
complex128 = _np.complex128
bool = _np.bool_  # noqa: A001  (tf.bool)

_builtin_bool = _b.bool


class Tensor(_np.ndarray):
    """ndarray with the handful of tf.Tensor / tf.Variable methods the reference calls."""

    def __new__(cls, value, dtype=None):
        return _np.asarray(value, dtype=dtype).view(cls)

    # keep 0-d results as tensors (numpy would hand back scalars, which have no .numpy())
    def __array_wrap__(self, array, context=None, return_scalar=False):
        return _np.asarray(array).view(type(self))

    def __getitem__(self, item):
        out = _np.ndarray.__getitem__(self, item)
        if not isinstance(out, _np.ndarray):
            out = _np.asarray(out).view(type(self))
        return out

    def numpy(self):
        arr = _np.array(self.view(_np.ndarray))
        return arr[()] if arr.ndim == 0 else arr

    # tf.Variable.assign, also on slices (``var[j, :].assign(v)`` writes through the view)
    def assign(self, value):
        self[...] = _np.asarray(value, dtype=self.dtype)
        return self

    # tf.Tensor is immutable: ``a += b`` rebinds the name, it never writes into the operand
    # (matters e.g. for ``p3 -= p0 * k``, singletop_lo_tf.py:239, whose p3 the caller keeps)
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def __ipow__(self, other):
        return self ** other

    def __hash__(self):
        return id(self)

    def __bool__(self):
        return _builtin_bool(_np.ndarray.__bool__(self.view(_np.ndarray)))


def _t(value, dtype=None):
    return Tensor(value, dtype=dtype)


def _np_of(value):
    return value.view(_np.ndarray) if isinstance(value, Tensor) else value


def _is_typed(v):
    return isinstance(v, (_np.ndarray, _np.generic))


def _default_dtype(v):
    """tf.convert_to_tensor on python objects: float -> float32, int -> int32."""
    probe = _np.asarray(v)
    if probe.dtype == _np.float64:
        return _np.float32
    if probe.dtype == _np.int64:
        return _np.int32
    return probe.dtype


def _conv(v, hint=None):
    """Convert like tf.convert_to_tensor: typed values keep their dtype, python objects take
    `hint` (the dtype of a tensor operand) or TensorFlow's defaults."""
    if _is_typed(v):
        return _np.asarray(v)
    if isinstance(v, (list, tuple)) and any(_is_typed(e) for e in v):
        return _np.asarray([_np.asarray(e) for e in v])
    return _np.asarray(v, dtype=hint if hint is not None else _default_dtype(v))


def _pair(x, y):
    hint = None
    for v in (x, y):
        if _is_typed(v):
            hint = _np.asarray(v).dtype
            break
    return _conv(x, hint), _conv(y, hint)


def convert_to_tensor(value, dtype=None, **_):
    if dtype is not None:
        return _t(_np.asarray(_np_of(value)), dtype=dtype)
    return _t(_conv(value))


def constant(value, dtype=None, shape=None, **_):
    out = convert_to_tensor(value, dtype=dtype)
    if shape is not None:
        out = _t(_np.broadcast_to(out, shape).copy())
    return out


def cast(x, dtype, **_):
    """Float -> int casts truncate toward zero (C semantics), like tf.cast."""
    arr = _np.asarray(_np_of(x))
    return _t(arr.astype(dtype))


def Variable(initial_value, dtype=None, **_):  # noqa: N802
    return _t(_np.array(_np_of(initial_value), dtype=dtype, copy=True))


class TensorSpec:
    def __init__(self, shape=None, dtype=float32, name=None):
        self.shape, self.dtype, self.name = shape, dtype, name

    def __repr__(self):
        return f"TensorSpec(shape={self.shape}, dtype={_np.dtype(self.dtype).name})"


# ---------------------------------------------------------------------------------------------
# tf.function: eager pass-through that honours input_signature (dtype conversion of the
# positional arguments) and carries the attributes monte_carlo.py:542-545 inspects.
# ---------------------------------------------------------------------------------------------
class Function:
    def __init__(self, python_function, input_signature=None):
        self.python_function = python_function
        self.input_signature = input_signature
        self.function_spec = _types.SimpleNamespace(input_signature=input_signature)
        self.__name__ = getattr(python_function, "__name__", "function")
        self.__doc__ = getattr(python_function, "__doc__", None)
        self.__wrapped__ = python_function

    def __call__(self, *args, **kwargs):
        if self.input_signature:
            args = list(args)
            for k, spec in enumerate(self.input_signature):
                if k < len(args) and args[k] is not None:
                    args[k] = _t(_np.asarray(_np_of(args[k])), dtype=spec.dtype)
        return self.python_function(*args, **kwargs)

    def __get__(self, obj, objtype=None):  # usable as a method decorator
        if obj is None:
            return self
        return _types.MethodType(self, obj)


def function(func=None, input_signature=None, **_):
    if func is None:
        return lambda f: Function(f, input_signature)
    return Function(func, input_signature)


# ---------------------------------------------------------------------------------------------
# elementwise
# ---------------------------------------------------------------------------------------------
def _unary(np_func):
    def op(x, name=None):
        return _t(np_func(_conv(x)))

    return op


square = _unary(_np.square)
sqrt = _unary(_np.sqrt)
exp = _unary(_np.exp)
sin = _unary(_np.sin)
cos = _unary(_np.cos)
sinh = _unary(_np.sinh)
cosh = _unary(_np.cosh)
acos = _unary(_np.arccos)
acosh = _unary(_np.arccosh)
abs = _unary(_np.abs)  # noqa: A001  (complex -> hypot, like std::abs)
_log = _unary(_np.log)
_floor = _unary(_np.floor)
_real = _unary(_np.real)


def pow(x, y, name=None):  # noqa: A001
    a, b = _pair(x, y)
    return _t(_np.power(a, b))


def maximum(x, y, name=None):
    a, b = _pair(x, y)
    return _t(_np.maximum(a, b))


def equal(x, y, name=None):
    a, b = _pair(x, y)
    return _t(_np.equal(a, b))


def complex(real, imag, name=None):  # noqa: A001
    a, b = _pair(real, imag)
    out = _np.empty(_np.broadcast(a, b).shape, dtype=_np.complex128)
    out.real = a
    out.imag = b
    return _t(out)


def where(condition, x=None, y=None, name=None):
    cond = _np.asarray(_np_of(condition))
    if x is None and y is None:
        return _t(_np.argwhere(cond).astype(_np.int64))
    a, b = _pair(x, y)
    return _t(_np.where(cond, a, b))


def zeros_like(x, dtype=None, **_):
    return _t(_np.zeros_like(_np.asarray(_np_of(x)), dtype=dtype))


def ones_like(x, dtype=None, **_):
    return _t(_np.ones_like(_np.asarray(_np_of(x)), dtype=dtype))


def fill(dims, value, name=None):
    return _t(_np.full(tuple(int(d) for d in dims), _conv(value)))


# ---------------------------------------------------------------------------------------------
# reductions -- sequential, left to right along the reduced axis (module docstring)
# ---------------------------------------------------------------------------------------------
def _sequential_reduce(arr, axis, op):
    arr = _np.asarray(arr)
    if axis is None:
        arr, axis = arr.reshape(-1), 0
    if arr.shape[axis] == 0:
        ident = 0 if op is _np.add else 1
        return _np.full(arr.shape[:axis] + arr.shape[axis + 1:], ident, dtype=arr.dtype)
    moved = _np.moveaxis(arr, axis, 0)
    acc = _np.array(moved[0], copy=True)
    for k in _b.range(1, moved.shape[0]):
        acc = op(acc, moved[k])
    return acc


def _sequential_sum(arr, axis):
    """Left-to-right sum.  For a long reduced axis use cumsum, which numpy evaluates as a plain
    sequential recurrence (every prefix is an output), instead of a python loop."""
    arr = _np.asarray(arr)
    if axis is None:
        arr, axis = arr.reshape(-1), 0
    if arr.shape[axis] > 64 and arr.dtype.kind in "fc":
        return _np.take(_np.cumsum(arr, axis=axis), -1, axis=axis)
    return _sequential_reduce(arr, axis, _np.add)


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None):
    arr = _conv(input_tensor)
    out = _sequential_sum(arr, axis)
    if keepdims and axis is not None:
        out = _np.expand_dims(out, axis)
    return _t(out)


def reduce_prod(input_tensor, axis=None, keepdims=False, name=None):
    arr = _conv(input_tensor)
    out = _sequential_reduce(arr, axis, _np.multiply)
    if keepdims and axis is not None:
        out = _np.expand_dims(out, axis)
    return _t(out)


def _segment_sum(data, segment_ids, name=None):
    data = _np.asarray(_np_of(data))
    ids = _np.asarray(_np_of(segment_ids)).astype(_np.int64)
    n_seg = int(ids[-1]) + 1 if ids.size else 0
    out = _np.zeros((n_seg,) + data.shape[1:], dtype=data.dtype)
    _np.add.at(out, ids, data)  # unbuffered: adds in index order, i.e. sequentially per segment
    return _t(out)


def _reduce_all(x, axis=None, **_):
    return _t(_np.all(_conv(x), axis=axis))


def _reduce_any(x, axis=None, **_):
    return _t(_np.any(_conv(x), axis=axis))


# ---------------------------------------------------------------------------------------------
# shape / indexing
# ---------------------------------------------------------------------------------------------
def transpose(a, perm=None, **_):
    return _t(_np.transpose(_conv(a), perm))


def reshape(tensor, shape, name=None):
    if isinstance(tensor, (list, tuple)):
        tensor = _np.asarray([_np.asarray(_np_of(e)) for e in tensor])
    return _t(_np.reshape(_np.asarray(_np_of(tensor)), tuple(int(s) for s in _np.atleast_1d(shape))))


def stack(values, axis=0, name=None):
    arrs = [_conv(v) for v in values]
    hint = next((a.dtype for a, v in zip(arrs, values) if _is_typed(v)), None)
    if hint is not None:
        arrs = [a if _is_typed(v) else a.astype(hint) for a, v in zip(arrs, values)]
    return _t(_np.stack(arrs, axis=axis))


def concat(values, axis, name=None):
    return _t(_np.concatenate([_conv(v) for v in values], axis=axis))


def gather(params, indices, axis=None, batch_dims=0, name=None):
    p = _np.asarray(_np_of(params))
    idx = _np.asarray(_np_of(indices))
    if batch_dims == 0:
        return _t(_np.take(p, idx, axis=0 if axis is None else axis))
    if batch_dims == 1 and p.ndim == 2 and idx.ndim == 2:
        # out[b, k] = params[b, indices[b, k]]  (vflow.py:70-71)
        return _t(_np.take_along_axis(p, idx.astype(_np.int64), axis=1))
    raise NotImplementedError("gather: only batch_dims 0, or 1 on matrices")


def pad(tensor, paddings, mode="CONSTANT", constant_values=0, name=None):
    pads = [tuple(int(v) for v in row) for row in _np.asarray(_np_of(paddings))]
    arr = _np.asarray(_np_of(tensor))
    return _t(_np.pad(arr, pads, mode="constant", constant_values=constant_values))


def range(start, limit=None, delta=1, dtype=None, name=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    vals = [_np.asarray(_np_of(v)) if _is_typed(v) else v for v in (start, limit, delta)]
    if dtype is None:
        typed = [_np.asarray(v).dtype for v in vals if _is_typed(v)]
        dtype = typed[0] if typed else _default_dtype(limit)
    return _t(_np.arange(vals[0], vals[1], vals[2]).astype(dtype))


def repeat(input, repeats, axis=None, name=None):  # noqa: A002
    return _t(_np.repeat(_np.asarray(_np_of(input)), _np.asarray(_np_of(repeats)), axis=axis))


def shape(input, out_type=int32, name=None):  # noqa: A002
    return _t(_np.asarray(_np.shape(_np_of(input)), dtype=out_type))


def while_loop(cond, body, loop_vars, parallel_iterations=10, **_):
    state = tuple(loop_vars)
    while _builtin_bool(cond(*state)):
        state = tuple(body(*state))
    return state


# ---------------------------------------------------------------------------------------------
# tf.math, tf.random, tf.config, tf.autograph, misc
# ---------------------------------------------------------------------------------------------
math = _types.SimpleNamespace(
    pow=pow, log=_log, floor=_floor, real=_real, exp=exp, sqrt=sqrt, square=square, abs=abs,
    segment_sum=_segment_sum, reduce_sum=reduce_sum, reduce_prod=reduce_prod,
    reduce_all=_reduce_all, reduce_any=_reduce_any, maximum=maximum,
    logical_and=lambda x, y, name=None: _t(_np.logical_and(_conv(x), _conv(y))),
    logical_or=lambda x, y, name=None: _t(_np.logical_or(_conv(x), _conv(y))),
)


class _Random:
    """tf.random with a seedable numpy generator and a feed queue (the parity seam)."""

    def __init__(self):
        self._rng = _np.random.default_rng(0)
        self._queue = []
        self.log = []  # every array handed out, newest last (read by the golden script)

    def set_seed(self, seed):
        self._rng = _np.random.default_rng(int(seed))

    def feed(self, array):
        """Queue `array`: the next tf.random.uniform call of the same shape returns it verbatim."""
        self._queue.append(_np.array(array, dtype=_np.float64, copy=True))

    def uniform(self, shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
        shape = tuple(int(s) for s in _np.atleast_1d(_np.asarray(_np_of(shape))))
        if self._queue and self._queue[0].shape == shape:
            out = self._queue.pop(0).astype(dtype)
        else:
            if maxval is None:
                maxval = 1
            u = self._rng.random(shape).astype(dtype)
            lo = _np.asarray(minval, dtype=dtype)
            hi = _np.asarray(maxval, dtype=dtype)
            out = u * (hi - lo) + lo  # python/ops/random_ops.py: rnd * (maxval - minval) + minval
        self.log.append(out)
        del self.log[:-4]
        return _t(out)


random = _Random()


class _LogicalDevice:
    def __init__(self, name, device_type):
        self.name, self.device_type = name, device_type


def _list_logical_devices(device_type=None):
    # a CPU-only TensorFlow reports no GPU: the reference then runs its sequential chunk loop
    # (monte_carlo.py:474-478)
    if device_type in (None, "CPU"):
        return [_LogicalDevice("/device:CPU:0", "CPU")]
    return []


config = _types.SimpleNamespace(
    run_functions_eagerly=lambda flag=True: None,
    experimental_run_functions_eagerly=lambda flag=True: None,
    list_logical_devices=_list_logical_devices,
    experimental=_types.SimpleNamespace(list_logical_devices=_list_logical_devices),
)
autograph = _types.SimpleNamespace(
    experimental=_types.SimpleNamespace(Feature=_types.SimpleNamespace(ALL="ALL")))


@contextlib.contextmanager
def device(name):
    yield


def get_logger():
    return _logging.getLogger("tensorflow")
