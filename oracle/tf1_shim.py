"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path.

An *eager* emulation of the small TensorFlow-1 surface that the reference
(khaotik/DaNet-Tensorflow, `/root/reference`) touches, backed by torch-CPU
tensors.  TensorFlow is not installable in this container, and
`app/hparams.py:9` imports it at module top, so without this shim none of the
reference's own Python can execute.  With it, `tests/golden/make_golden.py`
imports the UNMODIFIED reference modules (`main.Model`, `app.modules`,
`app.ops`, `app.utils`) and runs `Model.build()` on seeded inputs; the op
composition, axis orders and quirks are therefore the reference's own code, and
only the TF primitives (matmul, einsum, softmax, scan, ...) are restated here.

Semantics notes (what each primitive follows):
  * the graph is executed eagerly while `Model.build()` runs; `tf.placeholder`
    returns the array pre-bound in `FEEDS[name]`;
  * `tf.get_variable` re-uses an existing variable of the same scoped name
    (the reference's scan body is traced once in TF, but called T times here);
  * `tf.argmax/argmin` return the FIRST extremum on ties (Eigen CPU behaviour);
  * `AdamOptimizer.apply_gradients` follows TF1's formula
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).

This file is not a copy of any reference source: the reference contains no TF
implementation, only calls into it.
"""
import contextlib
import sys
import types

import numpy as np
import torch

FEEDS = {}            # placeholder name -> numpy array
VARIABLES = {}        # scoped name -> TT (insertion ordered)
GRADS = {}            # scoped name -> numpy gradient (filled by compute_gradients)
_SCOPE = []
_RNG = np.random.RandomState(1337)
WANT_GRADS = True
# dtype of get_variable() calls that pass none (TF: float32).  The golden generator
# sets float64 so that `anchors` (app/modules.py:503, no dtype) joins an fp64 graph.
DEFAULT_FLOAT = 'float32'


_LAYER_COUNT = {}


def reset(seed=1337):
    global _RNG
    FEEDS.clear()
    VARIABLES.clear()
    GRADS.clear()
    del _SCOPE[:]
    _LAYER_COUNT.clear()
    _RNG = np.random.RandomState(seed)


class DType(object):
    def __init__(self, name):
        self.name = name

    @property
    def is_complex(self):
        return self.name.startswith('complex')

    def __eq__(self, o):
        return _dtname(o) == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return 'tf.' + self.name


_TORCH = dict(
    float32=torch.float32, float64=torch.float64, int32=torch.int32,
    int64=torch.int64, complex64=torch.complex64, complex128=torch.complex128,
    bool=torch.bool)
_TORCH_INV = {v: k for k, v in _TORCH.items()}


def _dtname(d):
    if d is None:
        return None
    if isinstance(d, DType):
        return d.name
    if isinstance(d, str):
        return d
    if isinstance(d, torch.dtype):
        return _TORCH_INV[d]
    return np.dtype(d).name


def _td(d):
    return _TORCH[_dtname(d)]


class _Shape(object):
    def __init__(self, shp):
        self._s = [int(i) for i in shp]
        self.ndims = len(self._s)

    def as_list(self):
        return list(self._s)


class TT(object):
    """eager tensor: wraps a torch tensor, quacks like a TF1 Tensor/Variable"""
    __array_priority__ = 1000

    def __init__(self, v, name=None, trainable=False):
        self.v = v
        self.name = name
        self.trainable = trainable

    def get_shape(self):
        return _Shape(self.v.shape)

    @property
    def shape(self):
        return _Shape(self.v.shape)

    @property
    def dtype(self):
        return DType(_TORCH_INV[self.v.dtype])

    def numpy(self):
        return self.v.detach().cpu().numpy()

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        flips, new = [], []
        for ax, s in enumerate(idx):
            if isinstance(s, slice) and s.step is not None and s.step < 0:
                assert s.step == -1 and s.start is None and s.stop is None
                flips.append(ax)
                new.append(slice(None))
            else:
                new.append(_raw(s) if isinstance(s, TT) else s)
        out = self.v[tuple(new)]
        if flips:
            out = torch.flip(out, flips)
        return TT(out)

    def __len__(self):
        return self.v.shape[0]

    def __iter__(self):
        for i in range(self.v.shape[0]):
            yield TT(self.v[i])

    def __neg__(self):
        return TT(-self.v)

    def __add__(self, o): return _bin(torch.add, self, o)
    def __radd__(self, o): return _bin(torch.add, o, self)
    def __sub__(self, o): return _bin(torch.sub, self, o)
    def __rsub__(self, o): return _bin(torch.sub, o, self)
    def __mul__(self, o): return _bin(torch.mul, self, o)
    def __rmul__(self, o): return _bin(torch.mul, o, self)
    def __truediv__(self, o): return _bin(torch.div, self, o)
    def __rtruediv__(self, o): return _bin(torch.div, o, self)
    def __int__(self): return int(self.v)
    def __index__(self): return int(self.v)
    def __float__(self): return float(self.v)


def _raw(x, like=None):
    if isinstance(x, TT):
        return x.v
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(i, TT) for i in x):
        return torch.stack([_raw(i) for i in x])
    a = np.asarray(x)
    # python scalars stay float64/int64 until matched to the other operand
    t = torch.from_numpy(a.copy()) if a.ndim else torch.from_numpy(a.reshape(1).copy())[0]
    if like is not None and t.dtype.is_floating_point and like.dtype.is_floating_point:
        t = t.to(like.dtype)
    if like is not None and t.dtype.is_floating_point and like.dtype.is_complex:
        t = t.to(like.real.dtype)
    return t


def _bin(fn, a, b):
    ta = a.v if isinstance(a, TT) else None
    tb = b.v if isinstance(b, TT) else None
    if ta is None:
        ta = _raw(a, like=tb)
    if tb is None:
        tb = _raw(b, like=ta)
    return TT(fn(ta, tb))


def _ints(shape):
    if isinstance(shape, TT):
        return [int(i) for i in shape.v.reshape(-1).tolist()]
    return [int(i) for i in shape]


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


# --------------------------------------------------------------------------
# variables / scopes
# --------------------------------------------------------------------------
@contextlib.contextmanager
def variable_scope(name, regularizer=None, **kw):
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


name_scope = variable_scope


def _scoped(name):
    return '/'.join(_SCOPE + [name]) + ':0'


def constant_initializer(value=0., dtype=None, **kw):
    def init(shape, dt):
        a = np.asarray(value, dtype=np.float64)
        if a.ndim == 0:
            a = np.full(shape, float(a))
        return np.reshape(a, shape).astype(dt)
    return init


def random_uniform_initializer(minval=0., maxval=1., dtype=None, **kw):
    def init(shape, dt):
        return _RNG.uniform(minval, maxval, size=shape).astype(dt)
    return init


def random_normal_initializer(mean=0., stddev=1., **kw):
    def init(shape, dt):
        return (mean + stddev * _RNG.standard_normal(size=shape)).astype(dt)
    return init


def _glorot_uniform(shape, dt):
    fan_in, fan_out = shape[0], shape[-1]
    lim = np.sqrt(6. / (fan_in + fan_out))
    return _RNG.uniform(-lim, lim, size=shape).astype(dt)


def get_variable(name=None, shape=None, initializer=None, dtype=None,
                 trainable=True, **kw):
    full = _scoped(name)
    if full in VARIABLES:
        return VARIABLES[full]
    dt = _dtname(dtype) or DEFAULT_FLOAT
    init = initializer or _glorot_uniform
    val = init([int(s) for s in shape], dt)
    t = torch.from_numpy(np.ascontiguousarray(val))
    if trainable and WANT_GRADS:
        t.requires_grad_(True)
    v = TT(t, name=full, trainable=trainable)
    VARIABLES[full] = v
    return v


def Variable(initial_value, trainable=True, dtype=None, name='Variable'):
    full = _scoped(name)
    t = torch.tensor(initial_value, dtype=_td(dtype or 'float32'))
    v = TT(t, name=full, trainable=trainable)
    VARIABLES[full] = v
    return v


def trainable_variables():
    return [v for v in VARIABLES.values() if v.trainable]


def variables_initializer(var_list, **kw):
    return None


def global_variables_initializer():
    return None


def placeholder(dtype, shape=None, name=None):
    a = FEEDS[name]
    return TT(torch.from_numpy(np.ascontiguousarray(a)).to(_td(dtype)))


def assign(ref, value):
    ref.v = torch.tensor(float(value), dtype=ref.v.dtype)
    return None


# --------------------------------------------------------------------------
# math
# --------------------------------------------------------------------------
def _un(fn):
    return lambda x, name=None: TT(fn(_raw(x)))


sigmoid = _un(torch.sigmoid)
tanh = _un(torch.tanh)
cos = _un(torch.cos)
sin = _un(torch.sin)
log = _un(torch.log)
log1p = _un(torch.log1p)
square = _un(lambda t: t * t)
real = _un(torch.real)
imag = _un(torch.imag)
sqrt = _un(torch.sqrt)
exp = _un(torch.exp)


def abs_(x, name=None):
    return TT(torch.abs(_raw(x)))


def atan2(y, x, name=None):
    return TT(torch.atan2(_raw(y), _raw(x)))


def complex_(re, im, name=None):
    return TT(torch.complex(_raw(re), _raw(im)))


def maximum(a, b, name=None):
    return _bin(torch.maximum, a, b)


def less(a, b, name=None):
    return _bin(torch.lt, a, b)


def squared_difference(a, b, name=None):
    d = _raw(a) - _raw(b)
    return TT(d * d)


def clip_by_value(x, lo, hi, name=None):
    return TT(torch.clamp(_raw(x), float(lo), float(hi)))


def cast(x, dtype, name=None):
    return TT(_raw(x).to(_td(dtype)))


def ones_like(x, **kw):
    return TT(torch.ones_like(_raw(x)))


def zeros_like(x, **kw):
    return TT(torch.zeros_like(_raw(x)))


def constant(value, dtype=None, name=None, shape=None):
    t = _raw(value)
    if dtype is not None:
        t = t.to(_td(dtype))
    return TT(t)


def range_(*args, **kw):
    dtype = kw.get('dtype', 'int32')
    return TT(torch.arange(*[int(a) for a in args], dtype=_td(dtype)))


def shape_(x, **kw):
    return TT(torch.tensor(list(_raw(x).shape), dtype=torch.int64))


def _reduce(fn):
    def f(x, axis=None, keep_dims=False, keepdims=False, name=None):
        kd = bool(keep_dims or keepdims)
        t = _raw(x)
        ax = _axes(axis)
        if ax is None:
            ax = tuple(range(t.dim()))
        if isinstance(ax, tuple) and len(ax) == 0:
            return TT(t)
        return TT(fn(t, ax, kd))
    return f


reduce_sum = _reduce(lambda t, ax, kd: torch.sum(t, dim=ax, keepdim=kd))
reduce_mean = _reduce(lambda t, ax, kd: torch.mean(t, dim=ax, keepdim=kd))
reduce_max = _reduce(lambda t, ax, kd: torch.amax(t, dim=ax, keepdim=kd))


def reduce_prod(x, axis=None, **kw):
    t = _raw(x)
    return TT(torch.prod(t)) if axis is None else TT(torch.prod(t, dim=int(axis)))


def _first_arg(t, axis, largest):
    # numpy semantics: first extremum wins on ties
    a = t.detach().cpu().numpy()
    idx = np.argmax(a, axis=axis) if largest else np.argmin(a, axis=axis)
    return TT(torch.from_numpy(np.asarray(idx, dtype=np.int64)))


def argmax(x, axis=None, **kw):
    return _first_arg(_raw(x), int(axis), True)


def argmin(x, axis=None, **kw):
    return _first_arg(_raw(x), int(axis), False)


def matmul(a, b, name=None):
    return TT(torch.matmul(_raw(a), _raw(b)))


def einsum(eq, *ops):
    return TT(torch.einsum(eq, *[_raw(o) for o in ops]))


def tensordot(a, b, axes):
    return TT(torch.tensordot(_raw(a), _raw(b), dims=(list(axes[0]), list(axes[1]))))


def softmax(x, axis=-1, dim=None, name=None):
    return TT(torch.softmax(_raw(x), dim=(axis if dim is None else dim)))


def relu(x, name=None):
    return TT(torch.relu(_raw(x)))


def dropout(x, keep_prob=1., **kw):
    assert float(keep_prob) == 1., 'shim only supports keep_prob == 1 (SURVEY F5)'
    return x


# --------------------------------------------------------------------------
# shape ops
# --------------------------------------------------------------------------
def reshape(x, shape, name=None):
    return TT(_raw(x).reshape(_ints(shape)))


def transpose(x, perm=None, name=None):
    t = _raw(x)
    if perm is None:
        perm = list(range(t.dim()))[::-1]
    return TT(t.permute(*[int(p) for p in perm]))


def expand_dims(x, axis, name=None):
    return TT(_raw(x).unsqueeze(int(axis)))


def squeeze(x, axis=None, name=None):
    t = _raw(x)
    return TT(t.squeeze() if axis is None else t.squeeze(int(axis)))


def concat(values, axis, name=None):
    ts = [_raw(v) for v in values]
    if all(not t.dtype.is_floating_point and not t.dtype.is_complex for t in ts):
        ts = [t.to(torch.int64) for t in ts]
    return TT(torch.cat(ts, dim=int(axis)))


def stack(values, axis=0, name=None):
    ts = [_raw(v) for v in values]
    if all(not t.dtype.is_floating_point and not t.dtype.is_complex for t in ts):
        ts = [t.to(torch.int64) for t in ts]
    return TT(torch.stack(ts, dim=int(axis)))


def split(value, num_or_size_splits, axis=0, name=None):
    t = _raw(value)
    if isinstance(num_or_size_splits, int):
        size = t.shape[axis] // num_or_size_splits
        parts = torch.split(t, size, dim=axis)
    else:
        parts = torch.split(t, [int(s) for s in num_or_size_splits], dim=axis)
    return [TT(p) for p in parts]


def tile(x, multiples, name=None):
    return TT(_raw(x).repeat(*_ints(multiples)))


def gather(params, indices, name=None):
    return TT(_raw(params)[_raw(indices).to(torch.int64)])


def gather_nd(params, indices, name=None):
    p = _raw(params)
    idx = _raw(indices).to(torch.int64)
    k = idx.shape[-1]
    return TT(p[tuple(idx[..., i] for i in range(k))])


def one_hot(indices, depth, dtype='float32', **kw):
    idx = _raw(indices).to(torch.int64)
    return TT(torch.nn.functional.one_hot(idx, int(depth)).to(_td(dtype)))


def unsorted_segment_sum(data, segment_ids, num_segments, name=None):
    d = _raw(data)
    ids = _raw(segment_ids).to(torch.int64)
    out = torch.zeros((int(num_segments),) + tuple(d.shape[ids.dim():]), dtype=d.dtype)
    return TT(out.index_add(0, ids.reshape(-1), d.reshape((-1,) + tuple(d.shape[ids.dim():]))))


def map_fn(fn, elems, dtype=None, **kw):
    if isinstance(elems, (tuple, list)):
        n = _raw(elems[0]).shape[0]
        outs = [fn(tuple(TT(_raw(e)[i]) for e in elems)) for i in range(n)]
    else:
        n = _raw(elems).shape[0]
        outs = [fn(TT(_raw(elems)[i])) for i in range(n)]
    return TT(torch.stack([_raw(o) for o in outs]))


def scan(fn, elems, initializer=None, **kw):
    x = _raw(elems)
    acc = initializer
    outs = []
    for t in range(x.shape[0]):
        acc = fn(acc, TT(x[t]))
        outs.append(acc)
    if isinstance(acc, (tuple, list)):
        return tuple(TT(torch.stack([_raw(o[k]) for o in outs]))
                     for k in range(len(acc)))
    return TT(torch.stack([_raw(o) for o in outs]))


# --------------------------------------------------------------------------
# session / summaries / savers / optimizers
# --------------------------------------------------------------------------
def _np(x):
    if isinstance(x, TT):
        return x.numpy()
    if isinstance(x, dict):
        return {k: _np(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_np(v) for v in x]
    return x


class Session(object):
    graph = None

    def run(self, fetches, feed_dict=None):
        return _np(fetches)


class _Saver(object):
    def __init__(self, var_list=None, **kw):
        self.var_list = var_list

    def save(self, *a, **kw):
        pass

    def restore(self, *a, **kw):
        pass


class _Optimizer(object):
    def __init__(self, learning_rate=None, **kw):
        self.lr = learning_rate
        self.kw = kw

    def compute_gradients(self, loss, var_list=None):
        vs = var_list or trainable_variables()
        if not WANT_GRADS:
            return [(None, v) for v in vs]
        gs = torch.autograd.grad(_raw(loss), [v.v for v in vs], allow_unused=True)
        out = []
        for g, v in zip(gs, vs):
            if g is not None:
                GRADS[v.name] = g.detach().numpy().copy()
            out.append((TT(g) if g is not None else None, v))
        return out


class GradientDescentOptimizer(_Optimizer):
    def apply_gradients(self, grads_and_vars, **kw):
        lr = float(_raw(self.lr))
        self.updated = {v.name: (v.v.detach() - lr * _raw(g).detach()).numpy()
                        for g, v in grads_and_vars}
        LAST_OPTIMIZER[0] = self
        return None


class AdamOptimizer(_Optimizer):
    def apply_gradients(self, grads_and_vars, **kw):
        lr = float(_raw(self.lr))
        b1 = self.kw.get('beta1', 0.9)
        b2 = self.kw.get('beta2', 0.999)
        eps = self.kw.get('epsilon', 1e-8)
        t = 1
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        self.updated = {}
        self.clipped = {}
        for g, v in grads_and_vars:
            g_ = _raw(g).detach().to(torch.float64)
            m = (1 - b1) * g_
            vv = (1 - b2) * g_ * g_
            new = v.v.detach().to(torch.float64) - lr_t * m / (torch.sqrt(vv) + eps)
            self.updated[v.name] = new.to(v.v.dtype).numpy()
            self.clipped[v.name] = _raw(g).detach().numpy().copy()
        LAST_OPTIMIZER[0] = self
        return None


LAST_OPTIMIZER = [None]


# --------------------------------------------------------------------------
# tf.layers (used by the conv-bilstm-v1 encoder only, app/modules.py:263-379)
# --------------------------------------------------------------------------


def _layer_name(base):
    """tf.layers default naming inside the current scope: conv2d, conv2d_1, conv2d_2, ..."""
    key = '/'.join(_SCOPE + [base])
    n = _LAYER_COUNT.get(key, 0)
    _LAYER_COUNT[key] = n + 1
    return base if n == 0 else '%s_%d' % (base, n)


def _glorot_uniform_conv(shape, dt):
    # tf glorot_uniform on a [kh, kw, cin, cout] kernel: fans include the receptive field
    rf = int(np.prod(shape[:-2]))
    lim = np.sqrt(6. / (rf * shape[-2] + rf * shape[-1]))
    return _RNG.uniform(-lim, lim, size=shape).astype(dt)


def layers_conv2d(inputs, filters, kernel_size, strides=(1, 1), padding='valid', data_format='channels_last',
                  activation=None, use_bias=True, kernel_initializer=None, name=None, **kw):
    assert data_format == 'channels_first' and padding == 'same', 'only what app/modules.py:263-379 uses'
    kh, kw_ = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    assert kh % 2 == 1 and kw_ % 2 == 1
    x = _raw(inputs)
    cin = int(x.shape[1])
    with variable_scope(name or _layer_name('conv2d')):
        kernel = get_variable('kernel', [kh, kw_, cin, filters], initializer=kernel_initializer or _glorot_uniform_conv)
        bias = get_variable('bias', [filters], initializer=constant_initializer(0.)) if use_bias else None
    y = torch.nn.functional.conv2d(x, kernel.v.permute(3, 2, 0, 1), None if bias is None else bias.v,
                                   padding=(kh // 2, kw_ // 2))
    y = TT(y)
    return activation(y) if activation is not None else y


def layers_max_pooling2d(inputs, pool_size, strides, padding='valid', data_format='channels_last', name=None):
    assert data_format == 'channels_first' and padding == 'valid'
    return TT(torch.nn.functional.max_pool2d(_raw(inputs), tuple(pool_size), tuple(strides)))


def layers_dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, name=None, **kw):
    x = _raw(inputs)
    with variable_scope(name or _layer_name('dense')):
        kernel = get_variable('kernel', [int(x.shape[-1]), units], initializer=kernel_initializer)
        bias = get_variable('bias', [units], initializer=constant_initializer(0.)) if use_bias else None
    y = x @ kernel.v
    if bias is not None:
        y = y + bias.v
    y = TT(y)
    return activation(y) if activation is not None else y


def install():
    """register the fake `tensorflow` (and absent optional deps) in sys.modules"""
    tf = types.ModuleType('tensorflow')
    g = globals()
    for k in ('variable_scope name_scope get_variable Variable trainable_variables '
              'variables_initializer global_variables_initializer placeholder assign '
              'sigmoid tanh cos sin log log1p square real imag sqrt exp atan2 maximum '
              'less squared_difference clip_by_value cast ones_like zeros_like constant '
              'reduce_sum reduce_mean reduce_max reduce_prod argmax argmin matmul einsum '
              'tensordot reshape transpose expand_dims squeeze concat stack split tile '
              'gather gather_nd one_hot unsorted_segment_sum map_fn scan Session '
              'constant_initializer random_uniform_initializer random_normal_initializer'
              ).split():
        setattr(tf, k, g[k])
    tf.abs = abs_
    tf.complex = complex_
    tf.range = range_
    tf.shape = shape_
    for d in ('float32', 'float64', 'int32', 'int64', 'complex64', 'complex128'):
        setattr(tf, d, DType(d))
    tf.nn = types.SimpleNamespace(sigmoid=sigmoid, softmax=softmax, relu=relu,
                                  dropout=dropout, tanh=tanh)
    tf.train = types.SimpleNamespace(
        Saver=_Saver, AdamOptimizer=AdamOptimizer,
        GradientDescentOptimizer=GradientDescentOptimizer)
    tf.summary = types.SimpleNamespace(
        scalar=lambda *a, **k: None, merge=lambda *a, **k: None,
        FileWriter=lambda *a, **k: types.SimpleNamespace(add_summary=lambda *a, **k: None))
    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
        l1_regularizer=lambda s: (lambda _: None),
        l2_regularizer=lambda s: (lambda _: None)))
    tf.layers = types.SimpleNamespace(conv2d=layers_conv2d, max_pooling2d=layers_max_pooling2d, dense=layers_dense)
    sys.modules['tensorflow'] = tf
    # optional deps of dataset readers that are off the hot path
    for name in ('h5py', 'fuel', 'fuel.datasets', 'fuel.datasets.hdf5', 'fuel.schemes'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.H5PYDataset = object
            m.SequentialScheme = object
            sys.modules[name] = m
    # scipy >= 1.13 dropped scipy.signal.hann; default.json:7 still calls it
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, 'hann'):
        scipy.signal.hann = scipy.signal.windows.hann
    return tf
