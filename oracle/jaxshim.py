"""Functional stand-ins for the third-party modules the reference imports.

TEST INFRASTRUCTURE ONLY (same rule as bnf_oracle.py: nothing under ``bayesnf_b200/`` may import
this module).  Purpose: EXECUTE THE REFERENCE'S OWN SOURCE FILES -- ``models.py`` and
``inference.py`` under /root/reference/src/bayesnf, unmodified, imported from where they lie -- in
the build container, where jax / flax / optax / tensorflow-probability are not installable, and
commit what they compute as golden vectors (scripts/make_golden_numerics.py ->
tests/golden/numerics_*.npz).  The oracle restatement and the CUDA path are then checked against
values produced by the reference's code instead of against my reading of it.

What is the reference's and what is the shim's:

  reference code that runs unchanged   the model (`BayesianNeuralField1D.__call__`, the feature
                                       builders, `make_likelihood_model`, `prior_model_fn`),
                                       `make_model`, `make_prior`, `fit_map` (+ `_make_init_fn`,
                                       `num_splits`), `ensemble_map` (`_target_log_prob_fn`, `_run`,
                                       `_one_epoch`, `_one_step`, `_reshape_to_batches`,
                                       `permute_dataset`), `ensemble_vi`'s target / surrogate
                                       builders, `make_vi_init`, `predict_bnf`,
                                       `forecast_parameters_batched`, the Normal quantile functions
  restated here (third-party           array arithmetic (torch-CPU float64), `vmap` / `pmap` /
  semantics, each a few lines,         `lax.scan` as Python loops, `value_and_grad` (torch
  documented behaviour of the          autograd), pytrees (dict keys sorted, as jax does), Flax
  versions the reference pins)         `Module.param` / `init` / `apply`, `nn.Dense` (x @ kernel +
                                       bias, auto-named Dense_<n>), `optax.adam`, the TFP
                                       distributions' log_prob / cdf / quantile formulas,
                                       `JointDistributionCoroutine.log_prob` (sum over components),
                                       `fit_surrogate_posterior_stateless` (reverse-KL Monte-Carlo
                                       loss), threefry key handling (bayesnf_b200.jax_prng, pinned
                                       by KATs)
  NOT reproduced                       the random STREAMS of samplers (TruncatedNormal init draws,
                                       VI's reparameterisation noise): the shim draws them from
                                       numpy generators keyed by the threefry key words, and the
                                       golden files record the draws (float32-representable
                                       values) so that the oracle / CUDA path start from the
                                       same numbers

Arithmetic is float64 throughout (float32 inputs are promoted exactly): the goldens pin the
ALGORITHM; float32 rounding is judged by the tolerances the parity tests state.
"""

from __future__ import annotations

import collections
import math
import operator
import sys
import types
import zlib

import numpy as np
import torch

_F = torch.float64

# When a dict, the samplers append what they draw: 'perm' (jax.random.permutation results),
# 'eps' / 'vi_init' / 'vi_final' (fit_surrogate_posterior_stateless), 'normal_eps' (Normal.sample).
TRACE = None


# --------------------------------------------------------------------------
# arrays
# --------------------------------------------------------------------------
def _raw(x):
  """Arr / ndarray / numpy scalar / nested list -> torch tensor; python scalars pass through."""
  if isinstance(x, Arr):
    return x.t
  if isinstance(x, torch.Tensor):
    return x
  if isinstance(x, (np.ndarray, np.generic)):
    a = np.asarray(x)
    if a.dtype.kind == 'f':
      a = a.astype(np.float64)
    elif a.dtype.kind in 'ui':
      a = a.astype(np.int64)
    return torch.from_numpy(np.array(a, order='C', copy=True))   # (ascontiguousarray would make 0-d 1-d)
  if isinstance(x, (list, tuple)):
    if any(isinstance(v, (Arr, torch.Tensor)) for v in x):
      return torch.stack([torch.as_tensor(_raw(v)) for v in x])
    return _raw(np.asarray(x))
  return x


def _pair(a, b):
  """jax weak typing: a python float meeting an integer array gives a float array."""
  if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
    if a.is_floating_point() and not b.is_floating_point():
      b = b.to(_F)
    elif b.is_floating_point() and not a.is_floating_point():
      a = a.to(_F)
  elif isinstance(a, float) and isinstance(b, torch.Tensor) and not b.is_floating_point():
    b = b.to(_F)
  elif isinstance(b, float) and isinstance(a, torch.Tensor) and not a.is_floating_point():
    a = a.to(_F)
  return a, b


def _binop(op):
  def fwd(self, other):
    a, b = _pair(self.t, _raw(other))
    return Arr(op(a, b))

  def rev(self, other):
    a, b = _pair(_raw(other), self.t)
    return Arr(op(a, b))
  return fwd, rev


def _index(ix):
  if isinstance(ix, tuple):
    return tuple(_index(i) for i in ix)
  if isinstance(ix, Arr):
    return ix.t if ix.t.dtype == torch.bool else ix.t.long()
  if isinstance(ix, np.ndarray):
    return torch.from_numpy(ix.astype(bool if ix.dtype == bool else np.int64))
  if isinstance(ix, np.integer):
    return int(ix)
  return ix


class Arr:
  """The shim's jax.Array: a torch tensor behind the ndarray surface the reference touches."""

  __array_ufunc__ = None          # numpy operands defer their binary operators to this class
  __slots__ = ('t',)

  def __init__(self, t):
    t = _raw(t)
    if not isinstance(t, torch.Tensor):
      t = torch.as_tensor(t, dtype=_F if isinstance(t, float) else None)
    if t.is_floating_point() and t.dtype != _F:
      t = t.to(_F)
    self.t = t

  shape = property(lambda self: tuple(self.t.shape))
  ndim = property(lambda self: self.t.dim())
  size = property(lambda self: self.t.numel())
  dtype = property(lambda self: self.t.dtype)
  T = property(lambda self: Arr(self.t.T))

  __add__, __radd__ = _binop(operator.add)
  __sub__, __rsub__ = _binop(operator.sub)
  __mul__, __rmul__ = _binop(operator.mul)
  __truediv__, __rtruediv__ = _binop(operator.truediv)
  __floordiv__, __rfloordiv__ = _binop(operator.floordiv)
  __pow__, __rpow__ = _binop(operator.pow)
  __matmul__, __rmatmul__ = _binop(operator.matmul)
  __gt__ = _binop(operator.gt)[0]
  __ge__ = _binop(operator.ge)[0]
  __lt__ = _binop(operator.lt)[0]
  __le__ = _binop(operator.le)[0]
  __and__ = _binop(operator.and_)[0]
  __or__ = _binop(operator.or_)[0]

  def __eq__(self, other):
    return Arr(self.t == _raw(other))

  def __ne__(self, other):
    return Arr(self.t != _raw(other))

  __hash__ = object.__hash__

  def __neg__(self):
    return Arr(-self.t)

  def __invert__(self):
    return Arr(~self.t)

  def __getitem__(self, ix):
    return Arr(self.t[_index(ix)])

  def __len__(self):
    return self.t.shape[0]

  def __iter__(self):
    return (Arr(v) for v in self.t)

  def __float__(self):
    return float(self.t)

  def __int__(self):
    return int(self.t)

  def __bool__(self):
    return bool(self.t)

  def __index__(self):
    return int(self.t)

  def __array__(self, dtype=None, copy=None):
    a = self.t.detach().numpy()
    return a.astype(dtype) if dtype is not None else a

  def __repr__(self):
    return f'Arr({self.t!r})'

  def reshape(self, *shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
      shape = tuple(shape[0])
    return Arr(self.t.reshape(tuple(int(s) for s in shape)))

  def astype(self, dtype):
    return Arr(self.t.to(_torch_dtype(dtype)))

  def mean(self, axis=None):
    return _reduce(torch.mean, self, axis)

  def sum(self, axis=None):
    return _reduce(torch.sum, self, axis)

  def max(self, axis=None):
    return _reduce(torch.amax, self, axis)

  def min(self, axis=None):
    return _reduce(torch.amin, self, axis)


def _torch_dtype(dtype):
  if dtype is None:
    return None
  if isinstance(dtype, torch.dtype):
    return _F if dtype.is_floating_point else dtype
  kind = np.dtype(dtype).kind
  return _F if kind == 'f' else (torch.bool if kind == 'b' else torch.int64)


def _reduce(fn, x, axis):
  t = _raw(x)
  if not isinstance(t, torch.Tensor):
    t = torch.as_tensor(t, dtype=_F)
  if axis is None:
    axis = tuple(range(t.dim()))
  elif isinstance(axis, (int, np.integer)):
    axis = (int(axis),)
  if len(axis) == 0:
    return Arr(t)
  return Arr(fn(t, dim=tuple(int(a) for a in axis)))


def _unary(fn):
  def f(x):
    t = _raw(x)
    if not isinstance(t, torch.Tensor):
      t = torch.as_tensor(t, dtype=_F)
    if not t.is_floating_point():
      t = t.to(_F)
    return Arr(fn(t))
  return f


def _as_tensor(x, dtype=None):
  t = _raw(x)
  if not isinstance(t, torch.Tensor):
    t = torch.as_tensor(t, dtype=_F if isinstance(t, float) else None)
  if dtype is not None:
    t = t.to(dtype)
  return t


def _shape(shape):
  if isinstance(shape, (int, np.integer)):
    return (int(shape),)
  return tuple(int(s) for s in shape)


def _stack(vals):
  """Stack vmap / scan outputs: numpy key arrays stay numpy, everything else becomes Arr."""
  if all(isinstance(v, np.ndarray) for v in vals):
    return np.stack(vals)
  return Arr(torch.stack([_as_tensor(v) for v in vals]))


def _make_jnp():
  m = types.ModuleType('jax.numpy')
  m.ndarray = Arr
  m.pi = math.pi
  m.newaxis = None
  m.float32, m.float64, m.int32, m.int64 = np.float32, np.float64, np.int32, np.int64
  for name, fn in dict(cos=torch.cos, sin=torch.sin, tanh=torch.tanh, exp=torch.exp, log=torch.log,
                       sqrt=torch.sqrt, square=torch.square, ceil=torch.ceil, abs=torch.abs,
                       log1p=torch.log1p, expm1=torch.expm1, floor=torch.floor).items():
    setattr(m, name, _unary(fn))
  m.array = m.asarray = lambda x, dtype=None: Arr(_as_tensor(x, _torch_dtype(dtype)))
  m.reshape = lambda x, shape: Arr(_as_tensor(x)).reshape(shape)
  m.shape = lambda x: tuple(_as_tensor(x).shape)
  m.column_stack = lambda xs: Arr(torch.column_stack([_as_tensor(x) for x in xs]))
  m.concatenate = lambda xs, axis=0: Arr(torch.cat([_as_tensor(x) for x in xs], dim=axis))
  m.stack = lambda xs, axis=0: Arr(torch.stack([_as_tensor(x) for x in xs], dim=axis))
  m.tile = lambda x, reps: Arr(_as_tensor(x).repeat(reps))
  m.arange = lambda *a: Arr(torch.arange(*[int(v) for v in a]))
  m.prod = lambda x, axis=None: Arr(torch.prod(_as_tensor(x)) if axis is None
                                    else torch.prod(_as_tensor(x), dim=int(axis)))
  m.sum = lambda x, axis=None: _reduce(torch.sum, x, axis)
  m.mean = lambda x, axis=None: _reduce(torch.mean, x, axis)
  m.amin = lambda x, axis=None: _reduce(torch.amin, x, axis)
  m.amax = lambda x, axis=None: _reduce(torch.amax, x, axis)
  m.ones = lambda shape, dtype=None: Arr(torch.ones(_shape(shape), dtype=_torch_dtype(dtype) or _F))
  m.zeros = lambda shape, dtype=None: Arr(torch.zeros(_shape(shape), dtype=_torch_dtype(dtype) or _F))
  m.ones_like = lambda x: Arr(torch.ones_like(_as_tensor(x)))
  m.zeros_like = lambda x: Arr(torch.zeros_like(_as_tensor(x)))
  m.where = lambda c, a, b: Arr(torch.where(_as_tensor(c), *[
      torch.as_tensor(v, dtype=_F) if not isinstance(v, torch.Tensor) else v
      for v in _pair(_raw(a), _raw(b))]))
  m.maximum = lambda a, b: Arr(torch.maximum(*torch.broadcast_tensors(_as_tensor(a, _F), _as_tensor(b, _F))))
  m.minimum = lambda a, b: Arr(torch.minimum(*torch.broadcast_tensors(_as_tensor(a, _F), _as_tensor(b, _F))))
  return m


# --------------------------------------------------------------------------
# pytrees (jax.tree_util): tuples, lists, dicts with SORTED keys, None = empty node
# --------------------------------------------------------------------------
class _Leaf:
  pass


_LEAF = _Leaf()


def _flatten(tree, leaves):
  if tree is None:
    return None
  if isinstance(tree, dict):
    keys = sorted(tree)
    return ('dict', keys, [_flatten(tree[k], leaves) for k in keys])
  if isinstance(tree, tuple) and hasattr(tree, '_fields'):
    return ('namedtuple', type(tree), [_flatten(v, leaves) for v in tree])
  if isinstance(tree, (tuple, list)):
    return (type(tree).__name__, None, [_flatten(v, leaves) for v in tree])
  leaves.append(tree)
  return _LEAF


def _unflatten(treedef, it):
  if treedef is None:
    return None
  if treedef is _LEAF:
    return next(it)
  kind, meta, kids = treedef
  vals = [_unflatten(k, it) for k in kids]
  if kind == 'dict':
    return dict(zip(meta, vals))
  if kind == 'namedtuple':
    return meta(*vals)
  return tuple(vals) if kind == 'tuple' else list(vals)


def tree_flatten(tree):
  leaves = []
  treedef = _flatten(tree, leaves)
  return leaves, treedef


def tree_leaves(tree):
  return tree_flatten(tree)[0]


def tree_structure(tree):
  return tree_flatten(tree)[1]


def tree_unflatten(treedef, leaves):
  return _unflatten(treedef, iter(leaves))


def tree_map(fn, tree, *rest):
  leaves, treedef = tree_flatten(tree)
  others = [tree_leaves(r) for r in rest]
  return tree_unflatten(treedef, [fn(*vs) for vs in zip(leaves, *others)])


# --------------------------------------------------------------------------
# jax transforms
# --------------------------------------------------------------------------
def _slice_leaf(v, i):
  return v[i]


def vmap(fn, in_axes=0, out_axes=0):
  assert out_axes == 0

  def mapped(*args):
    axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    assert all(a in (0, None) for a in axes), 'the shim maps over leading axes only'
    n = None
    for a, ax in zip(args, axes):
      if ax is None:
        continue
      lv = tree_leaves(a)
      if lv:
        n = lv[0].shape[0]
        break
    outs = []
    for i in range(n):
      outs.append(fn(*[a if ax is None else tree_map(lambda v: _slice_leaf(v, i), a)
                       for a, ax in zip(args, axes)]))
    return tree_map(lambda *vs: _stack(vs), *outs)
  return mapped


def scan(fn, init, xs=None, length=None):
  n = length if xs is None else tree_leaves(xs)[0].shape[0]
  carry, ys = init, []
  for i in range(int(n)):
    carry, y = fn(carry, None if xs is None else tree_map(lambda v: _slice_leaf(v, i), xs))
    ys.append(y)
  return carry, (tree_map(lambda *vs: _stack(vs), *ys) if ys else None)


def value_and_grad(fn):
  def wrapped(params, *rest, **kw):
    leaves, treedef = tree_flatten(params)
    req = [Arr(_as_tensor(l).detach().clone().requires_grad_(True)) for l in leaves]
    out = fn(tree_unflatten(treedef, req), *rest, **kw)
    grads = torch.autograd.grad(out.t.sum(), [r.t for r in req], allow_unused=True)
    grads = [Arr(torch.zeros_like(r.t) if g is None else g) for g, r in zip(grads, req)]
    return Arr(out.t.detach()), tree_unflatten(treedef, grads)
  return wrapped


def _jit(fn=None, **_):
  return fn if fn is not None else (lambda f: f)


# --------------------------------------------------------------------------
# jax.random: threefry key handling from bayesnf_b200.jax_prng (KAT-pinned); sampler streams
# are NOT jax's (see the module docstring)
# --------------------------------------------------------------------------
def _key(key):
  return np.asarray(key.t.numpy() if isinstance(key, Arr) else key).astype(np.uint32)


def _np_rng(key):
  k = _key(key)
  return np.random.Generator(np.random.Philox(key=int(k[0]) << 32 | int(k[1])))


def _make_random():
  from bayesnf_b200 import jax_prng
  m = types.ModuleType('jax.random')
  m.PRNGKey = lambda seed: jax_prng.prng_key(seed)
  m.split = lambda key, num=2: jax_prng.split(_key(key), num)
  m.fold_in = lambda key, data: jax_prng.fold_in(_key(key), data)

  def permutation(key, x):
    n = int(x) if isinstance(x, (int, np.integer)) else x.shape[0]
    perm = jax_prng.permutation(_key(key), n)
    if TRACE is not None:
      TRACE.setdefault('perm', []).append(np.asarray(perm, dtype=np.int64))
    if isinstance(x, (int, np.integer)):
      return Arr(perm)
    return Arr(_as_tensor(x)[torch.from_numpy(perm.astype(np.int64))])
  m.permutation = permutation
  m.normal = lambda key, shape=(), dtype=None: Arr(_np_rng(key).standard_normal(_shape(shape)))
  return m


# --------------------------------------------------------------------------
# flax.linen
# --------------------------------------------------------------------------
class _Field:
  def __init__(self, default_factory):
    self.default_factory = default_factory


class _Scope:
  def __init__(self, mode, params, key=None):
    self.mode, self.params, self.key, self.counters, self.draws = mode, params, key, {}, 0


_module_stack = []


class Module:
  """flax.linen.Module, as far as BayesianNeuralField1D uses it: dataclass-style fields, `param`,
  `init`, `apply`, auto-named compact submodules."""

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    fields = []
    for klass in reversed(cls.__mro__):
      for name in getattr(klass, '__annotations__', {}):
        if name not in fields and not name.startswith('_'):
          fields.append(name)
    cls._fields_ = fields

  def __init__(self, *args, **kw):
    names = list(self._fields_)
    for name, v in zip(names, args):
      kw[name] = v
    for name in names:
      if name in kw:
        v = kw[name]
      else:
        v = getattr(type(self), name)
        if isinstance(v, _Field):
          v = v.default_factory()
      object.__setattr__(self, name, v)
    self._scope = None
    if _module_stack:          # constructed inside a parent's compact __call__: auto-name
      parent = _module_stack[-1]
      ps = parent._scope
      idx = ps.counters.get(type(self).__name__, 0)
      ps.counters[type(self).__name__] = idx + 1
      name = f'{type(self).__name__}_{idx}'
      if ps.mode == 'init':
        sub = ps.params.setdefault(name, {})
      else:
        sub = ps.params[name]
      self._scope = _Scope(ps.mode, sub, ps.key)

  def param(self, name, init_fn, shape=()):
    s = self._scope
    if s.mode == 'init':
      s.draws += 1
      s.params[name] = init_fn(np.array([s.draws, zlib.crc32(name.encode())], np.uint32), _shape(shape))
    return s.params[name]

  def _run(self, scope, args):
    self._scope = scope
    _module_stack.append(self)
    try:
      return type(self).__call__(self, *args)
    finally:
      _module_stack.pop()
      self._scope = None

  def init(self, key, *args):
    scope = _Scope('init', {}, key)
    self._run(scope, args)
    return {'params': scope.params}

  def apply(self, variables, *args):
    return self._run(_Scope('apply', variables['params']), args)


class Dense(Module):
  """flax.linen.Dense: y = x @ kernel + bias, kernel (in, features), bias (features,)."""
  features: int
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, x):
    kernel = self.param('kernel', self.kernel_init, (x.shape[-1], self.features))
    bias = self.param('bias', self.bias_init, (self.features,))
    return x @ kernel + bias


def _make_flax():
  flax = types.ModuleType('flax')
  nn = types.ModuleType('flax.linen')
  nn.Module, nn.Dense = Module, Dense
  nn.compact = lambda f: f
  nn.elu = lambda x: Arr(torch.where(x.t > 0, x.t, torch.expm1(torch.clamp(x.t, max=0.0))))
  nn.tanh = _unary(torch.tanh)
  nn.softplus = _softplus
  nn.sigmoid = _sigmoid
  init = types.ModuleType('flax.linen.initializers')
  init.normal = lambda stddev=1e-2: (
      lambda key, shape, dtype=None: Arr(_np_rng(key).standard_normal(_shape(shape)) * stddev))
  nn.initializers = init
  struct = types.ModuleType('flax.struct')
  struct.field = lambda default_factory=None, **_: _Field(default_factory)
  core = types.ModuleType('flax.core')
  fd = types.ModuleType('flax.core.frozen_dict')
  fd.FrozenDict = dict
  scope = types.ModuleType('flax.core.scope')
  scope.FrozenVariableDict = dict
  core.frozen_dict, core.scope = fd, scope
  flax.linen, flax.struct, flax.core = nn, struct, core
  return {'flax': flax, 'flax.linen': nn, 'flax.linen.initializers': init, 'flax.struct': struct,
          'flax.core': core, 'flax.core.frozen_dict': fd, 'flax.core.scope': scope}


def _softplus(x):
  t = _as_tensor(x, _F)
  return Arr(torch.logaddexp(t, torch.zeros_like(t)))


def _sigmoid(x):
  return Arr(torch.sigmoid(_as_tensor(x, _F)))


# --------------------------------------------------------------------------
# optax
# --------------------------------------------------------------------------
class _Adam:
  """optax.adam(learning_rate): scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0) with bias
  correction by the incremented count, then scale by -learning_rate."""

  def __init__(self, learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    self.lr, self.b1, self.b2, self.eps, self.eps_root = learning_rate, b1, b2, eps, eps_root

  def init(self, params):
    z = tree_map(lambda p: Arr(torch.zeros_like(_as_tensor(p))), params)
    return (0, z, tree_map(lambda p: Arr(torch.zeros_like(_as_tensor(p))), params))

  def update(self, grads, state, params=None):
    count, mu, nu = state
    count += 1
    mu = tree_map(lambda m, g: self.b1 * m + (1 - self.b1) * g, mu, grads)
    nu = tree_map(lambda v, g: self.b2 * v + (1 - self.b2) * (g * g), nu, grads)
    c1, c2 = 1 - self.b1 ** count, 1 - self.b2 ** count

    def upd(m, v):
      return -self.lr * ((m / c1) / (Arr(torch.sqrt((v / c2).t + self.eps_root)) + self.eps))
    return tree_map(upd, mu, nu), (count, mu, nu)


def _make_optax():
  m = types.ModuleType('optax')
  m.adam = _Adam
  m.apply_updates = lambda params, updates: tree_map(lambda p, u: p + u, params, updates)
  return m


# --------------------------------------------------------------------------
# tensorflow_probability.substrates.jax
# --------------------------------------------------------------------------
def _bshape(*xs):
  return tuple(torch.broadcast_shapes(*[tuple(_as_tensor(x).shape) for x in xs]))


class Distribution:
  name = None

  @property
  def batch_shape(self):
    return _bshape(self.loc, self.scale)

  def sample(self, sample_shape=(), seed=None):
    raise NotImplementedError

  def prob(self, x):
    return Arr(torch.exp(self.log_prob(x).t))


class Normal(Distribution):
  def __init__(self, loc, scale, name=None):
    self.loc, self.scale, self.name = Arr(_as_tensor(loc, _F)), Arr(_as_tensor(scale, _F)), name

  def log_prob(self, x):
    z = (Arr(_as_tensor(x, _F)) - self.loc) / self.scale
    return -0.5 * z * z - 0.5 * math.log(2 * math.pi) - Arr(torch.log(self.scale.t))

  def cdf(self, x):
    z = (Arr(_as_tensor(x, _F)) - self.loc) / self.scale
    return Arr(torch.special.ndtr(z.t))

  def quantile(self, q):
    return self.loc + self.scale * Arr(torch.special.ndtri(_as_tensor(q, _F)))

  def sample(self, sample_shape=(), seed=None, eps=None):
    shape = _shape(sample_shape) + _bshape(self.loc, self.scale)
    if eps is None:
      eps = Arr(_np_rng(seed).standard_normal(shape).astype(np.float32))     # float32-exact draws
      if TRACE is not None:
        TRACE.setdefault('normal_eps', []).append(eps)
    return self.loc + self.scale * eps


class Logistic(Distribution):
  def __init__(self, loc, scale, name=None):
    self.loc, self.scale, self.name = Arr(_as_tensor(loc, _F)), Arr(_as_tensor(scale, _F)), name

  def log_prob(self, x):
    z = (Arr(_as_tensor(x, _F)) - self.loc) / self.scale
    return -z - 2.0 * _softplus(-z) - Arr(torch.log(self.scale.t))

  def sample(self, sample_shape=(), seed=None):
    shape = _shape(sample_shape) + _bshape(self.loc, self.scale)
    return self.loc + self.scale * Arr(_np_rng(seed).logistic(size=shape))


class Deterministic(Distribution):
  def __init__(self, loc, name=None):
    self.loc, self.name = Arr(_as_tensor(loc, _F)), name

  batch_shape = property(lambda self: self.loc.shape)

  def log_prob(self, x):
    return Arr(torch.where(_as_tensor(x, _F) == self.loc.t, 0.0, -math.inf))

  def sample(self, sample_shape=(), seed=None):
    return Arr(self.loc.t.expand(_shape(sample_shape) + self.loc.shape).clone())


class TruncatedNormal(Distribution):
  def __init__(self, loc, scale, low, high, name=None):
    self.loc, self.scale = Arr(_as_tensor(loc, _F)), Arr(_as_tensor(scale, _F))
    self.low, self.high, self.name = float(low), float(high), name

  def sample(self, sample_shape=(), seed=None):
    """Inverse-CDF draw; the STREAM is the shim's, not TFP's (module docstring)."""
    shape = _shape(sample_shape) + _bshape(self.loc, self.scale)
    u = torch.from_numpy(_np_rng(seed).random(shape))
    lo = torch.special.ndtr((self.low - self.loc.t) / self.scale.t)
    hi = torch.special.ndtr((self.high - self.loc.t) / self.scale.t)
    z = torch.special.ndtri(lo + u * (hi - lo))
    # rounded to float32 so that a float32 implementation can start from exactly these numbers
    return Arr(torch.clamp(self.loc.t + self.scale.t * z, self.low, self.high).float().double())


class NegativeBinomial(Distribution):
  """tfd.NegativeBinomial(total_count, logits): pmf(k) = C(k+n-1, k) (1-p)^n p^k with
  p = sigmoid(logits) (k successes before the n-th failure)."""

  def __init__(self, total_count, logits, name=None):
    self.total_count, self.logits = Arr(_as_tensor(total_count, _F)), Arr(_as_tensor(logits, _F))
    self.name = name

  def log_prob(self, x):
    n, lg, k = self.total_count.t, self.logits.t, _as_tensor(x, _F)
    logsig = torch.nn.functional.logsigmoid
    unnorm = n * logsig(-lg) + k * logsig(lg)
    lognorm = torch.lgamma(1.0 + k) + torch.lgamma(n) - torch.lgamma(n + k)
    return Arr(unnorm - lognorm)

  def mean(self):
    return self.total_count * Arr(torch.exp(self.logits.t))

  def variance(self):
    return self.mean() / Arr(torch.sigmoid(-self.logits.t))

  def stddev(self):
    return Arr(torch.sqrt(self.variance().t))

  def cdf(self, x):
    # P(X <= k) = I_{1-p}(n, floor(k)+1); scipy's regularised incomplete beta, no autograd needed
    from scipy import special
    n, lg = np.asarray(self.total_count), np.asarray(self.logits)
    k = np.floor(np.asarray(Arr(_as_tensor(x, _F))))
    out = special.betainc(n, np.maximum(k, 0.0) + 1.0, special.expit(-lg))
    return Arr(np.where(k < 0, 0.0, out))


class ZeroInflatedNegativeBinomial(Distribution):
  """Mixture pi * delta_0 + (1 - pi) * NB."""

  def __init__(self, total_count, logits, inflated_loc_probs, name=None):
    self.nb = NegativeBinomial(total_count, logits)
    self.total_count, self.logits = self.nb.total_count, self.nb.logits
    self.inflated_loc_probs, self.name = Arr(_as_tensor(inflated_loc_probs, _F)), name

  def log_prob(self, x):
    pi, k = self.inflated_loc_probs.t, _as_tensor(x, _F)
    nb = self.nb.log_prob(x).t
    pi, nb = torch.broadcast_tensors(pi, nb)
    at0 = torch.logaddexp(torch.log(pi), torch.log1p(-pi) + nb)
    return Arr(torch.where(k == 0, at0, torch.log1p(-pi) + nb))

  def mean(self):
    return (1.0 - self.inflated_loc_probs) * self.nb.mean()

  def variance(self):
    pi, m, v = self.inflated_loc_probs, self.nb.mean(), self.nb.variance()
    return (1.0 - pi) * (v + m * m) - self.mean() * self.mean()

  def stddev(self):
    return Arr(torch.sqrt(self.variance().t))

  def cdf(self, x):
    k = _as_tensor(x, _F)
    pi = self.inflated_loc_probs
    return Arr(torch.where(k >= 0, 1.0, 0.0)) * pi + (1.0 - pi) * self.nb.cdf(x)


class Independent(Distribution):
  def __init__(self, distribution, reinterpreted_batch_ndims=None, name=None):
    self.distribution, self.nd, self.name = distribution, reinterpreted_batch_ndims, name

  def log_prob(self, x):
    lp = self.distribution.log_prob(x)
    return lp.sum(axis=tuple(range(lp.ndim - self.nd, lp.ndim)))


class JointDistributionCoroutine(Distribution):
  """With `use_vectorized_map=True` the model is written for ONE sample and `batch_ndims` says how
  many leading dimensions of every component are batch; everything else is event, so a
  component's log-prob is summed over its trailing dimensions."""

  def __init__(self, model, use_vectorized_map=False, batch_ndims=None, name=None):
    self.model, self.batch_ndims = model, int(batch_ndims or 0)

  def _walk(self, visit):
    gen = self.model()
    out = []
    try:
      d = next(gen)
      while True:
        v = visit(len(out), d)
        out.append(v)
        d = gen.send(v)
    except StopIteration:
      pass
    return out

  def log_prob(self, *value):
    if len(value) == 1 and isinstance(value[0], (tuple, list)):
      value = value[0]
    total = [0.0]

    def visit(i, d):
      lp = d.log_prob(value[i])
      event_ndims = len(d.batch_shape) - self.batch_ndims
      total[0] = total[0] + lp.sum(axis=tuple(range(lp.ndim - event_ndims, lp.ndim)))
      return value[i]
    self._walk(visit)
    return total[0]

  def sample(self, sample_shape=(), seed=None, **kw):
    from bayesnf_b200 import jax_prng
    counter = [jax_prng.prng_key(0) if seed is None else _key(seed)]

    names = []

    def visit(i, d):
      counter[0], sub = jax_prng.split(counter[0], 2)
      names.append(d.name or f'var{i}')
      return d.sample(sample_shape, seed=sub)
    values = self._walk(visit)
    # TFP returns a StructTuple (a namedtuple keyed by the component names); the reference relies on
    # `_fields` / `_replace` (spatiotemporal.py:459-462)
    return collections.namedtuple('StructTuple', names)(*values)

  def sample_with(self, eps):
    """Reparameterised draw of a surrogate of Normals from given standard-normal tensors."""
    return tuple(self._walk(lambda i, d: d.sample(eps=eps[i])))


class _Root:
  def __init__(self, root):
    self.estimated_root = root


def _find_root(fn, low, high, value_tolerance=1e-5, max_iterations=60, **_):
  """Stand-in for tfp.math.find_root_chandrupatla: plain bisection of a monotone function to
  float64 resolution; a Chandrupatla root stops anywhere within `value_tolerance` of it, which is
  why the parity tests judge root quantiles by the CDF residual."""
  lo = _as_tensor(low, _F)
  hi = _as_tensor(high, _F)
  f0 = fn(Arr(lo))
  lo, hi = lo.expand(f0.shape).clone(), hi.expand(f0.shape).clone()
  for _ in range(200):
    mid = 0.5 * (lo + hi)
    neg = fn(Arr(mid)).t < 0
    lo = torch.where(neg, mid, lo)
    hi = torch.where(neg, hi, mid)
  return _Root(Arr(0.5 * (lo + hi)))


def _fit_surrogate_posterior_stateless(target_log_prob_fn, build_surrogate_posterior_fn,
                                       initial_parameters, optimizer, num_steps, sample_size=1,
                                       jit_compile=False, seed=None, eps_hook=None, **_):
  """tfp.vi.fit_surrogate_posterior_stateless with the default reverse-KL divergence:
  loss = mean over `sample_size` reparameterised draws z ~ q of (log q(z) - target(z)),
  minimised by `optimizer`; the per-step seed is split into (sample, target) seeds.  `eps_hook`
  (shim extension) receives (step, shapes) and returns the standard-normal draws to use, so a
  test can replay them elsewhere."""
  from bayesnf_b200 import jax_prng
  params = initial_parameters
  state = optimizer.init(params)
  key = _key(seed)
  losses = []
  if TRACE is not None:
    TRACE.setdefault('vi_init', []).append(params)

  def loss_fn(p, eps, tseed):
    q = build_surrogate_posterior_fn(*p)
    z = q.sample_with(eps)
    return (q.log_prob(z) - target_log_prob_fn(*z, seed=tseed)).mean(axis=0)

  for step in range(int(num_steps)):
    key, sub = jax_prng.split(key, 2)
    sseed, tseed = jax_prng.split(sub, 2)
    shapes = [(sample_size,) + tuple(params[2 * i].shape) for i in range(len(params) // 2)]
    if eps_hook is not None:
      eps = [Arr(e) for e in eps_hook(step, shapes)]
    else:
      rng = _np_rng(sseed)
      eps = [Arr(rng.standard_normal(s).astype(np.float32)) for s in shapes]   # float32-exact draws
    if TRACE is not None:
      TRACE.setdefault('eps', []).append(eps)
    loss, grads = value_and_grad(loss_fn)(params, eps, tseed)
    updates, state = optimizer.update(grads, state)
    params = tree_map(lambda p, u: p + u, params, updates)
    losses.append(loss)
  if TRACE is not None:
    TRACE.setdefault('vi_final', []).append(params)
  return params, _stack(losses)


def _make_tfp():
  tfp = types.ModuleType('tensorflow_probability.substrates.jax')
  tfd = types.ModuleType('tensorflow_probability.substrates.jax.distributions')
  for cls in (Distribution, Normal, Logistic, Deterministic, TruncatedNormal, NegativeBinomial,
              ZeroInflatedNegativeBinomial, Independent, JointDistributionCoroutine):
    setattr(tfd, cls.__name__, cls)
  tfd.JointDistributionSequential = JointDistributionCoroutine
  mathm = types.ModuleType('tensorflow_probability.substrates.jax.math')
  mathm.softplus_inverse = lambda x: Arr(torch.log(torch.expm1(_as_tensor(x, _F))))
  mathm.find_root_chandrupatla = _find_root
  vi = types.ModuleType('tensorflow_probability.substrates.jax.vi')
  vi.fit_surrogate_posterior_stateless = _fit_surrogate_posterior_stateless
  tfp.distributions, tfp.math, tfp.vi = tfd, mathm, vi
  top = types.ModuleType('tensorflow_probability')
  sub = types.ModuleType('tensorflow_probability.substrates')
  sub.jax, top.substrates = tfp, sub
  return {'tensorflow_probability': top, 'tensorflow_probability.substrates': sub,
          'tensorflow_probability.substrates.jax': tfp}


# --------------------------------------------------------------------------
# install
# --------------------------------------------------------------------------
def install():
  """Registers the stand-ins in sys.modules (idempotent).  Returns the fake `jax` module."""
  if getattr(sys.modules.get('jax'), '_bnf_shim', False):
    return sys.modules['jax']
  jax = types.ModuleType('jax')
  jax._bnf_shim = True
  jnp = _make_jnp()
  jax.numpy = jnp
  jax.Array = Arr
  typing_m = types.ModuleType('jax.typing')
  typing_m.ArrayLike = object
  jax.typing = typing_m
  nn = types.ModuleType('jax.nn')
  nn.softplus, nn.sigmoid = _softplus, _sigmoid
  jax.nn = nn
  lax = types.ModuleType('jax.lax')
  lax.scan = scan
  lax.rsqrt = lambda x: Arr(torch.rsqrt(_as_tensor(x, _F)))
  jax.lax = lax
  tu = types.ModuleType('jax.tree_util')
  tu.tree_leaves, tu.tree_structure, tu.tree_unflatten = tree_leaves, tree_structure, tree_unflatten
  tu.tree_map, tu.tree_flatten = tree_map, tree_flatten
  jax.tree_util = tu
  jax.random = _make_random()
  jax.vmap = vmap
  jax.pmap = lambda fn=None, in_axes=0, **_: vmap(fn, in_axes=in_axes)
  jax.jit = _jit
  jax.value_and_grad = value_and_grad
  jax.grad = lambda fn: (lambda *a, **k: value_and_grad(fn)(*a, **k)[1])
  jax.device_count = lambda: 1
  jax.local_device_count = lambda: 1
  jax.devices = lambda: ['shim-cpu']
  mods = {'jax': jax, 'jax.numpy': jnp, 'jax.typing': typing_m, 'jax.nn': nn, 'jax.lax': lax,
          'jax.tree_util': tu, 'jax.random': jax.random, 'optax': _make_optax()}
  mods.update(_make_flax())
  mods.update(_make_tfp())
  jaxtyping = types.ModuleType('jaxtyping')
  jaxtyping.PyTree = object
  mods['jaxtyping'] = jaxtyping
  sys.modules.update(mods)
  return jax


def import_reference(ref_root='/root/reference'):
  """install() + import the reference's `bayesnf` package from its own source tree."""
  import os
  install()
  src = os.path.join(ref_root, 'src')
  if src not in sys.path:
    sys.path.insert(0, src)
  from bayesnf import inference, models, spatiotemporal  # pylint: disable=import-outside-toplevel
  return models, inference, spatiotemporal
