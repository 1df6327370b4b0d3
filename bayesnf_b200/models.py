"""Host-side model bookkeeping (mirror of the reference's ``bayesnf.models``).

The arithmetic of the model (feature encode, dense stack, likelihoods, prior)
runs in libbnf_sm100.so; this module only keeps the reference's host-side
bookkeeping -- which is pure numpy there too -- and wraps the C plan:

* ``LikelihoodDist``                models.py:30-33
* ``make_seasonal_frequencies``     models.py:36-59 (bit-exact numpy bookkeeping)
* ``ModelSpec``                     the static part of ``make_model`` /
  ``make_prior`` (inference.py:234-268): parameter-leaf table in
  ``jax.tree_util.tree_leaves`` order, flat <-> tuple conversion.
"""

from __future__ import annotations

import ctypes as C
import enum
from typing import Sequence

import numpy as np

from . import _lib


class LikelihoodDist(enum.Enum):
  NORMAL = 'NORMAL'
  NB = 'NB'
  ZINB = 'ZINB'


_LIK_CODE = {LikelihoodDist.NORMAL: _lib.NORMAL, LikelihoodDist.NB: _lib.NB,
             LikelihoodDist.ZINB: _lib.ZINB}


def make_seasonal_frequencies(seasonality_periods, num_harmonics):
  """Frequency table of the seasonal features: (frequencies, harmonic numbers).

  Contract of models.py:36-59, pinned bit for bit by tests/golden/bookkeeping.json: period i
  contributes the float32 quotients ``h / p_i`` for ``h = 1 .. H_i``; a frequency that was already
  produced by an earlier (period, harmonic) pair is dropped (a weekly 2nd harmonic and a
  half-weekly 1st one are the same feature); the harmonic number travels with its frequency.
  Same ValueErrors as the reference.
  """
  p = np.array(seasonality_periods, dtype=np.float32)
  n_h = np.asarray(num_harmonics)
  if np.any(n_h > p / 2):
    raise ValueError('Harmonic cannot exceed half seasonal period.')
  if p.shape != n_h.shape:
    raise ValueError('Number of seasonal periods and harmonics must be equal.')
  if n_h.ndim != 1:
    raise ValueError('Arguments `num_harmonics` and `seasonality_periods` must be rank 1.')
  if p.size == 0:
    return np.zeros(0), np.zeros(0)
  # one flat (harmonic, period) table instead of per-period pieces
  counts = np.array([np.arange(1, h + 1).size for h in n_h])
  owner = np.repeat(np.arange(p.size), counts)                  # period index of every entry
  starts = np.cumsum(counts) - counts
  harmonic = (np.arange(counts.sum()) - starts[owner] + 1).astype(np.float32)
  freq = harmonic / p[owner]                                    # float32 / float32
  seen, keep = set(), []
  for i, bits in enumerate(freq.view(np.uint32).tolist()):      # first occurrence wins
    if bits not in seen:
      seen.add(bits)
      keep.append(i)
  return freq[keep], harmonic[keep]


class ModelSpec:
  """Static model description + the C plan built from ``model_args``.

  ``model_args`` keys are the reference's (spatiotemporal.py:360-370):
  depth, width, input_scales, num_seasonal_harmonics, seasonality_periods,
  init_x, fourier_degrees, interactions.
  """

  def __init__(self, *, width, depth, input_scales, num_seasonal_harmonics,
               seasonality_periods, init_x, fourier_degrees, interactions,
               observation_model='NORMAL'):
    self.distribution = LikelihoodDist(observation_model)
    self.width, self.depth = int(width), int(depth)
    self.input_dim = int(init_x[-1]) if len(init_x) > 1 else 1
    self.input_scales = np.ascontiguousarray(input_scales, dtype=np.float64)
    self.fourier_degrees = np.ascontiguousarray(fourier_degrees).astype(np.int32)
    self.interactions = np.ascontiguousarray(
        np.asarray(interactions).reshape(-1, 2)).astype(np.int32)
    if self.fourier_degrees.shape[0] != self.input_dim:
      raise ValueError('fourier_degrees must have one entry per input dim.')
    if self.input_scales.shape[0] != self.input_dim:
      raise ValueError('input_scales must have one entry per input dim.')
    freqs, harm = make_seasonal_frequencies(
        seasonality_periods, np.asarray(num_seasonal_harmonics))
    self.seasonal_freq = np.ascontiguousarray(freqs, dtype=np.float32)
    self.seasonal_harm = np.ascontiguousarray(harm, dtype=np.float32)

    cfg = _lib.Config()
    cfg.abi_version = _lib.BNF_ABI_VERSION
    cfg.input_dim, cfg.width, cfg.depth = self.input_dim, self.width, self.depth
    cfg.likelihood = _LIK_CODE[self.distribution]
    cfg.n_seasonal = len(self.seasonal_freq)
    cfg.seasonal_freq = self.seasonal_freq.ctypes.data_as(C.POINTER(C.c_float))
    cfg.seasonal_harm = self.seasonal_harm.ctypes.data_as(C.POINTER(C.c_float))
    cfg.fourier_degrees = self.fourier_degrees.ctypes.data_as(C.POINTER(C.c_int32))
    cfg.n_interactions = len(self.interactions)
    cfg.interactions = self.interactions.ctypes.data_as(C.POINTER(C.c_int32))
    cfg.input_scales = self.input_scales.ctypes.data_as(C.POINTER(C.c_double))
    handle = C.c_void_p()
    _lib.check(_lib.lib.bnf_plan_create(C.byref(cfg), C.byref(handle)))
    self._plan = handle
    info = _lib.PlanInfo()
    _lib.check(_lib.lib.bnf_plan_info(self._plan, C.byref(info)))
    self.num_params = info.num_params
    self.num_features = info.num_features
    self.padded_features = info.padded_features
    self.num_feature_groups = info.num_feature_groups
    # leaves after the three scalar heads, in tree_leaves order
    self.leaf_names, self.leaf_offsets, self.leaf_shapes = [], [], []
    buf = C.create_string_buffer(64)
    for i in range(info.num_leaves):
      off, rows, cols = C.c_int64(), C.c_int32(), C.c_int32()
      _lib.check(_lib.lib.bnf_plan_leaf(self._plan, i, buf, 64, C.byref(off),
                                        C.byref(rows), C.byref(cols)))
      shape = () if rows.value == 0 else (
          (rows.value,) if cols.value == 0 else (rows.value, cols.value))
      self.leaf_names.append(buf.value.decode())
      self.leaf_offsets.append(off.value)
      self.leaf_shapes.append(shape)

  def __del__(self):
    plan, self._plan = getattr(self, '_plan', None), None
    if plan and _lib is not None and getattr(_lib, 'lib', None) is not None:
      _lib.lib.bnf_plan_destroy(plan)

  @property
  def plan(self):
    return self._plan

  # --- flat (.., P) array  <->  reference params tuple -------------------------
  def unflatten(self, flat: np.ndarray) -> tuple[np.ndarray, ...]:
    """(..., P) -> (log_noise_scale, shape, inflated_loc_probs, *leaves)."""
    lead = flat.shape[:-1]
    out = [flat[..., 0], flat[..., 1], flat[..., 2]]
    for off, shape in zip(self.leaf_offsets, self.leaf_shapes):
      n = int(np.prod(shape)) if shape else 1
      out.append(flat[..., off:off + n].reshape(lead + tuple(shape)))
    return tuple(out)

  def flatten(self, params: Sequence[np.ndarray]) -> np.ndarray:
    """Inverse of :meth:`unflatten`; leading dims are kept."""
    params = [np.asarray(p, dtype=np.float32) for p in params]
    if len(params) != 3 + len(self.leaf_shapes):
      raise ValueError(f'expected {3 + len(self.leaf_shapes)} parameter leaves, '
                       f'got {len(params)}')
    lead = params[0].shape
    return np.concatenate([p.reshape(lead + (-1,)) for p in params], axis=-1)

  # --- reference params tuple  <->  Flax variables dict --------------------------
  def to_flax_variables(self, params: Sequence[np.ndarray]) -> tuple[tuple[np.ndarray, ...], dict]:
    """Split a params tuple into what the reference's code paths consume: the three likelihood
    scalars `(log_noise_scale, shape, inflated_loc_probs)` and the Flax variables dict
    `{'params': {'Dense_0': {'bias', 'kernel'}, ..., 'feature_inv_sp_scale0': ..., ...}}` that
    `tree_unflatten(tree_structure(mlp_template), params[3:])` builds (models.py:158-160).
    Leading ensemble axes are kept on every leaf."""
    if len(params) != 3 + len(self.leaf_names):
      raise ValueError(f'expected {3 + len(self.leaf_names)} parameter leaves, got {len(params)}')
    tree: dict = {}
    for name, leaf in zip(self.leaf_names, params[3:]):
      node = tree
      *scopes, last = name.split('/')
      for scope in scopes:
        node = node.setdefault(scope, {})
      node[last] = np.asarray(leaf)
    return tuple(np.asarray(p) for p in params[:3]), {'params': tree}

  def from_flax_variables(self, heads: Sequence[np.ndarray], variables: dict) -> tuple[np.ndarray, ...]:
    """Inverse of :meth:`to_flax_variables`: leaves are taken in `jax.tree_util.tree_leaves` order
    of the dict, i.e. keys sorted as strings at every level ('...scale10' before '...scale2'),
    which is the order of the params tuple (models.py:99-103)."""
    tree = variables.get('params', variables)

    def leaves(node, prefix=''):
      for key in sorted(node):
        if isinstance(node[key], dict):
          yield from leaves(node[key], prefix + key + '/')
        else:
          yield prefix + key, np.asarray(node[key])

    named = list(leaves(tree))
    if [n for n, _ in named] != self.leaf_names:
      raise ValueError('Flax variables do not match this model: expected leaves '
                       f'{self.leaf_names}, got {[n for n, _ in named]}')
    for (name, leaf), shape in zip(named, self.leaf_shapes):
      if tuple(leaf.shape[leaf.ndim - len(shape):]) != tuple(shape):
        raise ValueError(f'leaf {name}: trailing shape {leaf.shape} does not end with {shape}')
    if len(heads) != 3:
      raise ValueError('heads must be (log_noise_scale, shape, inflated_loc_probs)')
    return tuple(np.asarray(h) for h in heads) + tuple(leaf for _, leaf in named)

