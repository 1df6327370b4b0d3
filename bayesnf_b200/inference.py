"""Inference drivers: the reference's function seam over libbnf_sm100.so.

Mirrors ``bayesnf.inference`` (src/bayesnf/inference.py) for the hot path:

* ``fit_map``      inference.py:376-458 (+ ``ensemble_map`` :510-623)
* ``fit_vi``       inference.py:336-373 (+ ``ensemble_vi`` :626-764)
* ``predict_bnf``  inference.py:461-507 (+ batched forecast :103-200,
                   mixture quantiles :42-100)

Same argument names/meaning, same return structure (host numpy arrays with
leading ``(num_devices, members_per_device, ...)`` axes), same error behaviour.
PyTorch is only the device-buffer container / stream / process-group plumbing:
every FLOP of the model runs in the CUDA kernels behind the C ABI.

Differences that are deliberate and documented in DESIGN.md:
* "device" = one process (torch.distributed rank) per GPU; each rank fits and
  returns ITS members (leading axis 1); ``predict_bnf`` all-gathers predictions.
* ``seed`` may be an int or a JAX-style ``uint32[2]`` key; random streams are
  device Philox, not threefry, so results are statistically -- not bitwise --
  those of the reference for the same seed.
"""

from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Sequence

import numpy as np
import torch

from . import _lib
from . import models
from . import parallel

ArrayT = np.ndarray

# 'bf16x3': tcgen05 tensor cores at f32-class accuracy (operands as bf16 triples, six products per
#           GEMM, f32 pre-activations) -- meets the 1e-5 parity bar; the default.
# 'bf16'  : single-pass bf16 tensor cores (throughput mode, ~1e-2 of scale).
# 'fp32'  : SIMT f32 FMA (no tensor cores; any width).  'bf16_simt': debug.
# 'tf32x3' is accepted as an alias of 'bf16x3' (the split-operand mode under its generic name).
_PRECISIONS = {'fp32': _lib.PREC_FP32, 'bf16': _lib.PREC_BF16,
               'bf16_simt': _lib.PREC_BF16_SIMT, 'bf16x3': _lib.PREC_BF16X3,
               'tf32x3': _lib.PREC_BF16X3}
_default_precision = os.environ.get('BAYESNF_B200_PRECISION', 'bf16x3')


def set_default_precision(name: str) -> None:
  """'bf16x3' (tensor cores, <=1e-5 parity), 'bf16' (tensor cores, throughput) or 'fp32' (SIMT)."""
  global _default_precision
  if name not in _PRECISIONS:
    raise ValueError(f'unknown precision {name!r}')
  _default_precision = name


def get_default_precision() -> str:
  return _default_precision


def precision_supported(spec: models.ModelSpec, name: str) -> bool:
  """Whether the arithmetic mode `name` covers this model shape (`bnf_precision_supported`)."""
  if name not in _PRECISIONS:
    raise ValueError(f'unknown precision {name!r}')
  return _lib.lib.bnf_precision_supported(spec.plan, _PRECISIONS[name]) == 0


def _device() -> torch.device:
  if not torch.cuda.is_available():
    raise _lib.BnfError(
        'bayesnf_b200 needs a CUDA (sm_100) device: there is no CPU fallback.')
  return torch.device('cuda', torch.cuda.current_device())


def _ptr(t: torch.Tensor | None):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def seed_to_int(seed) -> int:
  """Accept an int or a JAX PRNGKey-like uint32[2] array."""
  if isinstance(seed, (int, np.integer)):
    return int(seed) & 0xFFFFFFFFFFFFFFFF
  a = np.asarray(seed).astype(np.uint64).ravel()
  if a.size == 1:
    return int(a[0])
  if a.size != 2:
    raise ValueError('seed must be an int or a uint32[2] key')
  return (int(a[0]) << 32) | int(a[1])


def fold_in(seed: int, data: int) -> int:
  """Derive an independent 64-bit seed (splitmix64 finaliser)."""
  z = (seed + 0x9E3779B97F4A7C15 * (data + 1)) & 0xFFFFFFFFFFFFFFFF
  z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
  z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
  return z ^ (z >> 31)


class Engine:
  """Device buffers + C calls for one ModelSpec on the current CUDA device."""

  def __init__(self, spec: models.ModelSpec, precision: str | None = None):
    self.spec = spec
    self.precision_name = precision or _default_precision
    if self.precision_name not in _PRECISIONS:
      raise ValueError(f'unknown precision {self.precision_name!r}')
    self.prec = _PRECISIONS[self.precision_name]
    if precision is None and _lib.lib.bnf_precision_supported(spec.plan, self.prec) != 0:
      # the DEFAULT tensor-core mode does not cover this shape (width not in {64..1024 powers of
      # two} or > 128 features): the f32 SIMT CUDA kernels do.  An explicitly requested mode
      # fails loudly instead.
      self.precision_name, self.prec = 'fp32', _lib.PREC_FP32
    self.device = _device()
    self._ws: dict[tuple, torch.Tensor] = {}

  # Workspaces are cached per (mode, networks, rows).  fit() / predict() alternate between a few
  # shapes (training step, forecast slab, ragged last slab), so a handful stays alive; the oldest
  # go when the cache holds more than `_WS_MAX_ENTRIES` buffers or `_WS_MAX_FRACTION` of the
  # device memory.
  _WS_MAX_ENTRIES = 4
  _WS_MAX_FRACTION = 0.5

  def workspace(self, mode: int, n_net: int, rows: int) -> torch.Tensor:
    key = (mode, n_net, rows)
    ws = self._ws.pop(key, None)
    if ws is None:
      nbytes = _lib.lib.bnf_workspace_bytes(self.spec.plan, self.prec, n_net, rows, mode)
      if nbytes == 0:
        raise ValueError('invalid workspace request')
      cap = self._WS_MAX_FRACTION * torch.cuda.get_device_properties(self.device).total_memory
      while self._ws and (len(self._ws) >= self._WS_MAX_ENTRIES or
                          nbytes + sum(t.numel() for t in self._ws.values()) > cap):
        self._ws.pop(next(iter(self._ws)))          # least recently used first
      ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
    self._ws[key] = ws                               # (re)insert as most recently used
    return ws

  def forward_slab_rows(self, n_net: int, budget_bytes: int | None = None) -> int:
    """Rows per forecast slab so that the forward workspace of `n_net` networks stays within a
    memory budget (the reference forecasts in 1024-row batches for the same reason,
    inference.py:129-181): clamp(budget // bytes_per_row, 128, 65536).  (Up to 65 536 rows per call --
    the batch size of the largest training configuration: at 16 384 the 1 Mi-row forecast of the bench's
    predict block was 64 calls of ~150 us of kernels each and bound by the host's launch rate.)"""
    if budget_bytes is None:
      free, _ = torch.cuda.mem_get_info(self.device)
      budget_bytes = min(8 << 30, free // 4)
    probe = 1024
    per_row = _lib.lib.bnf_workspace_bytes(self.spec.plan, self.prec, n_net, probe, _lib.WS_FORWARD) / probe
    rows = int(budget_bytes // max(per_row, 1.0))
    return max(128, min(65536, rows // 128 * 128))

  # ---- mlp.apply over networks (forecast_inner, inference.py:103-126) ----
  def forward(self, params: torch.Tensor, x: torch.Tensor, slab: int | None = None) -> torch.Tensor:
    """params [M,P] f32, x [N,D] f32 -> loc [M,N] f32 (row slabs like :129-181)."""
    M, N = params.shape[0], x.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=self.device)
    if slab is None:
      slab = self.forward_slab_rows(M)
    slab = max(1, min(slab, N))
    if slab == N:                                     # one slab: the kernels write `out` directly
      ws = self.workspace(_lib.WS_FORWARD, M, N)
      _lib.check(_lib.lib.bnf_forward(
          self.spec.plan, self.prec, _ptr(params), M, _ptr(x), None, 0, N,
          _ptr(out), _ptr(ws), ws.numel(), _stream()))
      return out
    tmp = torch.empty((M, slab), dtype=torch.float32, device=self.device)
    for s in range(0, N, slab):
      rows = min(slab, N - s)
      ws = self.workspace(_lib.WS_FORWARD, M, rows)
      xs = x[s:s + rows]
      _lib.check(_lib.lib.bnf_forward(
          self.spec.plan, self.prec, _ptr(params), M, _ptr(xs), None, 0, rows,
          _ptr(tmp), _ptr(ws), ws.numel(), _stream()))
      out[:, s:s + rows] = tmp.view(-1)[:M * rows].view(M, rows)
    return out

  def loglik_grad(self, params, x, y, idx=None, rows=None, want_grad=True):
    """-> (loglik [M], grad [M,P] | None).  idx: int32 [M or 1, rows] or None."""
    M = params.shape[0]
    if idx is not None:
      rows = idx.shape[-1]
      stride = idx.shape[-1] if idx.shape[0] > 1 else 0
    else:
      rows = rows or x.shape[0]
      stride = 0
    ll = torch.empty(M, dtype=torch.float32, device=self.device)
    grad = torch.empty((M, self.spec.num_params), dtype=torch.float32,
                       device=self.device) if want_grad else None
    ws = self.workspace(_lib.WS_GRAD, M, rows)
    _lib.check(_lib.lib.bnf_loglik_grad(
        self.spec.plan, self.prec, _ptr(params), M, _ptr(x), _ptr(y), _ptr(idx), stride,
        rows, _ptr(ll), _ptr(grad), _ptr(ws), ws.numel(), _stream()))
    return ll, grad

  def map_steps(self, params, adam_m, adam_v, step_count, x, y, idx, rows, n_total,
                n_steps, lr, prior_weight) -> torch.Tensor:
    M = params.shape[0]
    losses = torch.empty((n_steps, M), dtype=torch.float32, device=self.device)
    ws = self.workspace(_lib.WS_MAP, M, rows)
    stride = idx.shape[-1] if (idx is not None and idx.shape[0] > 1) else 0
    _lib.check(_lib.lib.bnf_map_steps(
        self.spec.plan, self.prec, _ptr(params), _ptr(adam_m), _ptr(adam_v),
        _ptr(step_count), M, _ptr(x), _ptr(y), _ptr(idx), stride, rows, n_total, n_steps,
        lr, prior_weight, _ptr(losses), _ptr(ws), ws.numel(), _stream()))
    return losses

  def map_epochs(self, params, adam_m, adam_v, step_count, x, y, rows, n_total, n_epochs, lr,
                 prior_weight, seed, first_member) -> torch.Tensor:
    """Minibatch epochs with per-member permutations drawn on the device (bnf_map_epochs):
    -> losses [n_epochs * (n_total // rows), M]."""
    M = params.shape[0]
    steps = n_epochs * (n_total // rows)
    losses = torch.empty((steps, M), dtype=torch.float32, device=self.device)
    ws = self.workspace(_lib.WS_MAP, M, rows)
    _lib.check(_lib.lib.bnf_map_epochs(
        self.spec.plan, self.prec, _ptr(params), _ptr(adam_m), _ptr(adam_v), _ptr(step_count), M,
        _ptr(x), _ptr(y), rows, n_total, n_epochs, lr, prior_weight, C.c_uint64(seed), first_member,
        _ptr(losses), _ptr(ws), ws.numel(), _stream()))
    return losses

  def vi_steps(self, mu, rho, adam_m, adam_v, step_count, n_mc, seed, device_id, x, y, rows, n_total,
               n_steps, lr, kl_weight) -> torch.Tensor:
    """n_steps VI steps, eps and the per-step shared sub-batch drawn on the device (bnf_vi_steps):
    -> losses [n_steps, E] (not yet times kl_weight)."""
    E = mu.shape[0]
    losses = torch.empty((n_steps, E), dtype=torch.float32, device=self.device)
    ws = self.workspace(_lib.WS_VI, E * n_mc, rows)
    _lib.check(_lib.lib.bnf_vi_steps(
        self.spec.plan, self.prec, _ptr(mu), _ptr(rho), _ptr(adam_m), _ptr(adam_v), _ptr(step_count),
        E, n_mc, C.c_uint64(seed), device_id, _ptr(x), _ptr(y), rows, n_total, n_steps, lr, kl_weight,
        _ptr(losses), _ptr(ws), ws.numel(), _stream()))
    return losses

  def vi_step(self, mu, rho, adam_m, adam_v, step_count, n_mc, eps, seed, x, y, idx,
              rows, n_total, lr, kl_weight, out_loss):
    E = mu.shape[0]
    ws = self.workspace(_lib.WS_VI, E * n_mc, rows)
    _lib.check(_lib.lib.bnf_vi_step(
        self.spec.plan, self.prec, _ptr(mu), _ptr(rho), _ptr(adam_m), _ptr(adam_v),
        _ptr(step_count), E, n_mc, _ptr(eps), C.c_uint64(seed), _ptr(x), _ptr(y), _ptr(idx),
        rows, n_total, lr, kl_weight, _ptr(out_loss), _ptr(ws), ws.numel(), _stream()))

  def vi_sample(self, mu, rho, n_samples, seed, eps=None) -> torch.Tensor:
    E = mu.shape[0]
    out = torch.empty((n_samples, E, self.spec.num_params), dtype=torch.float32,
                      device=self.device)
    _lib.check(_lib.lib.bnf_vi_sample(self.spec.plan, _ptr(mu), _ptr(rho), E, n_samples,
                                      _ptr(eps), C.c_uint64(seed), _ptr(out), _stream()))
    return out

  def init_params(self, lns_init: float, seed: int, first_member: int, n: int) -> torch.Tensor:
    out = torch.empty((n, self.spec.num_params), dtype=torch.float32, device=self.device)
    _lib.check(_lib.lib.bnf_init_params(self.spec.plan, lns_init, C.c_uint64(seed),
                                        first_member, n, _ptr(out), _stream()))
    return out


def mixture_quantiles(means: torch.Tensor, scales: torch.Tensor, quantiles: Sequence[float],
                      approximate: bool) -> torch.Tensor:
  """means [M,N], scales [M] (device f32) -> [len(q), N]."""
  M, N = means.shape
  q = (C.c_double * len(quantiles))(*[float(v) for v in quantiles])
  out = torch.empty((len(quantiles), N), dtype=torch.float32, device=means.device)
  ws = torch.empty(256, dtype=torch.uint8, device=means.device)
  _lib.check(_lib.lib.bnf_mixture_quantiles(
      _ptr(means.contiguous()), _ptr(scales.contiguous()), M, N, q, len(quantiles),
      1 if approximate else 0, _ptr(out), _ptr(ws), ws.numel(), _stream()))
  return out


def nb_mixture_quantiles(loc: torch.Tensor, shape_raw: torch.Tensor, pi_logit: torch.Tensor | None,
                         quantiles: Sequence[float]) -> tuple[torch.Tensor, torch.Tensor]:
  """loc [M,N], shape_raw [M], pi_logit [M]|None -> (means [M,N], quantiles [len(q),N])."""
  M, N = loc.shape
  q = (C.c_double * len(quantiles))(*[float(v) for v in quantiles])
  means = torch.empty((M, N), dtype=torch.float32, device=loc.device)
  out = torch.empty((len(quantiles), N), dtype=torch.float32, device=loc.device)
  ws = torch.empty(256, dtype=torch.uint8, device=loc.device)
  _lib.check(_lib.lib.bnf_nb_mixture_quantiles(
      _ptr(loc.contiguous()), _ptr(shape_raw.contiguous()),
      _ptr(pi_logit.contiguous() if pi_logit is not None else None), M, N, q, len(quantiles),
      _ptr(means), _ptr(out), _ptr(ws), ws.numel(), _stream()))
  return means, out


def _to_device_data(features, target=None):
  dev = _device()
  # jnp.array(...) with x64 disabled: float64 pandas values -> float32 (inference.py:553-554)
  x = torch.as_tensor(np.ascontiguousarray(np.asarray(features, dtype=np.float64)
                                           ).astype(np.float32)).to(dev)
  if x.ndim == 1:
    x = x[:, None]
  y = None
  if target is not None:
    y = torch.as_tensor(np.asarray(target, dtype=np.float64).astype(np.float32)).to(dev)
  return x, y


def device_permutation(seed: int, member: int, epoch: int, n_rows: int) -> np.ndarray:
  """The row order the device draws for (seed; member, epoch) -- permute_dataset
  (inference.py:35-39) as `bnf_map_epochs` evaluates it, computed on the host (tests / replay)."""
  out = (C.c_int32 * n_rows)()
  _lib.check(_lib.lib.bnf_debug_permutation(C.c_uint64(seed), member, epoch, n_rows, out))
  return np.frombuffer(out, dtype=np.int32).copy()


def fit_map(
    features: ArrayT,
    target: ArrayT,
    seed,
    observation_model: str,
    model_args: dict[str, Any],
    num_particles: int,
    learning_rate: float,
    num_epochs: int,
    prior_weight: float = 1.0,
    batch_size: int | None = None,
    num_splits: int = 1,
    precision: str | None = None,
    init_params: np.ndarray | None = None,
    batch_indices: np.ndarray | None = None,
    batch_order: str = 'device',
) -> tuple[tuple[np.ndarray, ...], np.ndarray]:
  """Fit a BNF ensemble by MAP (prior_weight=1) or MLE (prior_weight=0).

  ``batch_order='jax'`` replays the reference's minibatch row orders for this ``seed``
  (threefry key tree of inference.py:571-618 restated in `jax_prng`; host-generated, so meant for
  comparisons, not for throughput); the default draws them on the device.

  Reference: inference.py:376-458.  ``init_params`` ([members, P]) and
  ``batch_indices`` ([epochs, members, N] row orders) are test hooks that inject
  the initial parameters / batch order so the run can be compared with the CPU
  oracle; by default both come from device RNG streams derived from ``seed``.
  Returns (params tuple with leading (1, members_on_this_rank * ...) axes,
  losses (1, members, num_epochs)).
  """
  spec = models.ModelSpec(**model_args, observation_model=observation_model)
  eng = Engine(spec, precision)
  x, y = _to_device_data(features, target)
  n_total = y.shape[0]
  if batch_size is None:
    batch_size = n_total
  n_dev, rank = parallel.device_count(), parallel.device_index()
  members = (num_particles // num_splits) // n_dev
  if members < 1:
    raise ValueError('ensemble_size cannot be smaller than device_count.')
  steps_per_epoch = n_total // batch_size
  if steps_per_epoch < 1:
    # the reference scans over zero steps here: initial parameters, NaN epoch losses
    # (mean of an empty array, inference.py:583-614).  Same result, with a warning.
    import warnings
    warnings.warn(f'{batch_size=} exceeds the number of rows {n_total}: no training step is taken '
                  '(as in the reference)', RuntimeWarning)
  seed = seed_to_int(seed)
  target_scale = float(np.nanstd(np.asarray(target, dtype=np.float64)))
  lns_init = math.log(target_scale / 2.0)

  if batch_order not in ('device', 'jax'):
    raise ValueError(f'{batch_order=}')
  params_out, losses_out = [], []
  for i in range(num_splits):
    if batch_order == 'jax' and batch_indices is None and batch_size < n_total:
      from . import jax_prng
      orders = jax_prng.map_batch_orders(jax_prng.prng_key(seed), n_dev, members, n_total, num_epochs,
                                         split_index=i if num_splits > 1 else None)
      jax_orders = orders[:, rank]                        # [epochs, members, N] of this rank
    else:
      jax_orders = None
    seed_i = fold_in(seed, i) if num_splits > 1 else seed
    init_seed, opt_seed = fold_in(seed_i, 0x1001), fold_in(seed_i, 0x1002)
    if init_params is not None:
      p = torch.as_tensor(np.asarray(init_params, dtype=np.float32)).to(eng.device)
      p = p.reshape(-1, spec.num_params)[i * members:(i + 1) * members].contiguous().clone()
      if p.shape[0] != members:
        raise ValueError('init_params has the wrong number of members')
    else:
      p = eng.init_params(lns_init, init_seed, rank * members, members)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    step_count = torch.zeros(1, dtype=torch.int32, device=eng.device)
    if steps_per_epoch < 1:
      epoch_losses = torch.full((num_epochs, members), float('nan'), device=eng.device)
    elif batch_size >= n_total and batch_indices is None:
      losses = eng.map_steps(p, m, v, step_count, x, y, None, batch_size, n_total,
                             num_epochs, learning_rate, prior_weight)       # [epochs, M]
      epoch_losses = losses
    elif batch_indices is None and jax_orders is None:
      # the production path: per-member permutations, batch windows and every step on the device,
      # one call (and one CUDA graph) for all epochs
      ls = eng.map_epochs(p, m, v, step_count, x, y, batch_size, n_total, num_epochs, learning_rate,
                          prior_weight, opt_seed, rank * members)
      epoch_losses = ls.view(num_epochs, steps_per_epoch, members).mean(dim=1)   # :614
    else:
      per_epoch = []
      for ep in range(num_epochs):
        if batch_indices is not None:
          perm = torch.as_tensor(np.asarray(batch_indices[ep], dtype=np.int32)).to(eng.device)
          perm = perm.reshape(-1, n_total)[i * members:(i + 1) * members].contiguous()
        else:
          perm = torch.as_tensor(jax_orders[ep]).to(eng.device).contiguous()
        ls = eng.map_steps(p, m, v, step_count, x, y, perm, batch_size, n_total,
                           steps_per_epoch, learning_rate, prior_weight)
        per_epoch.append(ls.mean(dim=0))                                   # :614
      epoch_losses = torch.stack(per_epoch)
    params_out.append(p.cpu().numpy()[None])                # (1, members, P)
    losses_out.append(epoch_losses.t().cpu().numpy()[None])  # (1, members, epochs)
  flat = np.concatenate(params_out, axis=1)
  losses = np.concatenate(losses_out, axis=1)
  return spec.unflatten(flat), losses


class SurrogatePosterior:
  """Mean-field Normal surrogate q = prod N(mu, 1e-4 + softplus(rho)).

  Stand-in for the tfd.JointDistribution the reference returns
  (inference.py:760-764); holds the variational parameters as reference-ordered
  tuples with leading (1, members) axes.
  """

  def __init__(self, spec, mu: np.ndarray, rho: np.ndarray):
    self.loc = spec.unflatten(mu)
    self.inv_softplus_scale = spec.unflatten(rho)

  def stddev(self):
    return tuple(1e-4 + np.logaddexp(r, 0.0) for r in self.inv_softplus_scale)

  def mean(self):
    return self.loc


def fit_vi(
    features: ArrayT,
    target: ArrayT,
    seed,
    observation_model: str,
    model_args: dict[str, Any],
    ensemble_size: int,
    learning_rate: float,
    num_epochs: int,
    sample_size_divergence: int,
    sample_size_posterior: int,
    kl_weight: float,
    batch_size: int | None = None,
    precision: str | None = None,
    init_params: tuple[np.ndarray, np.ndarray] | None = None,
    eps: np.ndarray | None = None,
    posterior_eps: np.ndarray | None = None,
    batch_indices: np.ndarray | None = None,
) -> tuple[SurrogatePosterior, np.ndarray, tuple[np.ndarray, ...]]:
  """Fit an ensemble of mean-field surrogate posteriors (inference.py:336-373,
  :626-764).  ``num_epochs`` is the number of optimisation STEPS, as in the
  reference.  Test hooks: ``init_params`` = (mu, rho) [members, P]; ``eps``
  [steps, S, members, P]; ``posterior_eps`` [num_samples, members, P]; ``batch_indices``
  [steps, batch_size] shared sub-batch rows.
  Returns (surrogate, losses (1, members, steps) already times kl_weight,
  posterior samples tuple with leading (1, num_samples, members)).
  """
  spec = models.ModelSpec(**model_args, observation_model=observation_model)
  eng = Engine(spec, precision)
  x, y = _to_device_data(features, target)
  n_total = y.shape[0]
  if batch_size is not None:
    assert n_total >= batch_size, f'{batch_size=} exceeds {n_total=}'
  rows = batch_size if batch_size is not None else n_total
  n_dev, rank = parallel.device_count(), parallel.device_index()
  members = ensemble_size // n_dev
  if members < 1:
    raise ValueError('ensemble_size cannot be smaller than device_count.')
  seed = seed_to_int(seed)
  init_seed, opt_seed = fold_in(seed, 0x2001), fold_in(seed, 0x2002)
  fit_seed, sample_seed = fold_in(opt_seed, 1), fold_in(opt_seed, 2)
  P = spec.num_params
  if init_params is not None:
    mu = torch.as_tensor(np.asarray(init_params[0], np.float32)).to(eng.device).reshape(members, P).clone()
    rho = torch.as_tensor(np.asarray(init_params[1], np.float32)).to(eng.device).reshape(members, P).clone()
  else:
    mu = eng.init_params(0.0, init_seed, rank * members, members)   # :203-231, lns mean = 0
    rho = torch.full((members, P), math.log(math.expm1(0.3)), dtype=torch.float32,
                     device=eng.device)
  am = torch.zeros((members, 2, P), dtype=torch.float32, device=eng.device)
  av = torch.zeros_like(am)
  step_count = torch.zeros(1, dtype=torch.int32, device=eng.device)
  if eps is None and batch_indices is None:
    # the production path: eps, the per-step shared sub-batch (inference.py:704-709) and every step
    # on the device, one call (and one CUDA graph) for all steps
    losses = eng.vi_steps(mu, rho, am, av, step_count, sample_size_divergence,
                          (fit_seed + rank) & 0xFFFFFFFFFFFFFFFF, rank, x, y, rows, n_total, num_epochs,
                          learning_rate, kl_weight)
  else:
    # test hooks: injected eps [steps, S, members, P] and / or sub-batch rows [steps, rows]
    losses = torch.empty((num_epochs, members), dtype=torch.float32, device=eng.device)
    for step in range(num_epochs):
      idx = None
      if batch_indices is not None:
        idx = torch.as_tensor(np.asarray(batch_indices[step], np.int32)).to(eng.device).reshape(1, -1).contiguous()
      elif rows < n_total:
        idx = torch.as_tensor(device_permutation((fit_seed + rank) & 0xFFFFFFFFFFFFFFFF ^ 0x5649424154434855,
                                                 rank, step, n_total)[:rows]).to(eng.device)[None].contiguous()
      eps_t = None
      if eps is not None:
        eps_t = torch.as_tensor(np.asarray(eps[step], np.float32)).to(eng.device).contiguous()
      eng.vi_step(mu, rho, am, av, step_count, sample_size_divergence, eps_t,
                  (fit_seed + rank) & 0xFFFFFFFFFFFFFFFF, x, y, idx, rows, n_total,
                  learning_rate, kl_weight, losses[step])
  pe = None
  if posterior_eps is not None:
    pe = torch.as_tensor(np.asarray(posterior_eps, np.float32)).to(eng.device).contiguous()
  samples = eng.vi_sample(mu, rho, sample_size_posterior,
                          (sample_seed + rank) & 0xFFFFFFFFFFFFFFFF, pe)
  surrogate = SurrogatePosterior(spec, mu.cpu().numpy()[None], rho.cpu().numpy()[None])
  # inference.py:758: transpose to (devices, members, steps) and multiply by kl_weight
  losses_np = losses.t().cpu().numpy()[None] * kl_weight
  return surrogate, losses_np, spec.unflatten(samples.cpu().numpy()[None])


def forward_bnf(features: ArrayT, observation_model: str, params: Sequence[np.ndarray],
                model_args: dict[str, Any], precision: str | None = None) -> np.ndarray:
  """`mlp.apply` of every member on `features` (the call inside models.make_likelihood_model,
  models.py:157-160): returns the network outputs with shape `params[0].shape + (N,)`."""
  spec = models.ModelSpec(**model_args, observation_model=observation_model)
  eng = Engine(spec, precision)
  x, _ = _to_device_data(features)
  flat = spec.flatten(params)
  nets = torch.as_tensor(flat.reshape(-1, spec.num_params)).to(eng.device).contiguous()
  loc = eng.forward(nets, x)
  return loc.reshape(tuple(flat.shape[:-1]) + (loc.shape[1],)).cpu().numpy()


def predict_bnf(
    features: ArrayT,
    observation_model: str,
    params: Sequence[np.ndarray],
    model_args: dict[str, Any],
    quantiles: Sequence[float],
    ensemble_dims: int = 2,
    approximate_quantiles: bool = False,
    precision: str | None = None,
) -> tuple[np.ndarray, list[np.ndarray]]:
  """Predict new data from an existing BNF fit (inference.py:461-507).

  Returns (means (num_devices, [num_samples,] members, N), [quantile (N,), ...]).
  The forward pass runs on this rank's members only; the predictive parameters
  are all-gathered once (NCCL) and the mixture quantiles are then computed over
  every member of every rank.
  """
  assert ensemble_dims >= 1
  spec = models.ModelSpec(**model_args, observation_model=observation_model)
  eng = Engine(spec, precision)
  x, _ = _to_device_data(features)
  flat = spec.flatten(params)                       # (1, [S,] E, P)
  lead = flat.shape[:-1]
  nets = torch.as_tensor(flat.reshape(-1, spec.num_params)).to(eng.device).contiguous()
  loc = eng.forward(nets, x)                        # [M_local, N]
  dist = spec.distribution
  if dist == models.LikelihoodDist.NORMAL:
    scales = 0.01 + torch.exp(nets[:, 0])           # models.py:163
    loc_all = parallel.all_gather_leading(loc)      # (world, M_local, N)  <- the collective
    sc_all = parallel.all_gather_leading(scales)
    world = loc_all.shape[0]
    q = mixture_quantiles(loc_all.reshape(-1, loc.shape[1]), sc_all.reshape(-1),
                          list(quantiles), approximate_quantiles)
    means = loc_all.reshape((world,) + tuple(lead[1:]) + (loc.shape[1],)).cpu().numpy()
    return means, [q[i].cpu().numpy() for i in range(len(quantiles))]
  # NB / ZINB (inference.py:271-333): distribution means + mixture quantiles from the gathered
  # network outputs and the per-member shape / zero-inflation parameters.
  loc_all = parallel.all_gather_leading(loc)
  shape_all = parallel.all_gather_leading(nets[:, 1].contiguous())
  pi_all = parallel.all_gather_leading(nets[:, 2].contiguous())
  world, n_pts = loc_all.shape[0], loc.shape[1]
  means_t, q = nb_mixture_quantiles(
      loc_all.reshape(-1, n_pts), shape_all.reshape(-1),
      pi_all.reshape(-1) if dist == models.LikelihoodDist.ZINB else None, list(quantiles))
  means = means_t.reshape((world,) + tuple(lead[1:]) + (n_pts,)).cpu().numpy()
  return means, [q[i].cpu().numpy() for i in range(len(quantiles))]
