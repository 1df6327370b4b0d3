"""CPU oracle: restatement of the BayesNF hot path in torch-CPU (float32 / float64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``bayesnf_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
reported CPU baseline -- never as the thing shipped.

PARITY PINNED against the reference's own code executed here.  The reference is pure
JAX/Flax/TFP/Optax and none of those packages is installable in the build container (no
network, wheelhouse has no jax), and its only numeric goldens
(tests/test_data/bnf-*.mini.pred.csv) are skipped upstream and depend on JAX threefry streams.
Instead, scripts/make_golden_numerics.py imports the reference's models.py / inference.py
UNMODIFIED from /root/reference and executes them over oracle/jaxshim.py, a float64 stand-in for
the third-party calls they make (array ops, vmap / scan / value_and_grad, Flax Module.param and
Dense, optax.adam, the TFP log_prob formulas, checked against scipy).  The committed outputs
(tests/golden/numerics_*.npz: model forward, NORMAL / NB / ZINB log-likelihoods, prior, their
gradients, fit_map / fit_vi / predict_bnf end to end) are reproduced by this module to float64
round-off (tests/test_reference_goldens.py), and the generator is re-run against the committed
files whenever /root/reference is present.  What that does NOT pin: the third-party primitives
themselves (restated in the shim from their documented behaviour) and the random streams of the
TFP samplers (initial draws and VI noise are recorded in the goldens instead).  Also pinned (see
tests/test_oracle_pins.py): the pure numpy/pandas bookkeeping of the reference executed from
/root/reference by scripts/make_golden.py (seasonal frequencies, harmonics, data-handler
outputs, parameter count, feature ordering) and the analytic sigma check on the reference's
mini MAP/MLE golden predictions.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Third-party semantics (Flax Dense, TFP log_probs,
optax.adam, tfp.vi.fit_surrogate_posterior_stateless) are restated from the
documented behaviour of the pinned versions (requirements.Python3.10.14.txt:
flax 0.8.3, jax 0.4.26, optax 0.2.2, tensorflow-probability 0.24.0).
"""

from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

NORMAL, NB, ZINB = 'NORMAL', 'NB', 'ZINB'
_TWO_PI = 2 * math.pi


# --------------------------------------------------------------------------
# models.py:36-59  make_seasonal_frequencies
# --------------------------------------------------------------------------
def make_seasonal_frequencies(seasonality_periods, num_harmonics):
  """f32 harmonics/period, de-duplicated keeping FIRST occurrences in order."""
  periods = np.array(seasonality_periods, dtype=np.float32)
  num_harmonics = np.asarray(num_harmonics)
  if np.any(num_harmonics > periods / 2):
    raise ValueError('Harmonic cannot exceed half seasonal period.')
  if periods.shape != num_harmonics.shape:
    raise ValueError('Number of seasonal periods and harmonics must be equal.')
  if num_harmonics.ndim != 1:
    raise ValueError('`num_harmonics` and `seasonality_periods` must be rank 1.')
  if periods.shape[0] == 0:
    return np.zeros(0), np.zeros(0)
  harmonics = [np.arange(1, h + 1, dtype=np.float32) for h in num_harmonics]
  freqs = np.concatenate([h / p for h, p in zip(harmonics, periods)])
  _, first = np.unique(freqs, return_index=True)
  keep = np.sort(first)
  return freqs[keep], np.concatenate(harmonics)[keep]


def _softplus(x):
  # jax.nn.softplus = logaddexp(x, 0) = max(x,0) + log1p(exp(-|x|)); torch's
  # logaddexp evaluates the same expression and has the smooth derivative
  # sigmoid(x) at x == 0 (a clamp/abs formulation would give a kink there).
  return torch.logaddexp(x, torch.zeros_like(x))


class OracleModel:
  """models.py:197-273 BayesianNeuralField1D + inference.py:234-258 make_model.

  Parameters are handled exactly like the reference: a list
  ``[log_noise_scale, shape, inflated_loc_probs, *leaves]`` with leaves in the
  order jax.tree_util.tree_leaves gives for the Flax dict (sorted keys).
  """

  def __init__(self, width, depth, input_scales, num_seasonal_harmonics,
               seasonality_periods, init_x, fourier_degrees, interactions,
               dtype=torch.float32):
    self.width, self.depth = int(width), int(depth)
    self.dtype = dtype
    self.D = int(init_x[-1]) if len(init_x) > 1 else 1
    self.input_scales = np.asarray(input_scales, dtype=np.float64)
    self.fourier_degrees = np.asarray(fourier_degrees).astype(int)
    self.interactions = np.asarray(interactions).astype(int).reshape(-1, 2)
    self.freqs, self.harmonics = make_seasonal_frequencies(
        seasonality_periods, np.asarray(num_seasonal_harmonics))
    # --- feature groups in models.py:242-251 order; index BEFORE the size filter.
    groups = [('x', self.D, None)]
    for i, deg in enumerate(self.fourier_degrees):
      if deg > 0:                       # models.py:230-234 (filtered first)
        groups.append(('fourier', 2 * int(deg), (i, int(deg))))
    groups.append(('seasonal', 2 * len(self.freqs), None))
    groups.append(('inter', len(self.interactions), None))
    self.groups = [(idx, kind, n, meta)
                   for idx, (kind, n, meta) in enumerate(groups) if n > 0]
    self.F = sum(n for _, _, n, _ in self.groups)
    # --- Flax param dict -> sorted leaves (models.py:99, tree_leaves).
    shapes = {}
    fan = self.F
    for l in range(self.depth):
      shapes[f'Dense_{l}/bias'] = (self.width,)
      shapes[f'Dense_{l}/kernel'] = (fan, self.width)
      shapes[f'inv_sp_layer_scale{l}'] = ()
      fan = self.width
    shapes[f'Dense_{self.depth}/bias'] = (1,)
    shapes[f'Dense_{self.depth}/kernel'] = (fan, 1)
    shapes['inv_sp_output_scale'] = ()
    shapes['log_scale_adjustment'] = (self.D,)
    shapes['logit_activation_weight'] = ()
    for idx, _, _, _ in self.groups:
      shapes[f'feature_inv_sp_scale{idx}'] = ()

    def sort_key(name):          # nested dict: sort top-level key, then leaf key
      parts = name.split('/')
      return (parts[0], parts[1] if len(parts) > 1 else '')
    self.leaf_names = sorted(shapes, key=sort_key)
    self.leaf_shapes = [shapes[n] for n in self.leaf_names]
    self.num_params = 3 + sum(int(np.prod(s)) for s in self.leaf_shapes)

  # ---- flat <-> list helpers (test convenience, not in the reference) ----
  def unflatten(self, flat):
    out = [flat[..., 0], flat[..., 1], flat[..., 2]]
    o = 3
    for s in self.leaf_shapes:
      n = int(np.prod(s))
      out.append(flat[..., o:o + n].reshape(flat.shape[:-1] + tuple(s)))
      o += n
    return out

  def flatten(self, params):
    lead = params[0].shape
    return torch.cat([p.reshape(lead + (-1,)) for p in params], -1)

  def leaf(self, params, name):
    return params[3 + self.leaf_names.index(name)]

  # ---- models.py:62-88 features ----
  def seasonal_features(self, t):
    if len(self.freqs) == 0:
      return t.new_zeros(t.shape[0], 0)
    # `2 * jnp.pi * frequencies` is a python float times the float32 numpy table of
    # make_seasonal_frequencies: numpy evaluates it in float32 (weak scalar), in the reference as
    # here; the float64 twin keeps that rounded table and is exact from there on (this is what
    # the reference's own code gives when executed in float64, tests/golden/numerics_model_*).
    w = torch.from_numpy(np.float32(_TWO_PI) * self.freqs.astype(np.float32))
    y = w.to(self.dtype) * t.reshape(-1, 1)          # (2*pi*f) * x, models.py:73
    feats = torch.cat([torch.cos(y), torch.sin(y)], 1)
    den = torch.from_numpy(np.tile(self.harmonics, 2)).to(self.dtype)
    return feats / den

  def fourier_features(self, sx, degree):
    d = torch.arange(degree)
    if self.dtype == torch.float32:
      c = torch.tensor(np.float32(_TWO_PI)) * (2.0 ** d).to(torch.float32)
    else:
      c = _TWO_PI * (2.0 ** d).to(torch.float64)
    y = c.to(self.dtype) * sx.reshape(-1, 1)          # models.py:85
    feats = torch.cat([torch.cos(y), torch.sin(y)], 1)
    den = torch.tile((d + 1).to(self.dtype), (2,))
    return feats / den

  def encode(self, params, x):
    """models.py:216-252 -> (B, F) feature matrix."""
    if x.ndim == 1:
      x = x[:, None]
    lsa = self.leaf(params, 'log_scale_adjustment')
    scales = torch.from_numpy(self.input_scales).to(self.dtype)
    sx = x / (scales * torch.exp(lsa))
    cols = []
    for idx, kind, _, meta in self.groups:
      if kind == 'x':
        f = sx
      elif kind == 'fourier':
        f = self.fourier_features(sx[:, meta[0]], meta[1])
      elif kind == 'seasonal':
        f = self.seasonal_features(x[:, 0])          # RAW time, models.py:223
      else:
        f = torch.prod(sx[:, torch.from_numpy(self.interactions)], -1)
      s = _softplus(self.leaf(params, f'feature_inv_sp_scale{idx}'))
      cols.append(f * s)
    return torch.cat(cols, -1)

  # ---- models.py:254-273 dense stack ----
  def forward(self, params, x):
    h = self.encode(params, x)
    w = torch.sigmoid(self.leaf(params, 'logit_activation_weight'))
    for l in range(self.depth):
      k = self.leaf(params, f'Dense_{l}/kernel')
      b = self.leaf(params, f'Dense_{l}/bias')
      s = _softplus(self.leaf(params, f'inv_sp_layer_scale{l}'))
      h = h / math.sqrt(h.shape[-1]) if self.dtype == torch.float64 else \
          h / torch.tensor(np.float32(np.sqrt(np.float32(h.shape[-1]))))
      z = s * (h @ k + b)
      elu = torch.where(z > 0, z, torch.expm1(torch.where(z > 0, 0 * z, z)))
      h = w * elu + (1 - w) * torch.tanh(z)
    k = self.leaf(params, f'Dense_{self.depth}/kernel')
    b = self.leaf(params, f'Dense_{self.depth}/bias')
    s = _softplus(self.leaf(params, 'inv_sp_output_scale'))
    h = h / math.sqrt(h.shape[-1]) if self.dtype == torch.float64 else \
        h / torch.tensor(np.float32(np.sqrt(np.float32(h.shape[-1]))))
    return s * (h @ k + b)[..., 0]


# --------------------------------------------------------------------------
# models.py:106-194 make_likelihood_model(...).log_prob(y)
# --------------------------------------------------------------------------
def _log_sigmoid(x):
  return -_softplus(-x)


def nb_log_prob(x, total_count, logits):
  """TFP 0.24 NegativeBinomial._log_prob (total_count=r failures, logits)."""
  lbeta = (torch.lgamma(1. + x) + torch.lgamma(total_count)
           - torch.lgamma(1. + x + total_count))
  return (total_count * _log_sigmoid(-logits) + x * _log_sigmoid(logits)
          - lbeta - torch.log(total_count + x))


def likelihood_params(model, params, x, distribution):
  """inference.py:103-126 forecast_inner: the distribution parameters."""
  pred = model.forward(params, x)
  if distribution == NORMAL:
    return pred, 0.01 + torch.exp(params[0])                 # models.py:162-164
  mean = _softplus(pred)
  shape = _softplus(params[1])
  total_count = 1 / shape
  logits = -torch.log(shape) - torch.log(mean)               # models.py:173-175
  if distribution == NB:
    return total_count, logits
  pi = 1 / (1 + torch.exp(-params[2]))                       # models.py:184
  return total_count, logits, pi * torch.ones_like(mean)


def log_likelihood(model, params, x, y, distribution):
  """Sum over the batch of the observation log-prob (tfd.Independent(..., 1))."""
  lp = likelihood_params(model, params, x, distribution)
  if distribution == NORMAL:
    loc, scale = lp
    # TFP Normal._log_prob: -0.5*sqdiff(x/s, loc/s) - (0.5*log(2pi) + log(s))
    return torch.sum(-0.5 * (y / scale - loc / scale) ** 2
                     - (0.5 * math.log(_TWO_PI) + torch.log(scale)))
  if distribution == NB:
    return torch.sum(nb_log_prob(y, lp[0], lp[1]))
  total_count, logits, pi = lp
  nb = nb_log_prob(y, total_count, logits)
  neg_inf = torch.full_like(nb, -float('inf'))
  point = torch.where(y == 0, torch.log(pi), neg_inf)        # log(pi) + log 1[y==0]
  return torch.sum(torch.logsumexp(
      torch.stack([torch.log1p(-pi) + nb, point]), 0))


# --------------------------------------------------------------------------
# models.py:91-103 prior; TFP Logistic(loc, 1).log_prob
# --------------------------------------------------------------------------
def _logistic_log_prob(x, loc):
  z = x - loc
  return -z - 2. * _softplus(-z)


def prior_log_prob(params):
  lp = _logistic_log_prob(params[0], 0.0) + _logistic_log_prob(params[1], -1.5) \
      + _logistic_log_prob(params[2], 0.0)
  for p in params[3:]:
    lp = lp + torch.sum(_logistic_log_prob(p, 0.0))
  return lp


# --------------------------------------------------------------------------
# inference.py:558-569 MAP / MLE objective, :599-606 value_and_grad + Adam
# --------------------------------------------------------------------------
def map_loss(model, params, x, y, n_total, prior_weight, distribution):
  scale = n_total / y.shape[0]
  ll = log_likelihood(model, params, x, y, distribution) * scale
  if prior_weight == 0.0:
    return -ll
  return -(ll + prior_log_prob(params) * prior_weight)


def map_loss_and_grad(model, flat, x, y, n_total, prior_weight, distribution):
  """flat: (P,) tensor. Returns (loss, grad (P,)) via torch autograd."""
  flat = flat.detach().clone().requires_grad_(True)
  loss = map_loss(model, model.unflatten(flat), x, y, n_total, prior_weight,
                  distribution)
  (g,) = torch.autograd.grad(loss, flat, allow_unused=True)
  if g is None:
    g = torch.zeros_like(flat)
  return loss.detach(), g


def adam_update(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
  """optax.adam (0.2.2) one step. ``t`` is the step count AFTER increment."""
  m = (1 - b1) * g + b1 * m
  v = (1 - b2) * (g * g) + b2 * v
  one = torch.tensor(1.0, dtype=p.dtype)
  mhat = m / (one - torch.tensor(b1, dtype=p.dtype) ** t)
  vhat = v / (one - torch.tensor(b2, dtype=p.dtype) ** t)
  p = p + (-lr) * (mhat / (torch.sqrt(vhat) + eps))
  return p, m, v


def init_map_params(model, target, generator, dtype=None):
  """inference.py:399-427: lns=log(nanstd/2); kernels~TruncNormal[-2,2]; else 0."""
  dtype = dtype or model.dtype
  out = [torch.tensor(math.log(np.nanstd(target) / 2.0), dtype=dtype),
         torch.zeros((), dtype=dtype), torch.zeros((), dtype=dtype)]
  for s in model.leaf_shapes:
    out.append(truncated_normal(s, generator, dtype) if len(s) == 2
               else torch.zeros(s, dtype=dtype))
  return out


def truncated_normal(shape, generator, dtype):
  """Std normal conditioned on [-2, 2] by rejection (any RNG: parity tests
  inject the initial parameters, seeds are not matched with JAX threefry)."""
  n = int(np.prod(shape))
  out = torch.empty(0, dtype=torch.float64)
  while out.numel() < n:
    z = torch.randn(2 * n + 16, generator=generator, dtype=torch.float64)
    out = torch.cat([out, z[z.abs() <= 2]])
  return out[:n].reshape(shape).to(dtype)


def fit_map_member(model, flat0, x, y, batch_index_fn, num_epochs, batch_size,
                   lr, prior_weight, distribution):
  """inference.py:577-619 ``_run`` for one member with injected batch order.

  ``batch_index_fn(epoch)`` returns the row permutation used in that epoch
  (identity for full batch).  Returns (flat params, per-epoch mean loss).
  """
  n = y.shape[0]
  steps = n // batch_size
  p = flat0.clone()
  m = torch.zeros_like(p)
  v = torch.zeros_like(p)
  t = 0
  losses = []
  for ep in range(num_epochs):
    perm = batch_index_fn(ep)
    ep_losses = []
    for s in range(steps):
      rows = perm[s * batch_size:(s + 1) * batch_size]
      loss, g = map_loss_and_grad(model, p, x[rows], y[rows], n, prior_weight,
                                  distribution)
      t += 1
      p, m, v = adam_update(p, g, m, v, t, lr)
      ep_losses.append(loss)
    losses.append(torch.stack(ep_losses).mean())
  return p, torch.stack(losses)


# --------------------------------------------------------------------------
# inference.py:626-764 ensemble_vi: mean-field surrogate, reparam MC ELBO
# --------------------------------------------------------------------------
SOFTPLUS_INV_0P3 = math.log(math.expm1(0.3))    # tfp.math.softplus_inverse(0.3)


def vi_loss(model, mu, rho, eps, x, y, n_total, kl_weight, distribution):
  """One member.  mu, rho: (P,) ; eps: (S, P).  TFP fit_surrogate_posterior_
  stateless default loss: mean_s [ log q(z_s) - target(z_s) ], z_s = mu+sigma*eps_s
  with target = prior + loglik*(N/B)/kl_weight  (inference.py:687-702, :711-720).
  """
  sigma = 0.0001 + _softplus(rho)
  total = 0.0
  for s in range(eps.shape[0]):
    z = mu + sigma * eps[s]
    logq = torch.sum(-0.5 * ((z - mu) / sigma) ** 2 - 0.5 * math.log(_TWO_PI)
                     - torch.log(sigma))
    params = model.unflatten(z)
    tgt = prior_log_prob(params) + log_likelihood(
        model, params, x, y, distribution) * (n_total / y.shape[0]) / kl_weight
    total = total + (logq - tgt)
  return total / eps.shape[0]


def vi_loss_and_grad(model, mu, rho, eps, x, y, n_total, kl_weight,
                     distribution):
  mu = mu.detach().clone().requires_grad_(True)
  rho = rho.detach().clone().requires_grad_(True)
  loss = vi_loss(model, mu, rho, eps, x, y, n_total, kl_weight, distribution)
  gmu, grho = torch.autograd.grad(loss, [mu, rho])
  return loss.detach(), gmu, grho


# --------------------------------------------------------------------------
# inference.py:42-100 mixture quantiles
# --------------------------------------------------------------------------
def _ndtr(z):
  return 0.5 * torch.erfc(-z / math.sqrt(2.0))


def approximate_normal_quantile(means, scales, q):
  """inference.py:55-84. means (..., N); scales (..., 1); reduce leading axes."""
  axes = tuple(range(means.ndim - 1))
  m = means.mean(axes)
  s = torch.sqrt((scales ** 2 + means ** 2).mean(axes) - m ** 2)
  ndtri = math.sqrt(2.0) * torch.erfinv(torch.tensor(2.0 * q - 1.0,
                                                     dtype=torch.float64))
  return m + s * ndtri.to(means.dtype)


def mixture_cdf_residual(means, scales, xq, q):
  """mean_e Phi((x - mu_e)/sigma_e) - q (the function the root-finder zeroes)."""
  axes = tuple(range(means.ndim - 1))
  return _ndtr((xq - means) / scales).mean(axes) - q


def normal_quantile_via_root(means, scales, q, value_tol=1e-5, max_iter=60):
  """inference.py:42-52 via Chandrupatla's bracketing method (Chandrupatla 1997;
  tfp.math.find_root_chandrupatla semantics: global bracket, stop when |f|<=tol).
  Parity is judged by the CDF residual, not by x."""
  lo = (means.min() - 5 * scales.max()).expand(means.shape[-1]).clone()
  hi = (means.max() + 5 * scales.max()).expand(means.shape[-1]).clone()
  f = lambda xx: mixture_cdf_residual(means, scales, xx, q)
  a, b = hi, lo
  fa, fb = f(a), f(b)
  c, fc = a.clone(), fa.clone()
  t = torch.full_like(a, 0.5)
  best, fbest = torch.where(fa.abs() < fb.abs(), a, b), torch.minimum(fa.abs(), fb.abs())
  for _ in range(max_iter):
    xt = a + t * (b - a)
    ft = f(xt)
    same = torch.sign(ft) == torch.sign(fa)
    c, fc = torch.where(same, a, b), torch.where(same, fa, fb)
    b, fb = torch.where(same, b, a), torch.where(same, fb, fa)
    a, fa = xt, ft
    better = ft.abs() < fbest
    best, fbest = torch.where(better, xt, best), torch.where(better, ft.abs(), fbest)
    if bool((fbest <= value_tol).all()):
      break
    xi = (a - b) / (c - b)
    phi = (fa - fb) / (fc - fb)
    iqi = (phi ** 2 < xi) & ((1 - phi) ** 2 < 1 - xi)
    t_iqi = fa / (fb - fa) * fc / (fb - fc) \
        + (c - a) / (b - a) * fa / (fc - fa) * fb / (fc - fb)
    t = torch.where(iqi, t_iqi, torch.full_like(t, 0.5))
    tl = 1e-8 / (b - a).abs().clamp_min(1e-30)
    t = torch.minimum(torch.maximum(t, tl), 1 - tl)
    t = torch.where(torch.isfinite(t), t, torch.full_like(t, 0.5))
  return best


# --------------------------------------------------------------------------
# inference.py:271-333 NB / ZINB predictive mean and quantiles
# --------------------------------------------------------------------------
def nb_predictive(loc, shape_raw, pi_logit, distribution):
  """loc (M, N) network outputs, shape_raw / pi_logit (M,) -> dict of float64 numpy arrays
  (total_count (M,1), logits (M,N), pi (M,1), mean, stddev, prob0), following
  models.py:166-191 literally and the TFP NegativeBinomial / mixture moments."""
  import scipy.special as sp
  loc = np.asarray(loc, np.float64)
  a = np.logaddexp(np.asarray(shape_raw, np.float64), 0.0)[:, None]        # softplus(params[1])
  mean_net = np.logaddexp(loc, 0.0)
  r = 1.0 / a
  logits = -np.log(a) - np.log(mean_net)
  nb_mean = r * np.exp(logits)                        # tfd.NegativeBinomial.mean
  nb_var = nb_mean / sp.expit(-logits)                # tfd.NegativeBinomial.variance
  nb_p0 = sp.expit(-logits) ** r                      # exp(log_prob(0))
  if distribution == NB:
    pi = np.zeros_like(a)
  else:
    pi = 1.0 / (1.0 + np.exp(-np.asarray(pi_logit, np.float64)))[:, None]
  mean = (1 - pi) * nb_mean
  var = (1 - pi) * (nb_var + nb_mean ** 2) - mean ** 2      # mixture of delta_0 and NB
  return dict(r=r, logits=logits, pi=pi, mean=mean, stddev=np.sqrt(var), prob0=pi + (1 - pi) * nb_p0)


def nb_mixture_cdf(k, pred):
  """Mean over components of the (ZI)NB CDF at integer(s) k >= 0 -> (N,)."""
  import scipy.special as sp
  k = np.floor(np.asarray(k, np.float64))
  cdf = sp.betainc(pred['r'], 1.0 + k, sp.expit(-pred['logits']))    # TFP NegativeBinomial._cdf
  return (pred['pi'] + (1 - pred['pi']) * cdf).mean(0)


def nb_quantiles(pred, q):
  """inference.py:298-333: ceil(root of mean-CDF(x) - q on [0, high]) with the
  zero-mass override, evaluated as the exact discrete quantile
  min{k >= 0 integer : mean-CDF(k) >= q}, capped at ceil(high); the reference's
  Chandrupatla root + ceil lands on the same integer unless a CDF plateau lies
  within its 1e-5 value tolerance of q."""
  high = pred['mean'].max() + 1.1 / math.sqrt(1 - q) * pred['stddev'].max()
  n = pred['logits'].shape[1]
  lo = np.full(n, -1.0)
  hi = np.full(n, math.ceil(high))
  while np.any(hi - lo > 1):
    mid = np.floor((lo + hi) / 2)
    ge = nb_mixture_cdf(mid, pred) >= q
    act = hi - lo > 1
    hi = np.where(act & ge, mid, hi)
    lo = np.where(act & ~ge, mid, lo)
  return hi
