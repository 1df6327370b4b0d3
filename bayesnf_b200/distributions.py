"""Host-side stand-in for the TFP distribution that the reference's
`BayesianNeuralFieldEstimator.likelihood_model()` returns (spatiotemporal.py:433-468,
models.py:104-194).

The reference hands back `tfd.Independent(<Normal | NegativeBinomial |
ZeroInflatedNegativeBinomial>, 1)` whose batch shape is the ensemble shape
`(num_devices, [num_samples,] members)` and whose event shape is `(N,)`.  TFP is
not a dependency of this package; `PredictiveDistribution` offers the methods a
user of that object calls (`mean`, `stddev`, `variance`, `log_prob`, `prob`,
`sample`, `batch_shape`, `event_shape`, `.distribution` for the per-point
distribution with `log_prob`, `cdf`, `quantile`) on numpy arrays, restating the
TFP formulas recorded in SURVEY.md section 9.  The network outputs it is built
from come from the CUDA forward pass (`inference.Engine.forward`); nothing here
is on the training hot path.
"""
from __future__ import annotations

import numpy as np
from scipy import special

from . import models


def _softplus(x):
  return np.logaddexp(x, 0.0)


def _log_sigmoid(x):
  return -_softplus(-x)


class _Pointwise:
  """The per-point distribution (what `tfd.Independent(...).distribution` is)."""

  def __init__(self, kind, predictions, log_noise_scale, shape_raw, pi_logit):
    self.kind = kind
    self._pred = np.asarray(predictions, dtype=np.float64)
    lead = self._pred.shape[:-1]
    self._lns = np.broadcast_to(np.asarray(log_noise_scale, np.float64), lead)[..., None]
    self._shape_raw = np.broadcast_to(np.asarray(shape_raw, np.float64), lead)[..., None]
    self._pi_logit = np.broadcast_to(np.asarray(pi_logit, np.float64), lead)[..., None]

  # ---- parameters (models.py:157-191) ----------------------------------------------------
  @property
  def loc(self):
    return self._pred

  @property
  def scale(self):
    return np.broadcast_to(0.01 + np.exp(self._lns), self._pred.shape)      # models.py:163

  @property
  def total_count(self):
    return np.broadcast_to(1.0 / _softplus(self._shape_raw), self._pred.shape)   # models.py:173

  @property
  def logits(self):
    # models.py:174: -log(shape) - log(mean), mean = softplus(prediction)
    return -np.log(_softplus(self._shape_raw)) - np.log(_softplus(self._pred))

  @property
  def inflated_loc_probs(self):
    return np.broadcast_to(special.expit(self._pi_logit), self._pred.shape)      # models.py:184

  # ---- moments -----------------------------------------------------------------------------
  def _nb_mean(self):
    return self.total_count * np.exp(self.logits)

  def _nb_variance(self):
    return self._nb_mean() / special.expit(-self.logits)

  def mean(self):
    if self.kind == models.LikelihoodDist.NORMAL:
      return self.loc
    if self.kind == models.LikelihoodDist.NB:
      return self._nb_mean()
    return (1.0 - self.inflated_loc_probs) * self._nb_mean()

  def variance(self):
    if self.kind == models.LikelihoodDist.NORMAL:
      return self.scale ** 2
    if self.kind == models.LikelihoodDist.NB:
      return self._nb_variance()
    pi, m, v = self.inflated_loc_probs, self._nb_mean(), self._nb_variance()
    return (1.0 - pi) * (v + m * m) - ((1.0 - pi) * m) ** 2       # mixture of delta_0 and NB

  def stddev(self):
    return np.sqrt(self.variance())

  # ---- densities ---------------------------------------------------------------------------
  def _nb_log_prob(self, x):
    r, l = self.total_count, self.logits
    return (r * _log_sigmoid(-l) + x * _log_sigmoid(l)
            - (special.gammaln(1.0 + x) + special.gammaln(r) - special.gammaln(1.0 + x + r))
            - np.log(r + x))

  def _nb_cdf(self, x):
    r, l = self.total_count, self.logits
    xf = np.floor(x)
    c = special.betainc(r, 1.0 + np.maximum(xf, 0.0), special.expit(-l))
    return np.where(xf < 0, 0.0, c)

  def log_prob(self, x):
    x = np.asarray(x, dtype=np.float64)
    if self.kind == models.LikelihoodDist.NORMAL:
      s = self.scale
      return -0.5 * (x / s - self.loc / s) ** 2 - 0.5 * np.log(2.0 * np.pi) - np.log(s)
    nb = self._nb_log_prob(x)
    if self.kind == models.LikelihoodDist.NB:
      return nb
    pi = self.inflated_loc_probs
    with np.errstate(divide='ignore'):
      zero_part = np.where(x == 0, np.log(pi), -np.inf)
    return np.logaddexp(np.log1p(-pi) + nb, zero_part)

  def prob(self, x):
    return np.exp(self.log_prob(x))

  def cdf(self, x):
    x = np.asarray(x, dtype=np.float64)
    if self.kind == models.LikelihoodDist.NORMAL:
      return special.ndtr((x - self.loc) / self.scale)
    nb = self._nb_cdf(x)
    if self.kind == models.LikelihoodDist.NB:
      return nb
    pi = self.inflated_loc_probs
    return np.where(x < 0, 0.0, pi + (1.0 - pi) * nb)

  def quantile(self, q):
    if self.kind != models.LikelihoodDist.NORMAL:
      raise NotImplementedError('quantile() is defined for the Normal observation model; use '
                                'predict(..., quantiles=...) for ensemble NB/ZINB quantiles')
    return self.loc + self.scale * special.ndtri(np.asarray(q, dtype=np.float64))

  def sample(self, rng: np.random.Generator, sample_shape=()):
    shp = tuple(sample_shape) + self._pred.shape
    if self.kind == models.LikelihoodDist.NORMAL:
      return self.loc + self.scale * rng.standard_normal(shp)
    lam = rng.gamma(np.broadcast_to(self.total_count, shp), np.broadcast_to(np.exp(self.logits), shp))
    x = rng.poisson(lam).astype(np.float64)
    if self.kind == models.LikelihoodDist.ZINB:
      x = np.where(rng.random(shp) < np.broadcast_to(self.inflated_loc_probs, shp), 0.0, x)
    return x


class PredictiveDistribution:
  """`tfd.Independent(pointwise, reinterpreted_batch_ndims=1)` on numpy arrays.

  Args:
    observation_model: 'NORMAL', 'NB' or 'ZINB'.
    predictions: network outputs `mlp.apply(params, x)`, shape `batch_shape + (N,)`.
    log_noise_scale, shape, inflated_loc_probs: `params_[0..2]`, each of shape `batch_shape`
      (the reference appends a broadcasting axis, spatiotemporal.py:457-460).
  """

  def __init__(self, observation_model, predictions, log_noise_scale, shape, inflated_loc_probs):
    kind = models.LikelihoodDist(observation_model)
    self.observation_model = kind
    self.distribution = _Pointwise(kind, predictions, log_noise_scale, shape, inflated_loc_probs)

  @property
  def batch_shape(self):
    return self.distribution._pred.shape[:-1]

  @property
  def event_shape(self):
    return self.distribution._pred.shape[-1:]

  def mean(self):
    return self.distribution.mean()

  def variance(self):
    return self.distribution.variance()

  def stddev(self):
    return self.distribution.stddev()

  def log_prob(self, y):
    """Sum of the per-point log-probabilities over the event axis (shape `batch_shape`)."""
    return self.distribution.log_prob(y).sum(axis=-1)

  def prob(self, y):
    return np.exp(self.log_prob(y))

  def sample(self, sample_shape=(), seed=None):
    """Draws of shape `sample_shape + batch_shape + event_shape`.  `seed`: int, a 2-word PRNG
    key array, or None."""
    if isinstance(sample_shape, int):
      sample_shape = (sample_shape,)
    if seed is None:
      rng = np.random.default_rng()
    else:
      words = np.asarray(seed).astype(np.uint64).ravel()
      rng = np.random.default_rng([int(w) for w in words])
    return self.distribution.sample(rng, sample_shape)
