"""bayesnf_b200.distributions.PredictiveDistribution: the TFP formulas of
models.py:157-191 / SURVEY.md section 9 against scipy.stats on random parameters (CPU only)."""
import numpy as np
import pytest
from scipy import stats

from bayesnf_b200 import distributions


def _params(seed=0, lead=(1, 3), n=7):
  rng = np.random.default_rng(seed)
  pred = rng.normal(size=lead + (n,)) * 2.0
  return pred, rng.normal(size=lead) * 0.5, rng.normal(size=lead) * 0.5 - 1.0, rng.normal(size=lead)


def test_normal_matches_scipy():
  pred, lns, shp, pil = _params()
  d = distributions.PredictiveDistribution('NORMAL', pred, lns, shp, pil)
  assert d.batch_shape == (1, 3) and d.event_shape == (7,)
  sigma = (0.01 + np.exp(lns))[..., None]
  y = np.random.default_rng(1).normal(size=7)
  np.testing.assert_allclose(d.mean(), pred)
  np.testing.assert_allclose(d.stddev(), np.broadcast_to(sigma, pred.shape))
  np.testing.assert_allclose(d.log_prob(y), stats.norm(pred, sigma).logpdf(y).sum(-1), rtol=1e-12)
  np.testing.assert_allclose(d.distribution.cdf(y), stats.norm(pred, sigma).cdf(y), rtol=1e-12)
  np.testing.assert_allclose(d.distribution.quantile(0.9), stats.norm(pred, sigma).ppf(0.9), rtol=1e-12)


@pytest.mark.parametrize('kind', ['NB', 'ZINB'])
def test_negative_binomial_matches_scipy(kind):
  pred, lns, shp, pil = _params(seed=2)
  d = distributions.PredictiveDistribution(kind, pred, lns, shp, pil)
  shape = np.logaddexp(shp, 0.0)[..., None]
  mean_net = np.logaddexp(pred, 0.0)
  r = 1.0 / shape
  logits = -np.log(shape) - np.log(mean_net)
  nb = stats.nbinom(n=r, p=1.0 - 1.0 / (1.0 + np.exp(-logits)))
  pi = (1.0 / (1.0 + np.exp(-pil)))[..., None] if kind == 'ZINB' else 0.0
  ks = np.array([0, 1, 2, 5, 9, 0, 3], dtype=np.float64)
  want_pmf = (1.0 - pi) * nb.pmf(ks) + pi * (ks == 0)
  np.testing.assert_allclose(d.distribution.prob(ks), want_pmf, rtol=1e-9)
  np.testing.assert_allclose(d.log_prob(ks), np.log(want_pmf).sum(-1), rtol=1e-9)
  np.testing.assert_allclose(d.distribution.cdf(ks), pi + (1.0 - pi) * nb.cdf(ks), rtol=1e-9)
  np.testing.assert_allclose(d.mean(), (1.0 - pi) * nb.mean(), rtol=1e-9)
  # literal restatement of the reference code: mean = 1/(shape^2 * mean_net) (SURVEY.md section 9 caution)
  np.testing.assert_allclose(nb.mean(), 1.0 / (shape ** 2 * mean_net), rtol=1e-9)
  var = (1.0 - pi) * (nb.var() + nb.mean() ** 2) - ((1.0 - pi) * nb.mean()) ** 2
  np.testing.assert_allclose(d.variance(), var, rtol=1e-9)


@pytest.mark.parametrize('kind', ['NORMAL', 'NB', 'ZINB'])
def test_samples_follow_the_distribution(kind):
  pred, lns, shp, pil = _params(seed=3, lead=(2,), n=4)
  d = distributions.PredictiveDistribution(kind, pred, lns, shp + 1.0, pil)
  x = d.sample(20000, seed=7)
  assert x.shape == (20000, 2, 4)
  np.testing.assert_allclose(x.mean(0), d.mean(), rtol=0.08, atol=0.05)
  np.testing.assert_allclose(x.std(0), d.stddev(), rtol=0.12, atol=0.05)
  again = d.sample(20000, seed=7)
  np.testing.assert_array_equal(x, again)
  if kind != 'NORMAL':
    assert (x >= 0).all() and (x == np.floor(x)).all()
