"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same inputs.

Tolerances (stated per test):
* fp32 (SIMT) and bf16x3 (tcgen05 tensor cores, split operands) modes: the north-star bar,
  1e-5 relative to the tensor's scale for forward values / log-likelihoods; gradients are
  compared against the float64 twin of the oracle with 1e-4 of the leaf's max-abs (the f32
  oracle's own autograd noise is of that order on 80k-parameter sums).  Every parity test of
  the f32 path is parametrised over BOTH modes with the SAME tolerance.
* bf16 modes: bf16 operand rounding (2^-9) -> 3e-2 of scale.
"""
import json
import math
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN
from oracle import bnf_oracle as O

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(GOLDEN, 'bookkeeping.json')))


@pytest.fixture(scope='module')
def cuda():
  assert torch.cuda.is_available(), 'gpu tests need a CUDA device (no fallback)'
  torch.cuda.set_device(0)
  return torch.device('cuda', 0)


def _cfgs():
  return {
      'small': dict(width=64, depth=2, input_scales=[199., 1, 1], num_seasonal_harmonics=[2, 3],
                    seasonality_periods=[7.0, 30.0], init_x=(200, 3), fourier_degrees=[3, 2, 2],
                    interactions=np.array([[1, 2]])),
      'chickenpox': dict(width=256, depth=2, input_scales=[99., 1, 1], num_seasonal_harmonics=[2, 10],
                         seasonality_periods=[4.0, 52.1775], init_x=(100, 3), fourier_degrees=[5, 5, 5],
                         interactions=np.zeros((0, 2), int)),
      'deep': dict(width=128, depth=4, input_scales=[299., 1, 1], num_seasonal_harmonics=[4, 4],
                   seasonality_periods=[24, 168], init_x=(300, 3), fourier_degrees=[5, 5, 5],
                   interactions=np.zeros((0, 2), int)),
      'odd': dict(width=40, depth=3, input_scales=[50., 2.0], num_seasonal_harmonics=[],
                  seasonality_periods=[], init_x=(77, 2), fourier_degrees=[0, 4],
                  interactions=np.array([[0, 1]])),
  }


def _data(cfg, n, seed=0, counts=False):
  rng = np.random.default_rng(seed)
  D = len(cfg['input_scales'])
  cols = [np.arange(n, dtype=np.float64)] + [rng.normal(size=n) for _ in range(D - 1)]
  x = np.stack(cols, 1)
  if counts:
    y = rng.poisson(3.0, size=n).astype(np.float64)
    y[rng.random(n) < 0.3] = 0
  else:
    y = 3 * np.sin(np.arange(n) / 5.0) + rng.normal(size=n)
  return x, y


def _random_params(om, n_net, y, seed=0, jitter=0.3):
  g = torch.Generator().manual_seed(seed)
  flats = []
  for _ in range(n_net):
    flat = om.flatten(O.init_map_params(om, y, g))
    flat = flat + jitter * torch.randn(flat.shape, generator=g) * (flat == 0)  # move zeros off 0
    flat[1] = 0.2 * torch.randn((), generator=g)
    flat[2] = 0.2 * torch.randn((), generator=g)
    flats.append(flat)
  return torch.stack(flats)


def _engine(cfg, dist='NORMAL', prec='fp32'):
  from bayesnf_b200 import inference, models
  spec = models.ModelSpec(**cfg, observation_model=dist)
  return inference.Engine(spec, prec), spec


PARITY_MODES = ['fp32', 'bf16x3']   # SIMT f32 and the tensor-core f32-parity mode: same tolerances


def _skip_unsupported(name, prec):
  if prec == 'bf16x3' and name == 'odd':
    pytest.skip('width 40 is not a tensor-core shape (bf16x3 needs width in {64..1024}); fp32 covers it')


@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('name', ['small', 'chickenpox', 'deep', 'odd'])
def test_forward_fp32(cuda, name, prec):
  """mlp.apply: <= 1e-5 of output scale vs the f32 oracle (and its f64 twin)."""
  from bayesnf_b200 import inference
  _skip_unsupported(name, prec)
  cfg = _cfgs()[name]
  n = cfg['init_x'][0]
  x, y = _data(cfg, n)
  om, om64 = O.OracleModel(**cfg), O.OracleModel(**cfg, dtype=torch.float64)
  P = _random_params(om, 3, y)
  eng, spec = _engine(cfg, prec=prec)
  xd, _ = inference._to_device_data(x, y)
  loc = eng.forward(P.to(cuda), xd, slab=64).cpu()
  for j in range(3):
    want = om.forward(om.unflatten(P[j]), xd.cpu())
    want64 = om64.forward(om64.unflatten(P[j].double()), xd.cpu().double())
    scale = float(want64.abs().max())
    assert float((loc[j] - want).abs().max()) <= 1e-5 * scale + 1e-6
    assert float((loc[j].double() - want64).abs().max()) <= 1e-5 * scale + 1e-6


@pytest.mark.parametrize('prec', ['bf16x3', 'bf16', 'fp32'])
def test_forward_slab_sizes_agree(cuda, prec):
  """Engine.forward cuts the test rows into slabs sized from a memory budget (the reference forecasts
  in 1024-row batches, inference.py:129-181): 70 000 rows as 65 536 + 4 464 (the default cap), as
  4 096-row slabs and with a tiny budget give the same means; forward_slab_rows clamps to [128, 65 536]."""
  from bayesnf_b200 import inference
  cfg = _cfgs()['chickenpox']
  n = 70000
  g = torch.Generator().manual_seed(5)
  x = torch.stack([torch.rand(n, generator=g) * 99.0, torch.randn(n, generator=g), torch.randn(n, generator=g)], 1)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 4, np.random.default_rng(0).normal(size=64))
  eng, spec = _engine(cfg, prec=prec)
  xd = x.to(cuda).float().contiguous()
  assert eng.forward_slab_rows(4) == 65536 and eng.forward_slab_rows(4, budget_bytes=1) == 128
  mid = eng.forward_slab_rows(4, budget_bytes=64 << 20)
  assert 128 <= mid < 65536 and mid % 128 == 0
  a = eng.forward(P.to(cuda), xd)
  b = eng.forward(P.to(cuda), xd, slab=4096)
  c = eng.forward(P.to(cuda), xd, slab=eng.forward_slab_rows(4, budget_bytes=1))
  torch.cuda.synchronize()
  assert torch.isfinite(a).all()
  # every row's mean is a function of that row alone: the same whatever slab it was computed in
  scale = float(a.abs().max())
  assert float((a - b).abs().max()) <= 1e-6 * scale and float((a - c).abs().max()) <= 1e-6 * scale
  want = om.forward(om.unflatten(P[1]), x[:2000].float())
  tol = 1e-5 if prec != 'bf16' else 3e-2
  assert float((a[1, :2000].cpu() - want).abs().max()) <= tol * float(want.abs().max()) + 1e-6


@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('dist', ['NORMAL', 'NB', 'ZINB'])
@pytest.mark.parametrize('name', ['small', 'chickenpox', 'odd'])
def test_loglik_and_grad_fp32(cuda, name, dist, prec):
  from bayesnf_b200 import inference
  _skip_unsupported(name, prec)
  cfg = _cfgs()[name]
  n = cfg['init_x'][0]
  x, y = _data(cfg, n, counts=dist != 'NORMAL')
  om64 = O.OracleModel(**cfg, dtype=torch.float64)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 2, y, seed=3)
  eng, spec = _engine(cfg, dist, prec)
  xd, yd = inference._to_device_data(x, y)
  ll, grad = eng.loglik_grad(P.to(cuda), xd, yd)
  ll, grad = ll.cpu(), grad.cpu()
  for j in range(2):
    loss64, g64 = O.map_loss_and_grad(om64, P[j].double(), xd.cpu().double(), yd.cpu().double(),
                                      n, 0.0, dist)
    assert abs(float(ll[j]) + float(loss64)) <= 2e-5 * abs(float(loss64)) + 1e-4
    # per leaf: max-abs error relative to the leaf's max-abs gradient
    parts = [(0, 1), (1, 2), (2, 3)] + [(o, o + (int(np.prod(s)) if s else 1))
                                        for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
    for (a, b) in parts:
      want = -g64[a:b]
      got = grad[j, a:b].double()
      tol = 1e-4 * float(want.abs().max()) + 1e-5 * float(g64.abs().max()) * 1e-2 + 1e-7
      assert float((got - want).abs().max()) <= tol, (name, dist, prec, a, b, float((got - want).abs().max()), tol)


@pytest.mark.parametrize('prec', PARITY_MODES)
def test_minibatch_index_gather(cuda, prec):
  """Per-member index rows (inference.py:593-595) and the shared VI sub-batch."""
  from bayesnf_b200 import inference
  cfg = _cfgs()['small']
  n = 200
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 2, y)
  eng, _ = _engine(cfg, prec=prec)
  xd, yd = inference._to_device_data(x, y)
  rng = np.random.default_rng(5)
  idx = np.stack([rng.permutation(n)[:64], rng.permutation(n)[:64]]).astype(np.int32)
  ll, _ = eng.loglik_grad(P.to(cuda), xd, yd, idx=torch.tensor(idx, device=cuda))
  ll1, _ = eng.loglik_grad(P.to(cuda), xd, yd, idx=torch.tensor(idx[:1], device=cuda))
  for j in range(2):
    rows = torch.tensor(idx[j].astype(np.int64))
    want = O.log_likelihood(om, om.unflatten(P[j]), xd.cpu()[rows], yd.cpu()[rows], 'NORMAL')
    assert abs(float(ll[j]) - float(want)) <= 2e-5 * abs(float(want))
    rows = torch.tensor(idx[0].astype(np.int64))
    want = O.log_likelihood(om, om.unflatten(P[j]), xd.cpu()[rows], yd.cpu()[rows], 'NORMAL')
    assert abs(float(ll1[j]) - float(want)) <= 2e-5 * abs(float(want))


@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('pw', [1.0, 0.0])
def test_map_steps_fp32(cuda, pw, prec):
  """k Adam steps from injected init / batch order == oracle fit (inference.py:577-619)."""
  from bayesnf_b200 import inference
  cfg = _cfgs()['small']
  n, B, epochs = 200, 64, 2
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P0 = _random_params(om, 2, y, seed=11)
  rng = np.random.default_rng(9)
  perms = np.stack([[rng.permutation(n) for _ in range(2)] for _ in range(epochs)]).astype(np.int32)
  params, losses = inference.fit_map(x, y, 0, 'NORMAL', cfg, num_particles=2, learning_rate=0.01,
                                     num_epochs=epochs, prior_weight=pw, batch_size=B,
                                     precision=prec, init_params=P0.numpy(), batch_indices=perms)
  from bayesnf_b200 import models
  spec = models.ModelSpec(**cfg)
  flat = spec.flatten(params)[0]
  assert losses.shape == (1, 2, epochs)
  xd, yd = inference._to_device_data(x, y)
  for j in range(2):
    pj, lj = O.fit_map_member(om, P0[j], xd.cpu(), yd.cpu(),
                              lambda ep: torch.tensor(perms[ep, j].astype(np.int64)), epochs, B,
                              0.01, pw, 'NORMAL')
    # 6 Adam steps of lr=.01: parameters move by <= .06; compare the MOVE to 1e-3 of lr-scale
    assert float(np.abs(flat[j] - pj.numpy()).max()) <= 2e-4
    np.testing.assert_allclose(losses[0, j], lj.numpy(), rtol=2e-5)


@pytest.mark.parametrize('prec', PARITY_MODES)
def test_full_batch_multi_epoch(cuda, prec):
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n = 200
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P0 = _random_params(om, 1, y, seed=2)
  params, losses = inference.fit_map(x, y, 0, 'NORMAL', cfg, 1, 0.005, 4, precision=prec,
                                     init_params=P0.numpy())
  xd, yd = inference._to_device_data(x, y)
  pj, lj = O.fit_map_member(om, P0[0], xd.cpu(), yd.cpu(), lambda ep: torch.arange(n), 4, n,
                            0.005, 1.0, 'NORMAL')
  flat = models.ModelSpec(**cfg).flatten(params)[0, 0]
  assert float(np.abs(flat - pj.numpy()).max()) <= 2e-4
  np.testing.assert_allclose(losses[0, 0], lj.numpy(), rtol=2e-5)


@pytest.mark.parametrize('prec', PARITY_MODES)
def test_vi_step_fp32(cuda, prec):
  """One VI step with injected eps == oracle autograd through the reparameterised ELBO."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n, S, E = 200, 3, 2
  x, y = _data(cfg, n)
  om64 = O.OracleModel(**cfg, dtype=torch.float64)
  om = O.OracleModel(**cfg)
  spec = models.ModelSpec(**cfg)
  P = spec.num_params
  g = torch.Generator().manual_seed(4)
  mu = _random_params(om, E, y, seed=8)
  mu[:, 0] = 0.0
  rho = torch.full((E, P), O.SOFTPLUS_INV_0P3) + 0.05 * torch.randn(E, P, generator=g)
  eps = torch.randn(1, S, E, P, generator=g)
  kl, lr = 0.1, 0.01
  sur, losses, samples = inference.fit_vi(
      x, y, 0, 'NORMAL', cfg, ensemble_size=E, learning_rate=lr, num_epochs=1,
      sample_size_divergence=S, sample_size_posterior=4, kl_weight=kl, precision=prec,
      init_params=(mu.numpy(), rho.numpy()), eps=eps.numpy(),
      posterior_eps=np.zeros((4, E, P), np.float32))
  xd, yd = inference._to_device_data(x, y)
  mu1 = spec.flatten(sur.loc)[0]
  rho1 = spec.flatten(sur.inv_softplus_scale)[0]
  for e in range(E):
    loss, gmu, grho = O.vi_loss_and_grad(om64, mu[e].double(), rho[e].double(), eps[0, :, e].double(),
                                         xd.cpu().double(), yd.cpu().double(), n, kl, 'NORMAL')
    assert abs(losses[0, e, 0] - float(loss) * kl) <= 2e-5 * abs(float(loss) * kl)
    # first Adam step = -lr*g/(|g|+eps): compare the implied update
    want_mu = mu[e].double() - lr * gmu / (gmu.abs() + 1e-8)
    want_rho = rho[e].double() - lr * grho / (grho.abs() + 1e-8)
    big = gmu.abs() > 1e-3 * gmu.abs().max()       # tiny gradients flip sign under f32 noise
    assert float((torch.tensor(mu1[e]).double() - want_mu)[big].abs().max()) <= 2e-5
    bigr = grho.abs() > 1e-3 * grho.abs().max()
    assert float((torch.tensor(rho1[e]).double() - want_rho)[bigr].abs().max()) <= 2e-5
  # posterior "samples" with eps = 0 are the updated means; shapes follow inference.py:741-753
  assert samples[0].shape == (1, 4, E)
  np.testing.assert_allclose(spec.flatten(samples)[0, 0], mu1, rtol=0, atol=1e-6)


def test_quantiles(cuda):
  from bayesnf_b200 import inference
  g = torch.Generator().manual_seed(0)
  means = (torch.randn(12, 500, generator=g) * 2).to(cuda)
  scales = (torch.rand(12, generator=g) + 0.2).to(cuda)
  qs = [0.5, 0.025, 0.975]
  out = inference.mixture_quantiles(means, scales, qs, approximate=False).cpu()
  for i, q in enumerate(qs):
    res = O.mixture_cdf_residual(means.cpu()[None], scales.cpu()[None, :, None], out[i], q)
    assert float(res.abs().max()) <= 1.2e-5          # value_tolerance of inference.py:49
  approx = inference.mixture_quantiles(means, scales, qs, approximate=True).cpu()
  for i, q in enumerate(qs):
    want = O.approximate_normal_quantile(means.cpu()[None], scales.cpu()[None, :, None], q)
    np.testing.assert_allclose(approx[i].numpy(), want.numpy(), rtol=1e-5, atol=1e-5)


def test_init_params_rule(cuda):
  """inference.py:399-427: leaf0 = given value, kernels in [-2,2] ~ truncated normal, rest 0."""
  eng, spec = _engine(_cfgs()['chickenpox'])
  p = eng.init_params(1.25, 99, 0, 3).cpu().numpy()
  tup = spec.unflatten(p)
  assert np.all(tup[0] == 1.25) and np.all(tup[1] == 0) and np.all(tup[2] == 0)
  for name, leaf in zip(spec.leaf_names, tup[3:]):
    if name.endswith('kernel'):
      assert np.abs(leaf).max() <= 2.0
      if leaf.size > 20000:   # std of TruncatedNormal(0,1,[-2,2]) is 0.8796
        assert abs(leaf.std() - 0.8796) < 0.02 and abs(leaf.mean()) < 0.02
    else:
      assert np.all(leaf == 0)
  assert not np.array_equal(p[0], p[1])
  p2 = eng.init_params(1.25, 99, 1, 2).cpu().numpy()   # member streams are keyed by global id
  np.testing.assert_array_equal(p2[0], p[1])


def _chickenpox_tables():
  train = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.train.csv'), index_col=0, parse_dates=['datetime'])
  test = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.test.csv'), index_col=0, parse_dates=['datetime'])
  return train, test


def _chickenpox_estimator(cls, **kw):
  c = G['chickenpox']
  dc, mc = c['dataset_config'], c['model_config']
  return cls(feature_cols=dc['feature_cols'], target_col=dc['target_col'], timetype=dc['timetype'],
             freq=dc['freq'], standardize=dc['standardize'], width=mc['width'], depth=mc['depth'],
             seasonality_periods=mc['seasonality_periods'],
             num_seasonal_harmonics=mc['num_seasonal_harmonics'],
             observation_model=mc['observation_model'], **kw)


@pytest.mark.parametrize('objective', ['map', 'mle'])
def test_estimator_api_against_reference_mini_golden(cuda, objective):
  """tests/test_evaluate_mini.py config (4 particles, 5 epochs, lr .005) through the public API.
  Seeds cannot match JAX threefry, so the pinned quantity is sigma: the golden
  (2.5%, 97.5%) half width on the train rows."""
  import bayesnf_b200
  cls = bayesnf_b200.BayesianNeuralFieldMAP if objective == 'map' else bayesnf_b200.BayesianNeuralFieldMLE
  train, test = _chickenpox_tables()
  est = _chickenpox_estimator(cls, precision='fp32')
  est.fit(train, seed=np.array([0, 0], dtype=np.uint32), ensemble_size=4, learning_rate=0.005, num_epochs=5)
  assert est.losses_.shape == (1, 4, 5) and len(est.params_) == 19
  assert est.params_[0].shape == (1, 4)
  assert est.params_[4].shape == (1, 4, 57, 256)       # Dense_0/kernel
  both = pd.concat([train, test])
  means, quantiles = est.predict(both, quantiles=(0.5, 0.025, 0.975))
  assert means.shape == (1, 4, 308) and len(quantiles) == 3 and quantiles[0].shape == (308,)
  gold = pd.read_csv(os.path.join(GOLDEN, f'bnf-{objective}.chickenpox.8.mini.pred.csv'), index_col=0)
  gold_half = ((gold['yhat_upper'] - gold['yhat_lower']) / 2).to_numpy()[:100]
  half = ((quantiles[2] - quantiles[1]) / 2)[:100]
  assert abs(np.median(half) - np.median(gold_half)) / np.median(gold_half) < 5e-3
  # approximate quantiles agree with the root-found ones for near-identical members
  _, qa = est.predict(both, quantiles=(0.5,), approximate_quantiles=True)
  assert np.abs(qa[0][:100] - quantiles[0][:100]).max() < 0.5


def test_vi_estimator_api(cuda):
  import bayesnf_b200
  train, test = _chickenpox_tables()
  est = _chickenpox_estimator(bayesnf_b200.BayesianNeuralFieldVI, precision='fp32')
  est.fit(train, seed=0, ensemble_size=2, learning_rate=0.01, num_epochs=2, sample_size_posterior=3,
          sample_size_divergence=5, kl_weight=0.1)
  assert est.losses_.shape == (1, 2, 2) and np.isfinite(est.losses_).all()
  assert est.params_[0].shape == (1, 3, 2)
  means, q = est.predict(train, quantiles=(0.5,))
  assert means.shape == (1, 3, 2, 100) and q[0].shape == (100,)


@pytest.mark.parametrize('prec', ['bf16_simt'])
def test_bf16_storage_path(cuda, prec):
  """bf16 activations / f32 SIMT GEMMs: 3e-2 of scale (bf16 rounding 2^-9 per stored value)."""
  from bayesnf_b200 import inference
  cfg = _cfgs()['chickenpox']
  n = 100
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 2, y, seed=3)
  eng, spec = _engine(cfg, 'NORMAL', prec)
  xd, yd = inference._to_device_data(x, y)
  ll, grad = eng.loglik_grad(P.to(cuda), xd, yd)
  for j in range(2):
    loss, g = O.map_loss_and_grad(om, P[j], xd.cpu(), yd.cpu(), n, 0.0, 'NORMAL')
    assert abs(float(ll[j]) + float(loss)) <= 3e-2 * abs(float(loss))
    err = float((grad[j].cpu() + g).abs().max() / g.abs().max())
    assert err < 3e-2, err


def test_training_reduces_loss_full_size(cuda):
  """Size-independent property at the benchmark shape (config 2: W256 L2 E8, N=10k):
  full-batch MAP training monotonically (on average) reduces the loss."""
  from bayesnf_b200 import inference
  cfg = dict(_cfgs()['chickenpox'])
  n = 10440
  cfg['init_x'] = (n, 3)
  cfg['input_scales'] = [521.0, 1, 1]
  rng = np.random.default_rng(0)
  t = np.repeat(np.arange(522.0), 20)
  x = np.stack([t, np.tile(rng.normal(size=20), 522), np.tile(rng.normal(size=20), 522)], 1)
  y = 10 * np.sin(2 * np.pi * t / 52.1775) + 3 * x[:, 1] + rng.normal(size=n)
  params, losses = inference.fit_map(x, y, 1, 'NORMAL', cfg, 8, 0.005, 60, precision='fp32')
  assert losses.shape == (1, 8, 60) and np.isfinite(losses).all()
  assert (losses[0, :, -1] < losses[0, :, 0]).all()
  assert (np.diff(losses[0], axis=1) < 0).mean() > 0.9


@pytest.mark.parametrize('dist', ['NB', 'ZINB'])
def test_nb_predictive_quantiles(cuda, dist):
  """predict_bnf for count models (inference.py:271-333): distribution means and the discrete
  mixture quantiles vs the scipy-based oracle (f64 betainc on both sides -> exact integers)."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n = 120
  x, y = _data(cfg, n, counts=True)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 5, y, seed=21)
  spec = models.ModelSpec(**cfg, observation_model=dist)
  params = spec.unflatten(P.numpy()[None])
  qs = (0.5, 0.025, 0.975)
  means, quantiles = inference.predict_bnf(x, dist, params, cfg, qs, precision='fp32')
  assert means.shape == (1, 5, n) and len(quantiles) == 3
  xd, _ = inference._to_device_data(x, y)
  loc = torch.stack([om.forward(om.unflatten(P[j]), xd.cpu()) for j in range(5)]).numpy()
  pred = O.nb_predictive(loc, P[:, 1].numpy(), P[:, 2].numpy(), dist)
  np.testing.assert_allclose(means[0], pred['mean'], rtol=2e-4)
  for q, got in zip(qs, quantiles):
    want = O.nb_quantiles(pred, q)
    # identical integers except where the f32 network output moves a CDF value across q
    assert np.mean(got == want) >= 0.97, (q, got[:10], want[:10])
    assert np.abs(got - want).max() <= max(2.0, 0.02 * want.max())
  assert np.all(quantiles[1] <= quantiles[0]) and np.all(quantiles[0] <= quantiles[2])


def test_nb_estimator_predict_api(cuda):
  """End to end: ZINB estimator fit + predict through the public API."""
  import bayesnf_b200
  train, test = _chickenpox_tables()
  est = _chickenpox_estimator(bayesnf_b200.BayesianNeuralFieldMAP, precision='fp32')
  est.observation_model = 'ZINB'
  est.fit(train, seed=3, ensemble_size=3, learning_rate=0.005, num_epochs=8)
  means, q = est.predict(train, quantiles=(0.5, 0.9))
  assert means.shape == (1, 3, 100) and np.isfinite(means).all()
  assert q[0].shape == (100,) and np.all(q[0] <= q[1]) and np.all(q[0] == np.floor(q[0]))


def test_fit_map_replays_reference_batch_orders(cuda):
  """batch_order='jax': the minibatch row orders are those of the reference's threefry key tree
  (jax_prng.map_batch_orders) -- identical to injecting them through the batch_indices hook."""
  from bayesnf_b200 import inference, jax_prng
  cfg = _cfgs()['small']
  n = 96
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 3, y, seed=5).numpy()
  seed = np.array([0, 9], dtype=np.uint32)
  kw = dict(num_particles=3, learning_rate=0.01, num_epochs=3, batch_size=32, precision='fp32', init_params=P)
  _, l_jax = inference.fit_map(x, y, seed, 'NORMAL', cfg, batch_order='jax', **kw)
  orders = jax_prng.map_batch_orders(seed, 1, 3, n, 3)[:, 0]
  _, l_inj = inference.fit_map(x, y, seed, 'NORMAL', cfg, batch_indices=orders, **kw)
  np.testing.assert_allclose(l_jax, l_inj, rtol=1e-5)     # same orders; f32 atomics reorder sums
  assert np.isfinite(l_jax).all() and l_jax.shape == (1, 3, 3)


@pytest.mark.parametrize('prec,tol', [('fp32', 5e-5), ('bf16', 2e-2)])
def test_full_size_additivity_and_permutation_invariance(cuda, prec, tol):
  """Size-independent properties at BASELINE configs[1] (W256 L2, 8 members, N = 10 440), where
  the oracle is too slow to be the checker: the log-likelihood and its gradient are SUMS over
  rows (models.py:157-164, Independent(..., 1)), so (i) the values on two disjoint index halves
  add up to the full-batch values and (ii) a permutation of the rows changes nothing.  fp32:
  reduction-order noise only; bf16: the per-row values are identical, only f32 sums reorder, but
  small gradient leaves are cancellation-prone, hence the tolerance relative to the leaf scale."""
  from bayesnf_b200 import inference
  cfg = dict(_cfgs()['chickenpox'])
  n = 10440
  cfg['init_x'] = (n, 3)
  cfg['input_scales'] = [521.0, 1, 1]
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P = _random_params(om, 8, y, seed=11).to(cuda)
  eng, spec = _engine(cfg, 'NORMAL', prec)
  xd, yd = inference._to_device_data(x, y)
  ll, g = eng.loglik_grad(P, xd, yd)
  rng = np.random.default_rng(1)
  perm = rng.permutation(n).astype(np.int32)
  full = torch.tensor(np.tile(perm, (8, 1)), device=cuda)
  ll_p, g_p = eng.loglik_grad(P, xd, yd, idx=full)
  half = n // 2
  ll_a, g_a = eng.loglik_grad(P, xd, yd, idx=full[:, :half].contiguous())
  ll_b, g_b = eng.loglik_grad(P, xd, yd, idx=full[:, half:].contiguous())
  scale = g.abs().amax(dim=1, keepdim=True)
  assert float(((ll_p - ll) / ll).abs().max()) <= tol
  assert float(((ll_a + ll_b - ll) / ll).abs().max()) <= tol
  assert float(((g_p - g).abs() / scale).max()) <= tol
  assert float(((g_a + g_b - g).abs() / scale).max()) <= tol



def _chickenpox_full(n=10440):
  """BASELINE configs[1] at full size: 8 members x 10 440 rows (20 sites x 522 weeks), W256 L2."""
  cfg = dict(_cfgs()['chickenpox'])
  cfg['init_x'] = (n, 3)
  cfg['input_scales'] = [521.0, 1, 1]
  rng = np.random.default_rng(0)
  t = np.repeat(np.arange(522.0), 20)
  x = np.stack([t, np.tile(rng.normal(size=20), 522), np.tile(rng.normal(size=20), 522)], 1)
  y = 10 * np.sin(2 * np.pi * t / 52.1775) + 3 * x[:, 1] + rng.normal(size=n)
  return cfg, x, y


# mode -> (rel. tol. on the loss, gradient leaf tol. as a fraction of the leaf's max-abs,
#          abs. tol. on the 99.9 % quantile of |param - oracle param| after 5 Adam steps of lr 0.005)
FULL_SIZE_TOL = {'fp32': (2e-5, 1e-4, 1e-4), 'bf16x3': (2e-5, 1e-4, 1e-4), 'bf16': (2e-2, 5e-2, 5e-3)}


@pytest.mark.parametrize('prec', ['fp32', 'bf16x3', 'bf16'])
def test_full_size_configs1_against_oracle(cuda, prec):
  """BASELINE configs[1] AT FULL SIZE (8 members x 10 440 rows, width 256, depth 2) directly against
  the CPU oracle: the log-likelihood, every gradient leaf (vs the oracle's float64 twin) and the
  parameters + loss curve after 5 full-batch MAP Adam steps (vs the f32 oracle's own
  value_and_grad + optax.adam restatement, inference.py:599-608), for the SIMT f32 mode, the
  tensor-core f32-parity mode (same tolerances) and the single-pass bf16 mode (its own, looser,
  stated tolerances: bf16 has 8 significand bits)."""
  from bayesnf_b200 import inference, models
  ll_tol, g_tol, p_tol = FULL_SIZE_TOL[prec]
  cfg, x, y = _chickenpox_full()
  n, E, steps, lr = len(y), 8, 5, 0.005
  om, om64 = O.OracleModel(**cfg), O.OracleModel(**cfg, dtype=torch.float64)
  P0 = _random_params(om, E, y, seed=17)
  eng, spec = _engine(cfg, 'NORMAL', prec)
  xd, yd = inference._to_device_data(x, y)
  ll, grad = eng.loglik_grad(P0.to(cuda), xd, yd)
  ll, grad = ll.cpu(), grad.cpu()
  parts = [(0, 1)] + [(o, o + (int(np.prod(s)) if s else 1)) for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
  xc, yc = xd.cpu(), yd.cpu()
  for j in range(E):
    loss64, g64 = O.map_loss_and_grad(om64, P0[j].double(), xc.double(), yc.double(), n, 0.0, 'NORMAL')
    assert abs(float(ll[j]) + float(loss64)) <= ll_tol * abs(float(loss64)), (prec, j, float(ll[j]), float(loss64))
    for (a, b) in parts:
      want, got = -g64[a:b], grad[j, a:b].double()
      # bf16: scalar leaves that are cancellation-prone sums get the usual floor of 1e-3 of the
      # network's largest gradient (as in tests/test_gpu_tc.py); the f32-class modes get none
      floor = 1e-3 * g_tol if prec == 'bf16' else 1e-7
      tol = g_tol * float(want.abs().max()) + floor * float(g64.abs().max()) + 1e-7
      err = float((got - want).abs().max())
      assert err <= tol, (prec, j, a, b, err, tol)
  # ---- 5 Adam steps (prior + optax.adam + restaged weights) from the same initial parameters
  params, losses = inference.fit_map(x, y, 0, 'NORMAL', cfg, num_particles=E, learning_rate=lr,
                                     num_epochs=steps, precision=prec, init_params=P0.numpy())
  flat = models.ModelSpec(**cfg).flatten(params)[0]
  for j in range(E):
    pj, lj = O.fit_map_member(om, P0[j], xc, yc, lambda ep: torch.arange(n), steps, n, lr, 1.0, 'NORMAL')
    np.testing.assert_allclose(losses[0, j], lj.numpy(), rtol=max(ll_tol, 2e-5))
    d = np.abs(flat[j] - pj.numpy())
    # Adam normalises the step: an entry whose gradient is pure summation noise may walk the
    # other way (|delta| up to 2*lr per step), hence a quantile for the bulk and a hard cap
    assert float(np.quantile(d, 0.999)) <= p_tol, (prec, j, float(np.quantile(d, 0.999)))
    assert float(d.max()) <= 2 * lr * steps + 1e-6


# ---------------------------------------------------------------------------------------------
# device-side minibatch machinery
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('prec', PARITY_MODES)
def test_device_shuffled_epochs_match_injected_orders(cuda, prec):
  """fit_map with batch_size < N draws the per-member permutations, the batch windows and every
  step on the device (bnf_map_epochs, one CUDA graph); replaying the SAME orders -- computed on
  the host by the same keyed permutation -- through the injected-index path (which
  test_map_steps_fp32 pins to the oracle) gives the same losses and parameters."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n, B, epochs, E = 200, 64, 3, 3
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P0 = _random_params(om, E, y, seed=13).numpy()
  seed = 1234
  kw = dict(num_particles=E, learning_rate=0.01, num_epochs=epochs, batch_size=B, precision=prec, init_params=P0)
  p_dev, l_dev = inference.fit_map(x, y, seed, 'NORMAL', cfg, **kw)
  opt_seed = inference.fold_in(inference.seed_to_int(seed), 0x1002)
  orders = np.stack([[inference.device_permutation(opt_seed, j, ep, n) for j in range(E)] for ep in range(epochs)])
  p_inj, l_inj = inference.fit_map(x, y, seed, 'NORMAL', cfg, batch_indices=orders, **kw)
  assert l_dev.shape == (1, E, epochs) and np.isfinite(l_dev).all()
  np.testing.assert_allclose(l_dev, l_inj, rtol=2e-5)
  spec = models.ModelSpec(**cfg)
  assert float(np.abs(spec.flatten(p_dev) - spec.flatten(p_inj)).max()) <= 2e-4
  # the orders are per member and per epoch, and the ragged tail (200 - 3*64 rows) is dropped
  assert not np.array_equal(orders[0, 0], orders[0, 1]) and not np.array_equal(orders[0, 0], orders[1, 0])


def test_vi_five_steps_with_subbatch_against_oracle(cuda):
  """Five VI steps (tfp.vi.fit_surrogate_posterior_stateless as driven by ensemble_vi,
  inference.py:687-739) with a fresh shared sub-batch per step and injected eps against the f64
  oracle's autograd + optax.adam on (mu, rho): losses to 2e-5, variational parameters to 1e-4
  (99.5 % quantile; they move by up to 5*lr = 0.05)."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n, B, S, E, steps = 200, 80, 3, 2, 5
  x, y = _data(cfg, n)
  om64 = O.OracleModel(**cfg, dtype=torch.float64)
  om = O.OracleModel(**cfg)
  spec = models.ModelSpec(**cfg)
  P = spec.num_params
  g = torch.Generator().manual_seed(5)
  mu = _random_params(om, E, y, seed=9)
  mu[:, 0] = 0.0
  rho = torch.full((E, P), O.SOFTPLUS_INV_0P3) + 0.05 * torch.randn(E, P, generator=g)
  eps = torch.randn(steps, S, E, P, generator=g)
  rng = np.random.default_rng(3)
  rows = np.stack([rng.permutation(n)[:B] for _ in range(steps)]).astype(np.int32)
  kl, lr = 0.1, 0.01
  for prec in PARITY_MODES:
    sur, losses, _ = inference.fit_vi(
        x, y, 0, 'NORMAL', cfg, ensemble_size=E, learning_rate=lr, num_epochs=steps, sample_size_divergence=S,
        sample_size_posterior=2, kl_weight=kl, batch_size=B, precision=prec, init_params=(mu.numpy(), rho.numpy()),
        eps=eps.numpy(), posterior_eps=np.zeros((2, E, P), np.float32), batch_indices=rows)
    xd, yd = inference._to_device_data(x, y)
    mu1, rho1 = spec.flatten(sur.loc)[0], spec.flatten(sur.inv_softplus_scale)[0]
    for e in range(E):
      m_, r_ = mu[e].double(), rho[e].double()
      am, av = torch.zeros(2 * P, dtype=torch.float64), torch.zeros(2 * P, dtype=torch.float64)
      for t in range(steps):
        sel = torch.tensor(rows[t].astype(np.int64))
        loss, gmu, grho = O.vi_loss_and_grad(om64, m_, r_, eps[t, :, e].double(), xd.cpu().double()[sel],
                                             yd.cpu().double()[sel], n, kl, 'NORMAL')
        assert abs(losses[0, e, t] - float(loss) * kl) <= 2e-5 * abs(float(loss) * kl), (prec, e, t)
        pq, am, av = O.adam_update(torch.cat([m_, r_]), torch.cat([gmu, grho]), am, av, t + 1, lr)
        m_, r_ = pq[:P], pq[P:]
      # entries whose gradient is summation noise may step the other way under Adam's normalisation
      dm = (torch.tensor(mu1[e]).double() - m_).abs()
      dr = (torch.tensor(rho1[e]).double() - r_).abs()
      assert float(torch.quantile(dm, 0.995)) <= 1e-4 and float(dm.max()) <= 2 * steps * lr
      assert float(torch.quantile(dr, 0.995)) <= 1e-4 and float(dr.max()) <= 2 * steps * lr


@pytest.mark.parametrize('steps', [6, 19])
def test_vi_device_steps_graph_equals_direct_launches(cuda, monkeypatch, steps):
  """bnf_vi_steps (eps + per-step sub-batch drawn on the device, one CUDA graph replayed) is
  deterministic in its seed and equals the same call issued as direct launches, and -- split
  into single-step calls -- the same sequence (device-side step counters)."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n, B, S, E = 200, 96, 4, 3           # 19 steps: two unrolled graph launches (8 steps each) + three single
  x, y = _data(cfg, n)
  spec = models.ModelSpec(**cfg)
  eng = inference.Engine(spec, 'fp32')
  xd, yd = inference._to_device_data(x, y)

  def run(chunks):
    mu = eng.init_params(0.0, 5, 0, E)
    rho = torch.full_like(mu, O.SOFTPLUS_INV_0P3)
    am = torch.zeros((E, 2, spec.num_params), dtype=torch.float32, device=cuda)
    av, sc = torch.zeros_like(am), torch.zeros(1, dtype=torch.int32, device=cuda)
    ls = torch.cat([eng.vi_steps(mu, rho, am, av, sc, S, 77, 0, xd, yd, B, n, k, 0.01, 0.1) for k in chunks])
    torch.cuda.synchronize()
    assert int(sc[0]) == steps
    return ls.cpu().numpy(), mu.cpu().numpy(), rho.cpu().numpy()
  monkeypatch.delenv('BNF_NO_GRAPH', raising=False)
  a = run([steps])
  b = run([steps])
  monkeypatch.setenv('BNF_NO_GRAPH', '1')
  c = run([steps])
  d = run([1] * steps)
  assert np.isfinite(a[0]).all() and a[0].shape == (steps, E)
  for other in (b, c, d):
    np.testing.assert_allclose(a[0], other[0], rtol=2e-5 * max(1, steps // 6))    # f32 atomics reorder sums
    for k in (1, 2):
      diff = np.abs(a[k] - other[k])
      # (Adam normalises the step: over many steps an entry whose gradient is summation noise may
      # walk the other way; everything else stays together)
      assert float(np.quantile(diff, 0.995)) <= 2e-4 and float(diff.max()) <= (2e-4 if steps <= 6 else 2 * steps * 0.01)
  assert a[0][-1].mean() < a[0][0].mean()                          # the ELBO loss goes down


def test_device_eps_are_independent_draws(cuda):
  """The reparameterisation noise / posterior samples drawn on the device are iid N(0,1): every
  element owns its Philox words, so neighbouring coordinates share nothing
  (corr(eps_i^2, eps_{i+1}^2) = 0; a shared uniform would give ~ -0.12)."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['chickenpox']
  spec = models.ModelSpec(**cfg)
  eng = inference.Engine(spec, 'fp32')
  E, n_s = 2, 8
  mu = torch.zeros((E, spec.num_params), dtype=torch.float32, device=cuda)
  rho = torch.full_like(mu, math.log(math.expm1(1.0 - 1e-4)))       # sigma = 1
  z = eng.vi_sample(mu, rho, n_s, 123).reshape(-1).double().cpu()
  assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
  assert abs(float((z ** 4).mean()) - 3.0) < 0.05
  for lag in (1, 2, 3, 4):
    a, b = z[:-lag], z[lag:]
    assert abs(float(torch.corrcoef(torch.stack([a, b]))[0, 1])) < 5e-3, lag
    assert abs(float(torch.corrcoef(torch.stack([a * a, b * b]))[0, 1])) < 5e-3, lag
  z2 = eng.vi_sample(mu, rho, n_s, 124).reshape(-1).double().cpu()
  assert abs(float(torch.corrcoef(torch.stack([z, z2]))[0, 1])) < 5e-3


def test_num_splits_are_independent_sequential_fits(cuda):
  """fit_map(num_splits=2) (inference.py:432-457): particles // splits members are trained per
  split, one after the other, and concatenated on axis 1 -- equal to two separate fits of the
  two halves of the injected initial parameters."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  n = 200
  x, y = _data(cfg, n)
  om = O.OracleModel(**cfg)
  P0 = _random_params(om, 4, y, seed=31).numpy()
  kw = dict(learning_rate=0.01, num_epochs=4, precision='fp32')
  p_all, l_all = inference.fit_map(x, y, 3, 'NORMAL', cfg, num_particles=4, num_splits=2, init_params=P0, **kw)
  spec = models.ModelSpec(**cfg)
  flat = spec.flatten(p_all)
  assert flat.shape == (1, 4, spec.num_params) and l_all.shape == (1, 4, 4)
  for i in range(2):
    p_i, l_i = inference.fit_map(x, y, 3, 'NORMAL', cfg, num_particles=2, init_params=P0[2 * i:2 * i + 2], **kw)
    np.testing.assert_allclose(l_all[0, 2 * i:2 * i + 2], l_i[0], rtol=2e-5)
    assert float(np.abs(flat[0, 2 * i:2 * i + 2] - spec.flatten(p_i)[0]).max()) <= 2e-4
    # and each half against the oracle's own fit
    xd, yd = inference._to_device_data(x, y)
    for j in range(2):
      pj, lj = O.fit_map_member(om, torch.tensor(P0[2 * i + j]), xd.cpu(), yd.cpu(), lambda ep: torch.arange(n), 4, n,
                                0.01, 1.0, 'NORMAL')
      np.testing.assert_allclose(l_all[0, 2 * i + j], lj.numpy(), rtol=2e-5)
  # floor division of particles over splits (inference.py:445): 5 // 2 = 2 per split
  p5, l5 = inference.fit_map(x, y, 3, 'NORMAL', cfg, num_particles=5, num_splits=2, **kw)
  assert l5.shape == (1, 4, 4)
  # different splits get different seeds (fold_in(seed, i), :436-437): the members differ
  f5 = spec.flatten(p5)[0]
  assert not np.allclose(f5[0], f5[2])


def test_fit_vi_default_init_is_make_vi_init(cuda):
  """make_vi_init (inference.py:203-231): surrogate means = 0 for every non-2-D leaf (including
  log_noise_scale -- NOT log(std/2)), TruncatedNormal kernels; inverse-softplus scales =
  softplus^-1(0.3) everywhere.  Checked through fit_vi with a learning rate of 0."""
  from bayesnf_b200 import inference, models
  cfg = _cfgs()['small']
  x, y = _data(cfg, 200)
  sur, losses, samples = inference.fit_vi(x, y, 11, 'NORMAL', cfg, ensemble_size=3, learning_rate=0.0, num_epochs=1,
                                          sample_size_divergence=2, sample_size_posterior=2, kl_weight=0.1,
                                          precision='fp32')
  spec = models.ModelSpec(**cfg)
  for r in sur.inv_softplus_scale:
    np.testing.assert_allclose(r, O.SOFTPLUS_INV_0P3, rtol=1e-6)
  for sd in sur.stddev():
    np.testing.assert_allclose(sd, 0.3 + 1e-4, rtol=1e-5)
  assert np.all(sur.loc[0] == 0) and np.all(sur.loc[1] == 0) and np.all(sur.loc[2] == 0)
  for name, leaf in zip(spec.leaf_names, sur.loc[3:]):
    if name.endswith('kernel'):
      assert np.abs(leaf).max() <= 2.0 and leaf.std() > 0.5
    else:
      assert np.all(leaf == 0), name
  assert losses.shape == (1, 3, 1) and np.isfinite(losses).all()
