"""Parity against values computed by the REFERENCE'S OWN CODE.

tests/golden/numerics_*.npz were produced by scripts/make_golden_numerics.py, which imports
/root/reference/src/bayesnf/{models,inference}.py unmodified and executes them over
oracle/jaxshim.py (a float64 stand-in for the jax / flax / optax / TFP calls they make; none of
those packages can be installed in the build container).  The model forward, the three
log-likelihoods, the prior, their gradients, `fit_map` end to end (full batch, ragged minibatches in
the threefry permutation order, num_splits, prior_weight 0 / 1), `fit_vi` and `predict_bnf` are the
reference's functions; only the third-party primitives underneath are restated.

CPU tests: the oracle restatement reproduces those values to float64 round-off, which is what pins
the oracle.  GPU tests: the CUDA path (SIMT f32 and the tensor-core bf16x3 mode) against the same
values directly, at the float32 tolerances of tests/test_gpu_parity.py.
"""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import bnf_oracle as O

T64 = torch.float64
MODEL_CASES = ['chickenpox', 'odd', 'deep']
MAP_CASES = ['normal_full', 'normal_minibatch_splits', 'zinb_mle_minibatch', 'nb_full']
VI_CASES = ['normal_full', 'normal_subbatch']
DISTS = ['NORMAL', 'NB', 'ZINB']


def _load(kind, name):
  g = np.load(os.path.join(GOLDEN, f'numerics_{kind}_{name}.npz'))
  args = json.loads(str(g['model_args']))
  args['init_x'] = tuple(args['init_x'])
  args['interactions'] = np.asarray(args['interactions'], dtype=int).reshape(-1, 2)
  meta = json.loads(str(g['meta'])) if 'meta' in g.files else None
  return g, args, meta


def _t(a):
  return torch.tensor(np.asarray(a, dtype=np.float64))


def _relmax(got, want):
  want = np.asarray(want, dtype=np.float64)
  return float(np.abs(np.asarray(got, dtype=np.float64) - want).max() / max(np.abs(want).max(), 1e-300))


# --------------------------------------------------------------------------
# CPU: the oracle against the reference-executed values (float64 round-off)
# --------------------------------------------------------------------------
@pytest.mark.parametrize('name', MODEL_CASES)
def test_parameter_layout_is_the_reference_template(name):
  """Leaf order and shapes of `make_model`'s template (tree_leaves of the Flax dict): the oracle's
  derivation and the C plan (`bnf_plan_create`) agree with it."""
  from bayesnf_b200 import models
  g, args, _ = _load('model', name)
  names = [str(n) for n in g['leaf_names']]
  shapes = [tuple(s) for s in json.loads(str(g['leaf_shapes']))]
  om = O.OracleModel(**args)
  assert om.leaf_names == names and [tuple(s) for s in om.leaf_shapes] == shapes
  if name == 'deep':
    assert names.index('Dense_10/bias') < names.index('Dense_2/bias')     # string sort of the keys
  spec = models.ModelSpec(**args)
  assert spec.num_params == g['NORMAL_params'].size
  assert [tuple(s) for s in spec.leaf_shapes] == shapes


@pytest.mark.parametrize('dist', DISTS)
@pytest.mark.parametrize('name', MODEL_CASES)
def test_oracle_model_against_reference_execution(name, dist):
  g, args, _ = _load('model', name)
  om = O.OracleModel(**args, dtype=T64)
  x, y, flat = _t(g[f'{dist}_x']), _t(g[f'{dist}_y']), _t(g[f'{dist}_params'])
  assert _relmax(om.forward(om.unflatten(flat), x), g[f'{dist}_pred']) <= 1e-11
  f = flat.clone().requires_grad_(True)
  ll = O.log_likelihood(om, om.unflatten(f), x, y, dist)
  (gl,) = torch.autograd.grad(ll, f, allow_unused=True)
  assert abs(float(ll.detach()) - float(g[f'{dist}_loglik'])) <= 1e-11 * abs(float(g[f'{dist}_loglik']))
  assert _relmax(gl, g[f'{dist}_loglik_grad']) <= 1e-10
  f = flat.clone().requires_grad_(True)
  lp = O.prior_log_prob(om.unflatten(f))
  (gp,) = torch.autograd.grad(lp, f)
  assert abs(float(lp.detach()) - float(g[f'{dist}_prior'])) <= 1e-11 * abs(float(g[f'{dist}_prior']))
  assert _relmax(gp, g[f'{dist}_prior_grad']) <= 1e-12


def _map_orders(meta, n, split):
  from bayesnf_b200 import jax_prng
  per = meta['particles'] // meta['num_splits']
  return jax_prng.map_batch_orders(meta['seed'], 1, per, n, meta['epochs'],
                                   split_index=split if meta['num_splits'] > 1 else None)[:, 0]


@pytest.mark.parametrize('name', MAP_CASES)
def test_oracle_fit_map_against_reference_execution(name):
  """inference.fit_map run by the reference's code: same initial draws -> same parameters and
  per-epoch losses, including the minibatch order the threefry key tree gives
  (`jax_prng.map_batch_orders`) and the fold_in of num_splits."""
  g, args, meta = _load('map', name)
  om = O.OracleModel(**args, dtype=T64)
  x, y = _t(g['x']), _t(g['y'])
  n = y.shape[0]
  bs = meta['batch_size'] or n
  per = meta['particles'] // meta['num_splits']
  assert g['losses'].shape == (1, meta['particles'], meta['epochs'])
  for split in range(meta['num_splits']):
    orders = _map_orders(meta, n, split)
    for m in range(per):
      e = split * per + m
      order = (lambda ep: torch.arange(n)) if bs >= n else \
          (lambda ep, m=m: torch.from_numpy(orders[ep, m].astype(np.int64)))
      p, losses = O.fit_map_member(om, _t(g['init'][0, e]), x, y, order, meta['epochs'], bs,
                                   meta['lr'], meta['prior_weight'], meta['dist'])
      assert float(np.abs(p.numpy() - g['final'][0, e]).max()) <= 1e-10
      assert _relmax(losses, g['losses'][0, e]) <= 1e-12
  assert float(np.abs(g['final'] - g['init']).max()) > 0.01          # the fit moved


def test_map_initial_draws_follow_make_init_fn():
  """`_make_init_fn` as executed: log_noise_scale = log(nanstd(y) / 2), 2-D kernels inside [-2, 2]
  and not zero, every other leaf exactly zero (inference.py:399-427)."""
  g, args, meta = _load('map', 'normal_full')
  om = O.OracleModel(**args)
  init = g['init'][0]
  assert np.allclose(init[:, 0], np.log(np.nanstd(g['y'].astype(np.float64)) / 2.0), rtol=1e-6)   # nanstd of the f32 target
  assert np.all(init[:, 1:3] == 0)
  o = 3
  for nm, s in zip(om.leaf_names, om.leaf_shapes):
    k = int(np.prod(s)) if len(s) else 1
    blk = init[:, o:o + k]
    if len(s) == 2:
      assert np.abs(blk).max() <= 2.0 and np.abs(blk).min() > 0 and 0.7 < blk.std() < 1.0, nm
    else:
      assert np.all(blk == 0), nm
    o += k


@pytest.mark.parametrize('name', VI_CASES)
def test_oracle_fit_vi_against_reference_execution(name):
  """inference.fit_vi (ensemble_vi's target / surrogate builders, make_vi_init) run by the
  reference's code with recorded noise and sub-batches: the oracle's reverse-KL loss + Adam on
  (mu, rho) lands on the same surrogate and losses (x kl_weight, inference.py:756)."""
  g, args, meta = _load('vi', name)
  om = O.OracleModel(**args, dtype=T64)
  x, y = _t(g['x']), _t(g['y'])
  n, bs, kl = y.shape[0], meta['batch_size'], meta['kl_weight']
  assert np.allclose(g['rho0'], O.SOFTPLUS_INV_0P3, rtol=1e-14)            # make_vi_init
  for e in range(meta['ensemble']):
    mu, rho = _t(g['mu0'][e]), _t(g['rho0'][e])
    sm = [torch.zeros_like(mu) for _ in range(4)]
    for t in range(meta['epochs']):
      rows = torch.arange(n) if bs is None else torch.from_numpy(g['perm'][t][:bs])
      loss, gmu, grho = O.vi_loss_and_grad(om, mu, rho, _t(g['eps'][t, :, e]), x[rows], y[rows], n,
                                           kl, meta['dist'])
      mu, sm[0], sm[1] = O.adam_update(mu, gmu, sm[0], sm[1], t + 1, meta['lr'])
      rho, sm[2], sm[3] = O.adam_update(rho, grho, sm[2], sm[3], t + 1, meta['lr'])
      assert abs(float(loss) * kl - g['losses'][0, e, t]) <= 1e-11 * abs(g['losses'][0, e, t])
    assert float(np.abs(mu.numpy() - g['mu'][e]).max()) <= 1e-10
    assert float(np.abs(rho.numpy() - g['rho'][e]).max()) <= 1e-10


def _predict_inputs(name):
  g, args, meta = _load('predict', name)
  om = O.OracleModel(**args, dtype=T64)
  x, P = _t(g['x']), _t(g['params'][0])
  loc = torch.stack([om.forward(om.unflatten(P[m]), x) for m in range(P.shape[0])])
  return g, args, meta, P, loc


def test_oracle_predict_normal_against_reference_execution():
  g, _, _, P, means = _predict_inputs('normal')
  scales = (0.01 + torch.exp(P[:, 0]))[:, None]
  assert _relmax(means, g['means'][0]) <= 1e-11
  for i, q in enumerate(g['quantiles']):
    assert _relmax(O.approximate_normal_quantile(means, scales, float(q)), g['q_approx'][i]) <= 1e-11
    # root quantiles: judged by the CDF residual (the reference stops within 1e-5 of the root)
    assert float(O.mixture_cdf_residual(means, scales, _t(g['q_root'][i]), float(q)).abs().max()) <= 1e-12
    mine = O.normal_quantile_via_root(means, scales, float(q))
    assert float(O.mixture_cdf_residual(means, scales, mine, float(q)).abs().max()) <= 1e-5


@pytest.mark.parametrize('dist', ['nb', 'zinb'])
def test_oracle_predict_counts_against_reference_execution(dist):
  g, _, _, P, loc = _predict_inputs(dist)
  pred = O.nb_predictive(loc.numpy(), P[:, 1].numpy(), P[:, 2].numpy(), dist.upper())
  assert _relmax(pred['mean'], g['means'][0]) <= 1e-11
  for i, q in enumerate(g['quantiles']):
    assert np.array_equal(O.nb_quantiles(pred, float(q)), g['q_root'][i])


# ---- the estimator classes (spatiotemporal.py) run by the reference on a 4-location table ----
def _estimator_golden(kind):
  import io
  import pandas as pd
  g = np.load(os.path.join(GOLDEN, f'numerics_estimator_{kind}.npz'))
  train = pd.read_csv(io.StringIO(str(g['train_csv'])), parse_dates=['datetime'])
  test = pd.read_csv(io.StringIO(str(g['test_csv'])), parse_dates=['datetime'])
  args = json.loads(str(g['model_args']))
  args['init_x'] = tuple(args['init_x'])
  args['interactions'] = np.asarray(args['interactions'], dtype=int).reshape(-1, 2)
  return g, train, test, args, json.loads(str(g['fit'])), json.loads(str(g['estimator_kwargs']))


def _my_estimator(kind, kwargs, precision=None):
  from bayesnf_b200 import spatiotemporal
  cls = {'map': spatiotemporal.BayesianNeuralFieldMAP, 'mle': spatiotemporal.BayesianNeuralFieldMLE,
         'vi': spatiotemporal.BayesianNeuralFieldVI}[kind]
  return cls(precision=precision, **kwargs)


@pytest.mark.parametrize('kind', ['map', 'mle', 'vi'])
def test_estimator_host_side_against_reference_execution(kind):
  """Data handler outputs and `_model_args` of the estimator mirror equal the reference's."""
  g, train, test, args, fit, kwargs = _estimator_golden(kind)
  est = _my_estimator(kind, kwargs)
  np.testing.assert_array_equal(est.data_handler.get_train(train), g['x_train'])
  np.testing.assert_array_equal(est.data_handler.get_target(train), g['y_train'])
  np.testing.assert_array_equal(est.data_handler.get_test(test), g['x_test'])
  mine = est._model_args((fit['batch_size'] or len(train), 3))
  assert set(mine) == set(args)
  for k in args:
    np.testing.assert_array_equal(np.asarray(mine[k], dtype=np.float64).reshape(-1),
                                  np.asarray(args[k], dtype=np.float64).reshape(-1), err_msg=k)


@pytest.mark.parametrize('kind', ['map', 'mle'])
def test_oracle_estimator_fit_predict_against_reference_execution(kind):
  """BayesianNeuralField{MAP,MLE}.fit(table, seed).predict(table, quantiles) as run by the
  reference: the oracle from the same initial draws and threefry batch orders."""
  from bayesnf_b200 import jax_prng
  g, train, test, args, fit, _ = _estimator_golden(kind)
  om = O.OracleModel(**args, dtype=T64)
  x, y = _t(g['x_train']), _t(g['y_train'])
  n, bs = y.shape[0], fit['batch_size']
  orders = jax_prng.map_batch_orders(fit['seed'], 1, fit['ensemble_size'], n, fit['num_epochs'])[:, 0]
  finals = []
  for e in range(fit['ensemble_size']):
    p, losses = O.fit_map_member(om, _t(g['init'][0, e]), x, y,
                                 lambda ep, e=e: torch.from_numpy(orders[ep, e].astype(np.int64)),
                                 fit['num_epochs'], bs, fit['learning_rate'],
                                 1.0 if kind == 'map' else 0.0, 'NORMAL')
    assert float(np.abs(p.numpy() - g['params'][0, e]).max()) <= 1e-10
    assert _relmax(losses, g['losses'][0, e]) <= 1e-12
    finals.append(p)
  P = torch.stack(finals)
  xt = _t(g['x_test'])
  # likelihood_model(test).log_prob(y_test): one log-likelihood per member (spatiotemporal.py:433-468)
  for m in range(P.shape[0]):
    ll = O.log_likelihood(om, om.unflatten(P[m]), xt, _t(g['y_test']), 'NORMAL')
    assert abs(float(ll) - g['lm_log_prob'][0, m]) <= 1e-9 * abs(g['lm_log_prob'][0, m])
  means = torch.stack([om.forward(om.unflatten(P[m]), xt) for m in range(P.shape[0])])
  scales = (0.01 + torch.exp(P[:, 0]))[:, None]
  assert _relmax(means, g['means'][0]) <= 1e-9
  for i, q in enumerate(g['quantiles']):
    assert _relmax(O.approximate_normal_quantile(means, scales, float(q)), g['q_approx'][i]) <= 1e-9
    assert float(O.mixture_cdf_residual(means, scales, _t(g['q_root'][i]), float(q)).abs().max()) <= 1e-9


def test_oracle_vi_estimator_against_reference_execution():
  """BayesianNeuralFieldVI: surrogate after the fit, posterior samples (`params_` = mu + sigma * eps
  with sigma = 1e-4 + softplus(rho), inference.py:711-716, :741-745) and the predictive means over
  (sample, member)."""
  g, train, test, args, fit, _ = _estimator_golden('vi')
  om = O.OracleModel(**args, dtype=T64)
  x, y = _t(g['x_train']), _t(g['y_train'])
  n, kl = y.shape[0], fit['kl_weight']
  for e in range(fit['ensemble_size']):
    mu, rho = _t(g['mu0'][e]), _t(g['rho0'][e])
    sm = [torch.zeros_like(mu) for _ in range(4)]
    for t in range(fit['num_epochs']):
      loss, gmu, grho = O.vi_loss_and_grad(om, mu, rho, _t(g['eps'][t, :, e]), x, y, n, kl, 'NORMAL')
      mu, sm[0], sm[1] = O.adam_update(mu, gmu, sm[0], sm[1], t + 1, fit['learning_rate'])
      rho, sm[2], sm[3] = O.adam_update(rho, grho, sm[2], sm[3], t + 1, fit['learning_rate'])
      assert abs(float(loss) * kl - g['losses'][0, e, t]) <= 1e-11 * abs(g['losses'][0, e, t])
    assert float(np.abs(mu.numpy() - g['mu'][e]).max()) <= 1e-10
    sigma = 1e-4 + torch.nn.functional.softplus(rho)
    for s_ in range(fit['sample_size_posterior']):
      z = mu + sigma * _t(g['posterior_eps'][s_, e])
      assert float(np.abs(z.numpy() - g['params'][0, s_, e]).max()) <= 1e-10
      pred = om.forward(om.unflatten(z), _t(g['x_test']))
      assert _relmax(pred, g['means'][0, s_, e]) <= 1e-9



@pytest.mark.skipif(not os.path.isdir('/root/reference/src/bayesnf'),
                    reason='the reference tree exists only in the build container')
def test_goldens_are_what_the_reference_computes(tmp_path):
  """Re-runs the generator (the reference's code over the shim) and compares with the committed
  files, so the fixtures cannot drift from the script or from the reference."""
  env = dict(os.environ, BNF_GOLDEN_OUT=str(tmp_path))
  subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'make_golden_numerics.py')],
                 check=True, env=env, capture_output=True, timeout=600)
  fresh = sorted(glob.glob(os.path.join(str(tmp_path), 'numerics_*.npz')))
  assert len(fresh) == len(glob.glob(os.path.join(GOLDEN, 'numerics_*.npz'))) == 15
  for f in fresh:
    a, b = np.load(f), np.load(os.path.join(GOLDEN, os.path.basename(f)))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
      if a[k].dtype.kind in 'fc':
        np.testing.assert_allclose(a[k], b[k], rtol=1e-11, atol=1e-12, err_msg=f'{f}:{k}')
      else:
        assert np.array_equal(a[k], b[k]), (f, k)


# --------------------------------------------------------------------------
# GPU: the CUDA path against the same reference-executed values
# --------------------------------------------------------------------------
PARITY_MODES = ['fp32', 'bf16x3']


@pytest.fixture(scope='module')
def cuda():
  assert torch.cuda.is_available(), 'gpu tests need a CUDA device (no fallback)'
  torch.cuda.set_device(0)
  return torch.device('cuda', 0)


def _engine(args, dist, prec):
  from bayesnf_b200 import inference, models
  spec = models.ModelSpec(**args, observation_model=dist)
  if prec == 'bf16x3' and not inference.precision_supported(spec, prec):
    pytest.skip(f'width {args["width"]} is not a tensor-core shape; fp32 covers it')
  return inference.Engine(spec, prec), spec


def _leaf_ranges(spec):
  return [(0, 1), (1, 2), (2, 3)] + [(o, o + (int(np.prod(s)) if len(s) else 1))
                                     for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]


@pytest.mark.gpu
@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('dist', DISTS)
@pytest.mark.parametrize('name', MODEL_CASES)
def test_cuda_model_against_reference_execution(cuda, name, dist, prec):
  """Forward <= 1e-5 of the output scale, log-likelihood <= 2e-5 relative, every gradient leaf
  <= 1e-4 of the leaf's scale -- against the reference's own float64 values; the only slack added
  is twice what the float32 oracle itself is away from them (float32 rounding of the inputs of
  sin / cos at large arguments is not the kernel's)."""
  from bayesnf_b200 import inference
  g, args, _ = _load('model', name)
  eng, spec = _engine(args, dist, prec)
  x, y = g[f'{dist}_x'], g[f'{dist}_y']
  P = torch.tensor(g[f'{dist}_params'])[None].repeat(2, 1)
  xd, yd = inference._to_device_data(x, y)
  loc = eng.forward(P.to(cuda), xd).cpu()[0].double().numpy()
  om32 = O.OracleModel(**args)
  f32 = P[0].clone().requires_grad_(True)
  pred32 = om32.forward(om32.unflatten(f32), torch.tensor(x))
  ll32 = O.log_likelihood(om32, om32.unflatten(f32), torch.tensor(x), torch.tensor(y), dist)
  (g32,) = torch.autograd.grad(ll32, f32, allow_unused=True)
  want = g[f'{dist}_pred']
  scale = np.abs(want).max()
  slack = 2 * np.abs(pred32.detach().double().numpy() - want).max()
  assert np.abs(loc - want).max() <= 1e-5 * scale + 1e-6 + slack, (np.abs(loc - want).max(), scale, slack)
  ll, grad = eng.loglik_grad(P.to(cuda), xd, yd)
  ll, grad = ll.cpu().double().numpy(), grad.cpu().double().numpy()
  wl = float(g[f'{dist}_loglik'])
  assert abs(ll[0] - wl) <= 2e-5 * abs(wl) + 1e-4 + 2 * abs(float(ll32.detach()) - wl), (ll[0], wl)
  wg = g[f'{dist}_loglik_grad']
  g32 = g32.detach().double().numpy()
  for a, b in _leaf_ranges(spec):
    tol = 1e-4 * np.abs(wg[a:b]).max() + 1e-7 * np.abs(wg).max() + 1e-7 + 2 * np.abs(g32[a:b] - wg[a:b]).max()
    err = np.abs(grad[0, a:b] - wg[a:b]).max()
    assert err <= tol, (name, dist, prec, a, b, err, tol)
  assert np.abs(grad[0] - grad[1]).max() <= 1e-5 * np.abs(wg).max()   # same member twice (f32 atomics reorder sums)


@pytest.mark.gpu
@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('name', MAP_CASES)
def test_cuda_fit_map_against_reference_execution(cuda, name, prec):
  """`fit_map` through the public API from the reference run's initial draws, minibatches in the
  reference's threefry order (`batch_order='jax'`): per-epoch losses to 5e-5, parameters to 2e-4
  (99.5 % quantile; an entry whose gradient is rounding noise may step the other way under
  Adam's normalisation, bounded by 2 * steps * lr)."""
  from bayesnf_b200 import inference, models
  g, args, meta = _load('map', name)
  spec = models.ModelSpec(**args, observation_model=meta['dist'])
  if prec == 'bf16x3' and not inference.precision_supported(spec, prec):
    pytest.skip(f'width {args["width"]} is not a tensor-core shape; fp32 covers it')
  n = g['y'].shape[0]
  seed = np.array([0, meta['seed']], dtype=np.uint32)
  params, losses = inference.fit_map(
      g['x'], g['y'], seed, meta['dist'], args, num_particles=meta['particles'],
      learning_rate=meta['lr'], num_epochs=meta['epochs'], prior_weight=meta['prior_weight'],
      batch_size=meta['batch_size'], num_splits=meta['num_splits'], precision=prec,
      init_params=g['init'][0].astype(np.float32), batch_order='jax')
  assert losses.shape == g['losses'].shape
  np.testing.assert_allclose(losses, g['losses'], rtol=5e-5)
  got = spec.flatten(params)[0].astype(np.float64)
  steps = meta['epochs'] * (n // (meta['batch_size'] or n))
  d = np.abs(got - g['final'][0])
  assert np.quantile(d, 0.995) <= 2e-4 and d.max() <= 2 * steps * meta['lr'], (np.quantile(d, 0.995), d.max())


@pytest.mark.gpu
@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('name', VI_CASES)
def test_cuda_fit_vi_against_reference_execution(cuda, name, prec):
  from bayesnf_b200 import inference, models
  g, args, meta = _load('vi', name)
  spec = models.ModelSpec(**args, observation_model=meta['dist'])
  if prec == 'bf16x3' and not inference.precision_supported(spec, prec):
    pytest.skip(f'width {args["width"]} is not a tensor-core shape; fp32 covers it')
  E, S, steps, bs = meta['ensemble'], meta['sample_size'], meta['epochs'], meta['batch_size']
  P = spec.num_params
  rows = None if bs is None else g['perm'][:, :bs].astype(np.int32)
  sur, losses, _ = inference.fit_vi(
      g['x'], g['y'], 0, meta['dist'], args, ensemble_size=E, learning_rate=meta['lr'], num_epochs=steps,
      sample_size_divergence=S, sample_size_posterior=2, kl_weight=meta['kl_weight'], batch_size=bs,
      precision=prec, init_params=(g['mu0'].astype(np.float32), g['rho0'].astype(np.float32)),
      eps=g['eps'], posterior_eps=np.zeros((2, E, P), np.float32), batch_indices=rows)
  np.testing.assert_allclose(losses, g['losses'], rtol=5e-5)
  mu1, rho1 = spec.flatten(sur.loc)[0], spec.flatten(sur.inv_softplus_scale)[0]
  for got, want in ((mu1, g['mu']), (rho1, g['rho'])):
    d = np.abs(got.astype(np.float64) - want)
    assert np.quantile(d, 0.995) <= 2e-4 and d.max() <= 2 * steps * meta['lr'], (np.quantile(d, 0.995), d.max())


@pytest.mark.gpu
@pytest.mark.parametrize('prec', PARITY_MODES)
def test_cuda_predict_normal_against_reference_execution(cuda, prec):
  from bayesnf_b200 import inference, models
  g, args, _ = _load('predict', 'normal')
  spec = models.ModelSpec(**args)
  params = spec.unflatten(g['params'])
  for approx, key in ((True, 'q_approx'), (False, 'q_root')):
    means, quants = inference.predict_bnf(g['x'], 'NORMAL', params, args, quantiles=tuple(g['quantiles']),
                                          approximate_quantiles=approx, precision=prec)
    scale = np.abs(g['means']).max()
    assert np.abs(np.asarray(means) - g['means']).max() <= 1e-5 * scale + 1e-6
    m64, s64 = _t(g['means'][0]), (0.01 + torch.exp(_t(g['params'][0][:, 0])))[:, None]
    for i, q in enumerate(g['quantiles']):
      if approx:
        assert np.abs(np.asarray(quants[i]) - g[key][i]).max() <= 2e-5 * scale + 1e-5
      else:   # inference.py:42-52 stops at |cdf - q| <= 1e-5
        r = O.mixture_cdf_residual(m64, s64, _t(np.asarray(quants[i])), float(q))
        assert float(r.abs().max()) <= 1.2e-5


@pytest.mark.gpu
@pytest.mark.parametrize('dist', ['nb', 'zinb'])
def test_cuda_predict_counts_against_reference_execution(cuda, dist):
  from bayesnf_b200 import inference, models
  g, args, _ = _load('predict', dist)
  spec = models.ModelSpec(**args, observation_model=dist.upper())
  params = spec.unflatten(g['params'])
  means, quants = inference.predict_bnf(g['x'], dist.upper(), params, args, quantiles=tuple(g['quantiles']),
                                        precision='fp32')
  assert _relmax(np.asarray(means), g['means']) <= 2e-5
  same = np.mean(np.asarray(quants) == g['q_root'])
  assert same >= 0.95, same      # integer quantiles; a CDF within float32 rounding of q may flip one


@pytest.mark.gpu
@pytest.mark.parametrize('prec', PARITY_MODES)
@pytest.mark.parametrize('kind', ['map', 'mle', 'vi'])
def test_cuda_estimator_against_reference_execution(cuda, kind, prec):
  """The estimator mirror end to end -- `.fit(table, seed, ...)` then `.predict(table, quantiles)`
  with the reference's call signature -- against what the reference's estimator computed from the
  same tables, initial draws, batch orders and noise."""
  g, train, test, args, fit, kwargs = _estimator_golden(kind)
  est = _my_estimator(kind, kwargs, precision=prec)
  E = fit['ensemble_size']
  seed = np.array([0, fit.pop('seed')], dtype=np.uint32)
  if kind == 'vi':
    est._fit_hooks = dict(init_params=(g['mu0'].astype(np.float32), g['rho0'].astype(np.float32)),
                          eps=g['eps'], posterior_eps=g['posterior_eps'])
  else:
    est._fit_hooks = dict(init_params=g['init'][0].astype(np.float32), batch_order='jax')
  est.fit(train, seed, **fit)
  np.testing.assert_allclose(est.losses_, g['losses'], rtol=5e-5)
  from bayesnf_b200 import models
  spec = models.ModelSpec(**args)
  got = spec.flatten(est.params_).astype(np.float64)
  d = np.abs(got - g['params'])
  steps = fit['num_epochs'] * (len(train) // (fit['batch_size'] or len(train)))
  assert d.max() <= 2 * steps * fit['learning_rate'] + 1e-3, d.max()
  # The reference's chickenpox configuration (period 4, 2 harmonics) contains sin(pi * t) at integer
  # weeks: zero in exact arithmetic, ~1e-6 of rounding in any float type.  Without a prior (MLE)
  # Adam normalises that column's gradient noise into full-size steps, in the reference as here, so
  # the Dense_0 kernel row of such a dead feature is compared by the step bound above only.
  om = O.OracleModel(**args, dtype=T64)
  feats = om.encode(om.unflatten(_t(g['params'].reshape(-1, spec.num_params)[0])), _t(g['x_train']))
  dead = np.where(feats.abs().max(0).values.numpy() < 1e-4)[0]
  assert len(dead) <= 1
  k0 = spec.leaf_offsets[spec.leaf_names.index('Dense_0/kernel')]
  live = np.ones(spec.num_params, dtype=bool)
  for r in dead:
    live[k0 + r * args['width']:k0 + (r + 1) * args['width']] = False
  assert np.quantile(d[..., live], 0.995) <= 2e-4, np.quantile(d[..., live], 0.995)
  scale = np.abs(g['means']).max()
  lm = est.likelihood_model(test)
  assert tuple(lm.batch_shape) == g['lm_log_prob'].shape and tuple(lm.event_shape) == (len(test),)
  np.testing.assert_allclose(lm.log_prob(g['y_test']), g['lm_log_prob'], rtol=2e-3)
  np.testing.assert_allclose(lm.distribution.scale, np.broadcast_to(g['lm_scale'], np.shape(lm.distribution.scale)), rtol=2e-3)
  assert np.abs(lm.distribution.loc - g['lm_loc']).max() <= 2e-3 * scale
  for approx, key in ((True, 'q_approx'), (False, 'q_root')):
    means, quants = est.predict(test, quantiles=tuple(g['quantiles']), approximate_quantiles=approx)
    assert np.asarray(means).shape == g['means'].shape
    assert np.abs(np.asarray(means) - g['means']).max() <= 2e-3 * scale     # parameters after the fit differ by ~1e-4
    spread = np.abs(g['q_approx'][2] - g['q_approx'][0]).max()
    for i in range(len(g['quantiles'])):
      assert np.abs(np.asarray(quants[i]) - g[key][i]).max() <= 2e-4 * spread + 2e-3 * scale
