"""Generate tests/golden/numerics_*.npz by EXECUTING the reference's own models.py / inference.py.

Runs only in the build container (needs /root/reference).  jax / flax / optax / TFP are not
installable here, so the reference's source files are imported over oracle/jaxshim.py -- a
functional float64 stand-in for the few third-party calls they make (see its docstring for what
is the reference's code and what is restated).  No reference source is copied; the outputs are
committed so that nothing at test / bench time reads /root/reference.

  numerics_model_<case>.npz    model forward, log-likelihoods (NORMAL / NB / ZINB), prior
                               log-prob and their gradients at random parameters
  numerics_map_<case>.npz      inference.fit_map end to end: initial draws, final parameters and
                               per-epoch losses (full batch, ragged minibatches in the JAX
                               permutation order, num_splits, prior_weight 0 / 1)
  numerics_vi_<case>.npz       inference.fit_vi: initial surrogate, the noise and sub-batches of
                               every step, final surrogate parameters and losses
  numerics_predict_<case>.npz  inference.predict_bnf: means and quantiles
  numerics_estimator_<kind>.npz  spatiotemporal.BayesianNeuralField{MAP,MLE,VI}: `.fit(table, seed)`
                               and `.predict(table, quantiles)` on the reference's chickenpox
                               fixture with its own dataset / model configuration (width 64)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle import jaxshim                                   # noqa: E402
models, inference, spatiotemporal = jaxshim.import_reference()
import jax                                                   # noqa: E402  (the shim)
from scipy import stats                                      # noqa: E402

OUT = os.environ.get('BNF_GOLDEN_OUT', os.path.join(ROOT, 'tests', 'golden'))
A = jaxshim.Arr


def model_cases():
  return {
      # the reference's chickenpox model arguments (tests/golden/bookkeeping.json) at width 64, the
      # narrowest tensor-core shape of the CUDA path
      'chickenpox': dict(width=64, depth=2, input_scales=np.array([399.0, 4.5, 3.2]),
                         num_seasonal_harmonics=np.array([2.0, 10.0]),
                         seasonality_periods=np.array([4.0, 52.1775]), init_x=(5, 3),
                         fourier_degrees=np.array([5.0, 5.0, 5.0]),
                         interactions=np.array([[0, 1], [0, 2], [1, 2]])),
      # a dimension without Fourier features, no seasonality, one hidden layer
      'odd': dict(width=8, depth=1, input_scales=np.array([2.0, 0.5, 1.5, 1.0]),
                  num_seasonal_harmonics=np.zeros((0,)), seasonality_periods=np.zeros((0,)),
                  init_x=(3, 4), fourier_degrees=np.array([0.0, 3.0, 2.0, 0.0]),
                  interactions=np.array([[0, 3], [1, 2]])),
      # eleven hidden layers: Dense_10 sorts before Dense_2; no interactions; 1-D input
      'deep': dict(width=4, depth=11, input_scales=np.array([3.0]),
                   num_seasonal_harmonics=np.array([3.0]), seasonality_periods=np.array([7.0]),
                   init_x=(2, 1), fourier_degrees=np.array([2.0]),
                   interactions=np.zeros((0, 2), dtype=int)),
  }


def args_json(args):
  return json.dumps({k: (np.asarray(v).tolist() if not isinstance(v, (int, tuple)) else v)
                     for k, v in args.items()})


def template_leaves(args):
  """Leaf paths + shapes of the reference's own parameter template."""
  _, template = inference.make_model(**args)
  names, shapes = [], []

  def walk(tree, prefix):
    for k in sorted(tree):
      if isinstance(tree[k], dict):
        walk(tree[k], prefix + [k])
      else:
        names.append('/'.join(prefix + [k]))
        shapes.append(tuple(tree[k].shape))
  walk(template['params'], [])
  leaves = jax.tree_util.tree_leaves(template)
  assert [tuple(l.shape) for l in leaves] == shapes
  return names, shapes


def flat(params):
  """Reference params tuple (one member) -> float64 vector in tuple order."""
  return np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1) for p in params])


def flat_members(params, lead):
  """Tuple of arrays with `lead` leading dims -> [*lead, P]."""
  return np.concatenate([np.asarray(p, dtype=np.float64).reshape(lead + (-1,)) for p in params], -1)


def f32_exact(a):
  """Store inputs that are float32-representable by construction as float32."""
  a = np.asarray(a, dtype=np.float64)
  assert np.array_equal(a.astype(np.float32).astype(np.float64), a)
  return a.astype(np.float32)


def random_params(shapes, rng, scale=0.5):
  out = [rng.normal(0.0, scale), rng.normal(-1.0, 0.3), rng.normal(0.0, 0.5)]
  out += [rng.normal(0.0, scale, size=s) for s in shapes]
  return tuple(A(np.asarray(p, dtype=np.float32)) for p in out)     # float32-representable values


def make_data(args, n, rng, counts):
  d = args['init_x'][-1]
  x = np.empty((n, d), dtype=np.float32)
  x[:, 0] = rng.integers(0, 400, size=n)
  for j in range(1, d):
    x[:, j] = rng.normal(0.0, 2.0 * args['input_scales'][j], size=n)
  if counts:
    y = rng.negative_binomial(2.0, 0.3, size=n).astype(np.float32)
    y[rng.random(n) < 0.25] = 0.0
  else:
    y = (np.sin(x[:, 0] / 9.0) + 0.1 * x[:, -1] + rng.normal(0, 0.3, size=n)).astype(np.float32)
  return x, y


def self_check_distributions(rng):
  """The shim's TFP formulas against scipy (independent implementations)."""
  k = rng.integers(0, 30, size=64).astype(float)
  n, lg = rng.uniform(0.2, 8.0), rng.normal(0.0, 1.5, size=64)
  nb = jaxshim.NegativeBinomial(n, lg)
  ref = stats.nbinom.logpmf(k, n, 1.0 / (1.0 + np.exp(lg)))     # scipy's p = P(failure) = 1 - p_tfp
  assert np.allclose(np.asarray(nb.log_prob(k)), ref, rtol=1e-10, atol=1e-10)
  assert np.allclose(np.asarray(nb.cdf(k)), stats.nbinom.cdf(k, n, 1.0 / (1.0 + np.exp(lg))), atol=1e-10)
  assert np.allclose(np.asarray(nb.mean()), stats.nbinom.mean(n, 1.0 / (1.0 + np.exp(lg))), rtol=1e-10)
  assert np.allclose(np.asarray(nb.stddev()), stats.nbinom.std(n, 1.0 / (1.0 + np.exp(lg))), rtol=1e-10)
  x = rng.normal(size=16)
  assert np.allclose(np.asarray(jaxshim.Normal(0.3, 1.7).log_prob(x)), stats.norm.logpdf(x, 0.3, 1.7))
  assert np.allclose(np.asarray(jaxshim.Logistic(-1.5, 1.0).log_prob(x)), stats.logistic.logpdf(x, -1.5, 1.0))


def gen_model(name, args, rng):
  names, shapes = template_leaves(args)
  mlp, template = inference.make_model(**args)
  treedef = jax.tree_util.tree_structure(template)
  prior = inference.make_prior(**dict(args))
  n = 37
  out = {'model_args': args_json(args), 'leaf_names': np.array(names),
         'leaf_shapes': np.array(json.dumps([list(s) for s in shapes]))}
  for dist in ('NORMAL', 'NB', 'ZINB'):
    x, y = make_data(args, n, rng, counts=dist != 'NORMAL')
    params = random_params(shapes, rng)
    pred = mlp.apply(jax.tree_util.tree_unflatten(treedef, params[3:]), A(x))

    def loglik(p):
      return models.make_likelihood_model(p, A(x), mlp, template, dist).log_prob(A(y))
    ll, gll = jax.value_and_grad(loglik)(params)
    lp, glp = jax.value_and_grad(lambda p: prior.log_prob(p))(params)
    out.update({f'{dist}_x': x, f'{dist}_y': y, f'{dist}_params': f32_exact(flat(params)),
                f'{dist}_pred': np.asarray(pred), f'{dist}_loglik': float(ll),
                f'{dist}_loglik_grad': flat(gll), f'{dist}_prior': float(lp),
                f'{dist}_prior_grad': flat(glp)})
  np.savez(os.path.join(OUT, f'numerics_model_{name}.npz'), **out)
  print('model', name, 'P =', out['NORMAL_params'].size, 'loglik', out['NORMAL_loglik'],
        out['NB_loglik'], out['ZINB_loglik'])


class InitRecorder:
  """Wraps inference.ensemble_map to note the initial draws of every call (they are a
  deterministic function of `seed` and `init_fn`; recomputing them leaves the call untouched)."""

  def __init__(self):
    self.inits, self.orig = [], inference.ensemble_map

  def __enter__(self):
    def wrapped(*a, **kw):
      init_seed, _ = jax.random.split(kw['seed'], 2)
      keys = jax.random.split(init_seed, (jax.device_count(), kw['ensemble_size']))
      self.inits.append(jax.vmap(jax.vmap(kw['init_fn']))(keys))
      return self.orig(*a, **kw)
    inference.ensemble_map = wrapped
    return self

  def __exit__(self, *exc):
    inference.ensemble_map = self.orig


def gen_map(name, args, rng, *, dist, n, particles, epochs, lr, prior_weight, batch_size, num_splits, seed):
  x, y = make_data(args, n, rng, counts=dist != 'NORMAL')
  with InitRecorder() as rec:
    params, losses = inference.fit_map(
        x, y, seed=jax.random.PRNGKey(seed), observation_model=dist, model_args=dict(args),
        num_particles=particles, learning_rate=lr, num_epochs=epochs, prior_weight=prior_weight,
        batch_size=batch_size, num_splits=num_splits)
  per = particles // num_splits
  init = np.concatenate([flat_members(i, (1, per)) for i in rec.inits], axis=1)
  final = flat_members(params, (1, particles))
  np.savez(os.path.join(OUT, f'numerics_map_{name}.npz'),
           model_args=args_json(args), x=x, y=y, init=init, final=final, losses=np.asarray(losses),
           meta=json.dumps(dict(dist=dist, particles=particles, epochs=epochs, lr=lr,
                                prior_weight=prior_weight, batch_size=batch_size,
                                num_splits=num_splits, seed=seed)))
  print('map', name, 'losses', np.asarray(losses)[0, :, [0, -1]].tolist())


def gen_vi(name, args, rng, *, dist, n, ensemble, epochs, lr, sample_size, kl_weight, batch_size, seed):
  x, y = make_data(args, n, rng, counts=dist != 'NORMAL')
  jaxshim.TRACE = {'eps': [], 'perm': [], 'vi_init': []}
  try:
    surrogate, losses, predictions = inference.fit_vi(
        x, y, seed=jax.random.PRNGKey(seed), observation_model=dist, model_args=dict(args),
        ensemble_size=ensemble, learning_rate=lr, num_epochs=epochs,
        sample_size_divergence=sample_size, sample_size_posterior=3, kl_weight=kl_weight,
        batch_size=batch_size)
    trace = jaxshim.TRACE
  finally:
    jaxshim.TRACE = None
  init = trace['vi_init'][0]                      # tuple (mu_0, rho_0, mu_1, rho_1, ...) of [E, *shape]
  mu0 = flat_members(init[0::2], (ensemble,))
  rho0 = flat_members(init[1::2], (ensemble,))
  final = trace['vi_final'][0]
  mu1 = flat_members(final[0::2], (ensemble,))
  rho1 = flat_members(final[1::2], (ensemble,))
  # eps: per step a list over components of [S, E, *shape] -> [steps, S, E, P]
  eps = np.stack([np.concatenate([np.asarray(e).reshape(sample_size, ensemble, -1) for e in step], -1)
                  for step in trace['eps']])
  perm = np.stack(trace['perm']) if trace['perm'] else np.zeros((0, n), dtype=np.int64)
  np.savez(os.path.join(OUT, f'numerics_vi_{name}.npz'),
           model_args=args_json(args), x=x, y=y, mu0=f32_exact(mu0), rho0=rho0, mu=mu1, rho=rho1, eps=f32_exact(eps),
           perm=perm, losses=np.asarray(losses), predictions_shape=np.array(
               [np.asarray(p).shape for p in predictions][0]),
           meta=json.dumps(dict(dist=dist, ensemble=ensemble, epochs=epochs, lr=lr,
                                sample_size=sample_size, kl_weight=kl_weight,
                                batch_size=batch_size, seed=seed)))
  print('vi', name, 'losses', np.asarray(losses)[0, :, [0, -1]].tolist(), 'perms', perm.shape)


def gen_predict(name, args, rng, *, dist, n, members):
  _, shapes = template_leaves(args)
  x, _ = make_data(args, n, rng, counts=False)
  per_member = [random_params(shapes, rng, scale=0.4) for _ in range(members)]
  params = tuple(A(np.stack([np.asarray(m[i]) for m in per_member])[None]) for i in range(len(per_member[0])))
  qs = (0.025, 0.5, 0.975)
  out = {'model_args': args_json(args), 'x': x, 'params': f32_exact(flat_members(params, (1, members))),
         'quantiles': np.array(qs), 'meta': json.dumps(dict(dist=dist))}
  if dist == 'NORMAL':
    for approx in (True, False):
      means, quants = inference.predict_bnf(x, dist, params, dict(args), quantiles=qs,
                                            approximate_quantiles=approx)
      out['means'] = np.asarray(means)
      out['q_approx' if approx else 'q_root'] = np.stack([np.asarray(q) for q in quants])
  else:
    means, quants = inference.predict_bnf(x, dist, params, dict(args), quantiles=qs)
    out['means'] = np.asarray(means)
    out['q_root'] = np.asarray(quants)
  np.savez(os.path.join(OUT, f'numerics_predict_{name}.npz'), **out)
  print('predict', name, {k: np.asarray(v).shape for k, v in out.items() if k.startswith(('means', 'q_'))})


def gen_estimator(kind):
  """The reference's estimator classes end to end on its own fixture (tests/test_data, already
  copied to tests/golden by scripts/make_golden.py) with the chickenpox entries of its
  scripts/dataset_config.py; width 64 and a few epochs keep the file small."""
  import pandas as pd
  sys.path.insert(0, '/root/reference/scripts')
  import dataset_config as ref_cfg                # pylint: disable=import-outside-toplevel
  dc = ref_cfg.DATASET_CONFIG['chickenpox']
  mc = dict(ref_cfg.MODEL_CONFIG['chickenpox']['map'])
  mc['width'] = 64
  # The reference's own fixture (chickenpox.8.*.csv) holds ONE location, so its standardised
  # latitude / longitude are (x - mu) / ~1e-14: fine for the reference's smoke tests, useless for a
  # float32-vs-float64 comparison.  A small synthetic table with the same columns instead:
  # 4 locations x 48 weeks, the last 8 weeks of every location held out.
  rng = np.random.default_rng(8)
  weeks = pd.date_range('2005-01-03', periods=48, freq='W-MON')
  locs = [(46.07, 18.23), (47.50, 19.04), (46.25, 20.15), (47.68, 17.63)]
  rows = []
  for lat, lon in locs:
    base = 40 + 25 * np.sin(2 * np.pi * (np.arange(48) + 3 * lat) / 52.0) + 4 * (lon - 18)
    for w, b in zip(weeks, base):
      rows.append(dict(datetime=w, latitude=lat, longitude=lon,
                       chickenpox=float(np.round(max(0.0, b + rng.normal(0, 6))))))
  table = pd.DataFrame(rows)
  train = table[table.datetime < weeks[40]].reset_index(drop=True)
  test = table[table.datetime >= weeks[40]].reset_index(drop=True)
  out_tables = {'train_csv': train.to_csv(index=False), 'test_csv': test.to_csv(index=False)}
  cls = {'map': spatiotemporal.BayesianNeuralFieldMAP, 'mle': spatiotemporal.BayesianNeuralFieldMLE,
         'vi': spatiotemporal.BayesianNeuralFieldVI}[kind]
  est = cls(feature_cols=dc['feature_cols'], target_col=dc['target_col'], timetype=dc['timetype'],
            freq=dc['freq'], standardize=dc['standardize'], **mc)
  qs = (0.025, 0.5, 0.975)
  out = {'quantiles': np.array(qs), 'width': 64, **{k: np.array(v) for k, v in out_tables.items()}}
  if kind == 'vi':
    fit = dict(ensemble_size=2, learning_rate=0.01, num_epochs=4, sample_size_posterior=3,
               sample_size_divergence=2, kl_weight=0.1, batch_size=None)
    jaxshim.TRACE = {}
    try:
      est.fit(train, seed=jax.random.PRNGKey(21), **fit)
      trace = jaxshim.TRACE
    finally:
      jaxshim.TRACE = None
    E, S = fit['ensemble_size'], fit['sample_size_divergence']
    init, final = trace['vi_init'][0], trace['vi_final'][0]
    out.update(mu0=f32_exact(flat_members(init[0::2], (E,))), rho0=flat_members(init[1::2], (E,)),
               mu=flat_members(final[0::2], (E,)), rho=flat_members(final[1::2], (E,)),
               eps=f32_exact(np.stack([np.concatenate([np.asarray(e).reshape(S, E, -1) for e in step], -1)
                                       for step in trace['eps']])),
               posterior_eps=f32_exact(np.concatenate(
                   [np.asarray(e).reshape(fit['sample_size_posterior'], E, -1)
                    for e in trace['normal_eps'][-len(init) // 2:]], -1)),
               params=flat_members(est.params_, (1, fit['sample_size_posterior'], E)))
  else:
    fit = dict(ensemble_size=2, learning_rate=0.005, num_epochs=3, batch_size=32, num_splits=1)
    with InitRecorder() as rec:
      est.fit(train, seed=jax.random.PRNGKey(21), **fit)
    out.update(init=flat_members(rec.inits[0], (1, 2)), params=flat_members(est.params_, (1, 2)))
  out['losses'] = np.asarray(est.losses_)
  out['fit'] = json.dumps(dict(fit, seed=21))
  # the reference's data-handler outputs and model arguments for these tables
  out['x_train'] = est.data_handler.get_train(train)
  out['y_train'] = est.data_handler.get_target(train)
  out['x_test'] = est.data_handler.get_test(test)
  out['model_args'] = args_json(est._model_args((fit['batch_size'] or len(train), 3)))
  out['estimator_kwargs'] = json.dumps(dict(
      feature_cols=dc['feature_cols'], target_col=dc['target_col'], timetype=dc['timetype'],
      freq=dc['freq'], standardize=dc['standardize'],
      **{k: (np.asarray(v).tolist() if not isinstance(v, (int, str)) else v) for k, v in mc.items()}))
  # likelihood_model(table): the predictive distribution object (spatiotemporal.py:433-468)
  y_test = test[dc['target_col']].to_numpy(dtype=np.float64)
  lm = est.likelihood_model(test)
  out['y_test'] = y_test
  out['lm_log_prob'] = np.asarray(lm.log_prob(A(y_test)))
  out['lm_loc'] = np.asarray(lm.distribution.loc)
  out['lm_scale'] = np.asarray(lm.distribution.scale)
  for approx in (True, False):
    means, quants = est.predict(test, quantiles=qs, approximate_quantiles=approx)
    out['means'] = np.asarray(means)
    out['q_approx' if approx else 'q_root'] = np.stack([np.asarray(q) for q in quants])
  np.savez(os.path.join(OUT, f'numerics_estimator_{kind}.npz'), **out)
  print('estimator', kind, 'losses', out['losses'].shape, 'means', out['means'].shape,
        'median quantile[:3]', out['q_root'][1][:3])


def main():
  rng = np.random.default_rng(20240517)
  self_check_distributions(rng)
  cases = model_cases()
  for name, args in cases.items():
    gen_model(name, args, rng)
  gen_map('normal_full', cases['chickenpox'], rng, dist='NORMAL', n=60, particles=2, epochs=6,
          lr=0.01, prior_weight=1.0, batch_size=None, num_splits=1, seed=7)
  gen_map('normal_minibatch_splits', cases['chickenpox'], rng, dist='NORMAL', n=50, particles=4,
          epochs=3, lr=0.005, prior_weight=1.0, batch_size=16, num_splits=2, seed=11)
  gen_map('zinb_mle_minibatch', cases['odd'], rng, dist='ZINB', n=45, particles=2, epochs=3,
          lr=0.01, prior_weight=0.0, batch_size=20, num_splits=1, seed=3)
  gen_map('nb_full', cases['deep'], rng, dist='NB', n=40, particles=2, epochs=4, lr=0.01,
          prior_weight=1.0, batch_size=None, num_splits=1, seed=5)
  gen_vi('normal_full', cases['odd'], rng, dist='NORMAL', n=30, ensemble=2, epochs=4, lr=0.01,
         sample_size=3, kl_weight=0.1, batch_size=None, seed=13)
  gen_vi('normal_subbatch', cases['chickenpox'], rng, dist='NORMAL', n=40, ensemble=2, epochs=3,
         lr=0.01, sample_size=2, kl_weight=0.5, batch_size=12, seed=17)
  gen_predict('normal', cases['chickenpox'], rng, dist='NORMAL', n=25, members=3)
  gen_predict('nb', cases['odd'], rng, dist='NB', n=20, members=3)
  gen_predict('zinb', cases['odd'], rng, dist='ZINB', n=20, members=3)
  for kind in ('map', 'mle', 'vi'):
    gen_estimator(kind)


if __name__ == '__main__':
  main()
