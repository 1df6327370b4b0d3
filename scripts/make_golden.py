"""Generate tests/golden/* from the REFERENCE's own pure numpy/pandas code.

Runs only in the build container (needs /root/reference); the outputs are
committed so that nothing at test/bench time reads /root/reference.

The reference imports jax / flax / tfp / optax at module import time and none of
them is installed here, so those third-party modules are replaced by inert
MagicMock stubs.  Only functions that never touch the stubs are then EXECUTED:
  spatiotemporal.seasonality_to_float / seasonalities_to_array   (:31-95)
  spatiotemporal.SpatiotemporalDataHandler                        (:114-192)
  BayesianNeuralFieldMAP._get_* / _model_args                     (:296-370)
  models.make_seasonal_frequencies                                (models.py:36-59)
No reference source is copied into the repo; CSV data fixtures are.
"""
import json
import os
import shutil
import sys
import types
from unittest import mock

import numpy as np
import pandas as pd

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

for name in ['jax', 'jax.numpy', 'jax.typing', 'flax', 'flax.linen', 'flax.core', 'flax.struct',
             'flax.core.frozen_dict', 'flax.core.scope', 'optax', 'jaxtyping',
             'tensorflow_probability', 'tensorflow_probability.substrates',
             'tensorflow_probability.substrates.jax']:
  sys.modules[name] = mock.MagicMock(name=name)
sys.path.insert(0, os.path.join(REF, 'src'))
sys.path.insert(0, os.path.join(REF, 'scripts'))
from bayesnf import models as ref_models            # noqa: E402
from bayesnf import spatiotemporal as ref_st        # noqa: E402
import dataset_config as ref_cfg                    # noqa: E402


def f32bits(a):
  return [int(v) for v in np.asarray(a, dtype=np.float32).view(np.uint32)]


def main():
  os.makedirs(OUT, exist_ok=True)
  g = {}
  pairs = [('Y', 'Y'), ('Q', 'Q'), ('Y', 'Q'), ('M', 'h'), ('Q', 'M'), ('Y', 'M'), ('M', 'D'),
           ('min', 's'), ('h', 's'), ('D', 's'), ('M', 's'), ('Q', 's'), ('Y', 's'), ('Y', 'W'),
           ('M', 'W'), ('Y', 'D'), ('W', 'D'), ('W', 'h'), ('D', 'h')]
  g['seasonality_to_float'] = [[s, f, ref_st.seasonality_to_float(s, f)] for s, f in pairs]
  g['seasonalities_to_array'] = {
      'args': [['D', 'W', 'M'], 'h'],
      'out': ref_st.seasonalities_to_array(['D', 'W', 'M'], 'h').tolist()}

  freq_cases = {
      'kat': ([4, 8], [2, 4]),
      'chickenpox': ([4.0, 52.1775], [2.0, 10]),
      'air_quality': ([24, 24 * 7], [4, 4]),
      'wind': ([7, 365.25 / 12, 365.25], [3, 10, 10]),
      'coprecip': ([12], [6]),
      'float_time': ([10, 12, .25], [.5, .5, .125]),
      'dups': ([6, 12, 24], [3, 6, 12]),
      'empty': ([], []),
  }
  g['make_seasonal_frequencies'] = {}
  for k, (p, h) in freq_cases.items():
    fr, hm = ref_models.make_seasonal_frequencies(np.asarray(p), np.asarray(h))
    g['make_seasonal_frequencies'][k] = {
        'periods': list(map(float, p)), 'harmonics_in': list(map(float, h)),
        'freq_bits': f32bits(fr), 'harm': [float(v) for v in hm]}

  # data handler on the reference's own fixture, with its own dataset config
  for f in ['chickenpox.8.train.csv', 'chickenpox.8.test.csv', 'bnf-map.chickenpox.8.mini.pred.csv',
            'bnf-mle.chickenpox.8.mini.pred.csv', 'bnf-vi.chickenpox.8.mini.pred.csv']:
    shutil.copy(os.path.join(REF, 'tests', 'test_data', f), os.path.join(OUT, f))
  dc = ref_cfg.DATASET_CONFIG['chickenpox']
  mc = dict(ref_cfg.MODEL_CONFIG['chickenpox']['map'])
  train = pd.read_csv(os.path.join(OUT, 'chickenpox.8.train.csv'), index_col=0, parse_dates=['datetime'])
  test = pd.read_csv(os.path.join(OUT, 'chickenpox.8.test.csv'), index_col=0, parse_dates=['datetime'])
  est = ref_st.BayesianNeuralFieldMAP(
      feature_cols=dc['feature_cols'], target_col=dc['target_col'], timetype=dc['timetype'],
      freq=dc['freq'], standardize=dc['standardize'], **mc)
  xtr = est.data_handler.get_train(train)
  ytr = est.data_handler.get_target(train)
  both = pd.concat([train, test])
  xte = est.data_handler.get_test(both)
  margs = est._model_args(xtr.shape)
  g['chickenpox'] = {
      'dataset_config': {k: v for k, v in dc.items() if k != 'series_id_fmt'},
      'model_config': {k: (np.asarray(v).tolist() if not isinstance(v, (int, str)) else v)
                       for k, v in mc.items()},
      'time_min': int(est.data_handler.time_min_),
      'time_scale': float(est.data_handler.time_scale_),
      'mu': est.data_handler.mu_.tolist(), 'std': est.data_handler.std_.tolist(),
      'input_scales': est.data_handler.get_input_scales().tolist(),
      'train_shape': list(xtr.shape), 'test_shape': list(xte.shape),
      'nanstd_y': float(np.nanstd(ytr)),
      'model_args': {
          'depth': margs['depth'], 'width': margs['width'],
          'input_scales': margs['input_scales'].tolist(),
          'num_seasonal_harmonics': np.asarray(margs['num_seasonal_harmonics']).tolist(),
          'seasonality_periods': np.asarray(margs['seasonality_periods']).tolist(),
          'init_x': list(margs['init_x']),
          'fourier_degrees': margs['fourier_degrees'].tolist(),
          'interactions_shape': list(margs['interactions'].shape)},
  }
  np.save(os.path.join(OUT, 'chickenpox_train_features.npy'), xtr.astype(np.float64))
  np.save(os.path.join(OUT, 'chickenpox_trainplustest_features.npy'), xte.astype(np.float64))
  np.save(os.path.join(OUT, 'chickenpox_train_target.npy'), ytr.astype(np.float64))

  # estimator bookkeeping cases from tests/test_spatiotemporal.py:49-74
  cases = []
  for p, h in [([], []), ([10, 15], [8, 6])]:
    m = ref_st.BayesianNeuralFieldMAP(freq='D', seasonality_periods=p, num_seasonal_harmonics=h,
                                      feature_cols=['t'], target_col='x', timetype='index')
    cases.append({'timetype': 'index', 'p': p, 'h': h,
                  'periods': np.asarray(m._get_seasonality_periods()).tolist(),
                  'harmonics': np.asarray(m._get_num_seasonal_harmonics()).tolist()})
  for p in [[], [10, 12, .25]]:
    m = ref_st.BayesianNeuralFieldMAP(seasonality_periods=p, feature_cols=['t'], target_col='x',
                                      timetype='float')
    cases.append({'timetype': 'float', 'p': p, 'h': None,
                  'periods': np.asarray(m._get_seasonality_periods()).tolist(),
                  'harmonics': np.asarray(m._get_num_seasonal_harmonics()).tolist()})
  g['estimator_bookkeeping'] = cases
  g['_provenance'] = ('generated by scripts/make_golden.py executing the pure numpy/pandas '
                      'functions of /root/reference (google/bayesnf v0.1.3) with pandas '
                      f'{pd.__version__}, numpy {np.__version__}')
  with open(os.path.join(OUT, 'bookkeeping.json'), 'w') as f:
    json.dump(g, f, indent=1)
  print('wrote', os.path.join(OUT, 'bookkeeping.json'))


if __name__ == '__main__':
  main()
