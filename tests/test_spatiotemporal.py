"""Host bookkeeping of the estimator API.

Replays the reference's own test cases (tests/test_spatiotemporal.py:21-120 of
google/bayesnf) against bayesnf_b200, plus the golden values produced by
scripts/make_golden.py from the reference's pure pandas code.
"""
import json
import os

import numpy as np
import pandas as pd
import pytest

from bayesnf_b200 import spatiotemporal
from conftest import GOLDEN

G = json.load(open(os.path.join(GOLDEN, 'bookkeeping.json')))


@pytest.mark.parametrize(
    'seasonality, freq, expected',
    [('Y', 'Y', 1), ('Q', 'Q', 1), ('Y', 'Q', 4), ('M', 'h', 730.5), ('Q', 'M', 3), ('Y', 'M', 12),
     ('M', 'D', 30.4375), ('min', 's', 60), ('h', 's', 3600), ('D', 's', 86400),
     ('M', 's', 2629800), ('Q', 's', 7889400), ('Y', 's', 31557600)])
def test_seasonality_to_float(seasonality, freq, expected):
  assert spatiotemporal.seasonality_to_float(seasonality, freq) == expected


@pytest.mark.parametrize('s, f, expected', G['seasonality_to_float'])
def test_seasonality_to_float_golden(s, f, expected):
  assert spatiotemporal.seasonality_to_float(s, f) == expected


def test_seasonalities_to_array():
  periods = spatiotemporal.seasonalities_to_array(['D', 'W', 'M'], 'h')
  np.testing.assert_allclose(periods, np.array([24, 168, 730.5]))
  assert periods.tolist() == G['seasonalities_to_array']['out']
  with pytest.raises(TypeError):
    spatiotemporal.seasonalities_to_array(['h'], 'D')
  with pytest.raises(TypeError):
    spatiotemporal.seasonalities_to_array([0.5], 'D')


@pytest.mark.parametrize('p, h', [([], []), ([10, 15], [8, 6])])
def test_get_seasonality_periods_index(p, h):
  model = spatiotemporal.BayesianNeuralFieldMAP(
      freq='D', seasonality_periods=p, num_seasonal_harmonics=h, feature_cols=['t'],
      target_col='x', timetype='index')
  assert np.all(model._get_seasonality_periods() == p)
  assert np.all(model._get_num_seasonal_harmonics() == h)


@pytest.mark.parametrize('p, h', [([], []), ([10, 12, .25], [.5, .5, .125])])
def test_get_seasonality_periods_float(p, h):
  model = spatiotemporal.BayesianNeuralFieldMAP(
      seasonality_periods=p, feature_cols=['t'], target_col='x', timetype='float')
  assert np.all(model._get_seasonality_periods() == p)
  assert np.all(model._get_num_seasonal_harmonics() == h)


def test_estimator_bookkeeping_golden():
  for case in G['estimator_bookkeeping']:
    kw = dict(seasonality_periods=case['p'], feature_cols=['t'], target_col='x',
              timetype=case['timetype'])
    if case['timetype'] == 'index':
      kw.update(freq='D', num_seasonal_harmonics=case['h'])
    m = spatiotemporal.BayesianNeuralFieldMAP(**kw)
    assert np.asarray(m._get_seasonality_periods()).tolist() == case['periods']
    assert np.asarray(m._get_num_seasonal_harmonics()).tolist() == case['harmonics']


def test_invalid_frequency():
  model = spatiotemporal.BayesianNeuralFieldMAP(feature_cols=['t'], target_col='x', timetype='index')
  with pytest.raises(ValueError):
    model._get_seasonality_periods()
  model = spatiotemporal.BayesianNeuralFieldMAP(freq='M', feature_cols=['t'], target_col='x',
                                                timetype='float')
  with pytest.raises(ValueError):
    model._get_seasonality_periods()


def test_invalid_seasonality_period():
  model = spatiotemporal.BayesianNeuralFieldMAP(
      seasonality_periods=['W'], feature_cols=['t'], target_col='x', timetype='float')
  with pytest.raises(ValueError):
    model._get_seasonality_periods()


def test_invalid_num_seasonal_harmonics():
  model = spatiotemporal.BayesianNeuralFieldMAP(
      seasonality_periods=[1, 5], num_seasonal_harmonics=[0.5, 1], feature_cols=['t'],
      target_col='x', timetype='float')
  with pytest.raises(ValueError):
    model._get_num_seasonal_harmonics()


def test_fourier_degrees_and_interactions_validation():
  m = spatiotemporal.BayesianNeuralFieldMAP(feature_cols=['t', 'a'], target_col='x', freq='D',
                                            fourier_degrees=[2, 3, 4])
  with pytest.raises(ValueError):
    m._get_fourier_degrees((10, 2))
  assert m._get_interactions().shape == (0, 2)
  m = spatiotemporal.BayesianNeuralFieldMAP(feature_cols=['t', 'a'], target_col='x', freq='D',
                                            interactions=[0, 1])
  with pytest.raises(ValueError):
    m._get_interactions()
  m = spatiotemporal.BayesianNeuralFieldMAP(feature_cols=['t', 'a'], target_col='x', freq='D')
  assert m._get_fourier_degrees((10, 2)).tolist() == [5, 5]


def _chickenpox_estimator(cls=spatiotemporal.BayesianNeuralFieldMAP):
  c = G['chickenpox']
  dc, mc = c['dataset_config'], c['model_config']
  return cls(feature_cols=dc['feature_cols'], target_col=dc['target_col'], timetype=dc['timetype'],
             freq=dc['freq'], standardize=dc['standardize'], width=mc['width'], depth=mc['depth'],
             seasonality_periods=mc['seasonality_periods'],
             num_seasonal_harmonics=mc['num_seasonal_harmonics'],
             observation_model=mc['observation_model'])


def test_data_handler_matches_reference_bit_exact():
  """get_train / get_test / get_target on the reference's chickenpox fixture."""
  c = G['chickenpox']
  train = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.train.csv'), index_col=0, parse_dates=['datetime'])
  test = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.test.csv'), index_col=0, parse_dates=['datetime'])
  est = _chickenpox_estimator()
  xtr = est.data_handler.get_train(train)
  ytr = est.data_handler.get_target(train)
  xte = est.data_handler.get_test(pd.concat([train, test]))
  h = est.data_handler
  assert int(h.time_min_) == c['time_min'] and float(h.time_scale_) == c['time_scale']
  assert h.mu_.tolist() == c['mu'] and h.std_.tolist() == c['std']
  assert h.get_input_scales().tolist() == c['input_scales']
  np.testing.assert_array_equal(xtr.astype(np.float64), np.load(os.path.join(GOLDEN, 'chickenpox_train_features.npy')))
  np.testing.assert_array_equal(xte.astype(np.float64), np.load(os.path.join(GOLDEN, 'chickenpox_trainplustest_features.npy')))
  np.testing.assert_array_equal(ytr.astype(np.float64), np.load(os.path.join(GOLDEN, 'chickenpox_train_target.npy')))
  assert float(np.nanstd(ytr)) == c['nanstd_y']
  ma = est._model_args(xtr.shape)
  gma = c['model_args']
  assert ma['depth'] == gma['depth'] and ma['width'] == gma['width']
  assert ma['input_scales'].tolist() == gma['input_scales']
  assert np.asarray(ma['num_seasonal_harmonics']).tolist() == gma['num_seasonal_harmonics']
  assert np.asarray(ma['seasonality_periods']).tolist() == gma['seasonality_periods']
  assert list(ma['init_x']) == gma['init_x']
  assert ma['fourier_degrees'].tolist() == gma['fourier_degrees']
  assert list(ma['interactions'].shape) == gma['interactions_shape']


def test_standardizing_time_column_is_an_error():
  train = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.train.csv'), index_col=0, parse_dates=['datetime'])
  est = spatiotemporal.BayesianNeuralFieldMAP(
      feature_cols=['datetime', 'latitude'], target_col='chickenpox', freq='W',
      standardize=['datetime'])
  with pytest.raises(TypeError):
    est.data_handler.get_train(train)


def test_nan_targets_are_dropped():
  train = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.train.csv'), index_col=0, parse_dates=['datetime'])
  train = train.copy()
  train.loc[train.index[:7], 'chickenpox'] = np.nan
  est = _chickenpox_estimator()
  assert est.data_handler.get_train(train).shape[0] == 93
  assert est.data_handler.get_target(train).shape[0] == 93


def test_class_attributes():
  assert spatiotemporal.BayesianNeuralFieldMAP._ensemble_dims == 2
  assert spatiotemporal.BayesianNeuralFieldMLE._ensemble_dims == 2
  assert spatiotemporal.BayesianNeuralFieldVI._ensemble_dims == 3
  assert spatiotemporal.BayesianNeuralFieldMAP._prior_weight == 1.0
  assert spatiotemporal.BayesianNeuralFieldMLE._prior_weight == 0.0
  assert spatiotemporal.BayesianNeuralFieldVI._scale_epochs_by_batch_size
  est = _chickenpox_estimator()
  assert est.params_ is None and est.losses_ is None and est.data_handler is not None
