"""scripts/evaluate.py: the reference's tests/test_evaluate_mini.py configuration
(chickenpox series 8, 4 particles, 5 epochs) through the runner, compared with the reference's
golden prediction files for format (index, columns, order) and for the one quantity a different
PRNG leaves comparable: the predictive half width on the training rows."""
import importlib.util
import json
import os

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _runner():
  spec = importlib.util.spec_from_file_location('bnf_evaluate', os.path.join(ROOT, 'scripts', 'evaluate.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def test_experiment_tables_match_reference_configs():
  """Spot values of scripts/dataset_config.py / evaluate.py (CPU only)."""
  ev = _runner()
  assert ev.DATASET_CONFIG['chickenpox']['freq'] == 'W' and ev.DATASET_CONFIG['sst']['feature_cols'][-1] == 'soi'
  assert ev.MODEL_CONFIG['chickenpox']['map']['width'] == 256
  np.testing.assert_allclose(ev.MODEL_CONFIG['wind']['vi']['seasonality_periods'], [7, 365.25 / 12, 365.25])
  assert ev.INFERENCE_CONFIG['air_quality']['map']['batch_size'] == 38096
  assert ev.INFERENCE_CONFIG['sst']['vi'] == dict(num_particles=16, num_epochs=600, learning_rate=0.005,
                                                  batch_size=8845, kl_weight=0.5, sample_size_divergence=5)
  assert 'batch_size' not in ev.INFERENCE_CONFIG['chickenpox']['mle']


@pytest.mark.gpu
@pytest.mark.parametrize('objective', ['map', 'mle', 'vi'])
def test_mini_experiment_outputs(tmp_path, objective):
  ev = _runner()
  inf = dict(num_particles=4, num_epochs=5, learning_rate=0.005)       # test_evaluate_mini.py:61-78
  if objective == 'vi':                                                 # test_evaluate_mini.py:82-88
    inf = dict(batch_size=None, kl_weight=0.1, learning_rate=0.01, num_epochs=2, num_particles=1,
               sample_size_divergence=5)
  losses, means, quantiles = ev.run_experiment(
      'chickenpox', GOLDEN, '8', str(tmp_path), objective, ev.DATASET_CONFIG['chickenpox'],
      ev.MODEL_CONFIG['chickenpox'][objective], inf, seed=np.array([0, 0], dtype=np.uint32), precision='fp32')
  stem = os.path.join(str(tmp_path), f'bnf-{objective}.chickenpox.8')
  gold = pd.read_csv(os.path.join(GOLDEN, f'bnf-{objective}.chickenpox.8.mini.pred.csv'), index_col=0)
  pred = pd.read_csv(stem + '.pred.csv', index_col=0)
  assert list(pred.columns) == list(gold.columns) == ['yhat', 'yhat_p50', 'yhat_lower', 'yhat_upper']
  assert pred.index.equals(gold.index) and np.isfinite(pred.to_numpy()).all()
  assert (pred['yhat_lower'] <= pred['yhat_p50']).all() and (pred['yhat_p50'] <= pred['yhat_upper']).all()
  loss = pd.read_csv(stem + '.loss.csv')
  assert loss.shape == (inf['num_epochs'], inf['num_particles']) and losses.shape[-1] == inf['num_epochs']
  log = json.load(open(stem + '.log.json'))
  assert set(log) == {'dataset', 'series_id', 'runtime', 'objective', 'dataset_config', 'model_config',
                      'inference_config'}
  assert quantiles.shape == (3, len(gold)) and means.shape[-1] == len(gold)
  if objective != 'vi':      # sigma after 5 Adam steps does not depend on the kernel draws
    train_idx = pd.read_csv(os.path.join(GOLDEN, 'chickenpox.8.train.csv'), index_col=0).index
    half = ((pred['yhat_upper'] - pred['yhat_lower']) / 2).loc[train_idx]
    gold_half = ((gold['yhat_upper'] - gold['yhat_lower']) / 2).loc[train_idx]
    assert abs(half.median() - gold_half.median()) / gold_half.median() < 5e-3
