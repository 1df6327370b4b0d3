#!/usr/bin/env python
"""Experiment runner for the spatiotemporal benchmarks.

Same command line, experiment tables and output files as the reference's runner
(/root/reference/scripts/evaluate.py:49-150, scripts/dataset_config.py), driving the
bayesnf_b200 estimators:

  python scripts/evaluate.py --data_root DIR --output_dir OUT --dataset chickenpox \
      --objective map [--start_id 5] [--stop_id 10] [--num_particles 8] [--precision bf16]

For every series it reads `<dataset>.<id>.train.csv` / `.test.csv` (index column 0, a
`datetime` column), fits, predicts train+test with quantiles (0.5, 0.025, 0.975) and writes

  bnf-<objective>.<dataset>.<id>.log.json   run metadata (evaluate.py:119-131)
  bnf-<objective>.<dataset>.<id>.loss.csv   one column per member, one row per epoch (:133-136)
  bnf-<objective>.<dataset>.<id>.pred.csv   yhat, yhat_p50, yhat_lower, yhat_upper indexed like
                                            the input tables, sorted by index (:138-150)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_LATLON = ['datetime', 'latitude', 'longitude']


def _dataset(target, freq, features=None):
  return dict(num_series=10, target_col=target, timetype='index', freq=freq,
              feature_cols=list(features or _LATLON), standardize=['latitude', 'longitude'])


# scripts/dataset_config.py:19-92
DATASET_CONFIG = {
    'air_quality': _dataset('pm10', 'h'),
    'wind': _dataset('wind', 'D'),
    'air': _dataset('pm10', 'D'),
    'chickenpox': _dataset('chickenpox', 'W'),
    'coprecip': _dataset('ppt', 'M'),
    'sst': _dataset('sst', 'M', _LATLON + ['soi']),
}


def _model(width, periods, harmonics):
  cfg = dict(width=width, depth=2, seasonality_periods=np.asarray(periods, dtype=float),
             num_seasonal_harmonics=np.asarray(harmonics), observation_model='NORMAL')
  return {'map': cfg, 'mle': cfg, 'vi': cfg}


# scripts/dataset_config.py:95-181
MODEL_CONFIG = {
    'air_quality': _model(512, [24, 24 * 7], [4, 4]),
    'wind': _model(512, [7, 365.25 / 12, 365.25], [3, 10, 10]),
    'air': _model(512, [7, 365.25 / 12, 365.25], [3, 10, 10]),
    'chickenpox': _model(256, [4.0, 52.1775], [2, 10]),
    'coprecip': _model(512, [12], [6]),
    'sst': _model(768, [12], [6]),
}


def _inference(particles, map_epochs, vi_epochs, vi_batch, kl, map_batch=None, vi_lr=0.01):
  point = dict(num_particles=particles, num_epochs=map_epochs, learning_rate=0.005)
  if map_batch:
    point['batch_size'] = map_batch
  vi = dict(num_particles=particles, num_epochs=vi_epochs, learning_rate=vi_lr, batch_size=vi_batch,
            kl_weight=kl, sample_size_divergence=5)
  return {'map': point, 'mle': dict(point), 'vi': vi}


# scripts/evaluate.py:199-307
INFERENCE_CONFIG = {
    'air_quality': _inference(16, 4000, 500, 3500, 0.2, map_batch=38096),
    'wind': _inference(64, 10000, 2000, 3944, 0.1),
    'air': _inference(8, 7500, 1000, 3800, 0.2),
    'chickenpox': _inference(64, 10000, 1000, 511, 0.1),
    'coprecip': _inference(16, 7500, 750, 3300, 0.2),
    'sst': _inference(16, 5000, 600, 8845, 0.5, map_batch=221127, vi_lr=0.005),
}


def run_experiment(dataset, data_root, series_id, output_dir, objective, dataset_config, model_config,
                   inference_config, seed, precision=None):
  """One fit + predict of one series; writes the three output files; returns
  (losses, means, quantiles) like the reference (evaluate.py:49-152)."""
  import bayesnf_b200

  def read(split):
    return pd.read_csv(os.path.join(data_root, f'{dataset}.{series_id}.{split}.csv'), index_col=0,
                       parse_dates=['datetime'])

  df_train, df_test = read('train'), read('test')
  os.makedirs(output_dir, exist_ok=True)
  stem = os.path.join(output_dir, f'bnf-{objective}.{dataset}.{series_id}')
  model_config = dict(model_config)
  model_config.update(feature_cols=dataset_config['feature_cols'], target_col=dataset_config['target_col'],
                      timetype=dataset_config['timetype'], freq=dataset_config.get('freq'),
                      standardize=dataset_config.get('standardize'))
  fit_args = dict(learning_rate=inference_config['learning_rate'], num_epochs=inference_config['num_epochs'],
                  batch_size=inference_config.get('batch_size'), ensemble_size=inference_config['num_particles'])
  if objective == 'vi':
    cls = bayesnf_b200.BayesianNeuralFieldVI
    fit_args.update(kl_weight=inference_config.get('kl_weight', 1.0),
                    sample_size_divergence=inference_config.get('sample_size_divergence', 10))
  elif objective in ('map', 'mle'):
    cls = bayesnf_b200.BayesianNeuralFieldMAP if objective == 'map' else bayesnf_b200.BayesianNeuralFieldMLE
    fit_args.update(num_splits=inference_config.get('num_particle_splits', 1))
  else:
    raise ValueError(f'objective={objective}')
  extra = {} if precision is None else {'precision': precision}

  t0 = time.perf_counter()
  model = cls(**model_config, **extra).fit(df_train, seed, **fit_args)
  both = pd.concat([df_train, df_test])
  means, quantiles = model.predict(both, quantiles=(0.5, 0.025, 0.975))
  losses = model.losses_
  runtime = time.perf_counter() - t0

  with open(stem + '.log.json', 'w') as f:
    json.dump(dict(dataset=dataset, series_id=series_id, runtime=runtime, objective=objective,
                   dataset_config=dataset_config, model_config=model_config, inference_config=inference_config),
              f, indent=2, default=repr)
  pd.DataFrame(losses.reshape((-1, losses.shape[-1])).T).to_csv(stem + '.loss.csv', index=False)
  index = model.data_handler.copy_and_filter_table(both).index
  pred = pd.DataFrame(dict(yhat=means.reshape(-1, means.shape[-1]).mean(axis=0), yhat_p50=quantiles[0],
                           yhat_lower=quantiles[1], yhat_upper=quantiles[2]), index=index)
  pred.sort_index(inplace=True)
  pred.to_csv(stem + '.pred.csv', index=True)
  return losses, means, np.asarray(quantiles)


def main(argv=None):
  ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  ap.add_argument('--output_dir', required=True)
  ap.add_argument('--data_root', required=True)
  ap.add_argument('--dataset', required=True, choices=sorted(DATASET_CONFIG))
  ap.add_argument('--objective', default='map', choices=['map', 'mle', 'vi'])
  ap.add_argument('--start_id', type=int, default=5, help='series with IDs >= this value')
  ap.add_argument('--stop_id', type=int, default=None, help='series with IDs < this value')
  ap.add_argument('--num_particles', type=int, default=None, help='override the ensemble size')
  ap.add_argument('--num_epochs', type=int, default=None, help='override the number of epochs')
  ap.add_argument('--precision', default=None, choices=['bf16', 'fp32', 'bf16_simt'])
  ap.add_argument('--seed', type=int, default=0)
  args = ap.parse_args(argv)
  inf = dict(INFERENCE_CONFIG[args.dataset][args.objective])
  if args.num_particles:
    inf['num_particles'] = args.num_particles
  if args.num_epochs:
    inf['num_epochs'] = args.num_epochs
  stop = args.stop_id or DATASET_CONFIG[args.dataset]['num_series']
  for series_id in range(args.start_id, stop):
    print(f'{args.dataset} series_id {series_id}', flush=True)
    run_experiment(args.dataset, args.data_root, str(series_id), args.output_dir, args.objective,
                   DATASET_CONFIG[args.dataset], MODEL_CONFIG[args.dataset][args.objective], inf,
                   seed=np.array([0, args.seed], dtype=np.uint32), precision=args.precision)


if __name__ == '__main__':
  main()
