"""Estimator API: drop-in surface of ``bayesnf.spatiotemporal`` for the hot path.

Same classes, constructor keywords, ``fit`` / ``predict`` signatures, attributes
(``params_``, ``losses_``, ``data_handler``) and error behaviour as the
reference (src/bayesnf/spatiotemporal.py:195-648); the numerical work is done by
``bayesnf_b200.inference`` on the GPU.  The pandas bookkeeping below (period
arithmetic, time indexing, standardisation) is host-side and integer/float64
exact with the reference -- pinned by tests/test_spatiotemporal.py, which
replays the reference's own test cases, and tests/golden/bookkeeping.json.
"""

from __future__ import annotations

from collections.abc import Sequence

import numpy as np
import pandas as pd

from . import inference
from . import parallel

_EPOCH = pd.Timestamp('2020-01-01')   # origin of the integer time index (spatiotemporal.py:101)


def _ordinal(when, freq: str) -> int:
  """Number of the ``freq`` period that contains ``when`` on pandas' period axis."""
  return pd.Period(when, freq=freq).ordinal


def seasonality_to_float(seasonality: str, freq: str) -> float:
  """Average number of ``freq`` periods per ``seasonality`` period.

  Both are counted between the ``seasonality`` period that holds 2020-01-01 and the one that holds
  2024-01-01 (four years, so one leap day is averaged in): ('Y', 'D') -> 365.25,
  ('M', 'D') -> 30.4375.  Values of spatiotemporal.py:31-59, pinned by the golden table in
  tests/golden/bookkeeping.json.
  """
  first = pd.Period(_EPOCH, freq=seasonality)
  last = pd.Period(_EPOCH + pd.DateOffset(years=4), freq=seasonality)
  n_seasons = last.ordinal - first.ordinal
  n_freq = _ordinal(last.start_time, freq) - _ordinal(first.start_time, freq)
  return n_freq / n_seasons


def seasonalities_to_array(seasonalities: Sequence[float | str], freq: str) -> np.ndarray:
  """Periods given as floats or pandas offset aliases -> float durations in units of ``freq``.

  A period shorter than one ``freq`` is a TypeError with the reference's messages
  (spatiotemporal.py:62-95).
  """
  def as_float(s):
    if not isinstance(s, str):
      if s < 1:
        raise TypeError(f'seasonality_float={s!r} should be larger than 1.')
      return s
    value = seasonality_to_float(s, freq)
    if value < 1:
      raise TypeError(
          f'seasonality={s!r} should represent a time span greater than '
          f'freq={freq!r}, but {s} is {value:.2f} of a {freq}')
    return value
  return np.array([as_float(s) for s in seasonalities])


class SpatiotemporalDataHandler:
  """DataFrame -> feature matrix, with the reference's conventions (spatiotemporal.py:98-192):

  * column 0 of ``feature_cols`` is time: ``timetype='index'`` turns datetimes into the number of
    ``freq`` periods since 2020-01-01, ``'float'`` takes the values as they are; either way the
    axis is shifted so that the first TRAINING time is 0 (test tables reuse that origin),
  * rows whose target is NaN are dropped from training tables,
  * the columns named in ``standardize`` are centred and scaled with the training mean / std,
  * ``get_input_scales`` reports the training time span for column 0 and 1 elsewhere.

  Nothing is written into the caller's DataFrame.  Outputs are equal, element for element, to the
  reference's (tests/test_spatiotemporal.py, tests/test_reference_goldens.py).
  """

  _time_idx = 0

  def __init__(self, feature_cols, target_col, timetype, freq, standardize=None):
    self.feature_cols = feature_cols
    self.target_col = target_col
    self.timetype = timetype
    self.freq = freq
    self.standardize = standardize
    self.mu_ = None
    self.std_ = None
    self.time_min_ = None
    self.time_scale_ = None

  @property
  def _time_column(self) -> str:
    return self.feature_cols[self._time_idx]

  # ---- rows ----
  def _labelled_rows(self, table: pd.DataFrame) -> pd.DataFrame:
    if self.target_col not in table.columns:
      return table
    return table.loc[table[self.target_col].notna()]

  def copy_and_filter_table(self, table: pd.DataFrame) -> pd.DataFrame:
    return self._labelled_rows(table).copy()

  def get_target(self, table: pd.DataFrame) -> np.ndarray:
    return self._labelled_rows(table)[self.target_col].values

  # ---- columns ----
  def _raw_time(self, column: pd.Series) -> np.ndarray:
    """Unshifted time axis of a table."""
    if self.timetype == 'index':
      periods = pd.PeriodIndex(column.dt.to_period(self.freq))
      return periods.asi8 - _ordinal(_EPOCH, self.freq)
    if self.timetype == 'float':
      return column.to_numpy(dtype=float)
    raise ValueError(f'Unknown timetype: {self.timetype}')

  def _assemble(self, table: pd.DataFrame, time: np.ndarray) -> np.ndarray:
    """Feature matrix with the shifted time axis in column 0 (dtype as DataFrame.values gives)."""
    cols = {c: (time if i == self._time_idx else table[c].to_numpy())
            for i, c in enumerate(self.feature_cols)}
    return pd.DataFrame(cols, columns=list(self.feature_cols)).values

  def _standardized(self, features: np.ndarray) -> np.ndarray:
    return (features - self.mu_) / self.std_ if self.standardize else features

  def get_train(self, table: pd.DataFrame) -> np.ndarray:
    """Training features; fixes the time origin / span and the standardisation statistics."""
    rows = self._labelled_rows(table)
    time = self._raw_time(rows[self._time_column])
    self.time_min_ = time.min()
    features = self._assemble(rows, time - self.time_min_)
    self.time_scale_ = features[:, self._time_idx].max()
    k = len(self.feature_cols)
    self.mu_, self.std_ = np.zeros(k), np.ones(k)
    if self.standardize:
      if self._time_column in self.standardize:
        raise TypeError('Do not standardize the time column!')
      which = [self.feature_cols.index(c) for c in self.standardize]
      block = features[:, which].astype(float)
      self.mu_[which], self.std_[which] = block.mean(axis=0), block.std(axis=0)
    return self._standardized(features)

  def get_test(self, table: pd.DataFrame) -> np.ndarray:
    """Features of a new table on the training time origin / statistics (after ``get_train``)."""
    time = self._raw_time(table[self._time_column]) - self.time_min_
    return self._standardized(self._assemble(table, time))

  def get_input_scales(self) -> np.ndarray:
    scales = np.ones(len(self.feature_cols))
    scales[self._time_idx] = self.time_scale_
    return scales


class BayesianNeuralFieldEstimator:
  """Base class; use BayesianNeuralFieldMAP / MLE / VI (spatiotemporal.py:195-468)."""

  _ensemble_dims: int
  _prior_weight: float = 1.0
  _scale_epochs_by_batch_size: bool = False
  # Test hook: extra keyword arguments for inference.fit_map / fit_vi (init_params, batch_order,
  # eps, ...), so that a fit can be replayed from the reference's recorded draws
  # (tests/test_reference_goldens.py).  Empty in normal use.
  _fit_hooks: dict = {}

  def __init__(
      self,
      *,
      feature_cols: Sequence[str],
      target_col: str,
      seasonality_periods: Sequence[float | str] | None = None,
      num_seasonal_harmonics: Sequence[int] | None = None,
      fourier_degrees: Sequence[float] | None = None,
      interactions: Sequence[tuple[int, int]] | None = None,
      freq: str | None = None,
      timetype: str = 'index',
      depth: int = 2,
      width: int = 512,
      observation_model: str = 'NORMAL',
      standardize: Sequence[str] | None = None,
      precision: str | None = None,
  ):
    """Arguments as in the reference (spatiotemporal.py:217-232).  ``precision``
    ('fp32' | 'bf16') is the only addition: arithmetic mode of the dense stack
    on the GPU (default: ``inference.get_default_precision()``)."""
    self.num_seasonal_harmonics = num_seasonal_harmonics
    self.seasonality_periods = seasonality_periods
    self.observation_model = observation_model
    self.depth = depth
    self.width = width
    self.feature_cols = feature_cols
    self.target_col = target_col
    self.timetype = timetype
    self.freq = freq
    self.fourier_degrees = fourier_degrees
    self.standardize = standardize
    self.interactions = interactions
    self.precision = precision
    self.losses_ = None
    self.params_ = None
    self.data_handler = SpatiotemporalDataHandler(
        self.feature_cols, self.target_col, self.timetype, self.freq,
        standardize=self.standardize)

  # ---- constructor keywords -> model_args (interface of spatiotemporal.py:296-370: the method
  # names, defaults and error messages are the reference's, its tests call them directly) ----
  def _get_fourier_degrees(self, batch_shape) -> np.ndarray:
    """One Fourier degree per input column: 5 unless given."""
    n_dims = batch_shape[-1]
    given = self.fourier_degrees
    degrees = np.full(n_dims, 5, dtype=int) if given is None else np.atleast_1d(given).astype(int)
    if degrees.shape[-1] != n_dims:
      raise ValueError(
          'The length of fourier_degrees ({}) must match the input dimension '
          'dimension ({}).'.format(degrees.shape[-1], n_dims))
    return degrees

  def _get_interactions(self) -> np.ndarray:
    """(N, 2) integer column pairs; none by default."""
    pairs = np.array([] if self.interactions is None else self.interactions).astype(int)
    if self.interactions is None:
      pairs = pairs.reshape(0, 2)
    if pairs.ndim != 2 or pairs.shape[-1] != 2:
      raise ValueError(
          'The argument for `interactions` should be a 2-d array of integers of '
          'shape (N, 2), indicating the column indices to interact (the passed '
          f'shape was {pairs.shape})')
    return pairs

  def _check_time_axis(self):
    """An integer time index needs the data frequency; a float axis must not name one."""
    has_freq = self.freq is not None
    if {'index': not has_freq, 'float': has_freq}.get(self.timetype, False):
      raise ValueError(f'Invalid {self.freq=} with {self.timetype=}.')

  def _get_seasonality_periods(self):
    self._check_time_axis()
    given = self.seasonality_periods
    if given is None:
      return np.zeros(0)
    assert self.timetype in ('index', 'float'), f'Impossible {self.timetype=}.'
    if self.timetype == 'float':
      return np.asarray(given, dtype=float)
    return seasonalities_to_array(given, self.freq)

  def _get_num_seasonal_harmonics(self):
    given = self.num_seasonal_harmonics
    assert self.timetype in ('index', 'float'), f'Impossible {self.timetype=}.'
    if self.timetype == 'index':        # discrete time: harmonics as given
      return np.zeros(0) if given is None else np.array(given)
    if given is not None:               # continuous time: exactly one harmonic per period
      raise ValueError(f'Cannot use num_seasonal_harmonics with {self.timetype=}.')
    # make_seasonal_frequencies builds arange(1, 1 + h): any h in (0, min(.5, p / 2)] gives [1]
    return np.fmin(.5, self._get_seasonality_periods() / 2)

  def _model_args(self, batch_shape):
    """The keyword set of inference.make_model for a batch of this shape."""
    # (harmonics first: with both a float axis and harmonics given, that is the error reported)
    seasonal = dict(num_seasonal_harmonics=self._get_num_seasonal_harmonics(),
                    seasonality_periods=self._get_seasonality_periods())
    features = dict(input_scales=self.data_handler.get_input_scales(),
                    fourier_degrees=self._get_fourier_degrees(batch_shape),
                    interactions=self._get_interactions())
    return dict(depth=self.depth, width=self.width, init_x=batch_shape, **seasonal, **features)

  def predict(self, table, quantiles=(0.5,), approximate_quantiles=False):
    """Predict the target at new times / locations (spatiotemporal.py:372-408).

    Returns ``(means, quantiles)``: ``means`` has shape (num_devices,
    ensemble_size // num_devices, len(table)) (VI: an extra posterior-sample
    axis after the device axis); ``quantiles`` is a list with one (len(table),)
    array per requested quantile.
    """
    test_data = self.data_handler.get_test(table)
    return inference.predict_bnf(
        test_data,
        self.observation_model,
        params=self.params_,
        model_args=self._model_args(test_data.shape),
        quantiles=quantiles,
        ensemble_dims=self._ensemble_dims,
        approximate_quantiles=approximate_quantiles,
        precision=self.precision,
    )

  def fit(self, table, seed):
    raise NotImplementedError('Should be implemented by subclass')

  def likelihood_model(self, table: pd.DataFrame):
    """Predictive distribution over new field values in `table` (spatiotemporal.py:433-468).

    NOTE: Must be called after `fit`.  The reference returns
    `tfd.Independent(Normal | NegativeBinomial | ZeroInflatedNegativeBinomial, 1)` with batch
    shape `(num_devices, [num_samples,] members)` and event shape `(len(table),)`; this returns a
    `distributions.PredictiveDistribution` with the same shapes and the methods `mean`, `stddev`,
    `variance`, `log_prob`, `prob`, `sample` and `.distribution` (per-point `log_prob`, `cdf`,
    `quantile`).  The network outputs come from the CUDA forward pass of this rank's members.
    """
    from . import distributions
    test_data = self.data_handler.get_test(table)
    predictions = inference.forward_bnf(
        test_data, self.observation_model, self.params_, self._model_args(test_data.shape),
        precision=self.precision)
    return distributions.PredictiveDistribution(
        self.observation_model, predictions, self.params_[0], self.params_[1], self.params_[2])


class BayesianNeuralFieldMAP(BayesianNeuralFieldEstimator):
  """Stochastic ensembles of maximum-a-posteriori estimates (spatiotemporal.py:471-541)."""

  _ensemble_dims = 2

  def fit(self, table, seed, ensemble_size=16, learning_rate=0.005,
          num_epochs=5_000, batch_size=None, num_splits=1):
    if ensemble_size < parallel.device_count():
      raise ValueError('ensemble_size cannot be smaller than device_count. '
                       'https://github.com/google/bayesnf/issues/28.')
    train_data = self.data_handler.get_train(table)
    train_target = self.data_handler.get_target(table)
    if batch_size is None:
      batch_size = train_data.shape[0]
    if self._scale_epochs_by_batch_size:
      num_epochs = num_epochs * (train_data.shape[0] // batch_size)
    model_args = self._model_args((batch_size, train_data.shape[-1]))
    self.params_, self.losses_ = inference.fit_map(
        train_data,
        train_target,
        seed=seed,
        observation_model=self.observation_model,
        model_args=model_args,
        num_particles=ensemble_size,
        learning_rate=learning_rate,
        num_epochs=num_epochs,
        prior_weight=self._prior_weight,
        batch_size=batch_size,
        num_splits=num_splits,
        precision=self.precision,
        **self._fit_hooks)
    return self


class BayesianNeuralFieldMLE(BayesianNeuralFieldMAP):
  """Maximum-likelihood ensembles: MAP with the prior switched off (:544-551)."""

  _prior_weight = 0.0


class BayesianNeuralFieldVI(BayesianNeuralFieldEstimator):
  """Ensembles of mean-field surrogate posteriors (spatiotemporal.py:554-648)."""

  _ensemble_dims = 3
  _scale_epochs_by_batch_size = True

  def fit(self, table, seed, ensemble_size=16, learning_rate=0.01,
          num_epochs=1_000, sample_size_posterior=30, sample_size_divergence=5,
          kl_weight=0.1, batch_size=None):
    train_data = self.data_handler.get_train(table)
    train_target = self.data_handler.get_target(table)
    if batch_size is None:
      batch_size = train_data.shape[0]
    if self._scale_epochs_by_batch_size:
      num_epochs = num_epochs * (train_data.shape[0] // batch_size)
    model_args = self._model_args((batch_size, train_data.shape[-1]))
    _, self.losses_, self.params_ = inference.fit_vi(
        train_data,
        train_target,
        seed=seed,
        observation_model=self.observation_model,
        model_args=model_args,
        ensemble_size=ensemble_size,
        learning_rate=learning_rate,
        num_epochs=num_epochs,
        sample_size_posterior=sample_size_posterior,
        sample_size_divergence=sample_size_divergence,
        kl_weight=kl_weight,
        batch_size=batch_size,
        precision=self.precision,
        **self._fit_hooks)
    return self
