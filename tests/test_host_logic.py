"""Host-side logic of bayesnf_b200 that needs no GPU: plan bookkeeping done in
the C library (parameter layout, feature columns), ABI surface, seeds, errors."""
import json
import os
import re

import numpy as np
import pytest

from bayesnf_b200 import _lib, inference, models
from conftest import GOLDEN, ROOT
from oracle import bnf_oracle as O

G = json.load(open(os.path.join(GOLDEN, 'bookkeeping.json')))

CONFIGS = {
    'chickenpox': dict(width=256, depth=2, input_scales=[99., 1, 1], num_seasonal_harmonics=[2, 10],
                       seasonality_periods=[4.0, 52.1775], init_x=(100, 3), fourier_degrees=[5, 5, 5],
                       interactions=np.zeros((0, 2), int)),
    'air_quality': dict(width=512, depth=4, input_scales=[1000., 1, 1], num_seasonal_harmonics=[4, 4],
                        seasonality_periods=[24, 168], init_x=(64, 3), fourier_degrees=[5, 5, 5],
                        interactions=np.zeros((0, 2), int)),
    'wind': dict(width=1024, depth=6, input_scales=[500., 1, 1], num_seasonal_harmonics=[3, 10, 10],
                 seasonality_periods=[7, 365.25 / 12, 365.25], init_x=(64, 3), fourier_degrees=[5, 5, 5],
                 interactions=np.zeros((0, 2), int)),
    'interactions': dict(width=32, depth=1, input_scales=[10., 1, 2, 3], num_seasonal_harmonics=[],
                         seasonality_periods=[], init_x=(8, 4), fourier_degrees=[0, 3, 0, 2],
                         interactions=np.array([[0, 1], [1, 3], [2, 3]])),
    'deep': dict(width=8, depth=12, input_scales=[1.], num_seasonal_harmonics=[1],
                 seasonality_periods=[7.0], init_x=(8, 1), fourier_degrees=[1],
                 interactions=np.zeros((0, 2), int)),
    'many_groups': dict(width=8, depth=1, input_scales=[1.] * 11, num_seasonal_harmonics=[2],
                        seasonality_periods=[12.0], init_x=(8, 11), fourier_degrees=[1] * 11,
                        interactions=np.array([[0, 1]])),
}


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_plan_layout_matches_oracle(name):
  """C-side tree_leaves order / offsets / F == the oracle's independent derivation."""
  cfg = CONFIGS[name]
  spec = models.ModelSpec(**cfg)
  om = O.OracleModel(**cfg)
  assert spec.num_params == om.num_params
  assert spec.num_features == om.F
  assert spec.leaf_names == om.leaf_names
  assert [tuple(s) for s in spec.leaf_shapes] == [tuple(s) for s in om.leaf_shapes]
  assert spec.padded_features % 64 == 0 and spec.padded_features >= spec.num_features
  assert spec.num_feature_groups == len(om.groups)
  # offsets are the running sum after the three scalars
  off = 3
  for o, s in zip(spec.leaf_offsets, spec.leaf_shapes):
    assert o == off
    off += int(np.prod(s)) if s else 1
  assert off == spec.num_params


def test_known_sizes():
  assert models.ModelSpec(**CONFIGS['chickenpox']).num_params == 80912
  assert models.ModelSpec(**CONFIGS['chickenpox']).num_features == 57
  assert models.ModelSpec(**CONFIGS['air_quality']).num_features == 49
  assert models.ModelSpec(**CONFIGS['air_quality']).num_params == 814098
  assert models.ModelSpec(**CONFIGS['wind']).num_features == 79
  assert models.ModelSpec(**CONFIGS['wind']).num_params == 5330964


def test_flatten_unflatten_roundtrip():
  spec = models.ModelSpec(**CONFIGS['interactions'])
  flat = np.random.default_rng(0).normal(size=(2, 3, spec.num_params)).astype(np.float32)
  tup = spec.unflatten(flat)
  assert tup[0].shape == (2, 3) and len(tup) == 3 + len(spec.leaf_names)
  k0 = tup[3 + spec.leaf_names.index('Dense_0/kernel')]
  assert k0.shape == (2, 3, spec.num_features, 32)
  np.testing.assert_array_equal(spec.flatten(tup), flat)
  with pytest.raises(ValueError):
    spec.flatten(tup[:-1])


@pytest.mark.parametrize('case', sorted(G['make_seasonal_frequencies']))
def test_product_seasonal_frequencies_bit_exact(case):
  c = G['make_seasonal_frequencies'][case]
  fr, hm = models.make_seasonal_frequencies(np.asarray(c['periods']), np.asarray(c['harmonics_in']))
  assert [int(v) for v in np.asarray(fr, np.float32).view(np.uint32)] == c['freq_bits']
  assert [float(v) for v in hm] == c['harm']


def test_config_errors_are_value_errors():
  bad = dict(CONFIGS['chickenpox'])
  bad['num_seasonal_harmonics'] = [3, 10]      # 3 > 4/2
  with pytest.raises(ValueError):
    models.ModelSpec(**bad)
  bad = dict(CONFIGS['chickenpox'])
  bad['fourier_degrees'] = [5, 5]
  with pytest.raises(ValueError):
    models.ModelSpec(**bad)
  bad = dict(CONFIGS['interactions'])
  bad['interactions'] = np.array([[0, 9]])
  with pytest.raises(ValueError):
    models.ModelSpec(**bad)
  with pytest.raises(ValueError):
    models.ModelSpec(**CONFIGS['chickenpox'], observation_model='POISSON')
  bad = dict(CONFIGS['deep'])
  bad['depth'] = 40
  with pytest.raises(ValueError):
    models.ModelSpec(**bad)


def test_header_symbols_exported_and_bound():
  """Every function declared in include/bnf.h is exported by the .so and bound."""
  hdr = open(os.path.join(ROOT, 'include', 'bnf.h')).read()
  hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
  declared = set(re.findall(r'\b(bnf_[a-z_0-9]+)\s*\(', hdr))
  assert declared, 'no declarations parsed'
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  for name in declared:
    assert hasattr(_lib.lib, name)
  assert _lib.lib.bnf_abi_version() == 1


def test_workspace_query_without_gpu():
  spec = models.ModelSpec(**CONFIGS['chickenpox'])
  fwd = _lib.lib.bnf_workspace_bytes(spec.plan, _lib.PREC_FP32, 8, 1024, _lib.WS_FORWARD)
  grad = _lib.lib.bnf_workspace_bytes(spec.plan, _lib.PREC_FP32, 8, 1024, _lib.WS_GRAD)
  mp = _lib.lib.bnf_workspace_bytes(spec.plan, _lib.PREC_FP32, 8, 1024, _lib.WS_MAP)
  assert 0 < fwd < grad < mp
  assert _lib.lib.bnf_workspace_bytes(spec.plan, _lib.PREC_FP32, 0, 1024, _lib.WS_MAP) == 0


def test_seeds():
  assert inference.seed_to_int(5) == 5
  assert inference.seed_to_int(np.array([0, 7], dtype=np.uint32)) == 7
  assert inference.seed_to_int(np.array([1, 2], dtype=np.uint32)) == (1 << 32) | 2
  a, b = inference.fold_in(3, 0), inference.fold_in(3, 1)
  assert a != b and 0 <= a < 2 ** 64 and 0 <= b < 2 ** 64
  with pytest.raises(ValueError):
    inference.seed_to_int(np.arange(3))


def test_no_cpu_fallback():
  """Without a GPU the compute entry points must raise, not fall back."""
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  feats = np.zeros((10, 3))
  with pytest.raises(_lib.BnfError):
    inference.fit_map(feats, np.arange(10.0), 0, 'NORMAL', CONFIGS['chickenpox'], 2, 0.01, 1)


def test_product_does_not_import_oracle():
  for root, _, files in os.walk(os.path.join(ROOT, 'bayesnf_b200')):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(root, f)).read()
        assert 'bnf_oracle' not in src and 'from oracle' not in src and 'import oracle' not in src, f


def test_flax_variables_round_trip():
  """params tuple <-> Flax variables dict: the dict's sorted-key leaf order
  IS the tuple order, including the string sort that puts '...scale10' before '...scale2'."""
  from bayesnf_b200 import models
  spec = models.ModelSpec(width=64, depth=2, input_scales=[99., 1, 1], num_seasonal_harmonics=[2, 3],
                          seasonality_periods=[7., 30.], init_x=(100, 3), fourier_degrees=[3, 2, 2],
                          interactions=[[1, 2]])
  rng = np.random.default_rng(0)
  flat = rng.normal(size=(1, 4, spec.num_params)).astype(np.float32)
  params = spec.unflatten(flat)
  heads, variables = spec.to_flax_variables(params)
  tree = variables['params']
  assert set(tree['Dense_0']) == {'bias', 'kernel'} and tree['Dense_0']['kernel'].shape == (1, 4, 28, 64)
  assert tree['Dense_2']['kernel'].shape == (1, 4, 64, 1) and tree['log_scale_adjustment'].shape == (1, 4, 3)
  assert tree['feature_inv_sp_scale5'].shape == (1, 4) and len(heads) == 3
  back = spec.from_flax_variables(heads, variables)
  assert len(back) == len(params)
  for a, b in zip(back, params):
    np.testing.assert_array_equal(a, b)
  np.testing.assert_array_equal(spec.flatten(back), flat)
  with pytest.raises(ValueError):
    spec.from_flax_variables(heads, {'params': {k: v for k, v in tree.items() if k != 'Dense_1'}})
  # 12 hidden layers: 'Dense_10' sorts before 'Dense_2' (string order of tree_leaves)
  deep = models.ModelSpec(width=64, depth=12, input_scales=[9.], num_seasonal_harmonics=[], seasonality_periods=[],
                          init_x=(10, 1), fourier_degrees=[2], interactions=np.zeros((0, 2), int))
  names = deep.leaf_names
  assert names.index('Dense_10/bias') < names.index('Dense_2/bias')
  assert names.index('inv_sp_layer_scale10') < names.index('inv_sp_layer_scale2')
  p2 = deep.unflatten(rng.normal(size=(2, deep.num_params)).astype(np.float32))
  h2, v2 = deep.to_flax_variables(p2)
  for a, b in zip(deep.from_flax_variables(h2, v2), p2):
    np.testing.assert_array_equal(a, b)



def test_philox_known_answers():
  """The device RNG's block function against the Random123 known-answer vectors
  (philox4x32-10), evaluated on the host through bnf_debug_philox."""
  import ctypes as C

  def ph(c, k):
    ca, ka, o = (C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), (C.c_uint32 * 4)()
    _lib.lib.bnf_debug_philox(ca, ka, o)
    return list(o)
  assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
  assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
  assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
      [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_device_permutation_is_a_uniform_bijection():
  """permute_dataset as the device evaluates it (keyed Feistel bijection + cycle walking,
  bnf_map_epochs): a permutation for every n incl. edge cases, different per member and epoch,
  first element uniform over keys (chi-square), no excess of consecutive neighbours."""
  for n in (1, 2, 3, 5, 100, 1023, 1024, 1025, 10440, 65536, 100003):
    p = inference.device_permutation(5, 0, 0, n)
    assert np.array_equal(np.sort(p), np.arange(n)), n
  a, b, c = (inference.device_permutation(5, 0, 0, 10440), inference.device_permutation(5, 1, 0, 10440),
             inference.device_permutation(5, 0, 1, 10440))
  assert (a == b).mean() < 0.01 and (a == c).mean() < 0.01 and (a == np.arange(10440)).mean() < 0.01
  np.testing.assert_array_equal(a, inference.device_permutation(5, 0, 0, 10440))     # deterministic
  n, keys = 64, 6400
  first, adjacent = np.zeros(n), 0
  for e in range(keys):
    p = inference.device_permutation(9, 3, e, n)
    first[p[0]] += 1
    adjacent += int((np.abs(np.diff(p)) == 1).sum())
  chi2 = float(((first - keys / n) ** 2 / (keys / n)).sum())
  assert chi2 < 110.0, chi2                       # df = 63: P(chi2 > 110) ~ 2e-4
  frac = adjacent / (keys * (n - 1))
  assert abs(frac - 2.0 / n) < 0.15 * 2.0 / n, frac


def test_default_precision_is_the_parity_tensor_core_mode():
  assert inference.get_default_precision() in ('bf16x3',) or 'BAYESNF_B200_PRECISION' in os.environ
  spec = models.ModelSpec(**CONFIGS['chickenpox'])
  assert _lib.lib.bnf_precision_supported(spec.plan, _lib.PREC_BF16X3) == 0
  odd = dict(CONFIGS['chickenpox'])
  odd['width'] = 40
  spec = models.ModelSpec(**odd)
  assert _lib.lib.bnf_precision_supported(spec.plan, _lib.PREC_BF16X3) != 0     # SIMT f32 covers it
  assert _lib.lib.bnf_precision_supported(spec.plan, _lib.PREC_FP32) == 0
