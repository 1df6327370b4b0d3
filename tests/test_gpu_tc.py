"""tcgen05 GEMM kernels (bf16 x bf16 -> f32 in TMEM): the GEMM alone against a
plain f32 matmul of the same bf16 inputs, then the whole bf16 path against the
oracle and against the bf16-storage SIMT path."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import bnf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda():
  assert torch.cuda.is_available(), 'gpu tests need a CUDA device (no fallback)'
  torch.cuda.set_device(0)
  return torch.device('cuda', 0)


def _gemm(mn_major, a, b, nets, m, n, k):
  from bayesnf_b200 import _lib
  c = torch.full((nets, m, n), float('nan'), dtype=torch.float32, device=a.device)
  _lib.check(_lib.lib.bnf_debug_gemm(
      mn_major, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()),
      nets, m, n, k, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
  torch.cuda.synchronize()
  return c


SHAPES = [(1, 128, 64, 64), (1, 128, 128, 128), (1, 128, 256, 256), (2, 256, 256, 512),
          (3, 200, 128, 192), (1, 100, 64, 64), (2, 1000, 512, 1024), (1, 77, 320, 64),
          (1, 4096, 1024, 1024), (8, 10440, 256, 256)]


@pytest.mark.parametrize('nets,m,n,k', SHAPES)
def test_gemm_k_major(cuda, monkeypatch, nets, m, n, k):
  """C = A[M,K] . B[N,K]^T  (operand layout of the forward and dgrad GEMMs), single-CTA tiles."""
  monkeypatch.setenv('BNF_CTA2', '0')
  g = torch.Generator(device='cuda').manual_seed(m * 7 + n)
  a = torch.randn(nets, m, k, generator=g, device=cuda).to(torch.bfloat16)
  b = torch.randn(nets, n, k, generator=g, device=cuda).to(torch.bfloat16)
  c = _gemm(0, a, b, nets, m, n, k)
  want = torch.bmm(a.float(), b.float().transpose(1, 2))
  err = float((c - want).abs().max())
  assert err <= 2e-3 * float(want.abs().max()), err    # f32 accumulation-order noise only


@pytest.mark.parametrize('nets,m,n,k', [(1, 128, 64, 64), (1, 128, 128, 128), (1, 256, 256, 256),
                                        (2, 128, 256, 1000), (1, 64, 64, 200), (3, 64, 128, 77),
                                        (2, 1024, 1024, 4096), (8, 256, 256, 10440)])
def test_gemm_mn_major(cuda, monkeypatch, nets, m, n, k):
  """C = A[K,M]^T . B[K,N]  (wgrad: reduction over batch rows, ragged K zero-filled by TMA)."""
  monkeypatch.setenv('BNF_CTA2', '0')
  g = torch.Generator(device='cuda').manual_seed(m * 3 + k)
  a = torch.randn(nets, k, m, generator=g, device=cuda).to(torch.bfloat16)
  b = torch.randn(nets, k, n, generator=g, device=cuda).to(torch.bfloat16)
  c = _gemm(1, a, b, nets, m, n, k)
  want = torch.bmm(a.float().transpose(1, 2), b.float())
  err = float((c - want).abs().max())
  assert err <= 2e-3 * float(want.abs().max()), err


def _cfg(width, depth, n):
  return dict(width=width, depth=depth, input_scales=[n - 1.0, 1, 1], num_seasonal_harmonics=[2, 10],
              seasonality_periods=[4.0, 52.1775], init_x=(n, 3), fourier_degrees=[5, 5, 5],
              interactions=np.zeros((0, 2), int))


def _setup(cfg, n, nets, dist='NORMAL'):
  from bayesnf_b200 import inference, models
  from test_gpu_parity import _data, _random_params
  x, y = _data(cfg, n, counts=dist != 'NORMAL')
  om = O.OracleModel(**cfg)
  P = _random_params(om, nets, y, seed=3)
  spec = models.ModelSpec(**cfg, observation_model=dist)
  xd, yd = inference._to_device_data(x, y)
  return om, spec, P, xd, yd


@pytest.mark.parametrize('width,depth,n', [(256, 2, 100), (256, 2, 1000), (128, 3, 333), (512, 2, 300),
                                           (64, 2, 200), (1024, 3, 700)])
def test_bf16_tc_matches_bf16_simt(cuda, width, depth, n):
  """Same bf16 storage, tensor cores vs SIMT f32 FMA: differences are only the
  accumulation order and the bf16 rounding of the staged weights and of dh -> 1e-2 of scale on values,
  5e-2 of the leaf scale on gradients (small leaves are cancellation-prone sums)."""
  from bayesnf_b200 import inference
  cfg = _cfg(width, depth, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3)
  out = {}
  for prec in ('bf16', 'bf16_simt'):
    eng = inference.Engine(spec, prec)
    loc = eng.forward(P.cuda(), xd)
    ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
    out[prec] = (loc.cpu(), ll.cpu(), grad.cpu())
  a, b = out['bf16'], out['bf16_simt']
  assert float((a[0] - b[0]).abs().max()) <= 1e-2 * float(b[0].abs().max())
  assert float(((a[1] - b[1]) / b[1]).abs().max()) <= 1e-2
  for j in range(3):
    parts = [(0, 3)] + [(o, o + (int(np.prod(s)) if s else 1))
                        for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
    for lo, hi in parts:
      scale = float(b[2][j, lo:hi].abs().max()) + 1e-3 * float(b[2][j].abs().max())
      assert float((a[2][j, lo:hi] - b[2][j, lo:hi]).abs().max()) <= 5e-2 * scale, (lo, hi)


@pytest.mark.parametrize('width,depth,n,nets', [(256, 2, 10440, 8), (512, 3, 20000, 4), (128, 2, 30000, 5)])
def test_bf16_tc_many_tiles_per_cta(cuda, width, depth, n, nets):
  """BASELINE configs[1] at full size (8 members x 10 440 rows) and two larger shapes: every CTA
  works through several tiles and several networks, so the epilogues' on-chip partial sums
  (shared-memory column sums, per-lane register sums) are flushed on network changes, with both
  tile->CTA mappings (contiguous at one n-tile, round-robin at W=512).  Same tolerances as
  test_bf16_tc_matches_bf16_simt."""
  from bayesnf_b200 import inference
  cfg = _cfg(width, depth, n)
  om, spec, P, xd, yd = _setup(cfg, n, nets)
  out = {}
  for prec in ('bf16', 'bf16_simt'):
    eng = inference.Engine(spec, prec)
    ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
    out[prec] = (ll.cpu(), grad.cpu())
  a, b = out['bf16'], out['bf16_simt']
  assert float(((a[0] - b[0]) / b[0]).abs().max()) <= 1e-2
  parts = [(0, 3)] + [(o, o + (int(np.prod(s)) if s else 1)) for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
  for j in range(nets):
    for lo, hi in parts:
      scale = float(b[1][j, lo:hi].abs().max()) + 1e-3 * float(b[1][j].abs().max())
      assert float((a[1][j, lo:hi] - b[1][j, lo:hi]).abs().max()) <= 5e-2 * scale, (j, lo, hi)


@pytest.mark.parametrize('dist', ['NORMAL', 'ZINB'])
def test_bf16_tc_vs_oracle(cuda, dist):
  """bf16 tensor-core path vs the f32 oracle: 3e-2 of scale (bf16 has 8 significand bits)."""
  from bayesnf_b200 import inference
  n = 500
  cfg = _cfg(256, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 2, dist)
  eng = inference.Engine(spec, 'bf16')
  ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
  for j in range(2):
    loss, g = O.map_loss_and_grad(om, P[j], xd.cpu(), yd.cpu(), n, 0.0, dist)
    assert abs(float(ll[j]) + float(loss)) <= 3e-2 * abs(float(loss))
    assert float((grad[j].cpu() + g).abs().max() / g.abs().max()) < 3e-2


def test_bf16_training_tracks_fp32(cuda):
  """60 full-batch MAP steps at the benchmark shape: the bf16 loss curve follows fp32's."""
  from bayesnf_b200 import inference
  n = 2088
  cfg = _cfg(256, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 4)
  x, y = xd.cpu().numpy().astype(np.float64), yd.cpu().numpy().astype(np.float64)
  res = {}
  for prec in ('fp32', 'bf16'):
    _, losses = inference.fit_map(x, y, 0, 'NORMAL', cfg, 4, 0.005, 60, precision=prec,
                                  init_params=P.numpy())
    res[prec] = losses[0]
  assert np.isfinite(res['bf16']).all()
  rel = np.abs(res['bf16'] - res['fp32']) / np.abs(res['fp32'])
  assert rel.max() < 2e-2, rel.max()
  assert (res['bf16'][:, -1] < res['bf16'][:, 0]).all()


@pytest.mark.parametrize('flag', ['BNF_NO_FUSED_ENCODE', 'BNF_FUSED_ENCODE', 'BNF_NO_BIAS0_WGRAD', 'BNF_NO_FUSED_ACT_BWD', 'BNF_NO_FUSED_ENC_BWD',
                                  'BNF_ENCODE_GENERIC', 'BNF_NO_FUSED_HEAD_EPI'])
def test_alternative_kernel_paths_agree(cuda, flag, monkeypatch):
  """The fused encode+Dense_0 kernel switched off (forward-only default) / on (training), the
  Dense_0 bias gradient by column sums instead of the wgrad GEMM, the unfused dgrad/act_bwd pair
  and the unfused dgrad_0 / encode-backward pair compute the same thing as the default path (same
  bf16 storage; only reduction order differs)."""
  from bayesnf_b200 import inference
  n = 700
  cfg = _cfg(256, 3, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3)
  eng = inference.Engine(spec, 'bf16')
  monkeypatch.delenv(flag, raising=False)
  loc0 = eng.forward(P.cuda(), xd).cpu()
  ll0, g0 = eng.loglik_grad(P.cuda(), xd, yd)
  monkeypatch.setenv(flag, '1')
  loc1 = eng.forward(P.cuda(), xd).cpu()
  ll1, g1 = eng.loglik_grad(P.cuda(), xd, yd)
  assert float((loc0 - loc1).abs().max()) <= 1e-2 * float(loc0.abs().max())
  assert float(((ll0 - ll1) / ll0).abs().max()) <= 5e-3
  g0, g1 = g0.cpu(), g1.cpu()
  for j in range(3):
    for lo, hi in [(0, 3)] + [(o, o + (int(np.prod(s)) if s else 1))
                              for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]:
      scale = float(g0[j, lo:hi].abs().max()) + 1e-3 * float(g0[j].abs().max())
      assert float((g0[j, lo:hi] - g1[j, lo:hi]).abs().max()) <= 3e-2 * scale, (flag, lo, hi)


@pytest.mark.parametrize('prec', ['bf16', 'bf16x3'])
@pytest.mark.parametrize('dist', ['NORMAL', 'ZINB'])
def test_head_rows_variant_agrees(cuda, monkeypatch, dist, prec):
  """BNF_HEAD_ROWS=1 (head + activation backward with one warp per row, h read once) computes what
  head_fused_kernel computes, on a ragged row count (W = 512; opt-in because it is slower)."""
  from bayesnf_b200 import inference
  n = 1237
  cfg = _cfg(512, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3, dist)
  eng = inference.Engine(spec, prec)
  monkeypatch.delenv('BNF_HEAD_ROWS', raising=False)
  ll0, g0 = eng.loglik_grad(P.cuda(), xd, yd)
  monkeypatch.setenv('BNF_HEAD_ROWS', '1')
  ll1, g1 = eng.loglik_grad(P.cuda(), xd, yd)
  tol = 2e-5 if prec == 'bf16x3' else 5e-3
  assert float(((ll0 - ll1) / ll0).abs().max()) <= tol
  g0, g1 = g0.cpu(), g1.cpu()
  for j in range(3):
    for lo, hi in [(0, 3)] + [(o, o + (int(np.prod(s)) if s else 1))
                              for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]:
      scale = float(g0[j, lo:hi].abs().max()) + 1e-3 * float(g0[j].abs().max())
      assert float((g0[j, lo:hi] - g1[j, lo:hi]).abs().max()) <= (1e-4 if prec == 'bf16x3' else 3e-2) * scale, (lo, hi)


def test_fused_head_matches_two_kernel_path(cuda, monkeypatch):
  """head_fused_kernel == head_kernel + act_bwd (fp32 mode: tight tolerance)."""
  from bayesnf_b200 import inference
  for dist in ('NORMAL', 'ZINB'):
    n = 333
    cfg = _cfg(128, 2, n)
    om, spec, P, xd, yd = _setup(cfg, n, 3, dist)
    eng = inference.Engine(spec, 'fp32')
    monkeypatch.delenv('BNF_NO_FUSED_HEAD', raising=False)
    ll0, g0 = eng.loglik_grad(P.cuda(), xd, yd)
    monkeypatch.setenv('BNF_NO_FUSED_HEAD', '1')
    ll1, g1 = eng.loglik_grad(P.cuda(), xd, yd)
    assert float(((ll0 - ll1) / ll1).abs().max()) <= 1e-5
    assert float((g0 - g1).abs().max()) <= 2e-4 * float(g1.abs().max())


@pytest.mark.parametrize('mn_major,nets,m,n,k', [
    (0, 1, 256, 256, 64), (0, 1, 256, 256, 256), (0, 2, 1000, 512, 1024), (0, 1, 130, 256, 128),
    (0, 3, 4096, 1024, 1024), (1, 1, 256, 256, 128), (1, 2, 512, 256, 1000), (1, 2, 1024, 1024, 4096)])
def test_gemm_cta_pair(cuda, monkeypatch, mn_major, nets, m, n, k):
  """cta_group::2 variant (default for 256-wide tiles; BNF_CTA2=0 disables): a CTA pair computes
  256-row tiles, each CTA staging half of the B tile; both variants against a f32 matmul."""
  monkeypatch.setenv('BNF_CTA2', '1')
  g = torch.Generator(device='cuda').manual_seed(m + n + k)
  if mn_major:
    a = torch.randn(nets, k, m, generator=g, device=cuda).to(torch.bfloat16)
    b = torch.randn(nets, k, n, generator=g, device=cuda).to(torch.bfloat16)
    want = torch.bmm(a.float().transpose(1, 2), b.float())
  else:
    a = torch.randn(nets, m, k, generator=g, device=cuda).to(torch.bfloat16)
    b = torch.randn(nets, n, k, generator=g, device=cuda).to(torch.bfloat16)
    want = torch.bmm(a.float(), b.float().transpose(1, 2))
  c = _gemm(mn_major, a, b, nets, m, n, k)
  err = float((c - want).abs().max())
  assert err <= 2e-3 * float(want.abs().max()), err


def test_cta_pair_training_path(cuda, monkeypatch):
  """Whole bf16 step with CTA pairs == single-CTA path (width 256 so every GEMM qualifies)."""
  from bayesnf_b200 import inference
  n = 1000
  cfg = _cfg(256, 3, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3)
  eng = inference.Engine(spec, 'bf16')
  monkeypatch.setenv('BNF_CTA2', '0')
  ll0, g0 = eng.loglik_grad(P.cuda(), xd, yd)
  monkeypatch.setenv('BNF_CTA2', '1')
  ll1, g1 = eng.loglik_grad(P.cuda(), xd, yd)
  assert float(((ll0 - ll1) / ll0).abs().max()) <= 1e-3
  assert float((g0 - g1).abs().max()) <= 1e-2 * float(g0.abs().max())


def test_estimators_bf16_end_to_end(cuda):
  """Public API in tensor-core mode: MAP with minibatches + per-member permutations, MLE full
  batch (CUDA-graph replay), VI with sub-batches, predict with quantiles."""
  import pandas as pd
  import bayesnf_b200
  rng = np.random.default_rng(0)
  T, S = 120, 12
  dates = pd.date_range('2021-01-04', periods=T, freq='W-MON')
  rows = []
  for s in range(S):
    lat, lon = rng.normal(), rng.normal()
    for t, d in enumerate(dates):
      rows.append((d, lat, lon, 5 * np.sin(2 * np.pi * t / 52.0) + lat + rng.normal(scale=0.3)))
  df = pd.DataFrame(rows, columns=['datetime', 'latitude', 'longitude', 'y'])
  kw = dict(feature_cols=['datetime', 'latitude', 'longitude'], target_col='y', freq='W',
            seasonality_periods=['Y'], num_seasonal_harmonics=[4], width=128, depth=2,
            standardize=['latitude', 'longitude'], precision='bf16')
  est = bayesnf_b200.BayesianNeuralFieldMAP(**kw).fit(df, seed=1, ensemble_size=4, num_epochs=6,
                                                     batch_size=256, learning_rate=0.01)
  assert est.losses_.shape == (1, 4, 6) and np.isfinite(est.losses_).all()
  assert est.losses_[0, :, -1].mean() < est.losses_[0, :, 0].mean()
  means, q = est.predict(df.iloc[:200], quantiles=(0.1, 0.5, 0.9))
  assert means.shape == (1, 4, 200) and np.all(q[0] <= q[1]) and np.all(q[1] <= q[2])
  # likelihood_model (spatiotemporal.py:433-468): batch (devices, members), event (rows,)
  lm = est.likelihood_model(df.iloc[:200])
  assert lm.batch_shape == (1, 4) and lm.event_shape == (200,)
  np.testing.assert_allclose(lm.mean(), means, rtol=1e-6)
  lp = lm.log_prob(df['y'].to_numpy()[:200])
  assert lp.shape == (1, 4) and np.isfinite(lp).all()
  np.testing.assert_allclose(lm.stddev()[..., 0], 0.01 + np.exp(est.params_[0]), rtol=1e-6)
  est = bayesnf_b200.BayesianNeuralFieldMLE(**kw).fit(df, seed=2, ensemble_size=2, num_epochs=40)
  assert est.losses_.shape == (1, 2, 40) and (est.losses_[0, :, -1] < est.losses_[0, :, 0]).all()
  est = bayesnf_b200.BayesianNeuralFieldVI(**kw).fit(df, seed=3, ensemble_size=2, num_epochs=2,
                                                    batch_size=480, sample_size_posterior=3)
  assert est.losses_.shape == (1, 2, 2 * (len(df) // 480)) and np.isfinite(est.losses_).all()
  means, q = est.predict(df.iloc[:50])
  assert means.shape == (1, 3, 2, 50) and np.isfinite(q[0]).all()
  lm = est.likelihood_model(df.iloc[:50])
  assert lm.batch_shape == (1, 3, 2) and lm.sample(5, seed=0).shape == (5, 1, 3, 2, 50)


@pytest.mark.parametrize('cta2', ['0', '1'])
@pytest.mark.parametrize('nets,m,n,k', [(1, 128, 64, 64), (1, 256, 256, 64), (2, 300, 256, 256),
                                        (3, 200, 128, 192), (2, 1000, 512, 1024), (8, 10440, 256, 256)])
def test_gemm_mixed_major(cuda, monkeypatch, cta2, nets, m, n, k):
  """C = A[M,K] . B[K,N]: A K-major, B MN-major -- the forward GEMM reading the natural
  (in,out) bf16 kernel copy, single-CTA and CTA-pair tiles."""
  monkeypatch.setenv('BNF_CTA2', cta2)
  g = torch.Generator(device='cuda').manual_seed(m + 3 * n + k)
  a = torch.randn(nets, m, k, generator=g, device=cuda).to(torch.bfloat16)
  b = torch.randn(nets, k, n, generator=g, device=cuda).to(torch.bfloat16)
  c = _gemm(2, a, b, nets, m, n, k)
  want = torch.bmm(a.float(), b.float())
  err = float((c - want).abs().max())
  assert err <= 2e-3 * float(want.abs().max()), err


@pytest.mark.parametrize('flag,val', [('BNF_LEGACY_STEP', '1'), ('BNF_PDL', '0'), ('BNF_FWD_WT', '1'),
                                      ('BNF_NO_GRAPH', '1')])
def test_fused_step_variants_agree(cuda, monkeypatch, flag, val):
  """The fused MAP step (map_update kernel, programmatic dependent launch, forward reading the
  natural-layout bf16 kernels, CUDA-graph replay) against the same training run with each piece
  switched off: same losses and parameters up to atomic-order noise."""
  from bayesnf_b200 import inference
  n = 1500
  cfg = _cfg(256, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3)
  x, y = xd.cpu().numpy().astype(np.float64), yd.cpu().numpy().astype(np.float64)
  monkeypatch.delenv(flag, raising=False)
  p0, l0 = inference.fit_map(x, y, 0, 'NORMAL', cfg, 3, 0.005, 12, precision='bf16', init_params=P.numpy())
  monkeypatch.setenv(flag, val)
  p1, l1 = inference.fit_map(x, y, 0, 'NORMAL', cfg, 3, 0.005, 12, precision='bf16', init_params=P.numpy())
  assert np.isfinite(l0).all() and l0.shape == (1, 3, 12)
  np.testing.assert_allclose(l0, l1, rtol=2e-3)
  f0, f1 = spec.flatten(p0)[0], spec.flatten(p1)[0]
  # Adam normalises the step, so an entry whose gradient is pure summation noise may walk the
  # other way; everything else must agree closely
  assert float(np.quantile(np.abs(f0 - f1), 0.999)) <= 2e-3


@pytest.mark.parametrize('prec', ['bf16x3', 'bf16', 'fp32'])
def test_unrolled_step_graph_equals_single_step_graph(cuda, monkeypatch, prec):
  """Long calls replay the step captured BNF_GRAPH_UNROLL (default 8) times per graph launch, so the
  programmatic-dependent-launch chain also spans the step boundary (fused update of step s -> first
  kernel of step s+1).  Device-drawn minibatches (the batch window is keyed by the device-side step
  count), 21 steps = two unrolled launches + five single-step launches: same losses and parameters
  as one graph launch per step, as direct launches, and as another unroll factor."""
  from bayesnf_b200 import inference
  n, B, epochs = 1500, 500, 7
  steps = epochs * (n // B)                       # 21
  cfg = _cfg(256, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 3)
  eng = inference.Engine(spec, prec)

  def run():
    p = P.cuda().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    sc = torch.zeros(1, dtype=torch.int32, device=cuda)
    ls = eng.map_epochs(p, m, v, sc, xd, yd, B, n, epochs, 0.005, 1.0, 11, 0)
    torch.cuda.synchronize()
    assert int(sc[0]) == steps
    return p.cpu().numpy(), ls.cpu().numpy()
  monkeypatch.delenv('BNF_GRAPH_UNROLL', raising=False)
  monkeypatch.delenv('BNF_NO_GRAPH', raising=False)
  p0, l0 = run()
  assert np.isfinite(l0).all() and l0.shape[0] == steps
  # (atomic-order noise through 21 Adam steps; a wrong batch window or step count moves a loss by percents)
  tol_l, tol_p = (5e-4, 2e-3) if prec != 'bf16' else (5e-3, 5e-3)
  for flag, val in (('BNF_GRAPH_UNROLL', '1'), ('BNF_GRAPH_UNROLL', '3'), ('BNF_NO_GRAPH', '1')):
    monkeypatch.setenv(flag, val)
    p1, l1 = run()
    monkeypatch.delenv(flag)
    np.testing.assert_allclose(l0, l1, rtol=tol_l, err_msg=f'{flag}={val}')
    # (Adam normalises the step: an entry whose gradient is summation noise may walk the other way)
    assert float(np.quantile(np.abs(p0 - p1), 0.999)) <= tol_p, (flag, val)


def test_single_step_calls_replay_cached_graph(cuda):
  """bnf_map_steps called one step at a time (what bench.py's e2e loop does) == one call with
  all the steps: the cached graph is keyed on its baked arguments and the device-side cursors
  (Adam count, loss row) are re-armed by every call's prologue."""
  from bayesnf_b200 import inference, _lib
  n = 900
  cfg = _cfg(128, 2, n)
  om, spec, P, xd, yd = _setup(cfg, n, 2)
  for prec in ('fp32', 'bf16'):
    eng = inference.Engine(spec, prec)
    outs = []
    for mode in ('single', 'batched'):
      p = P.cuda().clone()
      m, v = torch.zeros_like(p), torch.zeros_like(p)
      sc = torch.zeros(1, dtype=torch.int32, device=cuda)
      if mode == 'single':
        ls = torch.cat([eng.map_steps(p, m, v, sc, xd, yd, None, n, n, 1, 0.01, 1.0) for _ in range(7)])
      else:
        ls = eng.map_steps(p, m, v, sc, xd, yd, None, n, n, 7, 0.01, 1.0)
      torch.cuda.synchronize()
      assert int(sc[0]) == 7
      outs.append((p.cpu(), ls.cpu()))
    tol = 1e-5 if prec == 'fp32' else 2e-3
    np.testing.assert_allclose(outs[0][1].numpy(), outs[1][1].numpy(), rtol=tol)
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= (1e-5 if prec == 'fp32' else 2e-2)


def _split3(x):
  """f32 -> three bf16 planes concatenated along the last axis (the bf16x3 operand layout)."""
  x0 = x.to(torch.bfloat16)
  r1 = x - x0.float()
  x1 = r1.to(torch.bfloat16)
  x2 = (r1 - x1.float()).to(torch.bfloat16)
  return torch.cat([x0, x1, x2], dim=-1).contiguous()


@pytest.mark.parametrize('cta2', ['0', '1'])
@pytest.mark.parametrize('layout,nets,m,n,k', [
    (3, 1, 128, 64, 64), (3, 2, 300, 256, 256), (3, 2, 1000, 512, 1024), (3, 8, 10440, 256, 64),
    (4, 1, 64, 64, 200), (4, 2, 256, 256, 1000), (4, 8, 256, 256, 10440),
    (5, 1, 128, 64, 64), (5, 3, 200, 128, 192), (5, 2, 1000, 256, 256)])
def test_gemm_split_operands(cuda, monkeypatch, cta2, layout, nets, m, n, k):
  """bf16x3 GEMMs (six bf16 products of the operands' three planes, one TMEM accumulator) in the
  three operand layouts of the dense stack, against the f64 product of the f32 inputs:
  <= 2e-6 of the output scale per 1024 reduction elements (f32 rounding class; cuBLAS f32 measures
  5e-7 .. 1.2e-6 on the same shapes, profiles/experiments/README.md r2a-3).  The TMEM accumulation
  rounds toward zero once per 16 reduction elements, so the error of ONE accumulation grows with
  K; the real wgrad therefore splits the batch into <= 4096-row accumulations (tc_wgrad) whose
  partial sums are added by f32 atomics -- the whole-path tests check that."""
  monkeypatch.setenv('BNF_CTA2', cta2)
  g = torch.Generator(device='cuda').manual_seed(m + 3 * n + k + layout)
  if layout == 3:     # A [M,K] . B [K,N]
    a = torch.randn(nets, m, k, generator=g, device=cuda)
    b = torch.randn(nets, k, n, generator=g, device=cuda)
    want = torch.bmm(a.double(), b.double())
  elif layout == 4:   # A [K,M]^T . B [K,N]
    a = torch.randn(nets, k, m, generator=g, device=cuda)
    b = torch.randn(nets, k, n, generator=g, device=cuda)
    want = torch.bmm(a.double().transpose(1, 2), b.double())
  else:               # A [M,K] . B [N,K]^T
    a = torch.randn(nets, m, k, generator=g, device=cuda)
    b = torch.randn(nets, n, k, generator=g, device=cuda)
    want = torch.bmm(a.double(), b.double().transpose(1, 2))
  c = _gemm(layout, _split3(a), _split3(b), nets, m, n, k)
  err = float((c.double() - want).abs().max())
  assert err <= 2e-6 * max(1.0, k / 1024) * float(want.abs().max()), err


@pytest.mark.parametrize('prec,ll_tol,g_tol', [('bf16', 3e-2, 5e-2), ('bf16x3', 2e-5, 1e-4)])
def test_zinb_w512_l4_against_oracle(cuda, prec, ll_tol, g_tol):
  """BASELINE configs[3] shape (air_quality MLE: ZINB observation model, width 512, depth 4;
  models.py:166-191) on the tensor-core paths against the f64 oracle: at W > 256 the last layer
  runs the plain forward kernel + head_fused_kernel (row dots, ZINB log-pmf with lgamma/digamma,
  activation backward).  bf16: 3e-2 / 5e-2 of scale; bf16x3: the f32 parity tolerances."""
  from bayesnf_b200 import inference, models
  from test_gpu_parity import _data, _random_params
  n = 600
  cfg = dict(width=512, depth=4, input_scales=[n - 1.0, 1, 1], num_seasonal_harmonics=[4, 4],
             seasonality_periods=[24, 168], init_x=(n, 3), fourier_degrees=[5, 5, 5], interactions=np.zeros((0, 2), int))
  x, y = _data(cfg, n, counts=True)
  om, om64 = O.OracleModel(**cfg), O.OracleModel(**cfg, dtype=torch.float64)
  P = _random_params(om, 2, y, seed=23)
  spec = models.ModelSpec(**cfg, observation_model='ZINB')
  eng = inference.Engine(spec, prec)
  xd, yd = inference._to_device_data(x, y)
  ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
  ll, grad = ll.cpu(), grad.cpu()
  parts = [(0, 1), (1, 2), (2, 3)] + [(o, o + (int(np.prod(s)) if s else 1)) for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
  for j in range(2):
    loss64, g64 = O.map_loss_and_grad(om64, P[j].double(), xd.cpu().double(), yd.cpu().double(), n, 0.0, 'ZINB')
    assert abs(float(ll[j]) + float(loss64)) <= ll_tol * abs(float(loss64)), (prec, float(ll[j]), float(loss64))
    for a, b in parts:
      if a == 0:
        continue                                   # log_noise_scale is unused by ZINB: zero gradient
      want, got = -g64[a:b], grad[j, a:b].double()
      floor = 1e-3 * g_tol if prec == 'bf16' else 1e-7
      tol = g_tol * float(want.abs().max()) + floor * float(g64.abs().max()) + 1e-7
      assert float((got - want).abs().max()) <= tol, (prec, j, a, b, float((got - want).abs().max()), tol)
    assert float(grad[j, 0].abs()) == 0.0


@pytest.mark.parametrize('width,depth,n', [(1024, 3, 300), (512, 6, 260)])
def test_bf16x3_wide_and_deep_against_oracle(cuda, width, depth, n):
  """The f32-parity tensor-core mode at the widest supported layer (K = 1024: 64 round-toward-zero
  accumulation steps at full magnitude per GEMM) and at depth 6: forward <= 1e-5 of the output
  scale, log-lik <= 2e-5, every gradient leaf <= 1e-4 of its scale vs the f64 oracle -- the same
  bars as the small configurations."""
  from bayesnf_b200 import inference
  cfg = _cfg(width, depth, n)
  om, spec, P, xd, yd = _setup(cfg, n, 2)
  om64 = O.OracleModel(**cfg, dtype=torch.float64)
  eng = inference.Engine(spec, 'bf16x3')
  loc = eng.forward(P.cuda(), xd).cpu()
  ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
  ll, grad = ll.cpu(), grad.cpu()
  parts = [(0, 1)] + [(o, o + (int(np.prod(s)) if s else 1)) for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
  for j in range(2):
    want = om64.forward(om64.unflatten(P[j].double()), xd.cpu().double())
    assert float((loc[j].double() - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6
    loss64, g64 = O.map_loss_and_grad(om64, P[j].double(), xd.cpu().double(), yd.cpu().double(), n, 0.0, 'NORMAL')
    assert abs(float(ll[j]) + float(loss64)) <= 2e-5 * abs(float(loss64))
    for a, b in parts:
      w_, g_ = -g64[a:b], grad[j, a:b].double()
      tol = 1e-4 * float(w_.abs().max()) + 1e-7 * float(g64.abs().max()) + 1e-7
      assert float((g_ - w_).abs().max()) <= tol, (width, depth, j, a, b, float((g_ - w_).abs().max()), tol)


EDGE_CFGS = {
    # one hidden layer: TC_FWD_HEAD consumes the feature planes directly, no fused dgrad at all
    'depth1': dict(width=64, depth=1, input_scales=[99., 1, 1], num_seasonal_harmonics=[2, 3],
                   seasonality_periods=[7.0, 30.0], fourier_degrees=[3, 2, 2], interactions=np.array([[1, 2]])),
    # more than 64 features: Fp = 128 (two k-blocks in Dense_0, 128-wide dgrad_0 tile)
    'wide_features': dict(width=128, depth=2, input_scales=[99., 1, 1], num_seasonal_harmonics=[10, 10],
                          seasonality_periods=[30.0, 365.25], fourier_degrees=[9, 8, 8],
                          interactions=np.array([[0, 1], [1, 2]])),
    # F == Fp == 64 exactly: no pad column, so the Dense_0 bias gradient cannot ride on the wgrad GEMM
    'no_pad_column': dict(width=64, depth=2, input_scales=[99., 1, 1], num_seasonal_harmonics=[3, 10],
                          seasonality_periods=[7.0, 52.0], fourier_degrees=[6, 6, 5],
                          interactions=np.array([[1, 2]])),
}


@pytest.mark.parametrize('prec,ll_tol,g_tol', [('bf16x3', 2e-5, 1e-4), ('bf16', 3e-2, 6e-2)])
@pytest.mark.parametrize('rows,nets', [(1, 1), (130, 1), (257, 3)])
@pytest.mark.parametrize('name', sorted(EDGE_CFGS))
def test_edge_shapes_against_oracle(cuda, name, rows, nets, prec, ll_tol, g_tol):
  """Shapes at the edges of the tensor-core kernels (a single row, one row past a tile, one network;
  one hidden layer; 128 padded features; no pad column) in both tensor-core modes against the f64
  oracle: forward, log-likelihood and every gradient leaf.  With Fourier degrees above the
  reference's default 5 the f32 restatement ITSELF is up to 4e-4 away from its f64 twin on the
  cancellation-prone log_scale_adjustment gradient (2*pi*2^d*x reaches hundreds of radians in f32),
  so a leaf's tolerance is 1e-4 of its scale plus the f32 oracle's own distance from f64 -- measured,
  bf16x3 then sits within 4e-6 of the f32 oracle there."""
  from bayesnf_b200 import inference, models
  from test_gpu_parity import _data, _random_params
  cfg = dict(EDGE_CFGS[name], init_x=(rows, 3))
  x, y = _data(cfg, rows)
  om, om64 = O.OracleModel(**cfg), O.OracleModel(**cfg, dtype=torch.float64)
  if name == 'no_pad_column':
    assert om.F == 64
  if name == 'wide_features':
    assert 64 < om.F <= 128
  P = _random_params(om, nets, y if rows > 1 else np.array([1.0, 2.0]), seed=41)
  spec = models.ModelSpec(**cfg, observation_model='NORMAL')
  eng = inference.Engine(spec, prec)
  xd, yd = inference._to_device_data(x, y)
  loc = eng.forward(P.cuda(), xd).cpu()
  ll, grad = eng.loglik_grad(P.cuda(), xd, yd)
  ll, grad = ll.cpu(), grad.cpu()
  parts = [(0, 1)] + [(o, o + (int(np.prod(s)) if s else 1)) for o, s in zip(spec.leaf_offsets, spec.leaf_shapes)]
  for j in range(nets):
    want = om64.forward(om64.unflatten(P[j].double()), xd.cpu().double())
    f_tol = 1e-5 if prec == 'bf16x3' else 3e-2
    assert float((loc[j].double() - want).abs().max()) <= f_tol * float(want.abs().max()) + 1e-6, (name, rows, prec)
    loss64, g64 = O.map_loss_and_grad(om64, P[j].double(), xd.cpu().double(), yd.cpu().double(), rows, 0.0, 'NORMAL')
    _, g32 = O.map_loss_and_grad(om, P[j], xd.cpu(), yd.cpu(), rows, 0.0, 'NORMAL')
    assert abs(float(ll[j]) + float(loss64)) <= ll_tol * abs(float(loss64)) + 1e-5, (name, rows, prec)
    for a, b in parts:
      w_, g_ = -g64[a:b], grad[j, a:b].double()
      floor = 1e-3 * g_tol if prec == 'bf16' else 1e-7
      f32_slack = 2.0 * float((g32[a:b].double() - g64[a:b]).abs().max())
      tol = g_tol * float(w_.abs().max()) + floor * float(g64.abs().max()) + f32_slack + 1e-7
      assert float((g_ - w_).abs().max()) <= tol, (name, rows, nets, prec, a, b, float((g_ - w_).abs().max()), tol)
