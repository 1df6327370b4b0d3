#!/usr/bin/env python
"""bench.py -- BayesNF ensemble-training hot path on B200 (contract: see the task brief).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels)
  python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm

A "step" is one MAP training step (encode -> dense stack fwd -> Normal log-lik ->
backward -> prior + Adam) of every ensemble member over one full batch of
synthetic spatiotemporal data.  Default workload = BASELINE.json configs[1]:
chickenpox-shaped MAP, width 256, depth 2, 8 members per GPU, bf16 tensor cores,
N = 10 440 rows (20 sites x 522 weeks), full batch.  Multi-GPU: members shard
across ranks with NO data-path collective (weak scaling: 8 members per GPU).

The line printed by rank 0 carries: metric/value (whole-job samples/s with
inputs resident in HBM), e2e (same metric through the public Engine API with
pinned-host inputs copied every step and the loss read back every step -- `value`
with the reads pipelined behind the steps like a fit() call, `per_step_sync_value`
with the host blocking on every step's loss),
roofline (dominant kernel, CUDA-event timed inside this script),
cpu_baseline (the torch-CPU oracle on a bounded sample; "port": the reference's
JAX path cannot be installed here), clocks, gpu_launches.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train samples/sec (ensemble x batch rows per second, whole job)'

WORKLOADS = {
    # BASELINE.json configs[1]
    'chickenpox_map_e8': dict(width=256, depth=2, members_per_gpu=8, sites=20, times=522, batch=None,
                              periods=[4.0, 52.1775], harmonics=[2, 10], objective='map'),
    # BASELINE.json configs[4] per-GPU shard (roofline run): E=128/8 GPUs, B=65536
    'wind_map_e16': dict(width=1024, depth=6, members_per_gpu=16, sites=128, times=512, batch=65536,
                         periods=[7, 365.25 / 12, 365.25], harmonics=[3, 10, 10], objective='map'),
    # BASELINE.json configs[3]-shaped dense stack (W512 L4) with Normal obs, 8 members
    # BASELINE.json configs[2] per-GPU shard: 1M-row field, VI, W512 L4, 64 members / 8 GPUs,
    # S=5 Monte-Carlo draws, batch 65536 (one shared random sub-batch per step)
    'synthetic_vi_e8': dict(width=512, depth=4, members_per_gpu=8, sites=1000, times=1000, batch=65536,
                            periods=[24, 168], harmonics=[4, 4], objective='vi', mc_samples=5),
    'air_quality_map_e8': dict(width=512, depth=4, members_per_gpu=8, sites=64, times=512, batch=None,
                               periods=[24, 168], harmonics=[4, 4], objective='map'),
    # BASELINE.json configs[3] per-GPU shard: air_quality MLE (prior_weight 0), ZINB observation
    # model, W512 L4, 32 members / 4 GPUs, batch 38 096 (scripts/evaluate.py:199-204) out of
    # 76 224 rows (2 steps per epoch; per-member permutations drawn on the device every epoch)
    'air_quality_mle_zinb_e8': dict(width=512, depth=4, members_per_gpu=8, sites=64, times=1191, batch=38096,
                                    periods=[24, 168], harmonics=[4, 4], objective='mle', likelihood='ZINB'),
}


def synth(wl, seed=20240925):
  """Seeded synthetic field: T integer time steps x S sites."""
  rng = np.random.default_rng(seed)
  T, S = wl['times'], wl['sites']
  lat, lon = rng.normal(size=S), rng.normal(size=S)
  lat, lon = (lat - lat.mean()) / lat.std(), (lon - lon.mean()) / lon.std()
  t = np.repeat(np.arange(T, dtype=np.float64), S)
  la, lo = np.tile(lat, T), np.tile(lon, T)
  y = np.zeros(T * S)
  for p in wl['periods']:
    y += rng.normal() * np.sin(2 * np.pi * t / p + rng.uniform(0, 6.28))
  y += 0.7 * la - 0.4 * lo * la + rng.normal(scale=0.5, size=T * S)
  if wl.get('likelihood', 'NORMAL') != 'NORMAL':
    # counts: negative-binomial draws around exp(field), ~30 % structural zeros
    mean = np.exp(0.5 * y + 1.0)
    y = rng.negative_binomial(2.0, 2.0 / (2.0 + mean)).astype(np.float64)
    y[rng.random(T * S) < 0.3] = 0.0
  x = np.stack([t, la, lo], 1)
  margs = dict(width=wl['width'], depth=wl['depth'], input_scales=np.array([T - 1.0, 1.0, 1.0]),
               num_seasonal_harmonics=np.array(wl['harmonics']),
               seasonality_periods=np.array(wl['periods'], dtype=float),
               init_x=(wl['batch'] or T * S, 3), fourier_degrees=np.array([5, 5, 5]),
               interactions=np.zeros((0, 2), int))
  return x, y, margs


def workload_config(workload, world, F):
  """The `config` object of the JSON line -- identical for our arm and the reference arm."""
  wl = WORKLOADS[workload]
  E, S = wl['members_per_gpu'], wl.get('mc_samples', 1)
  n_total = wl['times'] * wl['sites']
  B = wl['batch'] or n_total
  act_bytes = E * S * B * wl['width'] * 2 * (2 * wl['depth'] + 2)     # bf16 activations of one step
  return {'workload': workload, 'width': wl['width'], 'depth': wl['depth'], 'features': int(F),
          'members_per_gpu': E, 'members_total': E * world, 'mc_samples': S, 'batch_rows': B, 'rows_total': n_total,
          'objective': wl['objective'], 'observation_model': wl.get('likelihood', 'NORMAL'),
          'parallelism': f'members sharded x{world}, no collective in training',
          'l2': (f'activation working set {act_bytes / 2**20:.0f} MiB per step > 126 MiB L2' if act_bytes > 126 * 2**20 else
                 f'activation working set {act_bytes / 2**20:.0f} MiB per step (fits L2; steps are data-dependent so '
                 'no flush is inserted)')}


def flops_per_sample(F, W, L):
  return 6 * (F * W + (L - 1) * W * W + W)      # fwd + dgrad + wgrad


class ClockSampler(threading.Thread):
  """nvidia-smi style clock / throttle-reason sampling DURING the timed region."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
    self.max_mhz = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:  # pylint: disable=broad-except
      self.nv = None

  def run(self):
    if self.nv is None:
      return
    nv = self.nv
    names = {nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown')
             else nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
             getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
             getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
             getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap'}
    while not self.stop_flag:
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if mask & bit:
            self.reasons.add(name)
      except Exception:  # pylint: disable=broad-except
        pass
      time.sleep(0.02)

  def summary(self):
    if not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
    return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons)}


# committed `ncu --set full` summaries of the same commands (profiles/), per (workload, precision)
NCU_SUMMARY = {('chickenpox_map_e8', 'bf16'): 'ncu_chickenpox_bf16_r2z_summary.csv',
               ('chickenpox_map_e8', 'bf16x3'): 'ncu_chickenpox_bf16x3_r2z_summary.csv',
               ('wind_map_e16', 'bf16'): 'ncu_wind_tc_gemm_r2_summary.csv',
               ('wind_map_e16', 'bf16x3'): 'ncu_wind_tc_gemm_bf16x3_r2_summary.csv'}
# kernel class -> template-argument substrings <BLOCK_N, A_MODE, MODE, CTA2, X3> of its instantiations
NCU_PATTERN = {'tc_gemm_fwd': (', 3, 0, 1, 0>', ', 3, 0, 0, 0>', ', 0, 0, 1, 0>', ', 0, 0, 0, 0>'),
               'tc_gemm_dgrad': (', 0, 5, 1, 0>', ', 0, 5, 0, 0>'), 'tc_gemm_wgrad': (', 1, 3, ',),
               'tc_fwd_head': (', 3, 7, 1, 0>', ', 3, 7, 0, 0>'), 'tc_dgrad0_enc': (', 0, 6, ',),
               'tc_gemm_fwd_x3': (', 3, 0, 1, 1>', ', 3, 0, 0, 1>'), 'tc_fwd_head_x3': (', 3, 7, 1, 1>', ', 3, 7, 0, 1>'),
               'tc_gemm_dgrad_x3': (', 0, 5, 1, 1>', ', 0, 5, 0, 1>')}


def algorithmic_work(rows, F, Fp, W, L, P_total, fused_head, precision='bf16'):
  """Algorithmic FLOPs and HBM bytes PER STEP of every kernel class:
  activations read/written once per producer/consumer, weights ignored (L2-resident).
  bf16: 2 bytes per activation element.  bf16x3: GEMM operands (feat, h, dU) are three bf16 planes
  = 6 bytes, pre-activations z are f32 = 4 bytes; the FLOPs are the ALGORITHMIC f32 ones (the
  tensor cores execute six bf16 products per f32 product)."""
  if precision == 'bf16x3':
    # feat / h: three bf16 planes (6 B), z: f32 (4 B), backpropagated dU: two bf16 planes (4 B)
    a6, a4, f6 = 6 * rows * W, 4 * rows * W, 6 * rows * Fp
    nf = L - 1 if fused_head else L          # layers run by the plain forward kernel
    return {
        'tc_gemm_fwd_x3': dict(flops=2.0 * rows * (F * W + (nf - 1) * W * W), bytes=f6 + nf * (a4 + a6) + (nf - 1) * a6),
        'tc_fwd_head_x3': dict(flops=2.0 * rows * (W * W + W), bytes=a6 + a4),          # h in, dU out
        'head_fused': dict(flops=2.0 * rows * W, bytes=a6 + a4 + a4),                   # h, z in, dU out
        'tc_gemm_dgrad_x3': dict(flops=2.0 * rows * (L - 1) * W * W, bytes=(L - 1) * 3 * a4),   # dU in, z in, dU out
        'tc_dgrad0_enc': dict(flops=2.0 * rows * F * W, bytes=a4),
        'tc_gemm_wgrad': dict(flops=2.0 * rows * (F * W + (L - 1) * W * W), bytes=f6 + a4 + (L - 1) * (a6 + a4)),
        'map_update': dict(flops=0.0, bytes=34.0 * P_total),                           # + three bf16 planes restaged
        'encode': dict(flops=0.0, bytes=float(f6)),
    }
  a = 2 * rows * W            # bytes of one bf16 activation tensor [rows, W]
  nf = L - 1 if fused_head else L          # launches of the plain forward kernel
  return {
      'tc_gemm_fwd': dict(flops=2.0 * rows * (F * W + (nf - 1) * W * W),
                          bytes=2 * rows * Fp + nf * 2 * a + (nf - 1) * a),     # feat, (z,h) out, h in
      'tc_fwd_head': dict(flops=2.0 * rows * (W * W + W), bytes=2 * a),          # h in, dU out
      'tc_gemm_dgrad': dict(flops=2.0 * rows * (L - 1) * W * W, bytes=(L - 1) * 3 * a),   # dU in, z in, dU out
      'tc_dgrad0_enc': dict(flops=2.0 * rows * F * W, bytes=a),
      'tc_gemm_wgrad': dict(flops=2.0 * rows * (F * W + (L - 1) * W * W), bytes=2 * rows * Fp + a + (L - 1) * 2 * a),
      'head_fused': dict(flops=2.0 * rows * W, bytes=3 * a),                     # h, z in, dU out
      'map_update': dict(flops=0.0, bytes=32.0 * P_total),                       # p,g,m,v in; p,m,v out; bf16 restage
      'encode': dict(flops=0.0, bytes=2.0 * rows * Fp),
  }


def ncu_traffic_gb(workload, kernel, precision='bf16'):
  """DRAM bytes (read+write) per launch of `kernel` from the committed `ncu --set full` summary
  of the same workload (profiles/), or None."""
  import csv
  path = os.path.join(ROOT, 'profiles', NCU_SUMMARY.get((workload, precision), ''))
  if not os.path.isfile(path) or kernel not in NCU_PATTERN:
    return None
  rows = list(csv.reader(open(path)))
  hdr, units = rows[0], rows[1]
  ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
  scale = {'Gbyte': 1.0, 'Mbyte': 1e-3, 'Kbyte': 1e-6, 'byte': 1e-9}[units[ir]]
  vals = [(float(r[ir]) + float(r[iw])) * scale for r in rows[2:]
          if any(pat in r[ik] for pat in NCU_PATTERN[kernel]) and 'tc_gemm_kernel<' in r[ik]]
  return max(vals) if vals else None      # the widest launch of that class (hidden layers)


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return dict(hbm_gbs=p['hbm_gbs'], tflops=p['bf16_tflops'], tflops_sustained=p['bf16_tflops_sustained'],
                source='measured (MEASURED_PEAKS.json)')
  return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def _oracle_stepper(x, y, margs, wl=None):
  """One oracle MAP/MLE step (value_and_grad + Adam) of ONE member on one batch."""
  lik = (wl or {}).get('likelihood', 'NORMAL')
  pw = 0.0 if (wl or {}).get('objective') == 'mle' else 1.0
  import torch
  from oracle import bnf_oracle as O
  om = O.OracleModel(**margs)
  g = torch.Generator().manual_seed(0)
  state = {'p': om.flatten(O.init_map_params(om, y, g)), 't': 0}
  state['m'], state['v'] = torch.zeros_like(state['p']), torch.zeros_like(state['p'])
  B = margs['init_x'][0]
  xt = torch.tensor(x[:B], dtype=torch.float32)
  yt = torch.tensor(y[:B], dtype=torch.float32)
  n_total = len(y)

  def step():
    state['t'] += 1
    loss, gr = O.map_loss_and_grad(om, state['p'], xt, yt, n_total, pw, lik)
    state['p'], state['m'], state['v'] = O.adam_update(state['p'], gr, state['m'], state['v'], state['t'], 0.005)
  return step, B


def _best_thread_count(step):
  """The CPU arm may use every host thread, but torch's intra-op pool gets SLOWER past a point
  on these GEMM sizes: time one step at a few pool sizes and keep the fastest."""
  import torch
  ncpu = os.cpu_count() or 1
  best, best_t = ncpu, float('inf')
  for n in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
    torch.set_num_threads(n)
    step()                                   # warm the pool
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    if dt < best_t:
      best, best_t = n, dt
  torch.set_num_threads(best)
  return best


def cpu_oracle_rate(x, y, margs, budget_s, wl=None):
  """The CPU baseline: oracle MAP steps for ~budget_s seconds.  Returns samples/s."""
  step, B = _oracle_stepper(x, y, margs, wl)
  threads = _best_thread_count(step)
  steps, t0 = 0, time.perf_counter()
  while True:
    step()
    steps += 1
    el = time.perf_counter() - t0
    if el > budget_s and steps >= 3:
      break
  return B * steps / el, steps, threads


def run_reference(args, wl, x, y, margs):
  """--impl reference: the reference's CPU implementation of the path.  JAX is not
  installable here (no network), so this is the torch-CPU oracle
  ("port") on all host cores.  One 'step' = one MAP step of one member on the full
  batch; K steps bounded to a few minutes."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import torch
  from oracle import bnf_oracle as O
  om_features = O.OracleModel(**margs).F
  step, B = _oracle_stepper(x, y, margs, wl)
  _best_thread_count(step)                   # torchrun exports OMP_NUM_THREADS=1: undo it
  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  el = time.perf_counter() - t0
  val = B * args.steps / el
  cores = torch.get_num_threads()
  sample = f'1 member x {B} rows x {args.steps} steps (of {wl["members_per_gpu"]} members per GPU)'
  print(json.dumps({
      'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'samples/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * el / args.steps,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic',
      'config': workload_config(args.workload, args.gpus, om_features),
      'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': cores, 'cores_available': os.cpu_count(),
                       'kind': 'port', 'sample': sample},
      'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }))


def run_workload(args, workload, precision, steps, warmup, env, do_e2e=True, do_profile=True, cpu_baseline=False,
                 repeats=0, min_window_ms=100.0, keep=None):
  """Times blocks of `steps` training steps of one workload in one arithmetic mode on this rank's
  GPU and returns the record rank 0 prints (None on the other ranks)."""
  import torch
  import torch.distributed as dist
  from bayesnf_b200 import _lib, inference, models
  rank, world, local, dev = env['rank'], env['world'], env['local'], env['dev']
  wl = WORKLOADS[workload]
  x, y, margs = synth(wl)
  lik = wl.get('likelihood', 'NORMAL')
  pw = 0.0 if wl['objective'] == 'mle' else 1.0            # MLE = MAP without the prior (spatiotemporal.py:544-551)
  spec = models.ModelSpec(**margs, observation_model=lik)
  eng = inference.Engine(spec, precision)
  E = wl['members_per_gpu']
  n_total = len(y)
  B = wl['batch'] or n_total
  xd, yd = inference._to_device_data(x, y)
  lns = float(np.log(np.nanstd(y) / 2))
  is_vi = wl['objective'] == 'vi'
  S = wl.get('mc_samples', 1)
  p = eng.init_params(0.0 if is_vi else lns, 1234, rank * E, E)
  sc = torch.zeros(1, dtype=torch.int32, device=dev)
  if is_vi:
    rho = torch.full_like(p, math.log(math.expm1(0.3)))
    m = torch.zeros((E, 2, spec.num_params), dtype=torch.float32, device=dev)
    v = torch.zeros_like(m)
  else:
    m, v = torch.zeros_like(p), torch.zeros_like(p)

  def run(k, xx=None, yy=None):
    """k training steps through the Engine (one C call; all draws on the device)."""
    xx = xd if xx is None else xx
    yy = yd if yy is None else yy
    if is_vi:     # tfp.vi steps: S reparameterised draws per member and a FRESH shared sub-batch every step
      return eng.vi_steps(p, rho, m, v, sc, S, 977 + rank, rank, xx, yy, B, n_total, k, 0.01, 0.1)
    if B >= n_total:
      return eng.map_steps(p, m, v, sc, xx, yy, None, B, n_total, k, 0.005, pw)
    # minibatches: per-member permutations and batch windows drawn on the device, k steps = k windows
    spe = n_total // B
    assert k % spe == 0, f'--steps must be a multiple of the {spe} steps per epoch of this workload'
    return eng.map_epochs(p, m, v, sc, xx, yy, B, n_total, k // spe, 0.005, pw, 4321, rank * E)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing -------------------------------------------------
  # Warm up with W steps AND one block of exactly the timed shape (K steps: graph capture,
  # workspace and loss-buffer allocation all happen here), then time R blocks of EXACTLY K steps,
  # each bracketed by barrier + synchronize on both sides and by CUDA events on the launching
  # stream; per block the MAX over ranks, over blocks the MEDIAN.  R >= 10 and R*K steps >= 100 ms
  # for the headline, so one host hiccup on one rank cannot set the number.
  spe_w = 1 if (is_vi or B >= n_total) else n_total // B
  run(-(-max(3, warmup) // spe_w) * spe_w)          # warm-up steps, rounded up to whole epochs for minibatch workloads
  run(steps)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  run(steps)
  e1.record()
  torch.cuda.synchronize()
  est_ms = max(e0.elapsed_time(e1), 1e-3)
  repeats = repeats or int(min(200, max(10 if min_window_ms >= 100.0 else 3, math.ceil(min_window_ms / est_ms))))
  rep = torch.tensor([repeats], dtype=torch.int64, device=dev)
  if world > 1:
    dist.all_reduce(rep, op=dist.ReduceOp.MAX)
  repeats = int(rep[0])
  sampler = ClockSampler(local)
  sampler.start()
  block_ms, launches = [], 0
  for _ in range(repeats):
    barrier()
    l0 = _lib.lib.bnf_debug_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = run(steps)
    e1.record()
    barrier()
    launches = _lib.lib.bnf_debug_launch_count() - l0
    block_ms.append(e0.elapsed_time(e1))
  sampler.stop_flag = True
  t_ms = torch.tensor(block_ms, dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)      # per block: the slowest rank
  ms = float(t_ms.median())
  ms_min, ms_max = float(t_ms.min()), float(t_ms.max())
  assert torch.isfinite(losses).all(), 'non-finite loss in the timed region'
  value = world * E * S * B * steps / (ms * 1e-3)   # network-rows per second (VI: x S draws)

  # ---- end to end through the public API, host buffers ------------------------
  e2e = None
  if do_e2e:
    xh = torch.tensor(x.astype(np.float32)).pin_memory()
    yh = torch.tensor(y.astype(np.float32)).pin_memory()
    # double-buffered device inputs: the H2D copy of step i+1 runs on a copy stream while step i
    # computes (an input pipeline that prefetches one step ahead); the compute stream waits for the
    # step's own copy, the copy stream waits until the step that last read the buffer is done.
    bufs = [(torch.empty_like(xd), torch.empty_like(yd)) for _ in range(2)]
    copy_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    ev_done = torch.cuda.Event()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    k_e2e = max(5, min(steps, 100))
    # Every step copies its inputs from pinned host memory and reads its loss back to the host.
    # Two ways to consume the result: (a) pipelined -- the D2H read of step i is enqueued behind
    # step i, the host goes on enqueueing step i+1 and synchronises once at the end (how a fit()
    # call behaves: the reference's jitted scan returns the losses after the loop, inference.py:
    # 608-614); (b) the host blocks on every step's loss before launching the next one.
    loss_host = torch.empty((k_e2e, E), dtype=torch.float32).pin_memory()
    if not is_vi and B < n_total:
      spe = n_total // B          # one call = one epoch of spe steps: copy the inputs once per call
    else:
      spe = 1
    main_stream = torch.cuda.current_stream()
    for ev in ev_free:
      ev.record(main_stream)

    def e2e_enqueue(i):
      b = i & 1
      xe, ye = bufs[b]
      with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev_free[b])
        xe.copy_(xh, non_blocking=True)
        ye.copy_(yh, non_blocking=True)
        ev_in[b].record(copy_stream)
      main_stream.wait_event(ev_in[b])
      ls = run(spe, xe, ye)
      ev_free[b].record(main_stream)
      ev_done.record(main_stream)
      with torch.cuda.stream(d2h_stream):            # the loss read does not hold up the next step either
        d2h_stream.wait_event(ev_done)
        ls.record_stream(d2h_stream)
        loss_host[i % k_e2e].copy_(ls.reshape(-1, E)[-1], non_blocking=True)

    def timed(fn):
      barrier()
      t0 = time.perf_counter()
      fn()
      barrier()
      t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
      if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return float(t[0])

    def pipelined():
      for i in range(k_e2e):
        e2e_enqueue(i)

    def blocking():
      for i in range(k_e2e):
        e2e_enqueue(i)
        torch.cuda.synchronize()                       # the loss of step i is on the host

    for i in range(4):
      e2e_enqueue(i)
    e2e_pipe_s = float(np.median([timed(pipelined) for _ in range(5)]))     # median of 5 blocks of k_e2e calls
    assert torch.isfinite(loss_host).all(), 'non-finite loss in the end-to-end region'
    e2e_sync_s = float(np.median([timed(blocking) for _ in range(3)]))
    e2e = {'value': world * E * S * B * spe * k_e2e / e2e_pipe_s, 'unit': 'samples/s', 'steps': k_e2e * spe,
           'h2d_bytes_per_step': int((xh.numel() * 4 + yh.numel() * 4) // spe),
           'd2h_bytes_per_step': int(E * 4),
           'mode': 'inputs H2D from pinned memory (double-buffered, copied on a second stream one step ahead) and '
                   'the loss D2H every step; reads pipelined behind the steps, one host synchronisation at the end',
           'per_step_sync_value': world * E * S * B * spe * k_e2e / e2e_sync_s}

  # ---- per-kernel CUDA-event timing (separate short run; not the headline) ----
  prof, roof = {}, None
  pk = peaks()
  if do_profile:
    _lib.check(_lib.lib.bnf_debug_profile(1))
    k_prof = 5 if (is_vi or B >= n_total) else (n_total // B) * max(1, 4 // (n_total // B))
    run(k_prof)
    buf = C.create_string_buffer(1 << 16)
    _lib.check(_lib.lib.bnf_debug_profile_report(buf, len(buf)))
    _lib.check(_lib.lib.bnf_debug_profile(0))
    for line in buf.value.decode().strip().splitlines():
      name, cnt, tot = line.split()
      prof[name] = {'launches_per_step': int(cnt) / k_prof, 'ms_per_step': float(tot) / k_prof}
    F, W, L = spec.num_features, wl['width'], wl['depth']
    Fp = spec.padded_features
    rows = E * S * B
    work = algorithmic_work(rows, F, Fp, W, L, E * S * spec.num_params, 'tc_fwd_head' in prof or 'tc_fwd_head_x3' in prof,
                            precision)
    for name, d in prof.items():
      if name in work and work[name]['flops'] > 0:
        d['tflops'] = work[name]['flops'] / (d['ms_per_step'] * 1e-3) / 1e12
      if name in work:
        d['algorithmic_gbs'] = work[name]['bytes'] / (d['ms_per_step'] * 1e-3) / 1e9
    # the dominant kernel class of the step and the roofline that bounds it: tensor pipe when its
    # algorithmic FLOPs at the measured bf16 peak take longer than its algorithmic bytes at the
    # measured HBM bandwidth, else HBM
    best = max((n for n in prof if n in work), key=lambda n: prof[n]['ms_per_step'], default=None)
    if best and precision in ('bf16', 'bf16x3'):
      d, wk = prof[best], work[best]
      n_l = d['launches_per_step']
      avg_ms = d['ms_per_step'] / n_l
      t_tensor = wk['flops'] / (pk['tflops_sustained'] * 1e12)
      t_hbm = wk['bytes'] / (pk['hbm_gbs'] * 1e9)
      traffic = ncu_traffic_gb(workload, best, precision)
      if t_tensor >= t_hbm:
        ach, peak, unit, bound = d['tflops'], pk['tflops_sustained'], 'TFLOP/s', 'tensor'
      else:
        ach, peak, unit, bound = d['algorithmic_gbs'], pk['hbm_gbs'], 'GB/s', 'hbm'
      roof = {'bound': bound, 'kernel': best, 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
              'traffic': traffic, 'traffic_unit': 'GB per launch (ncu dram read+write of the widest launch, profiles/)',
              'algorithmic_gb_per_launch': wk['bytes'] / n_l / 1e9,
              'algorithmic_gflop_per_launch': wk['flops'] / n_l / 1e9,
              'peak_source': pk['source'] + (', sustained bf16' if bound == 'tensor' else ', copy bandwidth'),
              'avg_launch_ms': avg_ms,
              'note': 'CUDA-event time per launch inside bench.py (profile pass); x3: executed tensor FLOPs are 6x '
                      'the algorithmic ones; see DESIGN.md section 9'}

  # ---- CPU baseline (rank 0, N=1 only) ----------------------------------------
  cpu = None
  if rank == 0 and world == 1 and cpu_baseline:
    val, timed_steps, cores = cpu_oracle_rate(x, y, margs, args.cpu_budget, wl)
    cpu = {'value': val, 'unit': 'samples/s', 'cores': cores, 'cores_available': os.cpu_count(), 'kind': 'port',
           'sample': f'1 member x {margs["init_x"][0]} rows x {timed_steps} {wl["objective"].upper()} steps '
                     '(torch-CPU oracle, f32)'}
  if keep is not None:
    keep.update(eng=eng, spec=spec, params=p, x=xd, E=E)
  if rank != 0:
    return None
  fps = flops_per_sample(spec.num_features, wl['width'], wl['depth'])
  return {
      'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': steps,
      'warmup': warmup, 'ms_per_step': ms / steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None,
      'dtype': {'bf16': 'bf16', 'fp32': 'f32', 'bf16_simt': 'bf16-storage/f32-fma',
                'bf16x3': 'f32 via bf16x3 split operands (tcgen05 kind::f16, f32 accumulate)'}[precision],
      'timing': {'blocks': repeats, 'steps_per_block': steps, 'block_ms_median': ms, 'block_ms_min': ms_min,
                 'block_ms_max': ms_max, 'rule': 'median over blocks of the max over ranks (CUDA events)'},
      'data': 'synthetic',
      'config': workload_config(workload, world, spec.num_features),
      'precision': precision,
      'per_gpu_samples_per_s': value / world,
      'algorithmic_tflops_per_gpu': value / world * fps / 1e12,
      'samples_definition': 'members x MC draws x batch rows per second' if is_vi else 'members x batch rows per second',
      'e2e': e2e,
      'gpu_launches': int(launches),
      'clocks': sampler.summary(),
      'roofline': roof,
      'kernels': prof,
      'cpu_baseline': cpu,
      'final_loss_mean': float(losses[-1].mean()),
  }


def predict_block(env, kept, n_rows=1 << 20):
  """predict_bnf on N GPUs (inference.py:461-507): forward of this rank's members over n_rows test
  rows, ONE NCCL all-gather of the predictive means over NVLink, mixture quantiles on every rank.
  Also checks that the gathered means of another rank's members equal a local recomputation."""
  import torch
  import torch.distributed as dist
  from bayesnf_b200 import inference, parallel
  rank, world, dev = env['rank'], env['world'], env['dev']
  eng, p, E = kept['eng'], kept['params'], kept['E']
  g = torch.Generator(device=dev).manual_seed(7)
  xt = torch.stack([torch.rand(n_rows, generator=g, device=dev) * 600.0,
                    torch.randn(n_rows, generator=g, device=dev), torch.randn(n_rows, generator=g, device=dev)], 1).contiguous()
  if world > 1:
    dist.broadcast(xt, 0)
  qs = [0.025, 0.5, 0.975]

  def once():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    loc = eng.forward(p, xt)
    scales = 0.01 + torch.exp(p[:, 0])
    ev[1].record()
    loc_all = parallel.all_gather_leading(loc)
    sc_all = parallel.all_gather_leading(scales)
    ev[2].record()
    q = inference.mixture_quantiles(loc_all.reshape(-1, n_rows), sc_all.reshape(-1), qs, False)
    ev[3].record()
    torch.cuda.synchronize()
    return loc_all, q, [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
  once()
  if world > 1:
    dist.barrier()
  loc_all, q, t = once()
  tt = torch.tensor(t, dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  t = [float(v) for v in tt]
  check = None
  if world > 1:
    # rank 0 recomputes the members of the LAST rank locally: the gathered block must be identical
    p_all = parallel.all_gather_leading(p)
    chk = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank == 0:
      sub = slice(0, 65536)
      ref = eng.forward(p_all[world - 1].contiguous(), xt[sub].contiguous())
      chk[0] = float((ref - loc_all[world - 1][:, sub]).abs().max())
    check = float(chk[0])
  assert torch.isfinite(q).all() and bool((q[0] <= q[1]).all()) and bool((q[1] <= q[2]).all())
  if rank != 0:
    return None
  recv = (world - 1) * E * n_rows * 4
  return {'rows': n_rows, 'members_total': E * world, 'quantiles': qs, 'precision': eng.precision_name,
          'forward_ms': t[0], 'all_gather_ms': t[1] if world > 1 else None, 'quantile_root_ms': t[2],
          'rows_per_s': n_rows / (sum(t) * 1e-3),
          'member_rows_per_s': E * world * n_rows / (sum(t) * 1e-3),
          'gathered_bytes_per_rank': E * world * n_rows * 4,
          'all_gather_recv_gbs': (recv / (t[1] * 1e-3) / 1e9) if world > 1 else None,
          'nvlink_peer_gbs_per_dir': 770.0,
          'gathered_equals_local_max_abs_diff': check,
          'collective': 'one all_gather_into_tensor (NCCL) of (members, rows) f32 means + one of the scales'}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=200)
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='chickenpox_map_e8', choices=sorted(WORKLOADS))
  ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3', 'tf32x3', 'fp32', 'bf16_simt'],
                  help="bf16: single-pass tcgen05 (BASELINE configs name bf16); bf16x3 (alias tf32x3): tcgen05 at "
                       "the 1e-5 parity tolerance (split operands); fp32: SIMT")
  ap.add_argument('--repeats', type=int, default=0, help='timed K-step blocks (0: >= 10 and >= 100 ms in total)')
  ap.add_argument('--cpu-budget', type=float, default=12.0, help='seconds of CPU-baseline work')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-profile', action='store_true')
  ap.add_argument('--no-extras', action='store_true',
                  help='skip the extra blocks of the default run (parity-mode line, wind roofline block, predict block)')
  args = ap.parse_args()
  if args.precision == 'tf32x3':
    args.precision = 'bf16x3'
  if args.impl == 'reference':
    wl = WORKLOADS[args.workload]
    x, y, margs = synth(wl)
    run_reference(args, wl, x, y, margs)
    return

  import torch
  import torch.distributed as dist

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
      os.environ['NCCL_DEBUG'] = 'WARN'   # keep stdout to the single JSON line
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
  torch.cuda.set_device(local)
  env = {'rank': rank, 'world': world, 'local': local, 'dev': torch.device('cuda', local)}

  kept = {}
  out = run_workload(args, args.workload, args.precision, args.steps, args.warmup, env, do_e2e=True,
                     do_profile=not args.no_profile, cpu_baseline=not args.no_cpu_baseline, repeats=args.repeats,
                     keep=kept)
  # ---- extra blocks of the default run: short, each a complete timed record of its own ----------
  extras = {}
  if not args.no_extras and args.workload == 'chickenpox_map_e8' and args.precision == 'bf16':
    # (1) the SAME workload in the tensor-core mode that meets the 1e-5 parity tolerance
    r = run_workload(args, args.workload, 'bf16x3', 20, 5, env, do_e2e=True, do_profile=not args.no_profile)
    if r:
      extras['parity_mode_bf16x3'] = {k: r[k] for k in ('value', 'unit', 'ms_per_step', 'dtype', 'timing', 'e2e',
                                                        'gpu_launches', 'roofline', 'kernels', 'algorithmic_tflops_per_gpu')}
    # (2) predict on all ranks: forward + one NCCL all-gather + mixture quantiles
    pb = predict_block(env, kept)
    if pb:
      extras['predict'] = pb
    kept.clear()
    torch.cuda.empty_cache()
    # (3) the tensor-bound regime: three steps of the wind shard (BASELINE configs[4] per GPU)
    r = run_workload(args, 'wind_map_e16', 'bf16', 3, 3, env, do_e2e=False, do_profile=not args.no_profile,
                     min_window_ms=0.0)
    if r:
      extras['roofline_wind_map_e16'] = {k: r[k] for k in ('value', 'unit', 'ms_per_step', 'config', 'timing', 'clocks',
                                                           'gpu_launches', 'roofline', 'kernels',
                                                           'algorithmic_tflops_per_gpu')}
  if rank == 0:
    out.update(extras)
    print(json.dumps(out), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
