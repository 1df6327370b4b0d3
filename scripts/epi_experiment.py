"""Epilogue ablation at the wind shard shape (needs a lib built with -DBNF_TC_EXPERIMENT).

For each BNF_TC_DBG mask run a few MAP steps under the per-kernel CUDA-event profile and print the
fwd / dgrad / wgrad times.  Results are numerically WRONG for non-zero masks - timing only.
bits: 1 skip z loads (dgrad), 2 skip bias butterfly+atomics (dgrad), 4 skip TMA stores (fwd+dgrad),
      8 skip activation math (fwd+dgrad).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bayesnf_b200 import _lib, inference, models  # noqa: E402


def main():
  wl = bench.WORKLOADS[os.environ.get('WL', 'wind_map_e16')]
  x, y, margs = bench.synth(wl)
  dev = torch.device('cuda', 0)
  spec = models.ModelSpec(**margs, observation_model='NORMAL')
  eng = inference.Engine(spec, os.environ.get('PREC', 'bf16'))
  E, n_total = wl['members_per_gpu'], len(y)
  B = wl['batch'] or n_total
  xd, yd = inference._to_device_data(x, y)
  p = eng.init_params(float(np.log(np.nanstd(y) / 2)), 1234, 0, E)
  m, v = torch.zeros_like(p), torch.zeros_like(p)
  sc = torch.zeros(1, dtype=torch.int32, device=dev)
  idx = None
  if B < n_total:      # fixed index rows (this script times one step shape): the device's own permutations
    idx = torch.tensor(np.stack([inference.device_permutation(0, e, 0, n_total)[:B] for e in range(E)]),
                       device=dev).contiguous()

  def run(k):
    for _ in range(k):
      eng.map_steps(p, m, v, sc, xd, yd, idx, B, n_total, 1, 0.0, 1.0)
    torch.cuda.synchronize()

  masks = [int(t) for t in os.environ.get('MASKS', '0,1,2,4,8,3,7,15,0').split(',')]
  run(2)
  for mask in masks:
    os.environ['BNF_TC_DBG'] = str(mask)
    run(1)
    _lib.check(_lib.lib.bnf_debug_profile(1))
    k = 3
    run(k)
    buf = C.create_string_buffer(1 << 16)
    _lib.check(_lib.lib.bnf_debug_profile_report(buf, len(buf)))
    _lib.check(_lib.lib.bnf_debug_profile(0))
    d = {}
    for line in buf.value.decode().strip().splitlines():
      name, cnt, tot = line.split()
      d[name] = float(tot) / k
    print('%s mask %2d  fwd %.4f  dgrad %.4f  wgrad %.4f  total %.4f ms/step' % (
        os.environ.get('TAG', ''), mask, d.get('tc_gemm_fwd', 0), d.get('tc_gemm_dgrad', 0),
        d.get('tc_gemm_wgrad', 0), sum(d.values())), flush=True)
    if os.environ.get('ALL_KERNELS'):
      print('   ', '  '.join('%s %.4f' % kv for kv in sorted(d.items())), flush=True)


if __name__ == '__main__':
  main()
