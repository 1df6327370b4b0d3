"""torchrun --nproc-per-node N scripts/multi_gpu_check.py
Member sharding + the single NCCL all-gather at predict:
every rank must return identical (world, E/world, N) means and quantiles, and rank
r's slice must equal what a 1-process fit of the same members produces."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesnf_b200 import inference, models  # noqa: E402


def main():
  rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
  local = int(os.environ.get('LOCAL_RANK', rank))
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  n, E = 600, 2 * world
  rng = np.random.default_rng(0)
  x = np.stack([np.arange(n, dtype=float), rng.normal(size=n), rng.normal(size=n)], 1)
  y = np.sin(np.arange(n) / 9.0) * 4 + rng.normal(size=n)
  margs = dict(width=128, depth=2, input_scales=np.array([n - 1.0, 1, 1]),
               num_seasonal_harmonics=np.array([2, 3]), seasonality_periods=np.array([7.0, 30.0]),
               init_x=(n, 3), fourier_degrees=np.array([3, 2, 2]), interactions=np.zeros((0, 2), int))
  spec = models.ModelSpec(**margs)
  init_all = np.random.default_rng(1).normal(size=(E, spec.num_params)).astype(np.float32) * 0.5
  per = E // world
  mine = init_all[rank * per:(rank + 1) * per]
  params, losses = inference.fit_map(x, y, 0, 'NORMAL', margs, num_particles=E, learning_rate=0.01,
                                     num_epochs=10, precision='fp32', init_params=mine)
  assert params[0].shape == (1, per) and losses.shape == (1, per, 10)
  means, qs = inference.predict_bnf(x[:200], 'NORMAL', params, margs, (0.5, 0.1), precision='fp32')
  assert means.shape == (world, per, 200), means.shape
  # every rank holds the same gathered result
  t = torch.tensor(means, device='cuda')
  ref = t.clone()
  dist.broadcast(ref, 0)
  assert torch.equal(t, ref)
  tq = torch.tensor(np.stack(qs), device='cuda')
  refq = tq.clone()
  dist.broadcast(refq, 0)
  assert torch.equal(tq, refq)
  if rank == 0:
    # recompute every member in ONE process (no process group semantics): slices must match
    os.environ['BNF_NO_GRAPH'] = '1'
    from bayesnf_b200 import parallel
    saved = (parallel.device_count, parallel.device_index)
    parallel.device_count, parallel.device_index = (lambda: 1), (lambda: 0)
    p_all, _ = inference.fit_map(x, y, 0, 'NORMAL', margs, num_particles=E, learning_rate=0.01,
                                 num_epochs=10, precision='fp32', init_params=init_all)
    parallel.all_gather_leading = lambda t_: t_[None]
    m_all, q_all = inference.predict_bnf(x[:200], 'NORMAL', p_all, margs, (0.5, 0.1), precision='fp32')
    np.testing.assert_allclose(m_all.reshape(world, per, 200), means, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(np.stack(q_all), np.stack(qs), rtol=2e-3, atol=2e-3)
    print(f'multi_gpu_check ok: world={world} means {means.shape} quantile[0][:3]={qs[0][:3]}')
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
