"""World-size-2 gloo tests of the member-sharding helpers (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from bayesnf_b200 import parallel
  assert parallel.device_count() == world and parallel.device_index() == rank
  assert parallel.members_per_device(7) == 3           # floor, inference.py:445
  local = torch.full((3, 5), float(rank)) + torch.arange(5.0)
  gathered = parallel.all_gather_leading(local)         # (world, 3, 5)
  scal = parallel.all_gather_leading(torch.tensor([10.0 * rank, 10.0 * rank + 1]))
  np.save(os.path.join(out_dir, f'g{rank}.npy'), gathered.numpy())
  np.save(os.path.join(out_dir, f's{rank}.npy'), scal.numpy())
  dist.destroy_process_group()


def test_all_gather_and_sharding(tmp_path):
  world = 2
  port = _free_port()
  mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  g0, g1 = np.load(tmp_path / 'g0.npy'), np.load(tmp_path / 'g1.npy')
  np.testing.assert_array_equal(g0, g1)
  assert g0.shape == (2, 3, 5)
  np.testing.assert_array_equal(g0[1] - g0[0], np.ones((3, 5)))
  s0 = np.load(tmp_path / 's0.npy')
  np.testing.assert_array_equal(s0, [[0, 1], [10, 11]])


def test_single_process_defaults():
  from bayesnf_b200 import parallel
  assert parallel.device_count() == 1 and parallel.device_index() == 0
  assert parallel.members_per_device(16) == 16
  t = torch.arange(6.0).reshape(2, 3)
  assert parallel.all_gather_leading(t).shape == (1, 2, 3)


def test_ensemble_smaller_than_world_is_value_error(monkeypatch):
  from bayesnf_b200 import parallel, spatiotemporal
  monkeypatch.setattr(parallel, 'device_count', lambda: 8)
  est = spatiotemporal.BayesianNeuralFieldMAP(feature_cols=['t'], target_col='x', freq='D')
  with pytest.raises(ValueError):
    est.fit(None, 0, ensemble_size=4)
