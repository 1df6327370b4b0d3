import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from oracle import bnf_oracle as O
from bayesnf_b200 import inference, models
from test_gpu_parity import _cfgs, _data, _random_params
cfg = _cfgs()['small']
n, B, epochs = 200, 64, 2
x, y = _data(cfg, n)
om = O.OracleModel(**cfg)
P0 = _random_params(om, 2, y, seed=11)
rng = np.random.default_rng(9)
perms = np.stack([[rng.permutation(n) for _ in range(2)] for _ in range(epochs)]).astype(np.int32)
spec = models.ModelSpec(**cfg)
xd, yd = inference._to_device_data(x, y)
res = {}
for mode in ('fused', 'legacy'):
  if mode == 'legacy': os.environ['BNF_LEGACY_STEP'] = '1'
  for pw in (1.0, 0.0):
    params, losses = inference.fit_map(x, y, 0, 'NORMAL', cfg, num_particles=2, learning_rate=0.01,
                                     num_epochs=epochs, prior_weight=pw, batch_size=B,
                                     precision='fp32', init_params=P0.numpy(), batch_indices=perms)
    flat = spec.flatten(params)[0]
    res[(mode, pw)] = flat
    for j in range(2):
      pj, lj = O.fit_map_member(om, P0[j], xd.cpu(), yd.cpu(),
                              lambda ep: torch.tensor(perms[ep, j].astype(np.int64)), epochs, B, 0.01, pw, 'NORMAL')
      d = np.abs(flat[j] - pj.numpy())
      k = np.argsort(d)[-3:]
      print(mode, pw, j, 'max', d.max(), 'idx', k, d[k], 'loss rel', np.abs(losses[0, j] - lj.numpy()).max() / np.abs(lj.numpy()).max())
for pw in (1.0, 0.0):
  print('fused vs legacy', pw, np.abs(res[('fused', pw)] - res[('legacy', pw)]).max())
