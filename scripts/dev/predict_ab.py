"""A/B of the forward-only (predict) path: fused encode + Dense_0 vs encode kernel + GEMM."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from bayesnf_b200 import inference, models
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
for wlname in ('chickenpox_map_e8', 'air_quality_map_e8'):
  wl = bench.WORKLOADS[wlname]
  x, y, margs = bench.synth(wl)
  spec = models.ModelSpec(**margs, observation_model='NORMAL')
  for prec in ('bf16', 'bf16x3'):
    eng = inference.Engine(spec, prec)
    E = wl['members_per_gpu']
    p = eng.init_params(1.0, 1, 0, E)
    g = torch.Generator(device=dev).manual_seed(7)
    n = 1 << 20
    xt = torch.stack([torch.rand(n, generator=g, device=dev) * 600.0, torch.randn(n, generator=g, device=dev),
                      torch.randn(n, generator=g, device=dev)], 1).contiguous()
    for flag in ('0', '1', '0', '1'):
      os.environ['BNF_NO_FUSED_ENCODE'] = flag
      eng.forward(p, xt)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(3):
        out = eng.forward(p, xt)
      e1.record()
      torch.cuda.synchronize()
      print(wlname, prec, 'NO_FUSED_ENCODE=' + flag, 'forward ms', round(e0.elapsed_time(e1) / 3, 3), 'slab', eng.forward_slab_rows(E), flush=True)
