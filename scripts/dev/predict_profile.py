"""Forward-only (predict) path at the bench's predict shape (1 Mi rows x 8 members, chickenpox model):
wall / device time of Engine.forward and the per-kernel CUDA-event profile of the library."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from bayesnf_b200 import _lib, inference, models
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
wl = bench.WORKLOADS['chickenpox_map_e8']
x, y, margs = bench.synth(wl)
spec = models.ModelSpec(**margs, observation_model='NORMAL')
tag = os.environ.get('TAG', '')
for prec in os.environ.get('PRECS', 'bf16,bf16x3').split(','):
  eng = inference.Engine(spec, prec)
  E = wl['members_per_gpu']
  p = eng.init_params(1.0, 1, 0, E)
  g = torch.Generator(device=dev).manual_seed(7)
  n = 1 << 20
  xt = torch.stack([torch.rand(n, generator=g, device=dev) * 600.0, torch.randn(n, generator=g, device=dev),
                    torch.randn(n, generator=g, device=dev)], 1).contiguous()
  eng.forward(p, xt)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  for _ in range(3):
    out = eng.forward(p, xt)
  e1.record()
  t1 = time.perf_counter()                      # host time to ENQUEUE the three forwards
  torch.cuda.synchronize()
  print(tag, prec, 'forward ms (device)', round(e0.elapsed_time(e1) / 3, 3), 'host enqueue ms', round((t1 - t0) / 3 * 1e3, 3),
        'slab', eng.forward_slab_rows(E), flush=True)
  _lib.check(_lib.lib.bnf_debug_profile(1))
  eng.forward(p, xt)
  torch.cuda.synchronize()
  buf = C.create_string_buffer(1 << 16)
  _lib.check(_lib.lib.bnf_debug_profile_report(buf, len(buf)))
  _lib.check(_lib.lib.bnf_debug_profile(0))
  tot = 0.0
  for line in buf.value.decode().strip().splitlines():
    name, cnt, ms = line.split()
    tot += float(ms)
    print('   ', tag, prec, name, cnt, 'launches', round(float(ms), 3), 'ms', flush=True)
  print('   ', tag, prec, 'sum of kernels', round(tot, 3), 'ms', flush=True)
