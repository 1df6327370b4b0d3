#!/usr/bin/env python
"""How does tcgen05 accumulate in TMEM?  Measures (through bnf_debug_gemm, bf16 operands):

1. pure accumulation error: exactly-representable positive bf16 inputs, result vs f64
   (signed mean relative error < 0 and growing ~K means round-toward-zero accumulation);
2. the split-operand emulation of an f32 GEMM: a = a0 + a1 + a2 (bf16 each), six products
   concatenated along K into ONE accumulator, and the two-accumulator variant (a0.b0 alone,
   the five small products together), both against the f64 product of the f32 inputs and
   against cuBLAS f32.
Printed numbers decide the design of the bf16x3 precision mode.
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bayesnf_b200 import _lib  # noqa: E402


def gemm(a, b):
  nets, m, k = a.shape
  n = b.shape[1]
  c = torch.empty((nets, m, n), dtype=torch.float32, device=a.device)
  _lib.check(_lib.lib.bnf_debug_gemm(0, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                     C.c_void_p(c.data_ptr()), nets, m, n, k,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
  torch.cuda.synchronize()
  return c


def split3(x):
  x0 = x.to(torch.bfloat16)
  r1 = x - x0.float()
  x1 = r1.to(torch.bfloat16)
  r2 = r1 - x1.float()
  x2 = r2.to(torch.bfloat16)
  return x0, x1, x2


def stats(tag, got, want):
  rel = (got.double() - want) / want.abs().clamp_min(1e-300)
  scale = want.abs().max()
  print(f'{tag:46s} mean_rel {float(rel.mean()):+.3e}  rms_rel {float(rel.pow(2).mean().sqrt()):.3e}  '
        f'max_abs/scale {float((got.double() - want).abs().max() / scale):.3e}', flush=True)


def main():
  torch.cuda.set_device(0)
  dev = torch.device('cuda', 0)
  g = torch.Generator(device=dev).manual_seed(1)
  os.environ['BNF_CTA2'] = '0'
  print('--- 1. accumulation of exact products (positive inputs)')
  for k in (64, 256, 1024, 4096):
    a = (torch.rand(1, 256, k, generator=g, device=dev) + 0.5).to(torch.bfloat16)
    b = (torch.rand(1, 256, k, generator=g, device=dev) + 0.5).to(torch.bfloat16)
    want = torch.bmm(a.double(), b.double().transpose(1, 2))
    stats(f'tcgen05 K={k}', gemm(a, b), want)
    stats(f'cuBLAS f32 K={k}', torch.bmm(a.float(), b.float().transpose(1, 2)), want)
  print('--- 1b. same with zero-mean inputs')
  for k in (256, 1024, 4096):
    a = torch.randn(1, 256, k, generator=g, device=dev).to(torch.bfloat16)
    b = torch.randn(1, 256, k, generator=g, device=dev).to(torch.bfloat16)
    want = torch.bmm(a.double(), b.double().transpose(1, 2))
    got = gemm(a, b)
    err = (got.double() - want)
    print(f'K={k}: rms err / rms value {float(err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()):.3e}; '
          f'corr(err, value) {float((err * want).mean() / (err.pow(2).mean().sqrt() * want.pow(2).mean().sqrt())):+.3f}')
  print('--- 2. split-operand f32 GEMM emulation')
  torch.backends.cuda.matmul.allow_tf32 = False
  for k, sign in ((64, 'randn'), (256, 'randn'), (256, 'pos'), (1024, 'randn'), (1024, 'pos')):
    if sign == 'pos':
      a = torch.rand(1, 256, k, generator=g, device=dev) + 0.5
      b = torch.rand(1, 256, k, generator=g, device=dev) + 0.5
    else:
      a = torch.randn(1, 256, k, generator=g, device=dev)
      b = torch.randn(1, 256, k, generator=g, device=dev)
    want = torch.bmm(a.double(), b.double().transpose(1, 2))
    a0, a1, a2 = split3(a)
    b0, b1, b2 = split3(b)
    # small products first
    A6 = torch.cat([a2, a0, a1, a1, a0, a0], dim=2).contiguous()
    B6 = torch.cat([b0, b2, b1, b0, b1, b0], dim=2).contiguous()
    A5 = torch.cat([a2, a0, a1, a1, a0], dim=2).contiguous()
    B5 = torch.cat([b0, b2, b1, b0, b1], dim=2).contiguous()
    A3 = torch.cat([a1, a0, a0], dim=2).contiguous()
    B3 = torch.cat([b0, b1, b0], dim=2).contiguous()
    tag = f'K={k} {sign}'
    print(tag)
    stats('  cuBLAS f32', torch.bmm(a, b.transpose(1, 2)), want)
    stats('  x3, 6 products, one accumulator', gemm(A6, B6), want)
    stats('  x3, 6 products, two accumulators', gemm(a0.contiguous(), b0.contiguous()) + gemm(A5, B5), want)
    stats('  x2, 3 products, one accumulator', gemm(A3, B3), want)
    stats('  bf16 x1', gemm(a0.contiguous(), b0.contiguous()), want)


if __name__ == '__main__':
  main()
