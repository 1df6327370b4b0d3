"""Ensemble-member sharding: one process per GPU, members are independent.

The reference's only parallelism is ``pmap`` over devices x ``vmap`` over
members with zero collectives (inference.py:577-578, :727); per-device count is
``ensemble_size // jax.device_count()`` (floor, :365, :445).  Here a "device" is
a ``torch.distributed`` rank (launched with torchrun, one per B200).  Training
needs no communication; ``predict`` does ONE all-gather of the per-member
predictive parameters (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def device_count() -> int:
  """World size (the reference's ``jax.device_count()``)."""
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def device_index() -> int:
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def members_per_device(total: int) -> int:
  """Floor division exactly as inference.py:365 / :445."""
  return total // device_count()


def all_gather_leading(t: torch.Tensor) -> torch.Tensor:
  """Stack ``t`` from every rank on a new leading axis -> (world, *t.shape).

  Single collective; with one process this is just ``t[None]``.
  """
  world = device_count()
  if world == 1:
    return t[None]
  t = t.contiguous()
  out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
  dist.all_gather_into_tensor(out.view(-1), t.view(-1))
  return out
