"""Multi-GPU plumbing: one process per GPU (torchrun), instances sharded by contiguous row ranges, ONE collective
per solve (all_gather of the packed solutions).  Instances are independent, so there is no
data-path collective and nothing to fuse with one."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
  """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
  return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous, balanced row range [lo, hi) of `total` instances owned by `rank`."""
  base, rem = divmod(total, world_size)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def pack_solution(z: torch.Tensor, lam: torch.Tensor, obj: torch.Tensor, cost: torch.Tensor, status: torch.Tensor,
                  iters: torch.Tensor) -> torch.Tensor:
  """[B, nvars + ncon + 4] fp64: z*, lambda, solver objective, re-integrated cost, status, iterations."""
  return torch.cat([z, lam, obj[:, None], cost[:, None], status.double()[:, None], iters.double()[:, None]], dim=1).contiguous()


def gather_solutions(packed: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
  """The single collective of the path: all ranks end up with every instance's packed solution, in global row order.
  Shards may differ in size by one row (shard_range): equal shards take one all_gather_into_tensor, unequal ones are
  padded to the largest shard and trimmed.  No-op without an initialised process group."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return packed
  w = dist.get_world_size()
  rows = torch.tensor([packed.shape[0]], dtype=torch.int64, device=packed.device)
  if out is not None and out.shape[0] == packed.shape[0] * w:
    dist.all_gather_into_tensor(out, packed)   # caller vouches for equal shards (the benchmark's weak-scaling layout)
    return out
  all_rows = torch.empty(w, dtype=torch.int64, device=packed.device)
  dist.all_gather_into_tensor(all_rows, rows)
  counts = [int(v) for v in all_rows.cpu()]
  mx = max(counts)
  if all(c == mx for c in counts):
    res = torch.empty(mx * w, packed.shape[1], dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(res, packed)
    return res
  pad = torch.zeros(mx, packed.shape[1], dtype=packed.dtype, device=packed.device)
  pad[:packed.shape[0]] = packed
  buf = torch.empty(mx * w, packed.shape[1], dtype=packed.dtype, device=packed.device)
  dist.all_gather_into_tensor(buf, pad)
  return torch.cat([buf[r * mx: r * mx + c] for r, c in enumerate(counts)], dim=0)
