"""Batch-sharded multi-GPU generation: one process per GPU, weights replicated,
images split contiguously across ranks, NO collective on the data path — sequences are
independent — and one all-gather of the generated ids at the end (NCCL over NVLink on GPUs;
gloo in the CPU tests).

The only cross-row coupling in the reference is the row-0 `</svg>` stop (D6), which refers to
GLOBAL row 0.  `merge_generated` reproduces the single-process rectangle from per-rank results:
only rank 0 arms the row-0 stop; every rank returns its own rectangle; the global length is the
row-0 stop step if it fired, else the longest rank; shorter ranks are padded (their rows had
finished, HF would emit pad), longer ones truncated (rows are independent, so the prefix is
identical to what a single process would have produced).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split; the first `global_batch % world` ranks take one extra image."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row0_stop_fired(ids_rank0: torch.Tensor, stop_ids: Sequence[int]) -> bool:
    n = len(stop_ids)
    return n > 0 and ids_rank0.shape[1] >= n and ids_rank0[0, -n:].tolist() == list(stop_ids)


def merge_generated(per_rank: List[torch.Tensor], stop_ids: Sequence[int], pad_token_id: int) -> torch.Tensor:
    """per_rank[r]: int tensor [B_r, n_r] of new tokens from rank r (rank 0 ran with the row-0 stop armed)."""
    if row0_stop_fired(per_rank[0], stop_ids):
        n = per_rank[0].shape[1]
    else:
        n = max(t.shape[1] for t in per_rank)
    rows = []
    for t in per_rank:
        if t.shape[1] >= n:
            rows.append(t[:, :n])
        else:
            pad = torch.full((t.shape[0], n - t.shape[1]), pad_token_id, dtype=t.dtype, device=t.device)
            rows.append(torch.cat([t, pad], dim=1))
    return torch.cat(rows, dim=0)


def all_gather_generated(local_ids: torch.Tensor, max_new: int, stop_ids: Sequence[int], pad_token_id: int,
                         global_batch: int) -> torch.Tensor:
    """Collective: every rank contributes [B_local, n_local]; every rank gets the global [B, n] rectangle."""
    world, rank = dist.get_world_size(), dist.get_rank()
    b_max = max(shard_range(global_batch, r, world)[1] - shard_range(global_batch, r, world)[0] for r in range(world))
    buf = torch.full((b_max, max_new + 1), pad_token_id, dtype=torch.int32, device=local_ids.device)
    buf[: local_ids.shape[0], : local_ids.shape[1]] = local_ids.to(torch.int32)
    buf[:, max_new] = local_ids.shape[1]                       # last column carries n_local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    per_rank = []
    for r, t in enumerate(out):
        lo, hi = shard_range(global_batch, r, world)
        n_r = int(t[0, max_new].item())
        per_rank.append(t[: hi - lo, :n_r])
    return merge_generated(per_rank, stop_ids, pad_token_id)
