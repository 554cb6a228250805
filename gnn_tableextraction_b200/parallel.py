"""Data parallelism by graph for the train step (pages are independent components of the
batched graph, /root/reference/src/models/model_train.py:297 -- no edge crosses pages, so no
feature exchange is ever needed; the reference itself is single device).

Each rank takes a contiguous slice of the global batch's pages, runs the same kernels on its
own batched graph, and the only exchange per step is ONE all-reduce (SUM) of the flat fp32 gradient
buffer (424 KB for the default model) whose tail carries [sum w*nll, sum w, #correct]: every rank
back-propagates the un-normalised local loss sum and the optimiser divides the summed gradient by the
GLOBAL label-weight sum, so the result equals the single-GPU gradients (not a mean of per-rank means)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of ``num_items`` for ``rank`` (first ``num_items % world`` ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(num_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_pages_balanced(page_sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy balance by node count for ragged pages: heaviest page to the lightest rank.
    Returns the page indices of every rank (each list ascending, so per-rank batches keep list order)."""
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for idx in sorted(range(len(page_sizes)), key=lambda i: (-page_sizes[i], i)):
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(idx)
        loads[r] += page_sizes[idx]
    return [sorted(v) for v in out]


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=group)
    return t
