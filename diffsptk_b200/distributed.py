"""Multi-GPU execution of the analysis path: shard the utterance axis, gather compact features.

Utterances are independent, so the only communication is one logical
all-gather of the per-rank feature tensors.  One process per GPU (``torch.distributed``, NCCL over
NVLink/NVSwitch; ``gloo`` in the CPU tests).  The gather is issued per utterance chunk with
``async_op=True`` so that the collective of chunk k overlaps the kernels of chunk k+1; every rank ends
up with the full ``[B, N, D]`` tensor laid out in global utterance order.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous balanced partition: the first ``n_items % world_size`` ranks get one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("invalid rank / world_size")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(x: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """This rank's utterances of a ``[B, ...]`` batch."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def sharded_features(fn: Callable[[torch.Tensor], torch.Tensor], x_local: torch.Tensor, *,
                     n_chunks: int = 4, gather: bool = True, group=None) -> torch.Tensor:
    """Apply ``fn`` to this rank's utterances and (optionally) all-gather the features.

    ``x_local`` is ``[B_local, T]`` with the same ``B_local`` on every rank (pad the batch if the
    global batch does not divide evenly).  Returns ``[B_local * world, N, D]`` in global utterance
    order when ``gather`` else the local ``[B_local, N, D]``.
    """
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return fn(x_local)
    world = dist.get_world_size(group)
    Bl = x_local.shape[0]
    n_chunks = max(1, min(n_chunks, Bl))
    edges = [shard_bounds(Bl, k, n_chunks) for k in range(n_chunks)]
    out = None
    pending = []
    for lo, hi in edges:
        y = fn(x_local[lo:hi])
        if out is None:
            out = torch.empty((world, Bl, *y.shape[1:]), device=y.device, dtype=y.dtype)
        views = [out[r, lo:hi] for r in range(world)]  # contiguous slabs of the final tensor
        pending.append((dist.all_gather(views, y.contiguous(), group=group, async_op=True), y))
    for work, _ in pending:
        work.wait()
    return out.reshape(world * Bl, *out.shape[2:])
