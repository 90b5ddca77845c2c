"""Multi-GPU execution of the analysis path: shard the utterance axis, gather compact features.

Utterances are independent, so the only communication is one logical
all-gather of the per-rank feature tensors.  One process per GPU (``torch.distributed``, NCCL over
NVLink/NVSwitch; ``gloo`` in the CPU tests).  Two implementations of the gather:

* ``sharded_features`` -- the utterances of a rank are cut into chunks; the features of chunk k are gathered
  with ONE in-place ``all_gather_into_tensor`` (send buffer = this rank's slot of the receive slab, no staging
  copy inside the process group) that overlaps the kernel of chunk k+1.  For the slabs to be contiguous the
  batch is partitioned block-cyclically (``shard_rows``): the gathered tensor is in global utterance order.
* ``FusedGatherMfcc`` -- no collective call at all: the output tensor lives in symmetric memory
  (``torch.distributed._symmetric_memory``), and the epilogue of the fused MFCC kernel stores every finished
  feature row into the output tensor of ALL ranks, through the NVSwitch multicast address (``multimem.st``) when
  the fabric offers one, else through the peers' unicast addresses.  The transfer overlaps the math quad by
  quad; one cross-rank barrier closes the step.
"""

from __future__ import annotations

import ctypes as C
import os
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
    """This rank's utterances of a ``[B, ...]`` batch (contiguous partition)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def shard_rows(n_items: int, rank: int, world_size: int, n_chunks: int = 1) -> torch.Tensor:
    """Global indices of the utterances rank ``rank`` owns under the block-cyclic partition that
    ``sharded_features(..., n_chunks)`` gathers in order: the batch is cut into ``n_chunks`` super-blocks, each
    super-block into ``world_size`` equal slices.  ``n_items`` must be a multiple of ``world_size``.  With
    ``n_chunks == 1`` this is the contiguous partition of ``shard``."""
    if n_items % world_size:
        raise ValueError("the global batch must be a multiple of world_size (pad it)")
    if not 0 <= rank < world_size:
        raise ValueError("invalid rank / world_size")
    local = n_items // world_size
    n_chunks = max(1, min(n_chunks, max(local, 1)))
    rows, off = [], 0
    for k in range(n_chunks):
        lo, hi = shard_bounds(local, k, n_chunks)
        cb = hi - lo
        rows.append(torch.arange(off + rank * cb, off + (rank + 1) * cb))
        off += world_size * cb
    return torch.cat(rows) if rows else torch.empty(0, dtype=torch.long)


def _supports_into_tensor(group) -> bool:
    try:
        return dist.get_backend(group) in ("nccl", "gloo")
    except Exception:
        return False


def sharded_features(fn: Callable[[torch.Tensor], torch.Tensor], x_local: torch.Tensor, *,
                     n_chunks: int = 4, gather: bool = True, group=None, sm_margin: Optional[int] = None) -> torch.Tensor:
    """Apply ``fn`` to this rank's utterances and (optionally) all-gather the features.

    ``x_local`` is ``[B_local, T]`` with the same ``B_local`` on every rank -- rows
    ``shard_rows(B_local * world, rank, world, n_chunks)`` of the global batch.  Returns
    ``[B_local * world, N, D]`` in global utterance order when ``gather`` else the local ``[B_local, N, D]``.

    ``sm_margin`` (default: 8 on CUDA with more than one chunk, else 0): SMs the library's persistent kernels leave
    free while the chunks are in flight, so that the NCCL kernels of chunk k can actually run beside the compute
    kernel of chunk k+1 (one CTA per SM otherwise occupies the whole device until it ends; round 2, N = 8: the
    gather was 1.05 ms exposed behind a 1.37 ms kernel).
    """
    if not gather or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return fn(x_local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Bl = x_local.shape[0]
    n_chunks = max(1, min(n_chunks, Bl))
    edges = [shard_bounds(Bl, k, n_chunks) for k in range(n_chunks)]
    into = _supports_into_tensor(group)
    out, off, pending = None, 0, []
    if sm_margin is None:
        sm_margin = 8 if (x_local.is_cuda and n_chunks > 1) else 0
    prev_margin = None
    if sm_margin and x_local.is_cuda:
        from . import _native
        prev_margin = _native.set_sm_margin(sm_margin)
    try:
        return _gather_chunks(fn, x_local, edges, world, rank, Bl, into, group)
    finally:
        if prev_margin is not None:
            from . import _native
            _native.set_sm_margin(prev_margin)


def _gather_chunks(fn, x_local, edges, world, rank, Bl, into, group):
    out, off, pending = None, 0, []
    for lo, hi in edges:
        cb = hi - lo
        y = fn(x_local[lo:hi])
        if out is None:
            out = torch.empty((world * Bl, *y.shape[1:]), device=y.device, dtype=y.dtype)
        slab = out[off:off + world * cb]              # contiguous [world, cb, N, D] region of the final tensor
        mine = slab[rank * cb:(rank + 1) * cb]        # in place: send buffer == receive buffer + rank * count
        mine.copy_(y)
        if into:
            pending.append(dist.all_gather_into_tensor(slab, mine, group=group, async_op=True))
        else:
            pending.append(dist.all_gather([slab[r * cb:(r + 1) * cb] for r in range(world)], mine.clone(),
                                           group=group, async_op=True))
        off += world * cb
    for work in pending:
        work.wait()
    return out


class FusedGatherMfcc:
    """MFCC of this rank's ``[B_local, T]`` waveforms with the all-gather fused into the kernel's stores.

    ``__call__`` returns the ``[world * B_local, N, D]`` feature tensor (contiguous partition: rank r owns rows
    ``[r * B_local, (r + 1) * B_local)``), identical on every rank.  The tensor is ONE symmetric-memory buffer
    that is reused by every call.  ``available`` is False (with ``reason``) when the process group, the driver
    or the fabric cannot map peer memory; callers then use ``sharded_features``.
    """

    def __init__(self, B_local: int, T: int, *, device, group=None, frame_length: int = 400, frame_period: int = 80,
                 fft_length: int = 512, mfcc_order: int = 13, n_channel: int = 40, sample_rate: int = 16000,
                 out_format: str | int = "y", mode: Optional[str] = None):
        from . import ops
        from .modules.mfcc import mfcc_format_id
        self.available, self.reason, self.mode = False, "", ""
        self.kw = dict(frame_length=frame_length, frame_period=frame_period, fft_length=fft_length,
                       mfcc_order=mfcc_order, n_channel=n_channel, sample_rate=sample_rate, out_format=out_format)
        if not (dist.is_available() and dist.is_initialized()):
            self.reason = "torch.distributed is not initialised"
            return
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > ops.MAX_GATHER_PEERS:
            self.reason = f"more than {ops.MAX_GATHER_PEERS} ranks"
            return
        self.Bl, self.T = B_local, T
        self.N = ops.num_frames(T, frame_period)
        self.D = ops._mfcc_dim(mfcc_order, mfcc_format_id(out_format))
        try:
            import torch.distributed._symmetric_memory as symm
            self.out = symm.empty((self.world * B_local, self.N, self.D), dtype=torch.float32, device=device)
            self.hdl = symm.rendezvous(self.out, self.group)
            self.peers = [int(p) for p in self.hdl.buffer_ptrs]
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)   # 0: no NVSwitch multicast object for this buffer
        except Exception as e:  # no peer access / no symmetric memory in this build
            self.reason = f"symmetric memory unavailable: {e!r}"[:200]
            return
        mode = mode or os.environ.get("DSB200_GATHER_MODE", "auto")
        self.mc_ptr = mc if mode in ("auto", "multicast") else 0
        if mode == "multicast" and not self.mc_ptr:
            self.reason = "no multicast address for this buffer"
            return
        self.mode = ("NVSwitch multicast stores, multimem.st" if self.mc_ptr
                     else f"unicast stores to the {self.world} peer mappings over NVLink")
        self.available = True

    def __call__(self, x_local: torch.Tensor) -> torch.Tensor:
        from . import fused
        if not self.available:
            raise RuntimeError("FusedGatherMfcc is not available: " + self.reason)
        if tuple(x_local.shape) != (self.Bl, self.T):
            raise ValueError(f"x_local must be [{self.Bl}, {self.T}]")
        self.hdl.barrier(channel=0)    # every rank is done reading the previous contents of its buffer
        fused.mfcc_from_waveform_gather(x_local, self.out, self.peers, self.mc_ptr, self.rank, **self.kw)
        self.hdl.barrier(channel=1)    # every rank's kernel has finished: all slabs have landed everywhere
        return self.out
