"""world_size-2 gloo test of the sharding / chunked all-gather logic (runs on CPU)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsptk_b200.distributed import shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 9, 8192):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_shard_rows_is_a_partition_in_gather_order():
    from diffsptk_b200.distributed import shard_rows
    for n, w, k in ((8, 2, 1), (8, 2, 2), (24, 4, 3), (6, 2, 3), (8192, 8, 4)):
        rows = [shard_rows(n, r, w, k) for r in range(w)]
        assert sorted(torch.cat(rows).tolist()) == list(range(n))
        assert all(len(x) == n // w for x in rows)
        if k == 1:
            assert all(rows[r].tolist() == list(range(r * n // w, (r + 1) * n // w)) for r in range(w))
    with pytest.raises(ValueError):
        shard_rows(7, 0, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffsptk_b200.distributed import shard, shard_rows, sharded_features
        from oracle import np_oracle as O

        g = torch.Generator().manual_seed(0)
        x = torch.randn(6, 1600, generator=g, dtype=torch.float64)  # same global batch on every rank

        def feat(xl):  # CPU stand-in for the CUDA pipeline (the helper is device-agnostic)
            P = O.stft(xl.numpy(), frame_length=400, frame_period=80, fft_length=512)
            return torch.from_numpy(O.mfcc(P, 13, 40, 16000))

        want = feat(x)
        xl = shard(x)
        assert xl.shape[0] == 3
        local = sharded_features(feat, xl, gather=False)
        ok = bool(torch.equal(local, want[rank * 3:(rank + 1) * 3]))
        # one chunk: the contiguous partition, one in-place all_gather_into_tensor
        ok = ok and bool(torch.equal(sharded_features(feat, xl, n_chunks=1), want))
        # chunked gathers overlap compute; their slabs are contiguous under the block-cyclic partition
        for n_chunks in (2, 3):
            rows = shard_rows(6, rank, world, n_chunks)
            full = sharded_features(feat, x[rows], n_chunks=n_chunks)
            ok = ok and bool(torch.equal(full, want))
        q.put((rank, ok, tuple(full.shape)))
    finally:
        dist.destroy_process_group()


def test_chunked_all_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (6, 20, 13)), (1, True, (6, 20, 13))]
