"""Two-GPU checks of the gather paths (skipped on a one-GPU box): the chunked in-place NCCL all-gather and the
all-gather fused into the MFCC kernel's stores over NVLink peer memory must both deliver, on every rank, exactly the
features of the whole batch."""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import diffsptk_b200.functional as F
        from diffsptk_b200.distributed import FusedGatherMfcc, shard_rows, sharded_features
        g = torch.Generator().manual_seed(0)
        x = torch.randn(8, 16000, generator=g)                 # the same global batch on every rank
        want = F.mfcc_from_waveform(x.to(dev))
        res = {}
        for k in (1, 2):
            rows = shard_rows(8, rank, world, k)
            got = sharded_features(lambda t: F.mfcc_from_waveform(t), x[rows].to(dev), n_chunks=k)
            res[f"nccl_chunks{k}"] = bool(torch.equal(got, want))
        fg = FusedGatherMfcc(8 // world, 16000, device=dev)
        if fg.available:
            lo = rank * (8 // world)
            got = fg(x[lo:lo + 8 // world].to(dev))
            torch.cuda.synchronize()
            res["fused"] = bool(torch.equal(got, want))
            res["fused_mode"] = fg.mode
            got = fg(x[lo:lo + 8 // world].to(dev) * 2.0)      # the buffer is reused: a second, different step
            torch.cuda.synchronize()
            res["fused_again"] = bool(torch.equal(got, F.mfcc_from_waveform(2.0 * x.to(dev))))
        else:
            res["fused"] = None
            res["fused_reason"] = fg.reason
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_feature_gathers_on_two_gpus():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out.items():
        assert res["nccl_chunks1"] and res["nccl_chunks2"], (rank, res)
        assert res["fused"] in (True, None), (rank, res)       # None: no peer mapping on this box (reason recorded)
        if res["fused"]:
            assert res["fused_again"], (rank, res)
