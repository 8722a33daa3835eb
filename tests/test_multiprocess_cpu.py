"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — contiguous env sharding by
global index and the statistics all-reduce (the engine's only collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import spacegym_b200 as sg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sg.shard_range(total, rank, world)
    # every rank "owns" its env range: fake per-env episode statistics keyed by the GLOBAL env index
    idx = torch.arange(start, start + count, dtype=torch.float64)
    local = torch.stack([(idx * 0.5).sum(), (idx % 7).sum(), torch.tensor(float(count)), (idx % 3 == 0).double().sum()])
    red = sg.allreduce_stats(local)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([start, count]))
    if rank == 0:
        out.put((red.tolist(), [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_and_stats_allreduce_gloo():
    total, world = 1000003, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    red, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards are contiguous, disjoint and cover [0, total)
    assert shards[0][0] == 0 and shards[0][0] + shards[0][1] == shards[1][0] and shards[1][0] + shards[1][1] == total
    assert abs(shards[0][1] - shards[1][1]) <= 1
    idx = torch.arange(total, dtype=torch.float64)
    want = [(idx * 0.5).sum().item(), (idx % 7).sum().item(), float(total), (idx % 3 == 0).double().sum().item()]
    assert all(abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(red, want))


def test_shard_range_properties():
    for total in (1, 7, 262144, 1000003):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                s, c = sg.shard_range(total, r, world)
                assert s == pos and c >= 0
                pos += c
            assert pos == total
