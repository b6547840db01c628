"""N>1 host logic on CPU: world_size-2 gloo processes partition a ragged batch and gather the results."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voicecraft_b200 import distributed as vd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = vd.partition(lengths, world, rank)
    local = []
    for i in mine:                      # stand-in for the decode: deterministic tokens per global utterance id
        g = torch.Generator().manual_seed(1000 + i)
        local.append(torch.randint(0, 2048, (4, lengths[i]), generator=g))
    full = vd.gather_token_lists(local, mine, len(lengths))
    ok = all(torch.equal(full[i], torch.randint(0, 2048, (4, lengths[i]), generator=torch.Generator().manual_seed(1000 + i)))
             for i in range(len(lengths)))
    q.put((rank, mine, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_a_balanced_exact_cover():
    lengths = [650, 120, 300, 300, 80, 900, 10, 450, 451]
    for world in (1, 2, 4, 8):
        parts = [vd.partition(lengths, world, r) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) <= sum(lengths) / world + max(lengths)


import pytest


@pytest.mark.parametrize("lengths", [[37, 5, 64, 12, 1, 29, 30], [17]])       # [17]: one utterance, rank 1 holds nothing
def test_world2_gloo_gather(lengths):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, _, ok in res)
    assert sorted(i for _, mine, _ in res for i in mine) == list(range(len(lengths)))
