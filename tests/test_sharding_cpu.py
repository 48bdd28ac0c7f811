"""Host-side logic of the multi-GPU path, on CPU with gloo (world_size 2).

The data path of the sharded mode needs GPUs; what is checked here is what does not: every rank derives the same
disjoint, covering name-rank stripes from kai_shard_range, and the peer-handle table travels through torch.distributed
the way bench.py does it (rank 0 exports, broadcast, every rank receives the same bytes).
"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from kai_scheduler_b200 import abi, engine


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_nodes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, count = engine.shard_range(n_nodes, world, rank)
    handles = [b"/kai_b200_test_segment".ljust(abi.PEER_HANDLE_BYTES, b"\0") if rank == 0 else b""]
    dist.broadcast_object_list(handles, src=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, base, count, handles[0]))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_nodes", [10_000, 50_001, 3])
def test_shard_ranges_agree_across_ranks(n_nodes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_nodes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    # name-rank stripes: shard r owns ranks r, r + world, ...: disjoint and covering
    assert got[0][1] == 0 and got[1][1] == 1
    assert got[0][2] + got[1][2] == n_nodes
    assert got[0][2] == (n_nodes + 1) // 2
    # both ranks hold rank 0's handle
    assert got[0][3] == got[1][3] and got[0][3].startswith(b"/kai_b200_test_segment")


def test_shard_range_many():
    for n in (0, 1, 7, 148, 50_000):
        for s in (1, 2, 4, 8):
            owned = np.zeros(n, dtype=np.int32)
            for r in range(s):
                b, c = engine.shard_range(n, s, r)
                assert b == r and c >= 0
                owned[b::s][:c] += 1
                assert len(owned[b::s]) == c
                mask = engine.shard_node_mask(np.arange(n), s, r)
                assert int(mask.sum()) == c
            assert (owned == 1).all()
