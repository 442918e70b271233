"""world_size-2 gloo test of the N>1 host logic (sharding + final summary gather) on the CPU."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from beamfe2_b200 import dist as bdist


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 100000, 10 ** 7):
        for world in (1, 2, 3, 4, 8):
            r = [bdist.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    r, _, w = bdist.init_from_env(backend='gloo')
    lo, hi = bdist.shard_range(total, r, w)
    g = torch.arange(lo, hi, dtype=torch.float64)
    info = torch.zeros(hi - lo, dtype=torch.int32)
    u = g.reshape(-1, 1, 1).expand(-1, 2, 3).contiguous()          # max |u| == global index
    lam = (g.reshape(-1, 1) + torch.arange(1, 5, dtype=torch.float64)) ** 2
    table = bdist.gather_summaries(bdist.summarize(u, lam, info), total)
    t = bdist.max_over_ranks(float(rank + 1), 'cpu')
    bdist.barrier()
    q.put((rank, table.numpy(), t))
    torch.distributed.destroy_process_group()


def test_two_rank_gather_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    total, world = 11, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, table, t in res:
        assert t == 2.0
        assert table.shape == (total, 5)
        assert np.array_equal(table[:, 1], np.arange(total))       # global batch order restored
        assert np.allclose(table[:, 2], np.arange(total) + 1)
