"""world_size-2 gloo test (CPU) of the only collective on the path: the per-step gather of results."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from vnect_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_results(stream):
    rng = np.random.default_rng(4000 + stream)
    return rng.uniform(0, 368, (21, 2)), rng.uniform(-500, 500, (21, 3)).astype(np.float32)


def _worker(rank, world, port, n_streams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.owned_streams(n_streams, rank, world)
    j2 = np.stack([_fake_results(s)[0] for s in mine])
    j3 = np.stack([_fake_results(s)[1] for s in mine])
    res = parallel.gather_results(parallel.pack_results(j2, j3), n_streams)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_streams):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want2 = np.stack([_fake_results(s)[0] for s in range(n_streams)])
    want3 = np.stack([_fake_results(s)[1] for s in range(n_streams)])
    for _, res in outs:
        g2, g3 = parallel.unpack_results(res)
        assert np.array_equal(g2, want2) and np.array_equal(g3, want3)


def test_gather_even_split():
    _run(2, 8)


def test_gather_ragged_split():
    _run(2, 7)
