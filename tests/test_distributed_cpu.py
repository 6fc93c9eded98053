"""Host-side sharding + the single all-gather, exercised with world_size 2 and 3 on the gloo backend (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntm_mpc import distributed as D


def test_shard_ranges_cover_the_batch_exactly():
    for S in (0, 1, 7, 8, 1024, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_range(S, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in spans) == D.padded_count(S, world) or S == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k_sim = 5
    x0 = np.arange(2 * S, dtype=np.float64).reshape(S, 2)
    params = np.zeros((S, 16))

    def compute(lo, hi):                       # stands in for the CUDA kernel: deterministic function of the scenario id
        ids = torch.arange(lo, hi, dtype=torch.float64)
        xk = ids[:, None, None] + torch.arange(k_sim + 1, dtype=torch.float64)[None, :, None] * 0.5 + torch.tensor([0.0, 0.25])[None, None, :]
        return xk, ids[:, None] * 2 + torch.arange(k_sim, dtype=torch.float64)[None, :], ids * 3

    r = D.closed_loop_sharded(x0, params, N=3, k_sim=k_sim, compute=compute)
    if rank == 0:
        q.put({k: v for k, v in r.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,S", [(2, 10), (2, 7), (3, 8), (2, 1)])
def test_all_gather_reassembles_the_full_batch(world, S):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ids = np.arange(S, dtype=np.float64)
    assert got["xk"].shape == (S, 6, 2) and got["uk"].shape == (S, 5) and got["cost"].shape == (S,)
    assert np.array_equal(got["cost"], ids * 3)
    assert got["status"].shape == (S,) and got["status"].dtype == np.int32 and not got["status"].any()
    assert np.array_equal(got["uk"], ids[:, None] * 2 + np.arange(5)[None, :])
    assert np.array_equal(got["xk"][:, :, 1], ids[:, None] + np.arange(6)[None, :] * 0.5 + 0.25)
