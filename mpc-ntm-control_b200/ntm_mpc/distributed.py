"""Scenario sharding across the GPUs of one box: one process per GPU (torchrun), contiguous blocks of
ceil(S/G) scenarios per rank, no communication inside the time loop, and ONE all-gather of the
per-scenario outputs at the end: the kernel writes one packed record
[xk | uk | cost | status] of 3*k_sim + 4 doubles per scenario (``ntm_mpc_closed_loop_rec_dev``), so a single
``all_gather_into_tensor`` call moves everything.  Scenarios never interact
(NTM_MPC_Sim.m:93-131 has no cross-scenario term), so the gathered result is bit-identical to a
single-GPU run of the whole batch.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np


def shard_range(S: int, world: int, rank: int) -> Tuple[int, int]:
    """Rank ``rank`` owns scenarios [lo, hi): contiguous blocks of ceil(S/world); trailing ranks may be short or empty."""
    per = -(-S // world)
    lo = min(rank * per, S)
    return lo, min(lo + per, S)


def padded_count(S: int, world: int) -> int:
    return -(-S // world)


def all_gather_scenarios(local, S: int, group=None):
    """``local``: torch tensor [n_local, ...] (scenario-major) on this rank; returns [S, ...] on every rank.
    Shards are padded to equal counts (the last ones may be short), gathered once, and trimmed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = padded_count(S, world)
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out[:S]


def closed_loop_sharded(x0: np.ndarray, params: np.ndarray, N: int, k_sim: int = 20, i_sim: int = 10, eps: float = 1e-14,
                        profile: int = 0, group=None, device=None, compute: Optional[Callable] = None,
                        state_rows: int = 0, xbounds=None) -> Dict[str, np.ndarray]:
    """Runs the fused closed loop on this rank's shard and all-gathers trajectories, costs and status words.
    ``state_rows`` / ``xbounds`` as in ``NtmMpc.closed_loop`` (getWLc's state rows kept in every QP).

    ``x0`` [S,2] and ``params`` [S,16] (or [16]) are the FULL batch on every rank (they are small: 144 B per
    scenario); only the shard is copied to the GPU.  ``compute(lo, hi) -> (xk, uk, cost)`` torch tensors can
    replace the CUDA path for host-logic tests (gloo on CPU; a 4-tuple adds the int32 status words); by default the
    sm_100a kernel runs.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = x0.shape[0]
    lo, hi = shard_range(S, world, rank)
    K = k_sim
    ld = 3 * K + 4                                       # NTM_REC_DOUBLES(k_sim)
    if compute is None:
        from .api import NtmMpc
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        mpc = NtmMpc(dev.index)
        mpc.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        n = hi - lo
        pb = params if params.ndim == 1 else params[lo:hi]
        d_x0 = torch.from_numpy(np.ascontiguousarray(x0[lo:hi])).to(dev)
        d_p = torch.from_numpy(np.ascontiguousarray(pb)).to(dev)
        rec = torch.zeros((n, ld), dtype=torch.float64, device=dev)
        if n:
            mpc.closed_loop_rec_dev(n, N, K, i_sim, eps, profile, d_x0.data_ptr(), d_p.data_ptr(),
                                    1 if params.ndim == 1 else n, rec.data_ptr(), 0, 0, state_rows, xbounds)
    else:
        res_c = compute(lo, hi)
        xk, uk, cost = res_c[:3]
        n = hi - lo
        status = res_c[3] if len(res_c) > 3 else torch.zeros((n,), dtype=torch.int32, device=xk.device)
        rec = torch.cat([xk.reshape(n, 2 * (K + 1)).double(), uk.reshape(n, K).double(), cost.reshape(n, 1).double(),
                         status.reshape(n, 1).double()], dim=1)
    g = all_gather_scenarios(rec, S, group)              # the ONE collective of the path
    out = dict(xk=g[:, :2 * (K + 1)].reshape(S, K + 1, 2), uk=g[:, 2 * (K + 1):3 * K + 2], cost=g[:, 3 * K + 2],
               status=g[:, 3 * K + 3].to(torch.int32))
    res = {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in out.items()}
    if compute is None:
        mpc.close()
    return res
