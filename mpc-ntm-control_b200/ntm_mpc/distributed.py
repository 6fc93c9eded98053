"""Scenario sharding across the GPUs of one box: one process per GPU (torchrun), contiguous blocks of
ceil(S/G) scenarios per rank, no communication inside the time loop, and ONE all-gather of the
per-scenario outputs {xk, uk, cost} at the end.  Scenarios never interact
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
    if compute is None:
        from . import LAYOUT_MATLAB
        from .api import NtmMpc
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        mpc = NtmMpc(dev.index)
        mpc.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        n = hi - lo
        pb = params if params.ndim == 1 else params[lo:hi]
        d_x0 = torch.from_numpy(np.ascontiguousarray(x0[lo:hi])).to(dev)
        d_p = torch.from_numpy(np.ascontiguousarray(pb)).to(dev)
        xk = torch.empty((n, k_sim + 1, 2), dtype=torch.float64, device=dev)
        uk = torch.empty((n, k_sim), dtype=torch.float64, device=dev)
        cost = torch.empty((n,), dtype=torch.float64, device=dev)
        status = torch.zeros((n,), dtype=torch.int32, device=dev)
        if n and state_rows:
            from .api import MC_STATE_BOX
            mpc.closed_loop_sc_dev(n, N, k_sim, i_sim, eps, profile, LAYOUT_MATLAB, d_x0.data_ptr(), d_p.data_ptr(),
                                   1 if params.ndim == 1 else n, state_rows, MC_STATE_BOX if xbounds is None else xbounds,
                                   xk.data_ptr(), uk.data_ptr(), 0, cost.data_ptr(), 0, 0, status.data_ptr())
        elif n:
            mpc.closed_loop_dev(n, N, k_sim, i_sim, eps, profile, LAYOUT_MATLAB, d_x0.data_ptr(), d_p.data_ptr(),
                                1 if params.ndim == 1 else n, xk.data_ptr(), uk.data_ptr(), 0, cost.data_ptr(), 0, 0,
                                status.data_ptr())
    else:
        res_c = compute(lo, hi)
        xk, uk, cost = res_c[:3]
        status = res_c[3] if len(res_c) > 3 else torch.zeros((hi - lo,), dtype=torch.int32, device=xk.device)
    out = dict(xk=all_gather_scenarios(xk, S, group), uk=all_gather_scenarios(uk, S, group),
               cost=all_gather_scenarios(cost, S, group), status=all_gather_scenarios(status, S, group))
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if compute is None:
        mpc.close()
    return res
