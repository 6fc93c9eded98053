"""torchrun --nproc-per-node G tools/check_multi_gpu.py : the sharded + all-gathered closed loop equals the
single-GPU run of the whole batch bit for bit (scenarios are independent)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
import numpy as np
import torch
import torch.distributed as dist

import ntm_mpc
from ntm_mpc import distributed as D, physics

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, x0, N = physics.batch_params(3, S=8191)                     # odd count: the last shard is short
prof = ntm_mpc.PROFILE_INNER_FIXED
r = D.closed_loop_sharded(x0, np.ascontiguousarray(P.T), N, profile=prof)
if dist.get_rank() == 0:
    one = ntm_mpc.NtmMpc(local).closed_loop(x0, P.T, N=N, profile=prof)
    ok = all(np.array_equal(r[k], one[k]) for k in ("xk", "uk", "cost"))
    print(f"multi-GPU check world={dist.get_world_size()} S={x0.shape[0]} bit-identical={ok}")
    assert ok
# the same with getWLc's state rows kept and the RK4 plant (EXT instantiations): NaN-aware comparison
XB = (0.05, 0.16, 2000.0, 12000.0)
prof2 = prof | ntm_mpc.PROFILE_PLANT_RK4
r = D.closed_loop_sharded(x0[:2047], np.ascontiguousarray(P.T[:2047]), N, profile=prof2, state_rows=ntm_mpc.STATE_ROWS_REFRESH, xbounds=XB)
if dist.get_rank() == 0:
    one = ntm_mpc.NtmMpc(local).closed_loop(x0[:2047], P.T[:2047], N=N, profile=prof2, state_rows=ntm_mpc.STATE_ROWS_REFRESH, xbounds=XB)
    ok = all(np.array_equal(r[k], one[k], equal_nan=True) for k in ("xk", "uk", "cost")) and np.array_equal(r["status"], one["status"])
    print(f"multi-GPU check (state rows + RK4) S=2047 bit-identical={ok} infeasible={int((one['status'] == 3).sum())}")
    assert ok
dist.barrier()
dist.destroy_process_group()
