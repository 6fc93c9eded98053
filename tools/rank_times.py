#!/usr/bin/env python
"""Per-rank kernel time of the strong-scaled config 3 shards: all ranks at once, then one rank at a time.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/rank_times.py"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-ntm-control_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
import ntm_mpc
from ntm_mpc import physics, distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Pf, x0f, N = physics.batch_params(3, S=65536)
mpc = None
for mode in (sys.argv[1:] or ['contiguous', 'interleaved']):
    if mode == "interleaved":
        idx = np.arange(rank, 65536, world)
    else:
        lo, hi = D.shard_range(65536, world, rank); idx = np.arange(lo, hi)
    P = np.ascontiguousarray(Pf[:, idx].T); x0 = np.ascontiguousarray(x0f[idx]); S = len(idx)
    if mpc is None:
        mpc = ntm_mpc.NtmMpc(local); stream = torch.cuda.current_stream(); mpc.set_stream(stream.cuda_stream)
    dP, dx = torch.from_numpy(P).to(dev), torch.from_numpy(x0).to(dev)
    xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    def once():
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); mpc.closed_loop_dev(S, N, 20, 10, 1e-14, ntm_mpc.PROFILE_INNER_FIXED, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr())
        e1.record(stream); torch.cuda.synchronize(); return e0.elapsed_time(e1)
    once(); once()
    dist.barrier(); torch.cuda.synchronize()
    together = []
    for _ in range(8):
        dist.barrier(); torch.cuda.synchronize(); together.append(once())
    alone = []
    for r in range(world):
        dist.barrier(); torch.cuda.synchronize()
        if r == rank: alone = [once() for _ in range(6)]
    dist.barrier()
    t = torch.tensor([statistics.median(together), max(together), statistics.median(alone), max(alone)], dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        print(f"{mode} shards of {S}, NTM_LPT={os.environ.get('NTM_LPT', 'default')}")
        for r, o in enumerate(out):
            print(f"  rank {r}: all ranks at once median {o[0]:.2f} max {o[1]:.2f} ms | alone median {o[2]:.2f} max {o[3]:.2f} ms")
        print(f"  step = max over ranks: at once {max(float(o[0]) for o in out):.2f} ms, alone {max(float(o[2]) for o in out):.2f} ms")
dist.destroy_process_group()
