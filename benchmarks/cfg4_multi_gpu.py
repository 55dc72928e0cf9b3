#!/usr/bin/env python
"""BASELINE cfg4: 10^6 SFCC DallAmico columns with per-column soil parameters (on-device Newton), adaptive dt, block-partitioned
over the GPUs of one node, NCCL all-reduce of the ensemble statistics.  Run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/cfg4_multi_gpu.py --columns 1000000 --days 4

No communication happens while stepping; every rank builds only its own shard (parameters drawn from a per-rank stream)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cryogrid_jl_b200 as cg  # noqa: E402
from cryogrid_jl_b200 import presets as common  # noqa: E402
from cryogrid_jl_b200 import distributed as D  # noqa: E402

DAY = 86400.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=1000000)
    ap.add_argument("--days", type=int, default=4)
    ap.add_argument("--solver", default="newton", choices=["newton", "lut"])
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    a, b = D.block_partition(args.columns, world)[rank]
    M = b - a
    kw = dict(grid=cg.DefaultGrid_5cm, top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0),
              bottom=cg.GeothermalHeatFlux(0.053), temp=cg.SamoylovDefault_tempprofile)
    if args.solver == "newton":
        tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="newton", per_column=True, seed=1000 + rank, **kw)
    else:
        tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="lut", alpha=2.0, n=1.6, **kw)
    tile.top_offset = np.random.default_rng(2000 + rank).normal(0.0, 2.0, M)
    eng = cg.Engine(tile, device=local)
    eng.initialcondition(0.0)
    eng.advance_to(1 * DAY)                              # spin-up day (dt grows from its initial 60 s)
    s0 = eng.stats()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.profile_enable(True); eng.profile_read()
    for d in range(2, 2 + args.days):
        eng.advance_to(d * DAY, sync=False)
    n_launch, ms = eng.profile_read()                    # synchronises; CUDA-event time of the stepping kernels on this rank
    t_end = (1 + args.days) * DAY
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    mean, var, n = D.ensemble_moments(eng, "T", t_end)   # k_ensemble_partial -> NCCL all-reduce of sum x, sum x^2
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s1 = eng.stats()
    buf = torch.tensor([ms, float(s1["cell_steps"] - s0["cell_steps"]), float(s1["status_or"])], dtype=torch.float64, device="cuda")
    mx = buf.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = buf.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        kernel_s = mx[0].item() / 1e3
        print(json.dumps(dict(config=f"cfg4: {args.columns} SFCC DallAmico columns, per-column parameters ({args.solver}), N={tile.grid.ncells}, "
                                     f"{args.days} days, {world} GPUs", n_gpus=world, columns=args.columns, cell_steps=sm[1].item(),
                              kernel_s_max_over_ranks=kernel_s, cell_steps_per_s=sm[1].item() / kernel_s, wall_s_rank0=t1 - t0,
                              ensemble_stats_s=t2 - t1, ensemble_n=n, mean_T_surface=float(mean[0]), max_std_T=float(np.sqrt(var).max()),
                              status_or=int(mx[2].item()), launches_per_rank=n_launch)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
