#!/usr/bin/env python
"""BASELINE.json north_star target: ">= 10^5 one-year SFCC columns on DefaultGrid_5cm integrated in under one minute on
8 x B200", and BASELINE cfg4 (10^6 columns with per-column soil parameters, NCCL ensemble statistics).

Integrates `--columns` SFCC DallAmico columns for one simulated year, block-partitioned over the ranks (one process per GPU,
no communication while stepping), with a DAILY save: the device-side output path keeps the selected depths and the thaw depth
(cgb_solve_saveat_select) and the ensemble mean / variance of T of every cell is all-reduced over NCCL at every save
(k_ensemble_partial -> dist.all_reduce of 2N+1 doubles).  Single GPU: `python benchmarks/target_sfcc_year.py`; one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/target_sfcc_year.py --columns 100000 --variant lut

Variants: lut = one soil class through the presolver LUT (k_heat_sfcc_fast); lut_n2 = PainterKarra n = 2; newton = per-column
alpha, n, porosity, organic fraction through the on-device Newton solver (cfg4 as specified).
Prints one JSON line (rank 0): wall time of the whole year incl. saves and statistics, stepping-kernel time per rank
(max, mean -> imbalance), all-reduce cost, throughput."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cryogrid_jl_b200 as cg  # noqa: E402
from cryogrid_jl_b200 import presets  # noqa: E402

DAY = 86400.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=100000, help="total over all ranks")
    ap.add_argument("--variant", default="lut", choices=["lut", "lut_n2", "newton"])
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--save-every", type=int, default=1, help="days between saves / ensemble statistics")
    ap.add_argument("--depths", default="0.025,0.5,1.0,2.0,5.0,10.0", help="depths [m] kept by the device-side output path")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cryogrid_jl_b200 import distributed as D
    a, b = D.block_partition(args.columns, world)[rank]
    M = b - a
    t_setup = time.perf_counter()
    tile = presets.cfg4_tile(M, args.variant, seed=4 + 17 * rank)
    eng = cg.Engine(tile, device=local)
    eng.initialcondition(0.0)                       # initial condition on the device (cgb_init_interp)
    eng.synchronize()
    t_setup = time.perf_counter() - t_setup
    N = tile.grid.ncells
    cells = np.unique([int(np.argmin(np.abs(tile.grid.cells - float(z)))) for z in args.depths.split(",")]).astype(np.int32)
    saves = list(range(args.save_every, args.days + 1, args.save_every))
    if not saves or saves[-1] != args.days:
        saves.append(args.days)
    sel = np.empty((len(saves), 1, cells.size, M))
    thaw = np.empty((len(saves), M))
    stats = np.empty((len(saves), 2, N))
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    eng.profile_enable(True); eng.profile_read()
    t_allreduce = 0.0
    t0 = time.perf_counter()
    for k, d in enumerate(saves):
        o, th = eng.solve_saveat_select([d * DAY], ("T",), cells=cells, thawdepth=True, out=sel[k:k + 1], thaw_out=thaw[k:k + 1])
        t1 = time.perf_counter()
        if dist is not None:
            mean, var, n = D.ensemble_moments(eng, "T", d * DAY)
        else:
            s, q = eng.ensemble_partial("T", d * DAY)
            mean = s / M; var = np.maximum(q / M - mean * mean, 0.0); n = M
        stats[k, 0], stats[k, 1] = mean, var
        t_allreduce += time.perf_counter() - t1
    eng.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n_launch, ms = eng.profile_read()
    s = eng.stats()
    vals = torch.tensor([wall, ms / 1e3, t_allreduce, float(s["cell_steps"]), float(s["column_steps"]), float(s["status_or"]), t_setup],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    if rank == 0:
        wall_max, kern_max, kern_mean = mx[0].item(), mx[1].item(), sm[1].item() / world
        print(json.dumps(dict(
            config=f"{args.columns} SFCC DallAmico columns ({args.variant}), DefaultGrid_5cm (N={N}), {args.days} simulated days, CGEuler dtmax 3600 s, "
                   f"save + NCCL ensemble mean/variance every {args.save_every} day(s), {world} GPU(s)",
            n_gpus=world, columns=args.columns, wall_s_max_over_ranks=wall_max, stepping_kernel_s_max_over_ranks=kern_max,
            stepping_kernel_s_mean_over_ranks=kern_mean, imbalance_max_over_mean=kern_max / kern_mean,
            stats_and_allreduce_s_rank_max=mx[2].item(), saves=len(saves), setup_s_max=mx[6].item(),
            cell_steps=sm[3].item(), cell_steps_per_s_device=sm[3].item() / kern_max, cell_steps_per_s_wall=sm[3].item() / wall_max,
            steps_per_column=sm[4].item() / args.columns, status_or=int(mx[5].item()), stepping_launches_per_rank=n_launch,
            ensemble_n=int(n), mean_T_top_last=float(stats[-1, 0, 0]), max_std_T_last=float(np.sqrt(stats[-1, 1]).max()),
            mean_thaw_depth_max_over_year=float(thaw.max(axis=0).mean()))))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
