"""Multi-GPU check (run under torchrun, one rank per GPU): columns are block-partitioned over the ranks, every rank
steps its own shard with no communication, then the ensemble mean/variance of T is all-reduced over NCCL and the saved
output is gathered; rank 0 compares both with a single-engine run over all columns.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/nccl_ensemble_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
import cryogrid_jl_b200 as cg  # noqa: E402
from cryogrid_jl_b200 import distributed as D  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    M = 1003                                       # ragged over the ranks on purpose
    tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="newton", per_column=True, grid=cg.DefaultGrid_5cm, seed=4,
                            top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0), bottom=cg.GeothermalHeatFlux(0.053),
                            temp=cg.SamoylovDefault_tempprofile)
    H0 = cg.initialcondition(tile)
    cols = D.my_columns(M)
    eng = cg.Engine(tile, device=local, columns=cols, dtmax=900.0)
    eng.set_state(np.ascontiguousarray(H0[:, cols]), 0.0)
    t_end = 2 * 86400.0
    eng.advance_to(t_end)
    mean, var, n = D.ensemble_moments(eng, "T", t_end)
    T_local = eng.diag("T", t_end)
    T_all = D.gather_saved(T_local, M, dst=0)
    ok = True
    if rank == 0:
        full = cg.Engine(tile, device=local, dtmax=900.0)
        full.set_state(H0, 0.0)
        full.advance_to(t_end)
        T_ref = full.diag("T", t_end)
        ok = (n == M and np.array_equal(T_all, T_ref) and np.allclose(mean, T_ref.mean(axis=1), rtol=1e-12, atol=1e-12)
              and np.allclose(var, T_ref.var(axis=1), rtol=1e-8, atol=1e-12))
        print(f"nccl ensemble check world={world}: {'OK' if ok else 'FAILED'}  (M={M}, mean T at surface {mean[0]:.4f}, "
              f"max std {np.sqrt(var).max():.4f})")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
