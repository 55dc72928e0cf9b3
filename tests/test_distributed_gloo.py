"""world_size-2 gloo test (CPU) of the multi-GPU host logic: column partitioning, ensemble-moment all-reduce and
the gather of saved outputs.  No collective runs while stepping (columns are independent), so this is all there is."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, N, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from cryogrid_jl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    full = rng.normal(size=(3, N, M))            # [nsave, N, M] "saved temperatures" of the whole ensemble
    cols = D.my_columns(M)
    local = full[..., cols]
    mean, var, n = D.allreduce_moments(local[-1].sum(axis=1), (local[-1] ** 2).sum(axis=1), cols.size)
    gathered = D.gather_saved(local, M, dst=0)
    ok = n == M and np.allclose(mean, full[-1].mean(axis=1)) and np.allclose(var, full[-1].var(axis=1))
    if rank == 0:
        ok = ok and gathered is not None and np.array_equal(gathered, full)
    else:
        ok = ok and gathered is None
    q.put((rank, bool(ok), int(cols[0]), int(cols[-1])))
    dist.destroy_process_group()


def test_partition_covers_everything():
    import cryogrid_jl_b200 as cg
    from cryogrid_jl_b200.distributed import block_partition
    for M, w in ((10, 3), (7, 8), (100000, 8), (5, 1)):
        parts = block_partition(M, w)
        assert parts[0][0] == 0 and parts[-1][1] == M
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_world2_gloo_moments_and_gather():
    world, M, N = 2, 11, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True], res
    assert res[0][2] == 0 and res[0][3] == 5 and res[1][2] == 6 and res[1][3] == 10
