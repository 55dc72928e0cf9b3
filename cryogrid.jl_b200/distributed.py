"""Multi-GPU seams: block partition of the columns over ranks (one process per GPU, no traffic while stepping),
all-reduce of ensemble moments and gather of saved outputs with torch.distributed (NCCL on GPUs, gloo in CPU tests).

The reference's analogue is the SciML EnsembleProblem reduction over threads/processes
(src/Solvers/DiffEq/ensemble.jl:33-45, examples/cglite_parameter_ensembles.jl:86-100)."""
import numpy as np
import torch
import torch.distributed as dist


def block_partition(M, world):
    """Contiguous block partition: rank r owns [starts[r], starts[r+1])."""
    base, rem = divmod(int(M), int(world))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(starts[r]), int(starts[r + 1])) for r in range(world)]


def my_columns(M, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = block_partition(M, world)[rank]
    return np.arange(a, b)


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_moments(local_sum, local_sumsq, local_count):
    """Σx, Σx², n over all ranks -> (mean, variance, n).  Inputs: arrays of length N (per cell) and a scalar count."""
    dev = _device()
    buf = torch.cat([torch.as_tensor(np.asarray(local_sum, dtype=np.float64)), torch.as_tensor(np.asarray(local_sumsq, dtype=np.float64)),
                     torch.tensor([float(local_count)], dtype=torch.float64)]).to(dev)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    buf = buf.cpu().numpy()
    N = (buf.size - 1) // 2
    n = buf[-1]
    mean = buf[:N] / n
    var = np.maximum(buf[N:2 * N] / n - mean * mean, 0.0)
    return mean, var, int(n)


def ensemble_moments(engine, var, t):
    """Ensemble mean/variance of a diagnostic over ALL ranks' columns.  On NCCL the per-rank partial sums are produced
    by the CUDA reduction kernel directly into a device tensor (no host round trip) and all-reduced over NVLink."""
    if dist.get_backend() == "nccl":
        buf = torch.empty(2 * engine.N + 1, dtype=torch.float64, device=_device())
        # the library writes all 2N sums on ITS stream (and synchronises it before returning); torch's stream must not have
        # pending work on the buffer either way
        torch.cuda.current_stream().synchronize()
        engine.ensemble_partial_device(var, t, buf.data_ptr())
        buf[-1] = float(engine.M)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        b = buf.cpu().numpy()
        n = b[-1]
        mean = b[:engine.N] / n
        return mean, np.maximum(b[engine.N:2 * engine.N] / n - mean * mean, 0.0), int(n)
    s, q = engine.ensemble_partial(var, t)
    return allreduce_moments(s, q, engine.M)


def gather_saved(local, M, dst=0):
    """Gather saved outputs [..., M_local] (column axis last) from every rank into [..., M] on rank dst."""
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = block_partition(M, world)
    dev = _device()
    mmax = max(b - a for a, b in parts)
    lead = local.shape[:-1]
    pad = np.zeros(lead + (mmax,), dtype=np.float64)
    pad[..., :local.shape[-1]] = local
    t = torch.as_tensor(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    if dist.get_backend() == "nccl":
        # NCCL has no gather-to-root in older torch: all_gather is uniform over NVSwitch anyway
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
    else:
        dist.gather(t, outs, dst=dst)
    if rank != dst:
        return None
    full = np.empty(lead + (M,), dtype=np.float64)
    for r, (a, b) in enumerate(parts):
        full[..., a:b] = outs[r].cpu().numpy()[..., :b - a]
    return full
