#!/usr/bin/env python
"""bench.py -- column-cell-steps/s of the batched heat-column hot path (BASELINE.json metric).

Workload (config.workload): BASELINE cfg2 -- 10^4 FreeWater enthalpy heat columns per GPU on DefaultGrid_5cm
(N=218), 5-layer Samoylov profile with per-column porosity jitter, synthetic sinusoidal Tair sampled hourly,
NFactor(0.6,0.9), GeothermalHeatFlux(0.053), CGEuler with MaxDelta(50 kJ) adaptive dt.  One bench "step" integrates
every column over `--days-per-step` simulated days (default 73; the default K=5 timed steps are one simulated
year, after W=3 steps of spin-up).  One "cell-step" = one grid cell advanced by one accepted Euler step.

  value     device-resident throughput: state already in HBM, K steps timed with CUDA events on the engine's stream.
  e2e       the same metric through the public API (solve(CryoGridProblem)) with HOST buffers: H2D of u0 and the
            D2H of the daily saved temperatures are inside the timed region.
  roofline  algorithmic bytes (16 B per cell-step: read H, write H) per stepping-kernel launch / its average duration
            (CUDA events around every launch), against the measured HBM copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline  the oracle restatement of the reference (C, OpenMP over columns) on a bounded column sample.

`--impl reference` times the reference's CPU implementation of the path (the oracle port -- CryoGrid.jl is Julia and
Julia is not installed here) with all host threads on the same config/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

DAY = 86400.0
BYTES_PER_CELL_STEP = 16.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--columns", type=int, default=10000, help="columns per GPU (weak scaling)")
    ap.add_argument("--days-per-step", type=float, default=73.0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--cpu-columns", type=int, default=1024, help="column sample of the CPU legs")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per launch of the headline kernel from the committed `ncu --set full` capture of THIS bench command on the
    final build (profiles/r3j_ncu_bench_fast_kernel.json, written by scripts/ncu_traffic.py); None when there is no capture."""
    path = os.path.join(ROOT, "profiles", "r3j_ncu_bench_fast_kernel.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return (int(d["dram_bytes_read"] + d["dram_bytes_write"]),
                f"dram__bytes_read.sum + dram__bytes_write.sum of the longest launch of {d['kernel'].split('(')[0]} in an ncu --set full capture of "
                f"`{d['command']}` (profiles/r3j_ncu_bench_fast_kernel.json)", d.get("fp64_pipe_pct"))
    except Exception:
        return None, "no ncu --set full capture of the bench command committed", None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region: one streaming `nvidia-smi -lms 100` process
    (started early, so its start-up latency is not part of the window), time-stamped rows, and a window set by
    mark_start()/mark_end().  Rows outside the window (warm-up under the same load) are the fallback if none fell inside."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                r = [x.strip() for x in line.strip().split(",")]
                if len(r) >= 6:
                    self.rows.append((time.perf_counter(), r))
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        return self

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.t1 is None:
            self.mark_end()
        time.sleep(0.15)                       # let the sample that covers the end of the window arrive
        if self.proc is not None:
            self.proc.terminate()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = self.t0 if self.t0 is not None else -1.0
        t1 = self.t1 if self.t1 is not None else float("inf")
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.1]
        rows = inside or [r for _, r in self.rows]
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(rows), "samples_in_timed_region": len(inside)}


def make_tile(columns, total_days, seed):
    from cryogrid_jl_b200 import presets
    return presets.samoylov_freew_tile(ncolumns=columns, seed=seed, forcing_days=int(total_days) + 3)


def bind_to_gpu_numa(local):
    """Pins this rank's CPU affinity to the NUMA node of its GPU BEFORE any pinned host buffer is allocated, so that the pages of
    the e2e output buffers are local to the GPU's PCIe root (with floating ranks the 8 x 33 GB/s of saved fields crossed the
    socket interconnect and the aggregate stalled near 90 GB/s: r1 SCALE e2e 0.88 at N = 8).  Best effort: returns a note."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return f"gpu {bdf}: local cpus {spec} not in this process's affinity mask"
        os.sched_setaffinity(0, allowed)
        try:
            with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
                node = f.read().strip()
        except Exception:
            node = "?"
        return f"gpu {bdf}: bound to {len(allowed)} cpus of NUMA node {node}"
    except Exception as e:      # no sysfs in the container, no such attribute, ...
        return f"not bound ({type(e).__name__}: {e})"


def host_threads():
    """Host cores this process may use.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently turn the CPU legs (all host threads, by contract) into single-thread runs."""
    try:
        n = max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        n = max(1, os.cpu_count() or 1)
    q = cgroup_cpu_quota()
    if q is not None:
        n = max(1, min(n, int(np.ceil(q))))      # more runnable threads than the quota only adds throttling stalls
    return n


def cgroup_cpu_quota():
    """CPUs' worth of time the container may use per period (cgroup v2 cpu.max / v1 cfs quota), or None when unlimited."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            a, b = f.read().split()[:2]
        if a != "max":
            return float(a) / float(b)
        return None
    except (OSError, ValueError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
            q = float(f.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            p = float(f.read())
        return q / p if q > 0 and p > 0 else None
    except (OSError, ValueError):
        return None


def pick_threads(tile, H0, cols, prepared, nmax):
    """Thread count for the CPU legs: the fastest of a few candidates up to the cores this process may use, measured on a short
    sample.  (What the scheduler reports is not always what runs: on a shared host, or under a CPU quota, more OpenMP threads than
    free cores was measured 4x SLOWER than the right count -- the reference arm must not be handicapped by that.)"""
    import oracle as orc
    cand = sorted({c for c in (nmax, (3 * nmax) // 4, nmax // 2, nmax // 4, 16, 8) if 1 <= c <= nmax}, reverse=True)
    best, best_rate, rates = cand[0], 0.0, {}
    for c in cand:
        rate = 0.0
        for _ in range(2):                      # the first call at a new thread count also pays for starting the threads
            H = np.ascontiguousarray(H0[:, cols].T); t = np.zeros(cols.size); dt = np.full(cols.size, 60.0)
            t0 = time.perf_counter()
            st, ns = orc.cgeuler_advance_batch(tile, cols, H, t, dt, 1.0 * DAY, nthreads=c, prepared=prepared)
            rate = max(rate, float(ns.sum()) / (time.perf_counter() - t0))
        rates[c] = rate
        if rate > best_rate:
            best, best_rate = c, rate
    return best, rates


def cpu_leg(tile, H0, seconds, nthreads=0, days=None, ncol=1024):
    """The reference's CPU path as restated by the oracle, timing copy (oracle/libcg_oracle_fast.so: -O3, AVX2 + FMA, hoisted
    invariants), OpenMP over columns on all host threads, on a bounded sample: `ncol` columns (>= 10^3 by default) for as many
    simulated days as fit the time budget.  Returns (cell_steps/s, dict)."""
    import oracle as orc
    orc.use_timing_copy(True)
    nmax = nthreads or host_threads()
    ncol = min(tile.M, max(ncol, 2 * nmax))
    cols = np.arange(ncol)
    prepared = orc.batch_columns(tile, cols)
    N = tile.grid.ncells
    nthreads, _ = pick_threads(tile, H0, cols, prepared, nmax) if nthreads == 0 else (nthreads, None)
    cores = nthreads

    def run(d):
        H = np.ascontiguousarray(H0[:, cols].T)
        t = np.zeros(ncol); dt = np.full(ncol, 60.0)
        t0 = time.perf_counter()
        st, ns = orc.cgeuler_advance_batch(tile, cols, H, t, dt, d * DAY, nthreads=nthreads, prepared=prepared)
        el = time.perf_counter() - t0
        return int(ns.sum()) * N, el
    if days is None:
        cs, el = run(1.0)                                   # calibrate on one simulated day
        days = max(1.0, min(365.0, 1.0 * seconds / max(el, 1e-3)))
    cs, el = run(days)
    return cs / el, {"cores": int(cores), "sample": f"{ncol} columns x {days:.2f} simulated days (start of the cfg2 run), oracle timing copy "
                     "(-O3 -march=x86-64-v3, OpenMP over columns)", "seconds": el}


def _timed_advance(eng, targets):
    eng.profile_enable(True); eng.profile_read()
    s0 = eng.stats()
    for tt in targets:
        eng.advance_to(tt, sync=False)
    eng.synchronize()
    n, ms = eng.profile_read()
    eng.profile_enable(False)
    s1 = eng.stats()
    return s1["cell_steps"] - s0["cell_steps"], ms / 1e3, s1


def side_configs(device, peak):
    """Short, driver-observable runs of the OTHER BASELINE.json configurations (the headline workload is cfg2): for each one the
    device throughput over a few simulated days at full column count, its HBM-equivalent fraction (SURVEY 8d bytes per cell-step)
    and a parity spot check of the same physics on 8 columns against the oracle (one RHS evaluation, relative error scaled as in
    SURVEY 8d; LiteImplicit: one daily step, relative enthalpy error)."""
    import cryogrid_jl_b200 as cg
    from cryogrid_jl_b200 import presets
    import oracle as orc
    orc.use_timing_copy(False)            # parity spot checks use the PARITY oracle (-O2 -ffp-contract=off)
    out = {}

    def heat_rhs_parity(tile, dtmax=3600.0):
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile, dtmax=dtmax, device=device); eng.set_state(H0, 0.0)
        dH = eng.rhs(0.0); eng.close()
        worst = 0.0
        for c in range(tile.M):
            oc = orc.OracleColumn(tile, c)
            dref, _ = oc.rhs(H0[:, c], 0.0)
            worst = max(worst, presets.rhs_error(dH[:, c], dref, oc.work.get("jH"), tile.grid.dedges))
        return worst

    def entry(name, cs, sec, bytes_per, columns, cells, parity, note, stats):
        out[name] = {"value": cs / sec, "unit": "cell-steps/s", "hbm_equiv_frac": cs / sec * bytes_per / (peak * 1e9), "bytes_per_cell_step": bytes_per,
                     "columns": columns, "cells": cells, "steps_per_column": stats["column_steps"] / columns, "status_or": stats["status_or"],
                     "parity_8col_vs_oracle": parity, "note": note}

    # cfg1b: examples/heat_sfcc_constantbc.jl with the SFCC wired in, ONE column (latency of a single column, not throughput)
    tile = presets.sfcc_tile(ncolumns=1, kind="painterkarra", solver="lut")
    eng = cg.Engine(tile, dtmax=900.0, device=device); eng.set_state(cg.initialcondition(tile), 0.0); eng.advance_to(DAY)
    cs, sec, st = _timed_advance(eng, [DAY * d for d in range(2, 12)]); eng.close()
    entry("cfg1b", cs, sec, 16.0, 1, tile.grid.ncells, heat_rhs_parity(presets.sfcc_tile(ncolumns=8, kind="painterkarra", solver="lut"), 900.0),
          "single SFCC column (PainterKarra a=1 n=2, presolver LUT), constant 10 W/m2 in/out, dtmax 900 s, 10 days; RHS parity", st)
    # cfg3: LiteImplicitEuler ensemble, 10^5 members, initial condition on the device
    M = 100000
    tile = presets.lite_ensemble_tile(ncolumns=M, seed=1234, forcing_days=60)
    eng = cg.Engine(tile, stepper="lite", device=device); eng.initialcondition(0.0); eng.advance_to(5 * DAY)
    cs, sec, st = _timed_advance(eng, [DAY * d for d in range(10, 41, 5)]); eng.close()
    t8 = presets.lite_ensemble_tile(ncolumns=8, seed=1234, forcing_days=60)
    H8 = cg.initialcondition(t8)
    e8 = cg.Engine(t8, stepper="lite", device=device); e8.set_state(H8, 0.0); e8.advance_to(DAY); Hg = e8.get_state(); e8.close()
    par = max(float(np.max(np.abs(Hg[:, c] - orc.OracleColumn(t8, c).lite_advance(H8[:, c], 0.0, DAY, DAY)[0]) / (np.abs(Hg[:, c]) + 1e3))) for c in range(8))
    entry("cfg3", cs, sec, 16.0, M, tile.grid.ncells, par,
          "LiteImplicitEuler parameter ensemble (cglite_parameter_ensembles-style), 35 daily steps, 5 per launch, ThermalSteadyStateInit on the device; "
          "parity = relative enthalpy error after one daily step", st)
    # cfg4: SFCC DallAmico -- presolver LUT of one soil class (SFCC fast path) and per-column parameters (on-device Newton)
    for name, variant, note in (("cfg4_lut", "lut", "10^5 SFCC DallAmico columns of one soil class through the presolver LUT (k_heat_sfcc_fast), 3 days"),
                                ("cfg4_newton", "newton", "10^5 SFCC DallAmico columns with per-column alpha, n, porosity, organic fraction (on-device Newton), 3 days")):
        tile = presets.cfg4_tile(M, variant)
        eng = cg.Engine(tile, device=device); eng.set_state(cg.initialcondition(tile), 0.0); eng.advance_to(DAY)
        cs, sec, st = _timed_advance(eng, [DAY * d for d in range(2, 5)]); eng.close()
        entry(name, cs, sec, 16.0, M, tile.grid.ncells, heat_rhs_parity(presets.cfg4_tile(8, variant)), note + "; RHS parity", st)
    # cfg5: coupled heat + salt
    tile = presets.cfg5_tile(M)
    eng = cg.SaltEngine(tile, device=device); eng.set_state(tile.initialcondition(), 0.0); eng.advance_to(DAY)
    cs, sec, st = _timed_advance(eng, [DAY * d for d in range(2, 5)]); eng.close()
    t8 = presets.cfg5_tile(8)
    u8 = t8.initialcondition(); rng = np.random.default_rng(1)
    u8[0] += rng.uniform(-3.0, 1.5, u8[0].shape); u8[1] += rng.uniform(0.0, 800.0, u8[1].shape)
    e8 = cg.SaltEngine(t8, device=device); e8.set_state(u8, 0.0); du = e8.rhs(0.0); e8.close()
    par = 0.0
    for c in range(8):
        dT, dc = orc.OracleSaltColumn(t8, c).rhs(u8[0, :, c], u8[1, :, c])
        par = max(par, float(np.max(np.abs(du[0, :, c] - dT)) / np.max(np.abs(dT))), float(np.max(np.abs(du[1, :, c] - dc)) / np.max(np.abs(dc))))
    entry("cfg5", cs, sec, 32.0, M, tile.grid.ncells, par, "10^5 coupled heat + salt columns (heat_sfcc_salt_constantbc-style), 3 days; "
          "parity = max |du - du_ref| / max |du_ref| of one RHS evaluation at a salty, partly frozen state", st)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cryogrid_jl_b200 as cg
    import oracle as orc
    orc.use_timing_copy(True)
    nmax = host_threads()
    ncol = max(args.cpu_columns, 2 * nmax)
    tile = make_tile(ncol, 2 + (args.steps + args.warmup) * args.days_per_step, seed=1)      # forcing covers every step of the run
    H0 = cg.initialcondition(tile)
    N = tile.grid.ncells
    cols = np.arange(ncol)
    prepared = orc.batch_columns(tile, cols)
    cores, thread_rates = pick_threads(tile, H0, cols, prepared, nmax)
    H = np.ascontiguousarray(H0.T); t = np.zeros(ncol); dt = np.full(ncol, 60.0)
    # size the per-step sample so that the whole run stays within ~2 minutes: the limiter-bound seasons of cfg2 take ~400 steps per
    # column and simulated day (the frozen ones 24), the throughput is the one just measured at the chosen thread count
    orc.cgeuler_advance_batch(tile, cols, H, t, dt, 1.0 * DAY, nthreads=cores, prepared=prepared)
    per_day = 400.0 * ncol / max(thread_rates[cores], 1.0)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    days = max(0.5, min(args.days_per_step, budget / max(per_day, 1e-6)))
    cur = 1.0
    for _ in range(args.warmup):
        cur += days
        orc.cgeuler_advance_batch(tile, cols, H, t, dt, cur * DAY, nthreads=cores, prepared=prepared)
    cs = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cur += days
        st, ns = orc.cgeuler_advance_batch(tile, cols, H, t, dt, cur * DAY, nthreads=cores, prepared=prepared)
        cs += int(ns.sum()) * N
    el = time.perf_counter() - t0
    val = cs / el
    sample = (f"{ncol} columns x {days:.2f} simulated days per step (bounded sample of cfg2), oracle port of CryoGrid.jl (Julia not installed), "
              "timing copy -O3 -march=x86-64-v3, OpenMP over columns")
    print(json.dumps({
        "impl": "reference", "metric": "column-cell-steps/sec", "value": val, "unit": "cell-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: FreeWater enthalpy heat columns, DefaultGrid_5cm (N=218), sinusoidal Tair, CGEuler adaptive dt",
                   "columns": ncol, "days_per_step": days},
        "cpu_baseline": {"value": val, "unit": "cell-steps/s", "cores": int(cores), "kind": "port", "sample": sample,
                         "host_threads": int(nmax), "threads_tried": {str(k): round(v * N, 1) for k, v in thread_rates.items()}},
        "e2e": {"value": val, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import cryogrid_jl_b200 as cg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cryogrid.jl_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local) if world > 1 else "single GPU: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W, D = args.steps, args.warmup, args.days_per_step
    total_days = (K + W) * D
    tile = make_tile(args.columns, total_days, seed=1 + rank)      # each rank owns its own block of columns
    N, M = tile.grid.ncells, tile.M
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, stepper="euler", device=local)
    eng.set_state(H0, 0.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    with ClockSampler(local) as clk:           # streaming sampler: started before the warm-up, windowed to the timed region
        for w in range(W):
            eng.advance_to((w + 1) * D * DAY)
        eng.synchronize()
        s0 = eng.stats()
        eng.profile_enable(True)
        eng.profile_read()
        barrier()
        clk.mark_start()
        t_wall = time.perf_counter()
        for k in range(K):
            flush.zero_()                      # L2 flush between timed steps (outside the CUDA-event brackets of the kernels)
            torch.cuda.synchronize()
            eng.advance_to((W + k + 1) * D * DAY, sync=True)
        barrier()
        wall = time.perf_counter() - t_wall
        clk.mark_end()
    n_launch, kern_ms = eng.profile_read()
    eng.profile_enable(False)
    s1 = eng.stats()
    cell_steps = s1["cell_steps"] - s0["cell_steps"]
    # device time of the timed region = sum of the stepping kernels (one per step) measured with CUDA events; max over ranks
    tt = torch.tensor([kern_ms / 1e3, wall, float(cell_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s, wall_s, total_cs = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        dev_s, wall_s, total_cs = float(tt[0]), float(tt[1]), float(tt[2])
    value = total_cs / dev_s
    peak, peak_src = peaks()
    ach = (cell_steps * BYTES_PER_CELL_STEP / max(n_launch, 1)) / (kern_ms / 1e3 / max(n_launch, 1)) / 1e9

    # ---- e2e: public API with host buffers (H2D of u0 + D2H of the daily saved T inside the timed region)
    e2e = None
    if not args.no_e2e:
        # same simulated window as the first timed step: start from the state after the spin-up (held in pinned HOST
        # memory), upload it, integrate D days with daily saves of T, and bring every save back to pinned host memory
        nsave = int(D)
        t_w = W * D * DAY
        eng_w = cg.Engine(tile, stepper="euler", device=local)
        eng_w.set_state(H0, 0.0)
        eng_w.advance_to(t_w)
        pinned_u0 = torch.empty((N, M), dtype=torch.float64, pin_memory=True).numpy()
        pinned_u0[:] = eng_w.get_state()
        eng_w.close()
        pinned_out = torch.empty((nsave + 1, 1, N, M), dtype=torch.float64, pin_memory=True).numpy()
        prob = cg.CryoGridProblem(tile, pinned_u0, (t_w, t_w + nsave * DAY), saveat=DAY, savevars=("T",))
        e2e_steps = max(1, min(K, 3))
        eng2 = cg.Engine(tile, stepper="euler", device=local)
        cg.solve(prob, cg.CGEuler(), engine=eng2, out=pinned_out, return_state=False)   # warm-up (allocates the staging buffers)
        barrier()
        t0 = time.perf_counter()
        cs2 = 0
        for _ in range(e2e_steps):
            sol = cg.solve(prob, cg.CGEuler(), engine=eng2, out=pinned_out, return_state=False)
            cs2 += sol.stats["cell_steps"]
        barrier()
        el = time.perf_counter() - t0
        te = torch.tensor([el, float(cs2)], dtype=torch.float64, device="cuda")
        if world > 1:
            a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            el, cs2 = float(a[0]), float(b[1])
        # context for the number above: what one pinned device-to-host copy of a saved field block achieves on this box
        probe_d = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
        probe_h = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, pin_memory=True)
        probe_h.copy_(probe_d, non_blocking=True); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(4):
            probe_h.copy_(probe_d, non_blocking=True)
        ev1.record(); torch.cuda.synchronize()
        d2h_gbs = 4 * probe_h.numel() * 8 / (ev0.elapsed_time(ev1) / 1e3) / 1e9
        del probe_d, probe_h
        e2e = {"value": cs2 / el, "unit": "cell-steps/s", "h2d_bytes_per_step": int(pinned_u0.nbytes),
               "pinned_d2h_gbs_probe": d2h_gbs, "d2h_gbs_achieved": (nsave + 1) * N * M * 8 * e2e_steps * world / el / 1e9 / world,
               "d2h_bytes_per_step": int((nsave + 1) * N * M * 8), "steps": e2e_steps, "ms_per_step": 1e3 * el / e2e_steps,
               "numa": numa_note,
               "note": "solve(CryoGridProblem(saveat=1 day, savevars=T)) on the first timed window, u0 uploaded from pinned host "
                       "memory, every saved T copied back to pinned host memory (overlapped on a second stream)"}
        # the device-side output path on the same window: only the depths users read from CryoGridOutput (out.T[Z(Near(zs))],
        # examples/cglite_parameter_ensembles.jl:96) and the thaw depth of every column leave the device
        cells = np.unique([int(np.argmin(np.abs(tile.grid.cells - z))) for z in (0.025, 0.5, 1.0, 2.0, 5.0, 10.0)]).astype(np.int32)
        sel_out = torch.empty((nsave, 1, cells.size, M), dtype=torch.float64, pin_memory=True).numpy()
        thaw_out = torch.empty((nsave, M), dtype=torch.float64, pin_memory=True).numpy()
        tsave = t_w + DAY * np.arange(1, nsave + 1)

        def run_select():
            eng2.set_state(pinned_u0, t_w)
            eng2.solve_saveat_select(tsave, ("T",), cells=cells, thawdepth=True, out=sel_out, thaw_out=thaw_out)
            return eng2.stats()["cell_steps"]
        run_select()
        barrier()
        t0 = time.perf_counter()
        cs3 = sum(run_select() for _ in range(e2e_steps))
        barrier()
        el3 = time.perf_counter() - t0
        te = torch.tensor([el3, float(cs3)], dtype=torch.float64, device="cuda")
        if world > 1:
            a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
            el3, cs3 = float(a[0]), float(b[1])
        e2e["select"] = {"value": cs3 / el3, "unit": "cell-steps/s", "h2d_bytes_per_step": int(pinned_u0.nbytes),
                         "d2h_bytes_per_step": int(sel_out.nbytes + thaw_out.nbytes), "ms_per_step": 1e3 * el3 / e2e_steps,
                         "note": "same window through cgb_solve_saveat_select: daily T at 6 depths + thaw depth per column (device-side output path)"}
        eng2.close()

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = side_configs(local, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, info = cpu_leg(tile, H0, args.cpu_seconds, ncol=args.cpu_columns)
        cpu = {"value": v, "unit": "cell-steps/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]}

    traffic, traffic_src, fp64_pct = measured_traffic()
    if rank == 0:
        out = {
            "metric": "column-cell-steps/sec", "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s / max(K, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2: 10^4 FreeWater enthalpy heat columns per GPU, DefaultGrid_5cm (N=218), synthetic sinusoidal Tair, "
                                   "CGEuler adaptive dt (MaxDelta 50 kJ), K steps = one simulated year",
                       "columns_per_gpu": M, "cells": N, "days_per_step": D, "l2_policy": "L2 flushed (256 MB memset) between timed steps; each step reads the "
                       "17 MB state once and keeps it in registers, so the kernel is FP64-bound either way (DESIGN.md)",
                       "wall_s_timed_region": wall_s},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_cell_step": BYTES_PER_CELL_STEP, "traffic_source": traffic_src,
                         "note": "algorithmic bytes (16 B per cell-step) far exceed the DRAM traffic because the state stays in registers across the Euler steps of a launch (temporal fusion); the kernel is bound by the FP64 pipe" +
                                 (f" ({fp64_pct:.0f} % busy in the capture)" if fp64_pct else "")},
            "cpu_baseline": cpu, "e2e": e2e, "configs": configs, "gpu_launches": int(s1["kernel_launches"] - s0["kernel_launches"]),
            "clocks": clk.summary(), "cell_steps": int(total_cs), "column_steps_per_column": (s1["column_steps"] - s0["column_steps"]) / M,
        }
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
