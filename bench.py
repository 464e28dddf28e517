#!/usr/bin/env python
"""bench.py -- RBCD iterations/s of the B200 hot path on the synthetic 3-D grid
pose graph with range edges (BASELINE.json configs[4]).

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

One step = one RBCD iteration: every agent exchanges its public states,
rebuilds its linear term G and takes one QuadraticOptimizer::optimize step
(default ROptParameters: <= 3 RTR outer x <= 50 tCG, tol 1e-2), colour by
colour of the robot graph, followed by the global cost / gradient-norm
evaluation the reference driver performs every iteration
(examples/MultiRobotExample.cpp:223-307).  The agents are partitioned over
the N GPUs (strong scaling: the problem is fixed).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=100, help="grid side S (S^3 poses)")
    ap.add_argument("--agents", type=int, default=256)
    ap.add_argument("--rank", type=int, default=5, help="relaxation rank r")
    ap.add_argument("--range-fraction", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-agents", type=int, default=2)
    ap.add_argument("--kernel-iters", type=int, default=20)
    ap.add_argument("--ttc-grid", type=int, default=12,
                    help="side of the single-agent grid of the time-to-certified leg (0: skip)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)
            v = d.get("hbm_gbs") or d.get("hbm_gb_s")
            if v:
                return float(v), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def k1_traffic(args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one full-batch K1 launch on the default
    workload, from the committed ncu --set full capture (profiles/k1_traffic.json); null for
    any other workload / GPU count (no capture)."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if world != 1 or (args.grid, args.agents, args.rank, args.range_fraction) != (100, 256, 5, 0.1):
        return None
    try:
        with open(path) as fh:
            d = json.load(fh)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(args, log):
    from dcora_b200 import graph as pg, synthetic
    t0 = time.perf_counter()
    meas, gt, idx = synthetic.grid3d(args.grid, args.range_fraction, seed=args.seed)
    part = pg.Partition.contiguous(meas.n, args.agents)
    gp = pg.GlobalProblem(meas, part)
    gp.Q()
    log(f"graph: n={meas.n} m={len(meas.pp_i)} ranges={len(meas.rg_i)} K={gp.K} "
        f"nnz(Q)={gp.Q().nnz} built in {time.perf_counter() - t0:.1f}s")
    return meas, gt, idx, gp


def initial_X(args, meas, idx, gp):
    """Spanning-tree guess lifted to rank r (dcora_b200.synthetic.lifted_initial_guess)."""
    from dcora_b200 import synthetic
    return synthetic.lifted_initial_guess(meas, idx, gp, args.grid, args.rank)


# --------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dcora_b200 as dc
    from dcora_b200 import _lib, rbcd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        assert world == 1 and args.gpus == 1, "launch with torchrun for --gpus > 1"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def log(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    _lib.load()
    meas, gt, idx, gp = build_problem(args, log)
    X0 = initial_X(args, meas, idx, gp)
    t0 = time.perf_counter()
    params = dc.ROptParameters(time_bound_s=0.0)
    S = rbcd.RBCD(gp, args.rank, rank=rank, world=world, device=local_rank, params=params)
    setup_s = time.perf_counter() - t0
    log(f"handle: agents/rank={S.A_local} colours={S.plan.num_colors} "
        f"nnz(L)={S.P.factor_nnz()} setup {setup_s:.1f}s")
    S.set_X(X0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg (value)
    hist = []
    for _ in range(args.warmup):
        S.iterate()
        hist.append(S.monitor())
    ext = torch.cuda.ExternalStream(S.P.stream(), device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = S.P.launch_count()
    u0 = S.agent_updates
    e0.record(ext)
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        S.iterate()
        hist.append(S.monitor())
    e1.record(ext)
    barrier()
    wall_ms = (time.perf_counter() - tw0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = S.P.launch_count() - l0
    updates = S.agent_updates - u0
    t = torch.tensor([max(dev_ms, wall_ms), float(launches), float(updates)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        total_ms, launches, updates = float(tm[0]), int(ts[1]), int(ts[2])
    else:
        total_ms = float(t[0])
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    log(f"value leg: {ms_per_step:.1f} ms/step  cost {hist[0][0]:.6e} -> {hist[-1][0]:.6e}  "
        f"gradnorm {hist[0][1]:.3e} -> {hist[-1][1]:.3e}")

    # ---- end-to-end leg: host buffers in, host buffers out, every step
    e2e = None
    if not args.no_e2e:
        p = S.plan
        # two pinned host buffers: the iterate goes host -> device at the start of every step
        # and device -> host at its end; the buffers swap roles between steps (the host
        # side of a caller keeps the result where the library wrote it)
        hbuf = [torch.from_numpy(np.ascontiguousarray(S.get_X().T)).pin_memory(), None]
        hbuf[1] = torch.empty_like(hbuf[0]).pin_memory()
        res_h = torch.zeros(2, dtype=torch.float64).pin_memory()
        xdev = torch.empty_like(S.X)
        xh = hbuf[0]

        def e2e_step():
            xdev.copy_(hbuf[0], non_blocking=True)      # H2D of this step's input
            S.set_device_X(xdev)
            torch.cuda.current_stream().synchronize()
            S.iterate()
            f, g = S.monitor()
            hbuf[1].copy_(S.X, non_blocking=True)       # D2H of this step's result
            res_h[0], res_h[1] = f, g
            torch.cuda.current_stream().synchronize()
            hbuf[0], hbuf[1] = hbuf[1], hbuf[0]
        for _ in range(min(1, args.warmup)):
            e2e_step()
        barrier()
        tw0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        ms = (time.perf_counter() - tw0) * 1e3
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt[0]) / args.steps
        nbytes = xh.numel() * 8
        e2e = {"value": 1e3 / e2e_ms, "unit": "RBCD iterations/s",
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes + 16),
               "ms_per_step": e2e_ms}
        log(f"e2e leg: {e2e_ms:.1f} ms/step")

    # ---- roofline of the hot-path kernels (rank 0's local batch, device-resident,
    #      operands > L2: Q + X are far larger than 126 MB at the default size)
    peak, peak_src = peaks()
    kern = {}
    for kind in ("spmm", "precond", "project", "retract", "dot", "axpby", "rhess"):
        ms, by = S.P.time_kernel(kind, args.kernel_iters)
        kern[kind] = {"ms": ms, "bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    log("kernels: " + "  ".join(f"{k}={v['gbs']:.0f}GB/s({v['frac']:.2f})" for k, v in kern.items()))

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, gp, S, X0, log)

    # ---- time-to-certified-solution (BASELINE.json metric, second half): single-agent
    #      Riemannian staircase + certificate on a small range-aided grid, everything counted
    #      (assembly, Cholesky + tape build + upload, solve, certification)
    ttc = None
    if rank == 0 and args.ttc_grid > 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ttc as ttc_mod
            ttc = ttc_mod.run(args.ttc_grid, args.rank, args.range_fraction, args.seed)
            ttc.pop("levels", None)
            log(f"time-to-certified ({ttc['poses']} poses): {ttc['seconds']:.2f} s "
                f"(setup {ttc['factor_upload_s']:.2f} s, solve+certify {ttc['solve_certify_s']:.2f} s), "
                f"certified={ttc['certified']} rank={ttc['rank']}")
        except Exception as e:  # pragma: no cover
            ttc = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": "RBCD iterations/s",
            "value": value,
            "unit": "RBCD iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "impl": "ours",
            "config": {
                "workload": f"synthetic {args.grid}^3-pose 3D grid, {int(args.range_fraction * 100)}% range edges, "
                            f"{args.agents} agents, r={args.rank} (BASELINE.json configs[4])",
                "poses": int(meas.n), "pose_edges": int(len(meas.pp_i)),
                "range_edges": int(len(meas.rg_i)), "agents": args.agents, "r": args.rank,
                "K": int(gp.K), "nnz_Q": int(gp.Q().nnz), "colours": int(S.plan.num_colors),
                "parallelism": f"agents partitioned over {world} GPU(s), all-gather of public states per colour step",
                "l2": "operands (Q 0.76 GB + X) exceed the 126 MB L2; no flush needed",
                "ropt": "default ROptParameters (3 RTR x 50 tCG, tol 1e-2), TimeBound disabled",
                "init": "spanning-tree composition of measurements, lifted with [I;0]",
            },
            "agent_updates_per_s": updates / (total_ms / 1e3),
            "cost": {"first": 2 * hist[0][0], "last": 2 * hist[-1][0]},
            "gradnorm": {"first": hist[0][1], "last": hist[-1][1]},
            "setup_s": setup_s,
            "timing": {"device_ms": dev_ms, "wall_ms": wall_ms},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": {"kernel": "bsr_spmm_kernel (Q-SpMM, K1)", "bound": "hbm",
                         "achieved": kern["spmm"]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kern["spmm"]["frac"], "traffic": k1_traffic(args, world),
                         "peak_source": peak_src,
                         "algorithmic_bytes": kern["spmm"]["bytes"],
                         "launch_ms": kern["spmm"]["ms"]},
            "kernels": kern,
        }
        if ttc is not None:
            out["time_to_certified"] = ttc
        if e2e is not None:
            out["e2e"] = e2e
        if cpu is not None:
            out["cpu_baseline"] = cpu
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
def _oracle_blocks(gp, agents, r):
    """Oracle agent problems for a sample of agents (same Q blocks)."""
    import scipy.sparse as sp
    from oracle import manifold as mf, rbcd as orbcd
    diag, cross = gp.split(agents=agents)
    blocks = {}
    for a in agents:
        L = mf.Layout(r, gp.d, int(gp.n_a[a]), int(gp.l_a[a]), int(gp.b_a[a]), ra=gp.ra)
        rowptr, col, val = diag[a]
        Q = sp.csr_matrix((val, col, rowptr), shape=(L.k, L.k))
        blocks[a] = orbcd.AgentBlock(L, Q, int(gp.coff[a]), reg=0.1)
    return blocks, cross


def _cpu_agent_update(args_tuple):
    """One Agent::updateX on the CPU oracle (worker for the reference arm)."""
    (rowptr, col, val, Ga, Xa, r, d, n, l, b, ra, reg) = args_tuple
    import scipy.sparse as sp
    from oracle import manifold as mf, solver
    try:
        # one BLAS thread per worker: the pool already uses every core (a forked worker
        # inherits the parent's multi-threaded BLAS pool and would oversubscribe 8x)
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    L = mf.Layout(r, d, n, l, b, ra=ra)
    Q = sp.csr_matrix((val, col, rowptr), shape=(L.k, L.k))
    P = solver.QuadraticProblem(L, Q, G=Ga, precond_reg=reg)
    P._factor()                       # set-up (excluded from the timed update)
    t0 = time.perf_counter()
    opt = solver.QuadraticOptimizer(P, solver.ROptParameters())
    Xn = opt.optimize(Xa)
    return time.perf_counter() - t0, opt.result.inner_iters


def _reg_for(gp, S, a):
    if S is not None and S.plan.a0 <= a < S.plan.a1:
        return S.P.precond_reg(a - S.plan.a0)
    return 0.1


def cpu_sample(args, gp, X, agents, S=None, procs=1):
    """Time Agent::updateX of the oracle for a sample of agents.  Returns the
    per-update seconds list."""
    diag, cross = gp.split(agents=agents)
    Ct = cross.T.tocsr()
    jobs = []
    for a in agents:
        c0, c1 = int(gp.coff[a]), int(gp.coff[a + 1])
        Ga = (Ct[c0:c1] @ X.T).T
        rowptr, col, val = diag[a]
        jobs.append((rowptr, col, val, Ga, np.ascontiguousarray(X[:, c0:c1]), args.rank, gp.d,
                     int(gp.n_a[a]), int(gp.l_a[a]), int(gp.b_a[a]), gp.ra, _reg_for(gp, S, a)))
    if procs <= 1:
        return [_cpu_agent_update(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_cpu_agent_update, jobs)


def cpu_baseline(args, gp, S, X0, log):
    """Oracle port, ONE thread, bounded sample: `cpu_sample_agents` agent
    updates of the same workload; extrapolated to a full iteration
    (= `agents` updates)."""
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:
        import contextlib
        ctx = contextlib.nullcontext()
    sample = list(np.linspace(0, args.agents - 1, args.cpu_sample_agents).astype(int))
    sample = sorted(set(int(a) for a in sample))
    with ctx:
        res = cpu_sample(args, gp, X0, sample, S=S, procs=1)
    per = float(np.mean([t for t, _ in res]))
    it_s = 1.0 / (per * args.agents)
    log(f"cpu baseline: {per:.2f} s per agent update (1 thread) -> {it_s:.5f} iterations/s")
    return {"value": it_s, "unit": "RBCD iterations/s", "cores": 1, "kind": "port",
            "sample": f"{len(sample)} of {args.agents} agent updates (agents {sample}) of the same "
                      f"workload from the same initial iterate, numpy/scipy oracle, 1 thread; "
                      f"one iteration = {args.agents} updates; factorisation excluded",
            "seconds_per_agent_update": per}


def run_reference(args):
    """The reference's CPU path (oracle port: the reference itself cannot be
    built here or on the GPU box -- Eigen/SuiteSparse/ROPTLIB/Spectra absent),
    all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return

    def log(msg):
        print("[bench-ref] " + msg, file=sys.stderr, flush=True)

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    meas, gt, idx, gp = build_problem(args, log)
    X0 = initial_X(args, meas, idx, gp)
    cores = os.cpu_count() or 1
    per_step = max(cores, 2)
    times = []
    total = args.warmup + args.steps
    stride = max(1, args.agents // (per_step * total))
    for s in range(total):
        agents = [(s * per_step + j) * stride % args.agents for j in range(per_step)]
        agents = sorted(set(agents))
        t0 = time.perf_counter()
        res = cpu_sample(args, gp, X0, agents, procs=cores)
        wall = time.perf_counter() - t0
        busy = float(np.sum([t for t, _ in res]))
        # throughput of the pool while it was solving: updates / (sum of solve time / cores)
        times.append(busy / cores / len(agents))
        log(f"step {s}: {len(agents)} updates, {busy / len(agents):.2f} s each, wall {wall:.1f}s")
    per = float(np.mean(times[args.warmup:]))
    ms_per_step = per * args.agents * 1e3
    value = 1e3 / ms_per_step
    out = {
        "metric": "RBCD iterations/s", "value": value, "unit": "RBCD iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {
            "workload": f"synthetic {args.grid}^3-pose 3D grid, {int(args.range_fraction * 100)}% range edges, "
                        f"{args.agents} agents, r={args.rank} (BASELINE.json configs[4])",
            "poses": int(meas.n), "agents": args.agents, "r": args.rank},
        "cpu_baseline": {"value": value, "unit": "RBCD iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} agent updates per step spread over {cores} processes "
                                   f"(one agent per core), extrapolated to {args.agents} updates per "
                                   f"iteration; numpy/scipy oracle; factorisation excluded"},
        "e2e": {"value": value, "unit": "RBCD iterations/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(out)


class _StdoutToStderr:
    """Everything written to fd 1 while the benchmark runs (e.g. the NCCL version banner
    of the first communicator) goes to stderr, so that stdout carries exactly ONE JSON
    line; `emit` restores fd 1 for that line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


_OUT = None


def emit_json(obj):
    text = json.dumps(obj)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


if __name__ == "__main__":
    a = parse()
    _OUT = _StdoutToStderr()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
