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
    ap.add_argument("--cpu-sample-agents", type=int, default=32)
    ap.add_argument("--kernel-iters", type=int, default=20)
    ap.add_argument("--ttc-agents", type=int, default=16)
    ap.add_argument("--ttc-max-iters", type=int, default=3000)
    ap.add_argument("--ttc-grid", type=int, default=47,
                    help="side S of the S^3-pose grid of the time-to-certified leg (0: skip); 47^3 = 103823 poses")
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


def kernel_traffic(name, args, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one whole-batch launch of a kernel on
    the default workload, from the committed ncu --set full capture
    (profiles/<name>_traffic.json); null for any other workload / GPU count (no capture)."""
    path = os.path.join(ROOT, "profiles", name + "_traffic.json")
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


def workload_config(args, meas, gp, ncolors):
    """The `config` object: identical in both arms (same workload, same algorithm)."""
    return {
        "workload": f"synthetic {args.grid}^3-pose 3D grid, {int(args.range_fraction * 100)}% range edges, "
                    f"{args.agents} agents, r={args.rank} (BASELINE.json configs[4])",
        "poses": int(meas.n), "pose_edges": int(len(meas.pp_i)),
        "range_edges": int(len(meas.rg_i)), "agents": args.agents, "r": args.rank,
        "K": int(gp.K), "nnz_Q": int(gp.Q().nnz), "colours": int(ncolors),
        "step": "one RBCD iteration: every agent once (colour by colour of the robot graph: "
                "exchange public states, rebuild G, one QuadraticOptimizer::optimize) + the "
                "global cost / gradient norm (examples/MultiRobotExample.cpp:223-307)",
        "ropt": "default ROptParameters (3 RTR x 50 tCG, tol 1e-2), TimeBound disabled",
        "init": "spanning-tree composition of measurements, lifted with [I;0]",
        "l2": "operands (Q 0.76 GB, factor 5.9 GB, X) exceed the 126 MB L2; no flush needed",
    }


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

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(vs):
        t = torch.tensor([float(v) for v in vs], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    # ---- warm-up, then the iterate both timed legs start from (kept in pinned host memory)
    hist = []
    for _ in range(args.warmup):
        S.iterate()
        hist.append(S.monitor())
    p = S.plan
    hstart = torch.empty((p.K, args.rank), dtype=torch.float64).pin_memory()
    hwork = torch.empty((p.K, args.rank), dtype=torch.float64).pin_memory()
    hstart_np, hwork_np = hstart.numpy().T, hwork.numpy().T     # (r, K) column-major views
    S.store_host_X(hstart_np)

    # ---- device-resident leg (value)
    ext = torch.cuda.ExternalStream(S.P.stream(), device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0, u0, i0, o0 = S.P.launch_count(), S.agent_updates, S.inner_iterations, S.outer_iterations
    e0.record(ext)
    tw0 = time.perf_counter()
    vhist = []
    for _ in range(args.steps):
        S.iterate()
        vhist.append(S.monitor())
    e1.record(ext)
    barrier()
    wall_ms = (time.perf_counter() - tw0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    total_ms = allmax(max(dev_ms, wall_ms))
    launches, updates, inner, outer = allsum([S.P.launch_count() - l0, S.agent_updates - u0,
                                              S.inner_iterations - i0, S.outer_iterations - o0])
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    hist += vhist
    log(f"value leg: {ms_per_step:.1f} ms/step  cost {hist[0][0]:.6e} -> {hist[-1][0]:.6e}  "
        f"gradnorm {hist[0][1]:.3e} -> {hist[-1][1]:.3e}  tCG iterations/step {inner / args.steps:.0f}")

    # ---- end-to-end leg: the same steps from the same iterate, but the iterate lives in
    #      (pinned) host memory between steps: uploaded through the C-ABI at the start of
    #      every step, read back with the step's cost / gradient norm at its end
    e2e = None
    if not args.no_e2e:
        hwork.copy_(hstart)
        res_h = np.zeros(2)
        ehist = []

        def e2e_step():
            S.load_host_X(hwork_np)                     # H2D of this step's input
            S.iterate()
            res_h[:] = S.monitor()
            S.store_host_X(hwork_np)                    # D2H of this step's result
            ehist.append(tuple(res_h))
        barrier()
        tw0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_ms = allmax((time.perf_counter() - tw0) * 1e3) / args.steps
        nbytes = hwork.numel() * 8
        same = all(abs(a[0] - b[0]) <= 1e-9 * abs(b[0]) for a, b in zip(ehist, vhist))
        e2e = {"value": 1e3 / e2e_ms, "unit": "RBCD iterations/s",
               "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes + 16),
               "ms_per_step": e2e_ms, "same_iterates_as_value_leg": bool(same)}
        log(f"e2e leg: {e2e_ms:.1f} ms/step (same trajectory as the value leg: {same})")

    # ---- roofline of the hot-path kernels (rank 0's local batch, device-resident,
    #      operands > L2: Q, the factor tapes and X are far larger than 126 MB at the default size)
    peak, peak_src = peaks()
    kern = {}
    for kind in ("spmm", "precond", "project", "retract", "dot", "axpby", "rhess"):
        ms, by = S.P.time_kernel(kind, args.kernel_iters)
        kern[kind] = {"ms": ms, "bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    log("kernels: " + "  ".join(f"{k}={v['gbs']:.0f}GB/s({v['frac']:.2f})" for k, v in kern.items()))

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        Xs = np.empty((args.rank, gp.K), order="F") if world == 1 else None
        if world == 1:
            Xs[:] = hstart_np
        cpu = cpu_baseline(args, gp, Xs if Xs is not None else X0, log)

    # ---- time-to-certified-solution (BASELINE.json metric, second half): the multi-agent
    # staircase (RBCD++ -> distributed certificate -> escape -> next rank; acceleration on like
    # the reference drivers, MultiRobotExample.cpp:121, with the number of colours as the block
    # count of the Nesterov sequences because a whole colour class moves per iteration) on a
    # partitioned range-aided grid, wall clock including assembly, factorisation and upload; the agents
    # are dealt to all ranks of the job (collective)
    ttc = None
    if args.ttc_grid > 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import ttc_multi
            barrier()
            ttc = ttc_multi.run(args.ttc_grid, args.ttc_agents, args.rank, args.range_fraction,
                                args.seed, acc=True, gradnorm_tol=0.1, eta=1e-3,
                                max_iters=args.ttc_max_iters, check_every=10,
                                log=None, dist_rank=rank, world=world, device=local_rank,
                                nesterov_blocks="colours", restart_interval="auto")
            barrier()
            ttc["levels"] = [{k: v for k, v in lv.items() if k in
                              ("r", "sweeps", "gradnorm", "lambda_min", "residual", "is_psd", "converged", "solve_s", "certify_s")}
                             for lv in ttc["levels"]]
            log(f"time-to-certified ({ttc['poses']} poses, {ttc['agents']} agents): "
                f"{ttc['seconds']:.2f} s, certified={ttc['certified']} rank={ttc['rank']} "
                f"gradnorm={ttc['gradnorm']:.3g}")
        except Exception as e:  # pragma: no cover
            ttc = {"error": repr(e)}

    if rank == 0:
        cfg = workload_config(args, meas, gp, S.plan.num_colors)
        dom = "precond"
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
            "config": cfg,
            "parallelism": f"agents partitioned over {world} GPU(s), all-gather of public states per colour step",
            "agent_updates_per_s": updates / (total_ms / 1e3),
            "inner_iterations_per_step": inner / args.steps,
            "outer_iterations_per_step": outer / args.steps,
            "k5_solves_per_step": inner / args.steps,
            "k5_solves_note": "agent-level preconditioner applications: one per tCG iteration "
                              "(the initial z = P^-1 r of a tCG run + one per continued iteration)",
            "cost": {"first": 2 * hist[0][0], "last": 2 * hist[-1][0]},
            "gradnorm": {"first": hist[0][1], "last": hist[-1][1]},
            "setup_s": setup_s,
            "timing": {"device_ms": dev_ms, "wall_ms": wall_ms},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": {"kernel": "trsv_tape_kernel (K5, preconditioner solve: the dominant kernel "
                                   "of an RBCD iteration), whole batch in one launch",
                         "bound": "hbm",
                         "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kern[dom]["frac"], "traffic": kernel_traffic("k5", args, world),
                         "peak_source": peak_src,
                         "algorithmic_bytes": kern[dom]["bytes"],
                         "launch_ms": kern[dom]["ms"]},
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
def cpu_baseline(args, gp, X, log):
    """The CPU path (oracle/native: C++ -O3 port of Agent::updateX, `kind: port`), ONE
    thread, on a bounded sample of the same workload: `cpu_sample_agents` agent updates from
    the iterate the timed GPU legs start from; one RBCD iteration = `agents` updates."""
    from oracle.cpu_rbcd import CpuRBCD
    sample = sorted(set(int(a) for a in np.linspace(0, args.agents - 1, args.cpu_sample_agents)))
    C = CpuRBCD(gp, args.rank, threads=1, agents=sample, log=log)
    C.set_X(X)
    t0 = time.perf_counter()
    for a in sample:                 # one at a time: each sees the others' previous state
        C.update_agents([a])
    dt = time.perf_counter() - t0
    per = dt / len(sample)
    it_s = 1.0 / (per * args.agents)
    log(f"cpu baseline: {per * 1e3:.0f} ms per agent update (1 thread, native oracle) -> "
        f"{it_s:.5f} iterations/s")
    return {"value": it_s, "unit": "RBCD iterations/s", "cores": 1, "kind": "port",
            "sample": f"{len(sample)} of {args.agents} agent updates (agents {sample[0]}..{sample[-1]}, evenly "
                      f"spaced) from the iterate the timed legs start from; C++ -O3 oracle "
                      f"(oracle/native: CSR SpMM, sparse Cholesky solve, RTR/tCG), 1 thread; one "
                      f"iteration = {args.agents} updates; factorisation and the per-iteration "
                      f"global cost evaluation excluded",
            "seconds_per_agent_update": per,
            "inner_iterations_per_update": C.stats["inner"] / max(1, C.stats["updates"])}


def run_reference(args):
    """The reference's CPU path -- the C++ -O3 oracle port of Agent::updateX driven by the
    reference's loop (the reference itself cannot be built here or on the GPU box:
    Eigen/SuiteSparse/ROPTLIB/Spectra absent) -- on all host cores, FULL steps of the same
    workload, same colouring / parameters / start as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return

    def log(msg):
        print("[bench-ref] " + msg, file=sys.stderr, flush=True)

    from oracle.cpu_rbcd import CpuRBCD
    meas, gt, idx, gp = build_problem(args, log)
    X0 = initial_X(args, meas, idx, gp)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    C = CpuRBCD(gp, args.rank, threads=cores, log=log)
    C.set_X(X0)
    hist, times = [], []
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            base = dict(C.stats)
        t0 = time.perf_counter()
        C.iterate()
        hist.append(C.monitor())
        times.append(time.perf_counter() - t0)
        log(f"step {s}: {times[-1]:.2f} s  cost {2 * hist[-1][0]:.6e}  gradnorm {hist[-1][1]:.3e}")
    per = float(np.mean(times[args.warmup:]))
    ms_per_step = per * 1e3
    value = 1.0 / per
    st = {k: C.stats[k] - base[k] for k in base}
    out = {
        "metric": "RBCD iterations/s", "value": value, "unit": "RBCD iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, meas, gp, int(C.colors.max()) + 1),
        "inner_iterations_per_step": st["inner"] / args.steps,
        "outer_iterations_per_step": st["outer"] / args.steps,
        "k5_solves_per_step": st["precond"] / args.steps,
        "cost": {"first": 2 * hist[0][0], "last": 2 * hist[-1][0]},
        "gradnorm": {"first": hist[0][1], "last": hist[-1][1]},
        "setup_s": C.setup_s,
        "step_seconds": times,
        "cpu_baseline": {"value": value, "unit": "RBCD iterations/s", "cores": cores, "kind": "port",
                         "sample": f"full RBCD iterations ({args.agents} agent updates + global cost / "
                                   f"gradient norm each): C++ -O3 oracle (oracle/native) per agent "
                                   f"update, the agents of a colour dealt to {cores} threads"},
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
