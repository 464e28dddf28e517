#!/usr/bin/env python
"""Time to a certified solution of a partitioned range-aided grid: RBCD(++) sweeps ->
distributed certificate -> escape -> next rank (dcora_b200.staircase.rbcd_staircase).
    python tools/ttc_multi.py --grid 47 --agents 16 [--no-acc]
Under torchrun the agents are dealt to the ranks (one GPU each)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def run(grid, agents, rank=5, range_fraction=0.1, seed=0, acc=True, gradnorm_tol=0.1, eta=1e-3,
        max_iters=1000, check_every=5, log=print, dist_rank=0, world=1, group=None, device=0,
        r_max=None, nesterov_blocks=None, restart_interval=30, certify_max_iters=400,
        agent_tol=None, stall_checks=0, adaptive_restart=False, stable_decrease=False):
    import dcora_b200 as dc
    from dcora_b200 import graph as pg, rbcd, synthetic
    from dcora_b200.staircase import rbcd_staircase
    t0 = time.perf_counter()
    meas, gt, idx = synthetic.grid3d(grid, range_fraction, seed=seed)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, agents))
    X0 = synthetic.lifted_initial_guess(meas, idx, gp, grid, rank)
    t1 = time.perf_counter()
    # every agent stops its own RTR at `agent_tol` (reference default 1e-2, QuadraticOptimizer):
    # the global gradient norm cannot fall below ~sqrt(A) * agent_tol, so many agents need a
    # proportionally smaller one to reach the drivers' RGradNormTol
    if agent_tol is None:
        agent_tol = min(1e-2, 0.5 * gradnorm_tol / np.sqrt(agents))
    S = rbcd.RBCD(gp, rank, rank=dist_rank, world=world, device=device, group=group,
                  params=dc.ROptParameters(time_bound_s=0.0, gradnorm_tol=agent_tol,
                                           stable_decrease=stable_decrease), acceleration=acc,
                  nesterov_blocks=nesterov_blocks, restart_interval=restart_interval)
    t2 = time.perf_counter()
    res = rbcd_staircase(S, X0, r_max=r_max or rank + 3, eta=eta, gradnorm_tol=gradnorm_tol,
                         max_iters=max_iters, check_every=check_every, log=log,
                         certify_max_iters=certify_max_iters, stall_checks=stall_checks,
                         adaptive_restart=adaptive_restart)
    t3 = time.perf_counter()
    return {"workload": f"synthetic {grid}^3-pose 3D grid, {int(range_fraction * 100)}% range edges, "
                        f"{agents} agents on {world} GPU(s), RBCD{'++' if acc else ''}"
                        f"{' (N = colours, restart every ' + str(restart_interval) + ' iterations)' if acc and nesterov_blocks else ''}, r0={rank}, "
                        f"gradnorm_tol={gradnorm_tol}, eta={eta}",
            "poses": int(meas.n), "K": int(gp.K), "agents": agents, "agent_gradnorm_tol": agent_tol,
            "certified": res.certified,
            "rank": res.rank, "cost_2f": 2 * res.cost, "gradnorm": res.gradnorm,
            "lambda_min": res.lambda_min, "assemble_s": t1 - t0, "factor_upload_s": t2 - t1,
            "solve_certify_s": t3 - t2, "seconds": t3 - t0, "levels": res.levels}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=47)
    ap.add_argument("--agents", type=int, default=16)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--tol", type=float, default=0.1)
    ap.add_argument("--max-iters", type=int, default=1000)
    ap.add_argument("--check-every", type=int, default=5)
    ap.add_argument("--no-acc", action="store_true")
    ap.add_argument("--nesterov-blocks", default=None, help="'colours' or an integer (default: agents)")
    ap.add_argument("--restart-interval", default="30", help="iterations (colour steps) or 'auto'")
    ap.add_argument("--certify-max-iters", type=int, default=400)
    ap.add_argument("--stall-checks", type=int, default=0)
    ap.add_argument("--adaptive-restart", action="store_true")
    ap.add_argument("--stable-decrease", action="store_true")
    ap.add_argument("--range-fraction", type=float, default=0.1,
                    help="0: pose graph only (the synthetic stand-in of the absent g2o100k, configs[3])")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    log = (lambda m: print("[ttc] " + m, file=sys.stderr, flush=True)) if rk == 0 else None
    out = run(a.grid, a.agents, a.rank, range_fraction=a.range_fraction, acc=not a.no_acc, gradnorm_tol=a.tol, max_iters=a.max_iters,
              check_every=a.check_every, log=log, dist_rank=rk, world=world, device=dev,
              nesterov_blocks=a.nesterov_blocks,
              restart_interval=a.restart_interval if a.restart_interval == "auto" else int(a.restart_interval),
              certify_max_iters=a.certify_max_iters, stall_checks=a.stall_checks,
              adaptive_restart=a.adaptive_restart, stable_decrease=a.stable_decrease)
    if rk == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
