#!/usr/bin/env python
"""Time-to-certified-solution of a synthetic range-aided 3-D grid, single agent:
python tools/ttc.py --grid 20"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def run(grid, rank=5, range_fraction=0.1, seed=0, tol=1e-4, eta=1e-3, log=print):
    import dcora_b200 as dc
    from dcora_b200 import graph as pg, synthetic
    from dcora_b200.staircase import riemannian_staircase
    t0 = time.perf_counter()
    meas, gt, idx = synthetic.grid3d(grid, range_fraction, seed=seed)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, 1))
    Q = gp.Q()
    t1 = time.perf_counter()
    X0 = synthetic.lifted_initial_guess(meas, idx, gp, grid, rank)
    ag = dc.AgentData(n=int(gp.n_a[0]), l=int(gp.l_a[0]), b=int(gp.b_a[0]), Q=Q,
                      precond_reg="auto_ra" if gp.ra else 0.1)
    P = dc.QuadraticProblem(ag, rank, gp.d, dc.ORDER_RA if gp.ra else dc.ORDER_SE)
    t2 = time.perf_counter()
    prm = dc.ROptParameters(RTR_iterations=200, RTR_tCG_iterations=200, gradnorm_tol=tol,
                            time_bound_s=0.0)
    X0 = P.project(X0)
    res = riemannian_staircase(P, X0, r_max=rank + 3, eta=eta, params=prm)
    t3 = time.perf_counter()
    out = {"workload": f"synthetic {grid}^3-pose 3D grid, {int(range_fraction * 100)}% range edges, "
                       f"1 agent, r0={rank}, gradnorm_tol={tol}, eta={eta}",
           "poses": int(meas.n), "K": int(gp.K), "certified": res.certified, "rank": res.rank,
           "cost_2f": 2 * res.cost, "gradnorm": res.gradnorm, "lambda_min": res.lambda_min,
           "assemble_s": t1 - t0, "factor_upload_s": t2 - t1, "solve_certify_s": t3 - t2,
           "seconds": t3 - t0, "levels": res.levels}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=20)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--tol", type=float, default=1e-4)
    a = ap.parse_args()
    print(json.dumps(run(a.grid, a.rank, tol=a.tol)))
