#!/usr/bin/env python
"""Full-batch hot-path kernels of the bench workload in isolation (for ncu captures and
kernel A/B runs):  python tools/kprofile.py [--grid S] [--agents A] [--iters N] kind [kind ...]
Prints one line per kernel kind with the CUDA-event time and algorithmic GB/s."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--agents", type=int, default=256)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--range-fraction", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rbcd-steps", type=int, default=0, help="RBCD iterations to run first")
    ap.add_argument("--split", action="store_true", help="per-colour wall-clock split of a sweep")
    ap.add_argument("kinds", nargs="*", default=["spmm", "precond"])
    a = ap.parse_args()
    import dcora_b200 as dc
    from dcora_b200 import rbcd
    log = lambda m: print("[kprofile] " + m, file=sys.stderr, flush=True)
    meas, gt, idx, gp = bench.build_problem(a, log)
    S = rbcd.RBCD(gp, a.rank, params=dc.ROptParameters(time_bound_s=0.0))
    S.set_X(bench.initial_X(a, meas, idx, gp))
    import time, torch
    for _ in range(a.rbcd_steps):
        t0 = time.perf_counter(); S.iterate(); torch.cuda.synchronize()
        log(f"rbcd step {1e3 * (time.perf_counter() - t0):.1f} ms")
    if a.split:
        # wall-clock split of one sweep per colour (synchronised: profiling aid)
        import numpy as np
        for c in range(S.plan.num_colors):
            act = (S.local_colors == c).astype(np.uint8)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            S.exchange(); S.P.update_G(S.X, S.H, act)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            S.X = S.opt.optimize(S.X, act)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            log(f"colour {c}: {int(act.sum())} agents  update_G {1e3 * (t1 - t0):.2f} ms  "
                f"optimize {1e3 * (t2 - t1):.2f} ms")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        S.monitor(); torch.cuda.synchronize()
        log(f"monitor {1e3 * (time.perf_counter() - t0):.2f} ms")
    peak, _ = bench.peaks()
    for kind in a.kinds:
        ms, by = S.P.time_kernel(kind, a.iters)
        print(json.dumps({"kernel": kind, "ms": ms, "bytes": by, "gbs": by / ms / 1e6,
                          "frac": by / ms / 1e6 / peak}), flush=True)


if __name__ == "__main__":
    main()
