#!/usr/bin/env python
"""Slot trace of the preconditioner kernel (K5) on the bench workload: prints the
duration of every barrier slot of one agent's solve -- alone on the GPU, with the
agents of its colour class active (what an RBCD colour step runs), and with the
whole batch active.

    python tools/trsv_trace.py [--grid 100 --agents 256 --agent 128]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--agents", type=int, default=256)
    ap.add_argument("--agent", type=int, default=128)
    ap.add_argument("--rank", type=int, default=5)
    a = ap.parse_args()
    import dcora_b200 as dc
    from dcora_b200 import _lib, graph as pg, rbcd, synthetic
    lib = _lib.load()
    meas, gt, idx = synthetic.grid3d(a.grid, 0.1, seed=0)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, a.agents))
    S = rbcd.RBCD(gp, a.rank, rank=0, world=1, device=0, params=dc.ROptParameters(time_bound_s=0.0))
    n = 1024
    alone = np.zeros(a.agents, np.uint8)
    colour = (S.local_colors == S.local_colors[a.agent]).astype(np.uint8)
    for name, act in (("alone", alone), (f"its colour class ({int(colour.sum())} agents)", colour),
                      ("whole batch", None)):
        st = (C.c_uint64 * n)()
        ptr = act.ctypes.data_as(_lib.c_uint8_p) if act is not None else None
        _lib.check(lib.dcora_debug_trsv_trace(S.P.h, a.agent, ptr, st, n))
        t = np.array(list(st), dtype=np.uint64)
        it = t[64:64 + 480].astype(np.int64)
        t = t[:64]
        end = int(np.argmax(t[1:] == 0)) + 1
        t = t[:end].astype(np.int64)
        dt = np.diff(t) / 1e3
        print(f"--- agent {a.agent}, {name}: total {(t[-1] - t[0]) / 1e3:.1f} us, "
              f"{len(dt) - 2} slots")
        print(f"  rhs load {dt[0]:.1f} | slots " + " ".join(f"{d:.1f}" for d in dt[1:-1]) +
              f" | write-back {dt[-1]:.1f}")
        if it[0] != 0:
            q = it[:480].reshape(-1, 4)
            q = q[q[:, 0] != 0]
            print("  iterations of warp 0 (cycles): wait | loads+fma | refill | gap to next")
            for k in range(min(len(q), 60)):
                gap = q[k + 1, 0] - q[k, 3] if k + 1 < len(q) else 0
                print(f"   {q[k,1]-q[k,0]:6d} {q[k,2]-q[k,1]:6d} {q[k,3]-q[k,2]:6d} {gap:6d}")


if __name__ == "__main__":
    main()
