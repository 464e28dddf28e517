#!/usr/bin/env python
"""Phase trace of the preconditioner kernel (K5) on the bench workload: prints
the duration of every phase A / phase B of one agent's forward and backward
walk, alone on the GPU and with every agent of the batch active.

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
    n = 256
    for all_agents in (0, 1):
        st = (C.c_uint64 * n)()
        _lib.check(lib.dcora_debug_trsv_trace(S.P.h, a.agent, all_agents, st, n))
        t = np.array(list(st), dtype=np.uint64)
        end = int(np.argmax(t[1:] == 0)) + 1
        t = t[:end].astype(np.int64)
        kinds = ["A" if (x & 1) else "B" for x in t[1:]]
        dt = np.diff(t) / 1e3
        print(f"--- agent {a.agent}, {'all agents active' if all_agents else 'alone'}: "
              f"total {(t[-1] - t[0]) / 1e3:.1f} us over {len(dt)} phases")
        print("  A total %.1f us, B total %.1f us" % (
            sum(d for d, k in zip(dt, kinds) if k == "A"), sum(d for d, k in zip(dt, kinds) if k == "B")))
        print("  " + " ".join(f"{k}{d:.1f}" for d, k in zip(dt, kinds)))


if __name__ == "__main__":
    main()
