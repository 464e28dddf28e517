"""CPU run of the multi-robot RBCD loop on the native oracle (TEST INFRASTRUCTURE).

The reference's driver (examples/MultiRobotExample.cpp:223-307) with the agents' updates
done by oracle/native (C++ -O3: Cholesky preconditioner factored once per agent, RTR/tCG
per update): the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).  The agents of
one colour of the robot graph share no measurement, so their updates are independent; they
are dealt to `threads` host threads (every update itself is single-threaded, like the
reference).  Same colouring, same parameters and same start as the GPU arm: the two arms
walk the same iterates.
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import manifold as mf, native


def ra_regularization(Q, ra):
    """Graph::computePreconditionerRegularization (src/Graph.cpp:1919-1960): 0.1 for pose
    graphs, lambda_max / (1e6 - 1) for range-aided graphs (fallback 0.1)."""
    if not ra:
        return 0.1
    try:
        lm = float(spla.eigsh(Q, k=1, which="LA", tol=1e-3, ncv=6, return_eigenvectors=False)[0])
        return lm / (1e6 - 1.0) if lm > 0 else 0.1
    except Exception:
        return 0.1


class CpuRBCD:
    def __init__(self, gp, r, threads=1, agents=None, log=None, ropt=None):
        """gp: the partitioned problem (column bookkeeping + global Q); `agents`: the
        agents to prepare (default all)."""
        from dcora_b200.graph import greedy_coloring   # host-side colouring shared by both arms
        self.gp, self.r, self.threads = gp, r, max(1, int(threads))
        self.ropt = dict(max_outer=3, max_inner=50, tol=1e-2, delta0=100.0)
        if ropt:
            self.ropt.update(ropt)
        t0 = time.perf_counter()
        which = list(range(gp.A)) if agents is None else [int(a) for a in agents]
        diag, cross = gp.split(agents=which)
        self.cross_t = cross.T.tocsr()
        self.colors = greedy_coloring(gp.agent_graph(cross))
        self.layout = {a: mf.Layout(r, gp.d, int(gp.n_a[a]), int(gp.l_a[a]), int(gp.b_a[a]),
                                    ra=gp.ra) for a in which}
        self.pool = ThreadPoolExecutor(self.threads)

        def prepare(a):
            rowptr, col, val = diag[a]
            L = self.layout[a]
            Q = sp.csr_matrix((val, col, rowptr), shape=(L.k, L.k))
            return a, native.AgentProblem(L, Q, ra_regularization(Q, gp.ra))
        self.problem = dict(self.pool.map(prepare, which))
        self.setup_s = time.perf_counter() - t0
        if log:
            log(f"cpu arm: {len(which)} agents factored in {self.setup_s:.1f} s on {self.threads} threads")
        self.stats = dict(inner=0, outer=0, precond=0, updates=0, solve_s=0.0)

    def set_X(self, X):
        self.X = np.asfortranarray(np.array(X, dtype=np.float64))

    def linear_terms(self, agents):
        """G of every agent in `agents` from the current iterate (Graph::constructLinearCostTerm)."""
        gp = self.gp
        rows = np.concatenate([np.arange(gp.coff[a], gp.coff[a + 1]) for a in agents])
        G = np.asarray(self.cross_t[rows] @ self.X.T).T          # (r, sum k_a)
        out, off = {}, 0
        for a in agents:
            k = int(gp.coff[a + 1] - gp.coff[a])
            out[a] = np.asfortranarray(G[:, off:off + k])
            off += k
        return out

    def update_agents(self, agents):
        """Agent::updateX for every agent in `agents` (mutually non-adjacent)."""
        gp = self.gp
        G = self.linear_terms(agents)

        def one(a):
            c0, c1 = int(gp.coff[a]), int(gp.coff[a + 1])
            return a, self.problem[a].update(G[a], self.X[:, c0:c1], **self.ropt)
        for a, (Xa, st) in self.pool.map(one, agents):
            self.X[:, int(gp.coff[a]):int(gp.coff[a + 1])] = Xa
            self.stats["inner"] += int(st["inner"]); self.stats["outer"] += int(st["outer"])
            self.stats["precond"] += int(st["precond"]); self.stats["updates"] += 1
            self.stats["solve_s"] += float(st["solve_s"])

    def iterate(self):
        """One RBCD iteration = every agent once, colour by colour."""
        for c in range(int(self.colors.max()) + 1):
            agents = [a for a in np.nonzero(self.colors == c)[0] if a in self.problem]
            if agents:
                self.update_agents(agents)

    def monitor(self):
        """Global cost and Riemannian gradient norm (the centralised evaluation of
        MultiRobotExample.cpp:264-281)."""
        gp = self.gp
        XQ = np.asarray(gp.Q() @ self.X.T).T
        f = 0.5 * float(np.sum(XQ * self.X))

        def gn2(a):
            c0, c1 = int(gp.coff[a]), int(gp.coff[a + 1])
            L = self.layout[a]
            rg = mf.tangent_project(L, self.X[:, c0:c1], XQ[:, c0:c1])
            return float(np.sum(rg * rg))
        g2 = sum(self.pool.map(gn2, list(self.problem)))
        return f, float(np.sqrt(g2))
