"""Oracle: multi-robot RBCD / RBCD++ loop (TEST INFRASTRUCTURE).

Restates, on the CPU with numpy/scipy,
  * the driver loop of examples/MultiRobotExample.cpp:223-307 (exchange public
    states -> selected agent iterate(true), others iterate(false) -> global
    cost / gradient norm),
  * Agent::iterate / updateX (src/Agent.cpp:535-596,1216-1278),
  * the Nesterov bookkeeping updateGamma / updateAlpha / updateY / updateV /
    shouldRestart / restartNesterovAcceleration (:1158-1214).

Agents are processed ONE AT A TIME, exactly like the reference.  `schedule`
is the list of agent groups selected per iteration; a group of mutually
non-adjacent agents (one colour of the robot graph) is equivalent to selecting
them one after another, because their subproblems do not see each other.

The linear term of agent a is G_a = (X * C)[:, cols_a] with C the cross-agent
part of the global data matrix; tests/test_host_graph.py pins this product
against the per-measurement restatement of
Graph::constructLinearCostTerm{PGO,RASLAM} (oracle/graph.py: linear_matrix).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import manifold as mf, solver


class AgentBlock:
    def __init__(self, layout: mf.Layout, Q, c0, reg=0.1, use_precond=True):
        self.L = layout
        self.c0, self.c1 = c0, c0 + layout.k
        self.P = solver.QuadraticProblem(layout, Q, precond_reg=reg, use_precond=use_precond)


class RBCD:
    def __init__(self, blocks, cross, params: solver.ROptParameters | None = None,
                 acceleration=False, restart_interval=30, nesterov_blocks=None):
        self.blocks = blocks
        self.C = sp.csr_matrix(cross)
        self.Ct = self.C.T.tocsr()        # row-slicing by output column
        self.params = params or solver.ROptParameters()
        self.acc = acceleration
        self.restart_interval = restart_interval
        self.A = len(blocks)
        # N of updateGamma / updateAlpha: number of agents (reference), or the number of
        # blocks of the schedule when a whole colour class moves per iteration
        self.N = int(nesterov_blocks) if nesterov_blocks else len(blocks)
        self.it = 0
        self.results = [None] * self.A

    def set_X(self, X):
        self.X = np.array(X, dtype=np.float64, copy=True)
        if self.acc:                      # initializeAcceleration (Agent.cpp:1178-1187)
            self.V = self.X.copy()
            self.Y = self.X.copy()
            self.gamma = 0.0
            self.alpha = 0.0

    def linear_term(self, a, Xsrc):
        b = self.blocks[a]
        return (self.Ct[b.c0:b.c1] @ Xsrc.T).T

    def _optimize(self, a, X0a, Xnbr):
        b = self.blocks[a]
        b.P.G = self.linear_term(a, Xnbr)
        opt = solver.QuadraticOptimizer(b.P, self.params)
        out = opt.optimize(X0a)
        self.results[a] = opt.result
        return out

    def iterate(self, selected):
        """One driver iteration: agents in `selected` optimise, all others
        call iterate(false)."""
        sel = set(int(a) for a in selected)
        self.it += 1
        if not self.acc:
            Xn = self.X.copy()
            for a in sel:
                b = self.blocks[a]
                Xn[:, b.c0:b.c1] = self._optimize(a, self.X[:, b.c0:b.c1], self.X)
            self.X = Xn
            return
        # RBCD++ (Agent.cpp:544-554).  The driver runs iterate(false) on every
        # non-selected agent BEFORE the selected agent fetches its neighbours'
        # public states (MultiRobotExample.cpp:228-258): the neighbour states that
        # enter G are the neighbours' new X, i.e. their Y -- or, on a restart
        # iteration (all agents share the counter), their XPrev (:1165-1176).
        N = self.N
        Xprev = self.X.copy()
        self.gamma = (1 + np.sqrt(1 + 4 * N ** 2 * self.gamma ** 2)) / (2 * N)
        self.alpha = 1 / (self.gamma * N)
        restart = (self.it + 1) % self.restart_interval == 0   # shouldRestart (:1158-1163)
        Y = np.empty_like(self.X)
        for b in self.blocks:
            s = slice(b.c0, b.c1)
            Y[:, s] = mf.project_manifold(b.L, (1 - self.alpha) * self.X[:, s]
                                          + self.alpha * self.V[:, s])
        # state of every agent after iterate(false)
        Xfalse = Xprev if restart else Y
        Xn = Xfalse.copy()
        for a in sel:
            b = self.blocks[a]
            if restart:
                # updateX(true, true) from Y, then restartNesterovAcceleration:
                # X = XPrev; updateX(true, false) -- the second solve is the one kept
                Xn[:, b.c0:b.c1] = self._optimize(a, Xprev[:, b.c0:b.c1], Xfalse)
            else:
                Xn[:, b.c0:b.c1] = self._optimize(a, Y[:, b.c0:b.c1], Xfalse)
        if restart:
            self.X = Xn
            self.V = Xn.copy()
            self.Y = Xn.copy()
            self.gamma = 0.0
            self.alpha = 0.0
            return
        V = np.empty_like(self.X)
        for b in self.blocks:
            s = slice(b.c0, b.c1)
            V[:, s] = mf.project_manifold(b.L, self.V[:, s] + self.gamma * (Xn[:, s] - Y[:, s]))
        self.X, self.V, self.Y = Xn, V, Y

    def sweep(self, colors):
        """One iteration per colour."""
        for c in range(int(np.max(colors)) + 1):
            self.iterate(np.nonzero(np.asarray(colors) == c)[0])

    def monitor(self, per_agent=False):
        """Global cost and Riemannian gradient norm, block by block
        (MultiRobotExample.cpp:264-281 evaluates the centralised problem)."""
        G = (self.Ct @ self.X.T).T
        f = 0.0
        norms = []
        for b in self.blocks:
            s = slice(b.c0, b.c1)
            Xa = self.X[:, s]
            XQ = b.P.XQ(Xa)
            f += 0.5 * np.sum(XQ * Xa) + 0.5 * np.sum(Xa * G[:, s])
            rg = mf.tangent_project(b.L, Xa, XQ + G[:, s])
            norms.append(np.sqrt(np.sum(rg * rg)))
        norms = np.array(norms)
        gn = np.sqrt(np.sum(norms ** 2))
        return (f, gn, norms) if per_agent else (f, gn)

    def run_selected(self, num_iters, selection="greedy", gradnorm_tol=0.0, seed=0,
                     integer_norms=False, first=0):
        """The driver loop itself (MultiRobotExample.cpp:223-307, block selection :288-305;
        MultiRobotExample_RASLAM.cpp:404-452 for the rule switch and its integer-valued
        std::map): one selected agent per iteration, next one greedy / uniform."""
        rng = np.random.default_rng(seed)
        has_nbr = np.array([self.Ct[b.c0:b.c1].nnz > 0 for b in self.blocks])
        selected, hist = int(first), []
        for _ in range(num_iters):
            self.iterate([selected])
            f, gn, norms = self.monitor(per_agent=True)
            hist.append((selected, f, gn))
            if gn < gradnorm_tol:
                break
            if not has_nbr[selected]:
                continue
            if selection == "greedy":
                selected = int(np.argmax(np.floor(norms) if integer_norms else norms))
            else:
                selected = int(rng.integers(0, self.A))
        return hist
