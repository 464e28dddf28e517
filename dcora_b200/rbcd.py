"""Multi-robot RBCD / RBCD++ driver over the C-ABI (host side).

Mirrors the loop of examples/MultiRobotExample.cpp:223-307 and
Agent::iterate / updateX (src/Agent.cpp:535-596,1189-1278): agents exchange
their public states, the selected agents rebuild G and take one
QuadraticOptimizer::optimize step, the driver monitors the global cost and
Riemannian gradient norm.  Differences to the reference, by design:

  * all agents resident on one GPU live in ONE batched problem handle and
    agents of one colour of the robot graph step concurrently (the reference
    steps one greedily selected agent at a time; agents that share no
    measurement commute, so a colour sweep equals a sequence of reference
    iterations with that selection order);
  * agents are partitioned over the ranks of a torch.distributed job (one
    process per GPU); the public states of every rank are exchanged with one
    all-gather per colour step (NCCL over NVLink on the GPU box, gloo on CPU
    for the plumbing tests) instead of PoseDict copies.

This module only plans index lists on the host and sequences C-ABI calls; all
solver arithmetic runs in libdcora_b200.so on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from .api import AgentData, ORDER_RA, ORDER_SE, QuadraticOptimizer, QuadraticProblem, ROptParameters
from .graph import GlobalProblem, greedy_coloring


def rank_agents(A, world):
    """Contiguous block of agents per rank: [a0, a1) for every rank."""
    base, rem = divmod(A, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    off = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(off[r]), int(off[r + 1])) for r in range(world)]


@dataclass
class RankPlan:
    """Host-side exchange plan of one rank (pure index bookkeeping)."""
    rank: int
    world: int
    a0: int
    a1: int
    c0: int                 # first global column of this rank
    K: int                  # local columns
    seg: tuple              # (n, l, b) of every rank's (padded) public segment
    pack_idx: np.ndarray    # local reference columns, in segment order (padded with 0)
    coupling: sp.csr_matrix  # (K + world*seg_cols, K)  [or (K, K) when world == 1]
    colors: np.ndarray      # colour of every GLOBAL agent
    num_colors: int
    row_key: np.ndarray = None   # (K + world*seg_cols): global column behind every coupling row

    @property
    def seg_cols(self):
        n, l, b = self.seg
        return self.dh * n + l + b

    dh: int = 4


def make_plans(gp: GlobalProblem, world: int, cross=None):
    """Exchange plans of all ranks (deterministic: every rank computes the same)."""
    if cross is None:
        _, cross = gp.split(agents=[])
    A, d, dh = gp.A, gp.d, gp.d + 1
    blocks = rank_agents(A, world)
    colors = greedy_coloring(gp.agent_graph(cross))
    ncol = int(colors.max()) + 1 if A else 0
    aoc = gp.agent_of_col()
    rank_of_agent = np.empty(A, np.int64)
    for r, (a0, a1) in enumerate(blocks):
        rank_of_agent[a0:a1] = r
    roc = rank_of_agent[aoc]                                  # rank of every global column
    if world == 1:
        a0, a1 = blocks[0]
        return [RankPlan(0, 1, a0, a1, 0, gp.K, (0, 0, 0), np.zeros(0, np.int32),
                         cross.tocsr(), colors, ncol, dh=dh,
                         row_key=np.arange(gp.K, dtype=np.int64))]
    # public columns per rank: rows of `cross` with an entry in another rank's column
    row_of = np.repeat(np.arange(gp.K), np.diff(cross.indptr))
    inter = roc[row_of] != roc[cross.indices]
    pub_cols = np.unique(row_of[inter])
    # classify public columns: pose / sphere / landmark  (global ids)
    kind, vid = _column_kinds(gp)
    pub = []
    for r in range(world):
        mine = pub_cols[roc[pub_cols] == r]
        poses = np.unique(vid[mine[kind[mine] == 0]])
        sph = np.unique(vid[mine[kind[mine] == 1]])
        lmk = np.unique(vid[mine[kind[mine] == 2]])
        pub.append((poses, sph, lmk))
    seg = (max(len(p[0]) for p in pub), max(len(p[1]) for p in pub), max(len(p[2]) for p in pub))
    n_s, l_s, b_s = seg
    seg_cols = dh * n_s + l_s + b_s
    # global column -> halo column (extended index relative to the rank's K is added later)
    halo_of = -np.ones(gp.K, np.int64)
    halo_key = -np.ones(world * seg_cols, np.int64)      # global column behind a halo column
    packs = []
    for r in range(world):
        poses, sph, lmk = pub[r]
        c0 = gp.coff[blocks[r][0]]
        seg_ref = np.zeros(seg_cols, np.int64)          # global column feeding each segment column
        valid = np.zeros(seg_cols, bool)
        j = np.arange(len(poses))
        if gp.ra:
            rot_pos = (j[:, None] * d + np.arange(d)).reshape(-1)
            tr_pos = d * n_s + l_s + j
            sp_pos = d * n_s + np.arange(len(sph))
            lm_pos = d * n_s + l_s + n_s + np.arange(len(lmk))
        else:
            rot_pos = (j[:, None] * dh + np.arange(d)).reshape(-1)
            tr_pos = j * dh + d
            sp_pos = lm_pos = np.zeros(0, np.int64)
        if len(poses):
            seg_ref[rot_pos] = gp.rot_cols(poses).reshape(-1); valid[rot_pos] = True
            seg_ref[tr_pos] = gp.trans_col(poses); valid[tr_pos] = True
        if len(sph):
            seg_ref[sp_pos] = gp.sphere_col(sph); valid[sp_pos] = True
        if len(lmk):
            seg_ref[lm_pos] = gp.lmk_col(lmk); valid[lm_pos] = True
        halo_of[seg_ref[valid]] = r * seg_cols + np.nonzero(valid)[0]
        halo_key[r * seg_cols + np.nonzero(valid)[0]] = seg_ref[valid]
        pk = np.where(valid, seg_ref - c0, 0)
        packs.append(pk.astype(np.int32))
    plans = []
    csc = cross.tocsc()
    for r, (a0, a1) in enumerate(blocks):
        c0, c1 = int(gp.coff[a0]), int(gp.coff[a1])
        K = c1 - c0
        sub = csc[:, c0:c1].tocoo()
        local = roc[sub.row] == r
        rows = np.where(local, sub.row - c0, K + halo_of[sub.row])
        assert (rows >= 0).all() and (halo_of[sub.row[~local]] >= 0).all()
        C = sp.csr_matrix((sub.data, (rows, sub.col)), shape=(K + world * seg_cols, K))
        key = np.concatenate([np.arange(c0, c1, dtype=np.int64), halo_key])
        plans.append(RankPlan(r, world, a0, a1, c0, K, seg, packs[r], C, colors, ncol, dh=dh,
                              row_key=key))
    return plans


def _column_kinds(gp: GlobalProblem):
    """kind (0 pose, 1 sphere, 2 landmark) and global variable id of every column."""
    kind = np.zeros(gp.K, np.int8)
    vid = np.zeros(gp.K, np.int64)
    n = gp.meas.n
    pc = gp.pose_cols(np.arange(n))
    vid[pc.reshape(-1)] = np.repeat(np.arange(n), gp.d + 1)
    m = len(gp.meas.rg_i)
    if m:
        sc = gp.sphere_col(np.arange(m))
        kind[sc] = 1
        vid[sc] = np.arange(m)
    if gp.meas.b:
        lc = gp.lmk_col(np.arange(gp.meas.b))
        kind[lc] = 2
        vid[lc] = np.arange(gp.meas.b)
    return kind, vid


class RBCD:
    """Batched multi-agent RBCD on one rank of a (possibly 1-process) job."""

    def __init__(self, gp: GlobalProblem, r: int, rank=0, world=1, device=0,
                 params: ROptParameters | None = None, precond=True, group=None,
                 plan: RankPlan | None = None, diag=None, acceleration=False,
                 restart_interval=30, _problem=None, k5_warps=None, nesterov_blocks=None):
        import torch
        self.torch = torch
        self.gp, self.r, self.rank, self.world, self.group = gp, r, rank, world, group
        self.precond = precond
        if _problem is not None:
            # another rank of the same partitioned problem (with_rank): the data matrices,
            # the Cholesky factors and the coupling are shared on the device
            self.plan = plan
            self.P = _problem
        else:
            import os, sys, time
            timing = os.environ.get("DCORA_SETUP_TIMING") and rank == 0
            t_lap = [time.perf_counter()]

            def lap(what):
                if timing:
                    now = time.perf_counter()
                    print(f"[dcora setup] {what:28s} {now - t_lap[0]:8.3f} s (python)", file=sys.stderr)
                    t_lap[0] = now
            if plan is None or diag is None:
                blocks = rank_agents(gp.A, world)
                diag, cross = gp.split(agents=range(*blocks[rank]))
                lap("split of Q by agent")
                plan = make_plans(gp, world, cross)[rank]
                lap("exchange plans + colouring")
            self.plan = plan
            reg = None if not precond else ("auto_ra" if gp.ra else 0.1)
            agents = [AgentData(n=int(gp.n_a[a]), l=int(gp.l_a[a]), b=int(gp.b_a[a]), Q=diag[a],
                                precond_reg=reg) for a in range(plan.a0, plan.a1)]
            # K5 is a latency chain per agent (261 us with 64 warps = 4 CTAs per agent, 212 us
            # with 96, 188 us with 128, 138 us with 256 on a configs[4] agent): when the colour
            # steps of this job solve few agents per GPU (several ranks), more CTAs per agent
            # shorten every step and the clusters still fit the machine in one wave.  The
            # choice follows the LARGEST class on any rank (all ranks compute the same value).
            # (> 64 warps serves ranks <= 8: pass k5_warps=64 for a staircase that goes higher)
            if k5_warps is None:
                biggest = 0
                for (b0, b1) in rank_agents(gp.A, world):
                    loc = plan.colors[b0:b1]
                    if len(loc):
                        biggest = max(biggest, int(np.bincount(loc).max()))
                k5_warps = 64
                if r <= 8 and biggest > 0:
                    k5_warps = 256 if biggest <= 8 else 128 if biggest <= 16 else 96 if biggest <= 24 else 64
            from . import _lib
            _lib.load().dcora_set_trsv_max_warps(int(k5_warps))
            try:
                self.P = QuadraticProblem(agents, r, gp.d, ORDER_RA if gp.ra else ORDER_SE, device)
            finally:
                _lib.load().dcora_set_trsv_max_warps(64)
            lap("dcora_problem_create (total)")
            segs = [plan.seg] * world if world > 1 else []
            self.P.set_coupling(plan.coupling, segs, row_key=plan.row_key)
            lap("set_coupling")
        # the iterate lives in ONE device buffer (self.X) that only the library writes
        self.P.set_iterate_tracking(True)
        self.params = params or ROptParameters()
        self.opt = QuadraticOptimizer(self.P, self.params)
        self.device = torch.device("cuda", device)
        self.A_local = plan.a1 - plan.a0
        self.local_colors = plan.colors[plan.a0:plan.a1]
        self.H = None
        if world > 1:
            self.H = torch.zeros((world * plan.seg_cols, r), dtype=torch.float64, device=self.device)
            self.sendbuf = torch.zeros((plan.seg_cols, r), dtype=torch.float64, device=self.device)
        self.agent_updates = 0
        self.inner_iterations = 0      # tCG iterations summed over the agent updates
        self.outer_iterations = 0
        # RBCD++ (Nesterov acceleration, Agent.cpp:535-596,1158-1214; AgentParameters
        # acceleration / restartInterval, Agent.h:40-185)
        self.acc = bool(acceleration)
        self.acc0 = self.acc          # disable_acceleration() is per staircase level
        # N of updateGamma / updateAlpha (Agent.cpp:1189-1200): the number of blocks one of which
        # is updated per iteration -- the number of agents in the reference; with the parallel
        # colour schedule a whole colour class moves per iteration, so the number of COLOURS
        # plays that role (nesterov_blocks='colours')
        self.nesterov_blocks = nesterov_blocks
        if restart_interval == "auto":
            # measured on the 47^3-pose grid: the best interval between Nesterov restarts is about
            # 5.5 sqrt(A) SWEEPS (22 sweeps for 16 agents, 45 for 64; profiles/r02_ttc_rbcdpp.txt)
            restart_interval = max(30, int(round(5.5 * np.sqrt(gp.A))) * max(1, self.plan.num_colors))
        self.restart_interval = int(restart_interval)
        self.it = 0
        if self.acc:
            # several iterate-sized buffers (X, Y, V, XPrev) are exchanged per step:
            # no single tracked buffer
            self.P.set_iterate_tracking(False)

    def with_rank(self, r):
        """The same partitioned problem at another relaxation rank (next staircase level):
        Q, the factors and the coupling stay on the device (dcora_problem_clone_rank)."""
        P = self.P.with_rank(r)
        P.KH = self.P.KH
        return RBCD(self.gp, r, self.rank, self.world, self.device.index, self.params,
                    self.precond, self.group, plan=self.plan, acceleration=self.acc0,
                    restart_interval=self.restart_interval, _problem=P,
                    nesterov_blocks=self.nesterov_blocks)

    # -- state ---------------------------------------------------------------
    def set_X(self, X_global):
        """X_global: numpy (r, K_global) in the agent-major reference ordering."""
        t = self.torch
        p = self.plan
        Xl = np.ascontiguousarray(X_global[:, p.c0:p.c0 + p.K].T)
        self.set_device_X(t.from_numpy(Xl).to(self.device))

    def set_device_X(self, X_dev):
        """Replace the iterate by a device tensor (K_local, r) written by the caller."""
        self.X = X_dev
        if self.acc:                            # initializeAcceleration (Agent.cpp:1178-1187)
            self.V = X_dev.clone()
            self.Y = X_dev.clone()
            self.gamma = 0.0
            self.alpha = 0.0
        else:
            self.P.set_iterate_tracking(True)   # forget the tracked copy

    def restart_acceleration(self):
        """restartNesterovAcceleration (Agent.cpp:1165-1176) on demand: V = Y = X, gamma = 0
        (the momentum is dropped, the iterate kept)."""
        if self.acc:
            self.V = self.X.clone()
            self.Y = self.X.clone()
            self.gamma = 0.0
            self.alpha = 0.0

    def disable_acceleration(self):
        """Continue with plain RBCD from the current iterate (used when the accelerated
        sequence stalls: RBCD++ is not monotone, RBCD is)."""
        if self.acc:
            self.acc = False
            self.V = self.Y = None
            self.P.set_iterate_tracking(True)

    def _count(self, active):
        """Bookkeeping of an optimize() call over the agents in `active`."""
        from ._lib import ROPT_RESULT_DTYPE
        res = np.frombuffer(self.opt._raw, dtype=ROPT_RESULT_DTYPE)
        m = active.astype(bool)
        self.agent_updates += int(m.sum())
        self.inner_iterations += int(res["inner_iterations"][m].sum())
        self.outer_iterations += int(res["outer_iterations"][m].sum())

    def load_host_X(self, X_host):
        """Upload the local iterate from a host buffer (numpy (r, K_local), column-major,
        ideally pinned) into the device buffer the solver works on."""
        if getattr(self, "X", None) is None:
            self.X = self.torch.empty((self.plan.K, self.r), dtype=self.torch.float64, device=self.device)
        self.P.load_iterate(X_host, self.X)
        self.set_device_X(self.X)

    def store_host_X(self, X_host):
        self.P.store_iterate(self.X, X_host)

    def get_X(self):
        """numpy (r, K_local)."""
        return self.X.cpu().numpy().T.copy()

    # -- exchange (Agent::getSharedStateDicts -> updateNeighborStates) ----------
    def exchange(self, src=None):
        """All-gather of the public columns of `src` (default: the iterate X)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        self.P.pack_columns(self.X if src is None else src, self.plan.pack_idx, out=self.sendbuf)
        dist.all_gather_into_tensor(self.H, self.sendbuf, group=self.group)
        self.torch.cuda.current_stream().synchronize()

    # -- one RBCD iteration = one sweep over the colours -------------------------
    def iterate(self):
        for c in range(self.plan.num_colors):
            active = (self.local_colors == c).astype(np.uint8)
            self.step(active)

    def step(self, active):
        """One driver iteration (examples/MultiRobotExample.cpp:223-258): the agents in
        `active` (local mask; they must share no measurement) take Agent::iterate(true),
        all others iterate(false)."""
        active = np.ascontiguousarray(active, dtype=np.uint8)
        if self.acc:
            self._accelerated_step(active)
            return
        self.exchange()
        if active.any():
            self.P.update_G(self.X, self.H, active)
            self.X = self.opt.optimize(self.X, active, out=self.X)
            self._count(active)

    def _accelerated_step(self, active):
        """One driver iteration of RBCD++ (Agent.cpp:535-596).  The non-selected agents run
        iterate(false) FIRST -- X <- Y, or X <- XPrev on a restart iteration (:1165-1176) --
        and only then does the selected agent fetch its neighbours' public states
        (MultiRobotExample.cpp:228-258), so its G is built from the neighbours' NEW X: their Y
        on an ordinary iteration, their XPrev on a restart iteration (every agent shares the
        iteration counter, so all restart together).  On a restart iteration the selected
        agent's first solve (from Y) is overwritten by restartNesterovAcceleration's solve
        from XPrev, so only that one is run here."""
        N = self.gp.A
        if self.nesterov_blocks == "colours":
            N = max(1, self.plan.num_colors)
        elif self.nesterov_blocks:
            N = int(self.nesterov_blocks)
        self.it += 1
        Xprev = self.X
        self.gamma = (1 + np.sqrt(1 + 4 * N ** 2 * self.gamma ** 2)) / (2 * N)   # updateGamma
        self.alpha = 1 / (self.gamma * N)                                       # updateAlpha
        restart = (self.it + 1) % self.restart_interval == 0                    # shouldRestart
        if restart:
            Xr = Xprev.clone()
            self.exchange(Xprev)
            if active.any():
                self.P.update_G(Xprev, self.H, active)
                Xr = self.opt.optimize(Xprev, active, out=Xr)
                self._count(active)
            self.X, self.V, self.Y = Xr, Xr.clone(), Xr.clone()
            self.gamma = 0.0
            self.alpha = 0.0
            return
        # updateY: Y = projectToManifold((1 - alpha) X + alpha V), every agent
        Y = self.P.combine_project(1 - self.alpha, self.X, self.alpha, self.V, None)
        Xn = Y.clone()                                    # iterate(false): X = Y
        self.exchange(Y)
        if active.any():
            self.P.update_G(Y, self.H, active)
            Xn = self.opt.optimize(Y, active, out=Xn)
            self._count(active)
        # updateV: V = projectToManifold(V + gamma (X - Y))
        V = self.P.combine_project(1.0, self.V, self.gamma, Xn - Y, None)
        self.X, self.V, self.Y = Xn, V, Y

    def monitor(self, per_agent=False):
        """Global cost f and Riemannian gradient norm (the centralised
        evaluation of MultiRobotExample.cpp:264-281, computed blockwise).  With
        `per_agent` also the gradient norm of every agent's block (all A agents of the job):
        the quantity the block selection rules rank (MultiRobotExample.cpp:288-305)."""
        self.exchange()
        self.P.update_G(self.X, self.H, None)
        quad, lin = self.P.cost_terms(self.X)
        gn = np.atleast_1d(self.P.RieGradNorm(self.X))
        f = float(np.sum(quad) + 0.5 * np.sum(lin))
        g2 = float(np.sum(gn ** 2))
        norms = np.asarray(gn, dtype=np.float64)
        if self.world > 1:
            import torch.distributed as dist
            v = self.torch.tensor([f, g2], dtype=self.torch.float64, device=self.device)
            dist.all_reduce(v, group=self.group)
            f, g2 = float(v[0]), float(v[1])
            if per_agent:
                full = self.torch.zeros(self.gp.A, dtype=self.torch.float64, device=self.device)
                full[self.plan.a0:self.plan.a1] = self.torch.from_numpy(norms).to(self.device)
                dist.all_reduce(full, group=self.group)
                norms = full.cpu().numpy()
        if per_agent:
            return f, float(np.sqrt(g2)), norms
        return f, float(np.sqrt(g2))

    # -- the reference drivers' one-agent-per-iteration loop ------------------------
    def run_selected(self, num_iters, selection="greedy", gradnorm_tol=0.0, seed=0,
                     integer_norms=False, first=0, log=None):
        """examples/MultiRobotExample.cpp:223-307 / MultiRobotExample_RASLAM.cpp:290-452:
        ONE selected agent takes iterate(true) per iteration, every other agent
        iterate(false); after the centralised evaluation the next agent is chosen by the
        BlockSelectionRule (include/DCORA/DCORA_types.h:54):

          "greedy"   the agent whose block of the Riemannian gradient is largest
                     (first of equals, std::max_element);
          "uniform"  an agent drawn uniformly (std::uniform_int_distribution).

        An agent without neighbours keeps the turn.  `integer_norms` reproduces the RASLAM
        example's std::map<unsigned, unsigned> (block norms truncated to integers before the
        comparison, MultiRobotExample_RASLAM.cpp:404).  Returns the list of
        (selected agent, f, gradnorm) per iteration."""
        assert selection in ("greedy", "uniform")
        rng = np.random.default_rng(seed)
        A = self.gp.A
        nbrs = np.array([len(a) for a in self.gp.agent_graph()])
        selected = int(first)
        hist = []
        for it in range(int(num_iters)):
            active = np.zeros(self.A_local, np.uint8)
            if self.plan.a0 <= selected < self.plan.a1:
                active[selected - self.plan.a0] = 1
            self.step(active)
            f, gn, norms = self.monitor(per_agent=True)
            hist.append((selected, f, gn))
            if log:
                log(f"iter {it} robot {selected} cost {2 * f:.6f} gradnorm {gn:.6g}")
            if gn < gradnorm_tol:
                break
            if nbrs[selected] == 0:
                continue
            if selection == "greedy":
                key = np.floor(norms) if integer_norms else norms
                selected = int(np.argmax(key))
            else:
                selected = int(rng.integers(0, A))
        return hist
