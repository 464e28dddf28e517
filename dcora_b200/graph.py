"""Host-side data plumbing of the hot path: measurements -> data matrices.

Mirrors what DCORA::Graph hands to QuadraticProblem (include/DCORA/Graph.h:
453-459): the quadratic data matrix Q of every agent, the cross-agent blocks
that Graph::constructLinearCostTerm{PGO,RASLAM} turn into the linear term G
(src/Graph.cpp:685-822,1190-1772), and the public-state index lists
(Graph::myPublicPoseIDs / ...UnitSphereIDs / ...LandmarkIDs, Graph.h:256-268).

Instead of one Graph object per agent, the whole multi-robot problem is
assembled once as a block-structured global matrix whose columns are ordered
agent by agent, each agent in the reference column ordering
(include/DCORA/manifold/LiftedVariable.h:27-30,148-152):

    SE: [Y1 p1 | Y2 p2 | ...]               k = (d+1) n
    RA: [Y1..Yn | r1..rl | p1..pn | L1..Lb] k = (d+1) n + l + b

Agent a's Q (Graph.cpp:579-683, 824-1188) is then the diagonal block of that
matrix (a shared edge contributes exactly the owned endpoint's diagonal block),
and G_a = X_others * (off-diagonal block column of a).  Everything here is
vectorised numpy/scipy set-up code; no solver arithmetic happens on the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


def _f64(x, shape):
    return np.ascontiguousarray(np.asarray(x, np.float64).reshape(shape))


def _i64(x):
    return np.ascontiguousarray(np.asarray(x, np.int64).reshape(-1))


@dataclass
class Measurements:
    """Global measurement arrays (include/DCORA/Measurements.h), indices are
    GLOBAL pose / landmark numbers.  Range endpoints are translation-like
    variables: value < n is a pose, value >= n is landmark (value - n)."""
    d: int
    n: int
    b: int = 0
    pp_i: np.ndarray = None
    pp_j: np.ndarray = None
    pp_R: np.ndarray = None
    pp_t: np.ndarray = None
    pp_kappa: np.ndarray = None
    pp_tau: np.ndarray = None
    pp_w: np.ndarray = None
    pl_i: np.ndarray = None
    pl_j: np.ndarray = None
    pl_t: np.ndarray = None
    pl_tau: np.ndarray = None
    pl_w: np.ndarray = None
    rg_i: np.ndarray = None
    rg_j: np.ndarray = None
    rg_range: np.ndarray = None
    rg_prec: np.ndarray = None
    rg_w: np.ndarray = None

    def __post_init__(self):
        d = self.d
        self.pp_i, self.pp_j = _i64(self.pp_i if self.pp_i is not None else []), \
            _i64(self.pp_j if self.pp_j is not None else [])
        m = len(self.pp_i)
        self.pp_R = _f64(self.pp_R if self.pp_R is not None else np.zeros((0, d, d)), (m, d, d))
        self.pp_t = _f64(self.pp_t if self.pp_t is not None else np.zeros((0, d)), (m, d))
        self.pp_kappa = _f64(self.pp_kappa if self.pp_kappa is not None else [], (m,))
        self.pp_tau = _f64(self.pp_tau if self.pp_tau is not None else [], (m,))
        self.pp_w = _f64(self.pp_w, (m,)) if self.pp_w is not None else np.ones(m)
        self.pl_i, self.pl_j = _i64(self.pl_i if self.pl_i is not None else []), \
            _i64(self.pl_j if self.pl_j is not None else [])
        m = len(self.pl_i)
        self.pl_t = _f64(self.pl_t if self.pl_t is not None else np.zeros((0, d)), (m, d))
        self.pl_tau = _f64(self.pl_tau if self.pl_tau is not None else [], (m,))
        self.pl_w = _f64(self.pl_w, (m,)) if self.pl_w is not None else np.ones(m)
        self.rg_i, self.rg_j = _i64(self.rg_i if self.rg_i is not None else []), \
            _i64(self.rg_j if self.rg_j is not None else [])
        m = len(self.rg_i)
        self.rg_range = _f64(self.rg_range if self.rg_range is not None else [], (m,))
        self.rg_prec = _f64(self.rg_prec if self.rg_prec is not None else [], (m,))
        self.rg_w = _f64(self.rg_w, (m,)) if self.rg_w is not None else np.ones(m)

    @property
    def range_aided(self):
        return len(self.rg_i) > 0 or len(self.pl_i) > 0 or self.b > 0


@dataclass
class Partition:
    """Ownership of the variables: robot of every pose / landmark.  Unit
    spheres belong to the robot of the range's source endpoint, numbered by a
    per-source-robot counter in measurement order (DCORA_utils.cpp:1092-1094)."""
    num_robots: int
    robot_of_pose: np.ndarray
    robot_of_lmk: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))

    @staticmethod
    def contiguous(n, num_robots):
        """examples/MultiRobotExample.cpp:56-83: n/num_robots poses per robot,
        the last robot takes the remainder."""
        per = max(1, n // num_robots)
        rob = np.minimum(np.arange(n) // per, num_robots - 1)
        return Partition(num_robots, rob.astype(np.int64))


class GlobalProblem:
    """Column bookkeeping + global data matrix of a partitioned problem."""

    def __init__(self, meas: Measurements, part: Partition, ra: bool | None = None):
        meas = drop_duplicate_edges(meas)
        self.meas, self.part = meas, part
        d = self.d = meas.d
        self.ra = meas.range_aided if ra is None else ra
        if not self.ra:
            assert not meas.range_aided, "pose-graph ordering cannot hold ranges / landmarks"
        A = self.A = part.num_robots
        rp = _i64(part.robot_of_pose)
        rl = _i64(part.robot_of_lmk) if meas.b else np.zeros(0, np.int64)
        assert len(rp) == meas.n and len(rl) == meas.b
        self.robot_of_pose, self.robot_of_lmk = rp, rl
        self.n_a = np.bincount(rp, minlength=A)
        self.b_a = np.bincount(rl, minlength=A) if meas.b else np.zeros(A, np.int64)
        # local indices = rank of the global index inside its robot
        self.local_pose = _rank_within(rp, A)
        self.local_lmk = _rank_within(rl, A) if meas.b else np.zeros(0, np.int64)
        # spheres
        tv_robot = np.concatenate([rp, rl])
        self.sph_robot = tv_robot[meas.rg_i] if len(meas.rg_i) else np.zeros(0, np.int64)
        self.l_a = np.bincount(self.sph_robot, minlength=A)
        self.local_sph = _rank_within(self.sph_robot, A)
        dh = d + 1
        self.k_a = dh * self.n_a + self.l_a + self.b_a
        self.coff = np.concatenate([[0], np.cumsum(self.k_a)]).astype(np.int64)
        self.K = int(self.coff[-1])
        self._Q = None

    # ---- reference column indices (global, agent-major) -------------------
    def rot_cols(self, i):
        """(len(i), d) columns of the rotation block of global pose i."""
        i = np.asarray(i)
        a = self.robot_of_pose[i]
        li = self.local_pose[i]
        d = self.d
        base = self.coff[a] + (li * d if self.ra else li * (d + 1))
        return base[..., None] + np.arange(d)

    def trans_col(self, i):
        i = np.asarray(i)
        a = self.robot_of_pose[i]
        li = self.local_pose[i]
        d = self.d
        if self.ra:
            return self.coff[a] + d * self.n_a[a] + self.l_a[a] + li
        return self.coff[a] + li * (d + 1) + d

    def pose_cols(self, i):
        """(len(i), d+1): rotation columns then the translation column."""
        return np.concatenate([self.rot_cols(i), np.asarray(self.trans_col(i))[..., None]], -1)

    def sphere_col(self, e):
        """column of the unit sphere of range measurement e."""
        e = np.asarray(e)
        a = self.sph_robot[e]
        return self.coff[a] + self.d * self.n_a[a] + self.local_sph[e]

    def lmk_col(self, j):
        j = np.asarray(j)
        a = self.robot_of_lmk[j]
        return self.coff[a] + self.d * self.n_a[a] + self.l_a[a] + self.n_a[a] + self.local_lmk[j]

    def tvar_col(self, t):
        """translation-like variable (pose translation or landmark)."""
        t = np.asarray(t)
        out = np.empty(t.shape, np.int64)
        isp = t < self.meas.n
        out[isp] = self.trans_col(t[isp])
        if (~isp).any():
            out[~isp] = self.lmk_col(t[~isp] - self.meas.n)
        return out

    def agent_of_col(self):
        return np.repeat(np.arange(self.A), self.k_a)

    # ---- global data matrix ------------------------------------------------
    def Q(self):
        """Global connection-Laplacian-type data matrix, scipy CSR, K x K."""
        if self._Q is None:
            self._Q = self._assemble()
        return self._Q

    def _assemble(self):
        me, d = self.meas, self.d
        dh = d + 1
        n = me.n
        K = self.K
        parts = []
        m = len(me.pp_i)
        if m:
            # T = [R t; 0 1], Omega = diag(w kappa .. , w tau)   (Graph.cpp:579-683)
            T = np.zeros((m, dh, dh))
            T[:, :d, :d] = me.pp_R
            T[:, :d, d] = me.pp_t
            T[:, d, d] = 1.0
            om = np.empty((m, dh))
            om[:, :d] = (me.pp_w * me.pp_kappa)[:, None]
            om[:, d] = me.pp_w * me.pp_tau
            TO = T * om[:, None, :]
            Bii = TO @ np.transpose(T, (0, 2, 1))
            Bij = -TO
            Bji = np.transpose(Bij, (0, 2, 1))
            Bjj = np.zeros((m, dh, dh))
            Bjj[:, np.arange(dh), np.arange(dh)] = om
            bi = np.concatenate([me.pp_i, me.pp_i, me.pp_j, me.pp_j])
            bj = np.concatenate([me.pp_i, me.pp_j, me.pp_i, me.pp_j])
            B = np.concatenate([Bii, Bij, Bji, Bjj])
            del T, TO, Bii, Bij, Bji, Bjj
            # merge duplicate blocks
            key = bi * n + bj
            order = np.argsort(key, kind="stable")
            key = key[order]
            first = np.concatenate([[True], key[1:] != key[:-1]])
            starts = np.nonzero(first)[0]
            B = np.add.reduceat(B[order], starts, axis=0)
            ub = key[starts]
            bi, bj = ub // n, ub % n
            del key, order
            # pose-interleaved global order: block p -> columns [p*dh, p*dh+dh)
            Qb = sp.bsr_matrix((B, bj, np.searchsorted(bi, np.arange(n + 1))),
                               shape=(n * dh, n * dh)).tocsr()
            del B
            Qb.eliminate_zeros()
            # relabel pose-interleaved columns -> reference agent-major columns
            cmap = self.pose_cols(np.arange(n)).reshape(-1)
            parts.append(_relabel(Qb, cmap, K))
            del Qb
        rows, cols, vals = [], [], []
        m = len(me.pl_i)
        if m:
            # pose i observes landmark j at t (in frame i), precision tau:
            # residual p_j - p_i - Y_i t          (Graph.cpp:824-1188, TT / ATau)
            w = me.pl_w * me.pl_tau
            ri = self.rot_cols(me.pl_i)             # (m, d)
            ti = np.asarray(self.trans_col(me.pl_i))
            lj = self.lmk_col(me.pl_j)
            t = me.pl_t
            # variables z = [Y_i (d cols), p_i, L_j] with coefficient row c = [-t, -1, +1]
            cidx = np.concatenate([ri, ti[:, None], lj[:, None]], 1)      # (m, d+2)
            coef = np.concatenate([-t, -np.ones((m, 1)), np.ones((m, 1))], 1)
            rows.append(np.repeat(cidx, d + 2, axis=1).reshape(-1))
            cols.append(np.tile(cidx, (1, d + 2)).reshape(-1))
            vals.append((w[:, None, None] * coef[:, :, None] * coef[:, None, :]).reshape(-1))
        m = len(me.rg_i)
        if m:
            # || p_j - p_i - rho r ||^2 w  -> z = [p_i, p_j, r] with c = [-1, +1, -rho]...
            # reference sign convention (DCORA_utils.cpp:1136-1137, Graph.cpp:1137-1141):
            # Q22 = w rho^2, Q23(l, src) = -w rho, Q23(l, dst) = +w rho,
            # Q33: (src,src)=(dst,dst)=w, (src,dst)=-w
            w = me.rg_w * me.rg_prec
            ci = self.tvar_col(me.rg_i)
            cj = self.tvar_col(me.rg_j)
            cs = self.sphere_col(np.arange(m))
            cidx = np.stack([ci, cj, cs], 1)
            coef = np.stack([-np.ones(m), np.ones(m), me.rg_range], 1)
            rows.append(np.repeat(cidx, 3, axis=1).reshape(-1))
            cols.append(np.tile(cidx, (1, 3)).reshape(-1))
            vals.append((w[:, None, None] * coef[:, :, None] * coef[:, None, :]).reshape(-1))
        if rows:
            E = sp.coo_matrix((np.concatenate(vals),
                               (np.concatenate(rows), np.concatenate(cols))),
                              shape=(K, K)).tocsr()
            parts.append(E)
        if not parts:
            return sp.csr_matrix((K, K))
        Q = parts[0]
        for P in parts[1:]:
            Q = Q + P
        Q.sort_indices()
        return Q

    # ---- per-agent matrices and coupling --------------------------------------
    def split(self, agents=None):
        """Returns (diag, cross): `diag[a]` = (rowptr, col, val) of agent a's Q
        in agent-local reference columns for a in `agents` (default all);
        `cross` = scipy CSR (K x K) holding only the entries that couple
        different agents."""
        Q = self.Q()
        want = set(range(self.A)) if agents is None else set(int(a) for a in agents)
        indptr, indices, data = Q.indptr, Q.indices, Q.data
        # agents own contiguous column ranges: an entry of a row of agent a is "same" iff its
        # column lies in [coff[a], coff[a+1]).  One pass per agent over its own rows (cache
        # sized pieces, no K- or nnz-sized gathers).
        xcount = np.empty(self.K, np.int64)
        xi, xd = [], []
        diag = {}
        for a in range(self.A):
            c0, c1 = int(self.coff[a]), int(self.coff[a + 1])
            p0, p1 = int(indptr[c0]), int(indptr[c1])
            idx = indices[p0:p1]
            same = (idx >= c0) & (idx < c1)
            ptr = indptr[c0:c1 + 1] - p0
            csum = np.concatenate([[0], np.cumsum(same, dtype=np.int64)])
            dptr = csum[ptr]
            xcount[c0:c1] = np.diff(ptr) - np.diff(dptr)
            other = ~same
            xi.append(idx[other])
            xd.append(data[p0:p1][other])
            if a in want:
                diag[a] = (dptr.astype(np.int32),
                           np.ascontiguousarray(idx[same] - c0, dtype=np.int32),
                           np.ascontiguousarray(data[p0:p1][same]))
        xptr = np.concatenate([[0], np.cumsum(xcount)])
        cross = sp.csr_matrix((np.concatenate(xd) if xd else np.zeros(0),
                               np.concatenate(xi) if xi else np.zeros(0, indices.dtype),
                               xptr.astype(indptr.dtype)), shape=Q.shape)
        return diag, cross

    def agent_graph(self, cross=None):
        """Adjacency (set per agent) of the robot-level graph."""
        if cross is None:
            _, cross = self.split(agents=[])
        aoc = self.agent_of_col()
        ra = np.repeat(aoc, np.diff(cross.indptr))
        ca = aoc[cross.indices]
        pairs = np.unique(ra.astype(np.int64) * self.A + ca)
        adj = [set() for _ in range(self.A)]
        for p in pairs:
            adj[p // self.A].add(int(p % self.A))
        return adj


def drop_duplicate_edges(meas: Measurements) -> Measurements:
    """Graph keeps the FIRST measurement of every EdgeID (src, dst) and silently drops
    later duplicates (src/Graph.cpp:139-141,162-169; duplicate ranges are fatal there,
    :174-181).  Returns `meas` itself when there is nothing to drop."""
    import dataclasses

    def first_of(i, j):
        if len(i) == 0:
            return None
        key = np.stack([i, j], 1)
        _, first = np.unique(key, axis=0, return_index=True)
        return None if len(first) == len(i) else np.sort(first)

    kp = first_of(meas.pp_i, meas.pp_j)
    kl = first_of(meas.pl_i, meas.pl_j)
    if len(meas.rg_i):
        kr = first_of(meas.rg_i, meas.rg_j)
        assert kr is None, "range measurements must be unique (Graph.cpp:174-181)"
    if kp is None and kl is None:
        return meas
    kw = {f.name: getattr(meas, f.name) for f in dataclasses.fields(meas)}
    if kp is not None:
        for nm in ("pp_i", "pp_j", "pp_R", "pp_t", "pp_kappa", "pp_tau", "pp_w"):
            kw[nm] = kw[nm][kp]
    if kl is not None:
        for nm in ("pl_i", "pl_j", "pl_t", "pl_tau", "pl_w"):
            kw[nm] = kw[nm][kl]
    return Measurements(**kw)


def greedy_coloring(adj):
    """Colours of the agent graph (largest degree first): agents of one colour
    share no measurement and can take their RBCD step concurrently without
    losing the block-coordinate-descent guarantee."""
    A = len(adj)
    color = -np.ones(A, np.int64)
    for a in sorted(range(A), key=lambda x: -len(adj[x])):
        used = {color[b] for b in adj[a] if color[b] >= 0}
        c = 0
        while c in used:
            c += 1
        color[a] = c
    return balance_coloring(adj, color)


def balance_coloring(adj, color):
    """Even out the colour classes without adding colours: repeatedly move an agent
    of the largest class into the smallest class that holds none of its neighbours.
    Every colour step of an RBCD sweep runs its agents concurrently (one thread-block
    cluster per agent in K5), so its cost is set by the class size: classes of
    [32 35 37 34 33 34 32 18 1] agents become ~[29 .. 28]."""
    color = np.array(color, np.int64)
    A = len(adj)
    if A == 0:
        return color
    ncol = int(color.max()) + 1
    for _ in range(4 * A):
        size = np.bincount(color, minlength=ncol)
        big = int(np.argmax(size))
        moved = False
        for tgt in np.argsort(size, kind="stable"):
            if size[tgt] + 1 >= size[big]:
                break
            for a in np.nonzero(color == big)[0]:
                if all(color[b] != tgt for b in adj[a]):
                    color[a] = tgt
                    moved = True
                    break
            if moved:
                break
        if not moved:
            break
    return color


def _rank_within(group, A):
    """rank of every element among the elements of its group (stable)."""
    group = np.asarray(group, np.int64)
    if len(group) == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(group, kind="stable")
    counts = np.bincount(group, minlength=A)
    start = np.concatenate([[0], np.cumsum(counts)])[:-1]
    out = np.empty(len(group), np.int64)
    out[order] = np.arange(len(group)) - np.repeat(start, counts)
    return out


def _relabel(Qc, cmap, K):
    """P Qc P^T for the injective relabelling old index -> cmap[old] (O(nnz),
    no sort): rows are gathered, column indices relabelled."""
    cmap = np.asarray(cmap, np.int64)
    nold = Qc.shape[0]
    counts = np.zeros(K, np.int64)
    counts[cmap] = np.diff(Qc.indptr)
    indptr = np.concatenate([[0], np.cumsum(counts)])
    inv = -np.ones(K, np.int64)
    inv[cmap] = np.arange(nold)
    src_rows = inv[np.nonzero(counts)[0]]           # old row of every non-empty new row
    lens = np.diff(Qc.indptr)[src_rows]
    # source positions: concatenated ranges [indptr_old[r], indptr_old[r]+len)
    starts = Qc.indptr[src_rows].astype(np.int64)
    offs = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
    src = np.repeat(starts, lens) + offs
    idx_dtype = np.int32 if K < 2**31 and indptr[-1] < 2**31 else np.int64
    out = sp.csr_matrix((Qc.data[src], cmap[Qc.indices[src]].astype(idx_dtype),
                         indptr.astype(idx_dtype)), shape=(K, K))
    return out
