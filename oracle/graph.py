"""Oracle: data matrices Q, G of one agent's graph (TEST INFRASTRUCTURE).

Restates src/Graph.cpp of the reference:
  * addMeasurement / duplicate dropping / n,l,b bookkeeping   :82-300
  * constructQuadraticCostTermPGO                              :579-683
  * constructLinearCostTermPGO                                 :685-822
  * constructQuadraticCostTermRASLAM                           :824-1188
  * constructLinearCostTermRASLAM                              :1190-1772
  * setIndicesFromStateOwnership                               :1774-1816
  * constructPreconditioner / regularisation                   :1901-1960

Column orderings (include/DCORA/manifold/LiftedVariable.h:27-30,148-152):
  SE: [Y1 p1 | Y2 p2 | ...],               k = (d+1) n
  RA: [Y1..Yn | r1..rl | p1..pn | L1..Lb], k = (d+1) n + l + b
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .dataio import POSE, LANDMARK, Measurements, PosePose, PoseLandmark, Range


def _dedupe(keys):
    """Indices of first occurrences (Graph drops duplicate EdgeIDs, :139-141,
    :162-169)."""
    if len(keys) == 0:
        return np.zeros(0, np.int64)
    _, first = np.unique(keys, axis=0, return_index=True)
    return np.sort(first)


class Graph:
    """One agent's graph.  ``ra=False`` is GraphType::PoseGraph (SE ordering),
    ``ra=True`` is GraphType::RangeAidedSLAMGraph (RA ordering)."""

    def __init__(self, agent_id, r, d, ra=False):
        self.id, self.r, self.d, self.ra = agent_id, r, d, ra
        self.n = self.l = self.b = 0
        self.pp = PosePose.empty(d)
        self.pl = PoseLandmark.empty(d)
        self.rg = Range.empty()
        self.prior = {}            # pose index -> (r x (d+1)) prior (setPrior)
        self.prior_kappa = 10000.0  # include/DCORA/Graph.h prior weights
        self.prior_tau = 100.0

    # -- measurements ------------------------------------------------------
    def set_measurements(self, meas: Measurements):
        me = self.id
        pp, pl, rg = meas.pp, meas.pl, meas.rg
        keep = (pp.r1 == me) | (pp.r2 == me)
        pp = pp.take(np.nonzero(keep)[0])
        pp = pp.take(_dedupe(np.stack([pp.r1, pp.p1, pp.r2, pp.p2], 1)))
        keep = (pl.r1 == me) | (pl.r2 == me)
        pl = pl.take(np.nonzero(keep)[0])
        pl = pl.take(_dedupe(np.stack([pl.r1, pl.p1, pl.r2, pl.p2], 1)))
        keep = (rg.r1 == me) | (rg.r2 == me)
        rg = rg.take(np.nonzero(keep)[0])
        if len(rg):
            k = np.stack([rg.type1, rg.r1, rg.p1, rg.type2, rg.r2, rg.p2], 1)
            assert len(_dedupe(k)) == len(rg), "range measurements must be unique"
        self.pp, self.pl, self.rg = pp, pl, rg
        n = b = l = 0
        for r_, p_, ty in ((pp.r1, pp.p1, None), (pp.r2, pp.p2, None),
                           (pl.r1, pl.p1, None)):
            s = r_ == me
            if s.any():
                n = max(n, int(p_[s].max()) + 1)
        s = pl.r2 == me
        if s.any():
            b = max(b, int(pl.p2[s].max()) + 1)
        for r_, p_, ty in ((rg.r1, rg.p1, rg.type1), (rg.r2, rg.p2, rg.type2)):
            s = (r_ == me) & (ty == POSE)
            if s.any():
                n = max(n, int(p_[s].max()) + 1)
            s = (r_ == me) & (ty == LANDMARK)
            if s.any():
                b = max(b, int(p_[s].max()) + 1)
        s = rg.r1 == me
        if s.any():
            l = int(rg.l[s].max()) + 1
        self.n, self.l, self.b = n, l, b
        if not self.ra:
            assert l == 0 and b == 0, "pose graph cannot contain spheres/landmarks"
        return self

    @property
    def k(self):
        return (self.d + 1) * self.n + self.l + self.b

    # -- column helpers ----------------------------------------------------
    def rot_cols(self, i):
        d = self.d
        i = np.asarray(i)
        base = i * d if self.ra else i * (d + 1)
        return base[..., None] + np.arange(d)

    def trans_col(self, i):
        d = self.d
        i = np.asarray(i)
        return (d * self.n + self.l + i) if self.ra else (i * (d + 1) + d)

    def sphere_col(self, i):
        return self.d * self.n + np.asarray(i)

    def landmark_col(self, i):
        return self.d * self.n + self.l + self.n + np.asarray(i)

    # -- Q -------------------------------------------------------------------
    def quadratic_matrix(self):
        return (self._Q_ra() if self.ra else self._Q_pgo()).tocsr()

    def _Q_pgo(self):
        """Q = AbT * Omega * AbT^T (Graph.cpp:579-683)."""
        d, n, me, pp = self.d, self.n, self.id, self.pp
        dh, m = d + 1, len(pp)
        own_i, own_j = pp.r1 == me, pp.r2 == me
        k = np.arange(m)
        rows, cols, vals = [], [], []
        # AT(i,k) = -[R t; 0 1]
        ki = k[own_i]
        i = pp.p1[own_i]
        for c in range(d):
            for r in range(d):
                rows.append(i * dh + r); cols.append(ki * dh + c)
                vals.append(-pp.R[own_i, r, c])
        for r in range(d):
            rows.append(i * dh + r); cols.append(ki * dh + d)
            vals.append(-pp.t[own_i, r])
        rows.append(i * dh + d); cols.append(ki * dh + d)
        vals.append(-np.ones(len(i)))
        # AT(j,k) = +I
        kj = k[own_j]
        j = pp.p2[own_j]
        for r in range(dh):
            rows.append(j * dh + r); cols.append(kj * dh + r)
            vals.append(np.ones(len(j)))
        AbT = sp.csr_matrix((np.concatenate(vals),
                             (np.concatenate(rows), np.concatenate(cols))),
                            shape=(dh * n, dh * m))
        om = np.empty(dh * m)
        for r in range(d):
            om[r::dh] = pp.weight * pp.kappa
        om[d::dh] = pp.weight * pp.tau
        return AbT @ sp.diags(om) @ AbT.T

    def _Q_ra(self):
        """Graph.cpp:824-1188."""
        d, n, l, b, me = self.d, self.n, self.l, self.b, self.id
        pp, pl, rg = self.pp, self.pl, self.rg
        mpp, mpl, mrg = len(pp), len(pl), len(rg)
        mpose = mpp + mpl

        def coo(shape, trip):
            if not trip:
                return sp.csr_matrix(shape)
            r_ = np.concatenate([t[0] for t in trip])
            c_ = np.concatenate([t[1] for t in trip])
            v_ = np.concatenate([np.broadcast_to(t[2], t[0].shape) for t in trip])
            return sp.csr_matrix((v_, (r_, c_)), shape=shape)

        arho, atau, tt = [], [], []
        k = np.arange(mpp)
        oi, oj = pp.r1 == me, pp.r2 == me
        i, ki = pp.p1[oi], k[oi]
        for c in range(d):
            for r in range(d):
                arho.append((i * d + r, ki * d + c, -pp.R[oi, r, c]))
        for r in range(d):
            tt.append((i * d + r, ki, -pp.t[oi, r]))
        atau.append((i, ki, -1.0))
        j, kj = pp.p2[oj], k[oj]
        for r in range(d):
            arho.append((j * d + r, kj * d + r, 1.0))
        atau.append((j, kj, 1.0))
        k = np.arange(mpl) + mpp
        oi, oj = pl.r1 == me, pl.r2 == me
        i, ki = pl.p1[oi], k[oi]
        for r in range(d):
            tt.append((i * d + r, ki, -pl.t[oi, r]))
        atau.append((i, ki, -1.0))
        atau.append((pl.p2[oj] + n, k[oj], 1.0))
        ARhoT = coo((d * n, d * mpp), arho)
        ATauT = coo((n + b, mpose), atau)
        TT = coo((d * n, mpose), tt)
        om_rho = np.repeat(pp.weight * pp.kappa, d)
        om_tau = np.concatenate([pp.weight * pp.tau, pl.weight * pl.tau])

        k = np.arange(mrg)
        oi, oj = rg.r1 == me, rg.r2 == me
        ct = [(rg.p1[oi] + n * (rg.type1[oi] == LANDMARK), k[oi], -1.0),
              (rg.p2[oj] + n * (rg.type2[oj] == LANDMARK), k[oj], 1.0)]
        CT = coo((n + b, mrg), ct)
        PT = coo((l, mrg), [(rg.l[oi], k[oi], 1.0)])
        om_rg = rg.weight * rg.precision
        D = sp.diags(rg.range) if mrg else sp.csr_matrix((0, 0))
        ORho, OTau = sp.diags(om_rho), sp.diags(om_tau)
        ORg = sp.diags(om_rg) if mrg else sp.csr_matrix((0, 0))

        Q11 = ARhoT @ ORho @ ARhoT.T + TT @ OTau @ TT.T
        Q13 = TT @ OTau @ ATauT.T
        Q22 = PT @ ORg @ D @ D @ PT.T
        Q23 = PT @ D @ ORg @ CT.T
        Q33 = ATauT @ OTau @ ATauT.T + CT @ ORg @ CT.T
        Z12 = sp.csr_matrix((d * n, l))
        return sp.bmat([[Q11, Z12, Q13],
                        [Z12.T, Q22, Q23],
                        [Q13.T, Q23.T, Q33]], format="csr")

    # -- G ---------------------------------------------------------------------
    def linear_matrix(self, nbr):
        """Per-measurement restatement of constructLinearCostTerm{PGO,RASLAM}
        (Graph.cpp:685-822, 1190-1772).  ``nbr`` maps
        ('pose',robot,frame) -> r x (d+1), ('lmk',robot,frame) -> r,
        ('sph',robot,idx) -> r.  Returns None when a neighbour state is
        missing (the reference returns false and the iteration is skipped)."""
        d, r, me = self.d, self.r, self.id
        G = np.zeros((r, self.k))
        pp, pl, rg = self.pp, self.pl, self.rg

        def add_pose(i, L):
            G[:, self.rot_cols(i)] += L[:, :d]
            G[:, self.trans_col(i)] += L[:, d]

        for e in np.nonzero(pp.r1 != pp.r2)[0]:
            T = np.eye(d + 1)
            T[:d, :d], T[:d, d] = pp.R[e], pp.t[e]
            Om = np.diag([pp.weight[e] * pp.kappa[e]] * d + [pp.weight[e] * pp.tau[e]])
            if pp.r1[e] == me:      # outgoing: AbT=-T, AcT=I  (:766-777)
                Xc = nbr.get(("pose", int(pp.r2[e]), int(pp.p2[e])))
                if Xc is None:
                    return None
                add_pose(pp.p1[e], Xc @ Om @ (-T).T)
            else:                   # incoming: AbT=I, AcT=-T   (:778-791)
                Xc = nbr.get(("pose", int(pp.r1[e]), int(pp.p1[e])))
                if Xc is None:
                    return None
                add_pose(pp.p2[e], Xc @ (-T) @ Om)
        for idx, P in self.prior.items():       # :805-816
            Om = np.diag([self.prior_kappa] * d + [self.prior_tau])
            add_pose(idx, -P @ Om)
        if not self.ra:
            return G

        for e in np.nonzero(pl.r1 != pl.r2)[0]:
            w = pl.weight[e] * pl.tau[e]
            t = pl.t[e]
            if pl.r1[e] == me:      # own pose i, neighbour landmark j (:1558-1572)
                Xc = nbr.get(("lmk", int(pl.r2[e]), int(pl.p2[e])))
                if Xc is None:
                    return None
                # Q_cb_31 = AcTau*w*Tb = (+1) w (-t)^T ; Q_cb_33 = (+1) w (-1)
                L = np.zeros((r, d + 1))
                L[:, :d] = np.outer(Xc, -w * t)
                L[:, d] = -w * Xc
                add_pose(pl.p1[e], L)
            else:                   # own landmark j, neighbour pose i (:1573-1589)
                Xc = nbr.get(("pose", int(pl.r1[e]), int(pl.p1[e])))
                if Xc is None:
                    return None
                # L = Xrot*Q_cb_13 + Xtr*Q_cb_33 ; Q_cb_13 = (-t) w (+1), Q_cb_33=(-1)w(+1)
                G[:, self.landmark_col(pl.p2[e])] += Xc[:, :d] @ (-w * t) - w * Xc[:, d]

        for e in np.nonzero(rg.r1 != rg.r2)[0]:
            w, rho = rg.weight[e] * rg.precision[e], rg.range[e]
            if rg.r1[e] == me:      # own src i + own sphere l, neighbour dst (:1611-1668)
                key = ("pose" if rg.type2[e] == POSE else "lmk",
                       int(rg.r2[e]), int(rg.p2[e]))
                Xc = nbr.get(key)
                if Xc is None:
                    return None
                xt = Xc[:, d] if rg.type2[e] == POSE else Xc
                # Q_cb_33 = Cc w Cb = (+1) w (-1); Q_cb_32 = Cc w D Pb = w rho
                col = (self.trans_col(rg.p1[e]) if rg.type1[e] == POSE
                       else self.landmark_col(rg.p1[e]))
                G[:, col] += -w * xt
                G[:, self.sphere_col(rg.l[e])] += w * rho * xt
            else:                   # own dst j, neighbour src + neighbour sphere (:1669-1760)
                key = ("pose" if rg.type1[e] == POSE else "lmk",
                       int(rg.r1[e]), int(rg.p1[e]))
                Xc = nbr.get(key)
                Xs = nbr.get(("sph", int(rg.r1[e]), int(rg.l[e])))
                if Xc is None or Xs is None:
                    return None
                xt = Xc[:, d] if rg.type1[e] == POSE else Xc
                col = (self.trans_col(rg.p2[e]) if rg.type2[e] == POSE
                       else self.landmark_col(rg.p2[e]))
                # Q_cb_33 = Cc w Cb = (-1) w (+1); Q_cb_23 = Pc D w Cb = rho w (+1)
                G[:, col] += -w * xt + rho * w * Xs
        return G

    # -- ownership lists (Graph.h:256-268) -------------------------------------
    def public_ids(self):
        """(pose frames, landmark frames, sphere idx) of MY variables that
        appear in shared measurements (loc_shared_*_ids_)."""
        me = self.id
        pp, pl, rg = self.pp, self.pl, self.rg
        poses, lmks, sph = set(), set(), set()
        sh = pp.r1 != pp.r2
        poses.update(pp.p1[sh & (pp.r1 == me)].tolist())
        poses.update(pp.p2[sh & (pp.r2 == me)].tolist())
        sh = pl.r1 != pl.r2
        poses.update(pl.p1[sh & (pl.r1 == me)].tolist())
        lmks.update(pl.p2[sh & (pl.r2 == me)].tolist())
        sh = rg.r1 != rg.r2
        s = sh & (rg.r1 == me)
        poses.update(rg.p1[s & (rg.type1 == POSE)].tolist())
        lmks.update(rg.p1[s & (rg.type1 == LANDMARK)].tolist())
        sph.update(rg.l[s].tolist())
        s = sh & (rg.r2 == me)
        poses.update(rg.p2[s & (rg.type2 == POSE)].tolist())
        lmks.update(rg.p2[s & (rg.type2 == LANDMARK)].tolist())
        return sorted(poses), sorted(lmks), sorted(sph)

    def neighbor_ids(self):
        me = self.id
        out = set()
        for a in (self.pp, self.pl, self.rg):
            out.update(a.r1[a.r1 != me].tolist())
            out.update(a.r2[a.r2 != me].tolist())
        return sorted(out)


def se_to_ra_perm(d, n):
    """perm such that X_ra = X_se[:, perm] for l=b=0 (Elements.cpp:170-202)."""
    rot = (np.arange(n)[:, None] * (d + 1) + np.arange(d)).ravel()
    tr = np.arange(n) * (d + 1) + d
    return np.concatenate([rot, tr])


def precond_regularization(Q, ra):
    """Graph.cpp:1901-1960: 0.1 for PGO; lambda_max(Q)/(1e6-1) for RA
    (Spectra SymEigsSolver nev=1 ncv=6 tol 1e-3 -> ARPACK here)."""
    if not ra:
        return 0.1
    import scipy.sparse.linalg as spla
    try:
        lam = spla.eigsh(Q, k=1, which="LA", ncv=6, tol=1e-3, maxiter=10000,
                         return_eigenvectors=False)[0]
    except Exception:
        return 0.1
    return lam / (1e6 - 1)
