"""Initial guesses: the reference's host-side initialisation routines,
restated on the product's global measurement arrays.

  odometry_initialization   src/DCORA_solver.cpp:270-302: compose the odometry chain
                            T_{i+1} = T_i * T_{i,i+1} from T_0 = I (or from a partial trajectory)
  chordal_initialization    src/DCORA_solver.cpp:218-268: rotations from the linear least squares
                            min sum kappa ||R_j - R_i R_ij||_F^2 with R_0 = I (B3 of
                            src/DCORA_utils.cpp:1543-1630, solved there with SPQR), projected to
                            SO(d) (projectToRotationGroup); translations from
                            min sum tau ||t_j - t_i - R_i t_ij||^2 with t_0 = 0 (recoverTranslations,
                            :1632-1659)
  lift                      X = YLift * T (Agent.cpp:502)

The reference runs these once per agent on the host (Eigen + SPQR); here they are host code too
(scipy sparse normal equations).  Both return (R (n, d, d), t (n, d))."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .graph import Measurements


def _project_rotations(M):
    """projectToRotationGroup (DCORA_utils.cpp:1661-1675) of a stack (n, d, d)."""
    U, _, Vt = np.linalg.svd(M)
    neg = np.linalg.det(U) * np.linalg.det(Vt) < 0
    U[neg, :, -1] *= -1
    return U @ Vt


def odometry_initialization(meas: Measurements, partial=None):
    """`meas` must hold an odometry edge (i, i+1) for every i; other edges are ignored."""
    d, n = meas.d, meas.n
    odo = {}
    for e in np.nonzero(meas.pp_j == meas.pp_i + 1)[0]:
        odo.setdefault(int(meas.pp_i[e]), int(e))
    R = np.empty((n, d, d))
    t = np.empty((n, d))
    start = 1
    if partial is not None and len(partial[0]):
        m = min(len(partial[0]), n)
        R[:m], t[:m] = partial[0][:m], partial[1][:m]
        start = m
    else:
        R[0], t[0] = np.eye(d), 0.0
    for dst in range(start, n):
        src = dst - 1
        assert src in odo, f"no odometry edge {src} -> {dst}"
        e = odo[src]
        R[dst] = R[src] @ meas.pp_R[e]
        t[dst] = t[src] + R[src] @ meas.pp_t[e]
    return R, t


def chordal_initialization(meas: Measurements):
    d, n, m = meas.d, meas.n, len(meas.pp_i)
    i, j = meas.pp_i, meas.pp_j
    # rotations: unknown rows of R_p stacked as x_p (d x d, one column per row of R);
    # residual of edge e: sqrt(kappa) (x_j - R_ij^T x_i)
    sk = np.sqrt(meas.pp_w * meas.pp_kappa)
    rows = (np.arange(m)[:, None] * d + np.arange(d))                         # (m, d)
    Aj = sp.csr_matrix((np.repeat(sk, d), (rows.reshape(-1), (j[:, None] * d + np.arange(d)).reshape(-1))),
                       shape=(m * d, n * d))
    RT = np.transpose(meas.pp_R, (0, 2, 1)) * sk[:, None, None]               # sqrt(kappa) R_ij^T
    Ai = sp.csr_matrix((RT.reshape(-1),
                        (np.repeat(rows, d, axis=1).reshape(-1),
                         np.tile(i[:, None] * d + np.arange(d), (1, d)).reshape(-1))),
                       shape=(m * d, n * d))
    A = (Aj - Ai).tocsc()
    # R_0 = I: move its columns to the right-hand side
    b = -(A[:, :d] @ np.eye(d))
    Ared = A[:, d:]
    N = (Ared.T @ Ared).tocsc()
    x = spla.splu(N).solve(Ared.T @ b)                                         # ((n-1) d, d)
    Rall = np.empty((n, d, d))
    Rall[0] = np.eye(d)
    Rall[1:] = np.transpose(x.reshape(n - 1, d, d), (0, 2, 1))                # rows back to matrices
    Rall[1:] = _project_rotations(Rall[1:])
    # translations: sqrt(tau) (t_j - t_i - R_i t_ij), t_0 = 0
    st = np.sqrt(meas.pp_w * meas.pp_tau)
    e = np.arange(m)
    B = sp.csr_matrix((np.concatenate([st, -st]), (np.concatenate([e, e]), np.concatenate([j, i]))),
                      shape=(m, n)).tocsc()
    c = st[:, None] * np.einsum("mab,mb->ma", Rall[i], meas.pp_t)
    Bred = B[:, 1:]
    t = np.zeros((n, d))
    t[1:] = spla.splu((Bred.T @ Bred).tocsc()).solve(Bred.T @ c)
    return Rall, t


def lift(gp, R, t, r, landmarks=None, Ylift=None):
    """Rank-r iterate (r, K) in the agent-major reference ordering of `gp` from a trajectory:
    X = YLift * [R_i t_i] (Agent.cpp:502; YLift = [I; 0] by default).  Unit spheres get the
    direction between the initialised end points of their range (DCORA_utils.cpp:1136-1137)."""
    d, n = gp.d, gp.meas.n
    X = np.zeros((d, gp.K))
    idx = np.arange(n)
    X[:, gp.rot_cols(idx).reshape(-1)] = np.transpose(R, (1, 0, 2)).reshape(d, n * d)
    X[:, gp.trans_col(idx)] = t.T
    me = gp.meas
    pos = t if landmarks is None else np.concatenate([t, landmarks])
    if me.b and landmarks is None:
        pos = np.concatenate([t, np.zeros((me.b, d))])
    if me.b:
        X[:, gp.lmk_col(np.arange(me.b))] = pos[n:].T
    if len(me.rg_i):
        u = pos[me.rg_i] - pos[me.rg_j]
        nrm = np.linalg.norm(u, axis=1)
        u[nrm == 0] = np.eye(d)[0]
        nrm[nrm == 0] = 1.0
        X[:, gp.sphere_col(np.arange(len(me.rg_i)))] = (u / nrm[:, None]).T
    Y = np.eye(r, d) if Ylift is None else np.asarray(Ylift)
    return Y @ X
