"""Shared helpers of the test-suite: golden fixture loading, oracle problems,
and the GPU problem built through the C-ABI from the same Q."""
from __future__ import annotations

import os

import numpy as np

from oracle import dataio, graph, manifold as mf, solver

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, prefix=""):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = int(z["d"])

    def get(grp, key):
        return z[f"{prefix}{grp}_{key}"]
    pp = dataio.PosePose(d, get("pp", "r1"), get("pp", "p1"), get("pp", "r2"),
                         get("pp", "p2"), get("pp", "R"), get("pp", "t"),
                         get("pp", "kappa"), get("pp", "tau"), get("pp", "weight"))
    pl = dataio.PoseLandmark(d, get("pl", "r1"), get("pl", "p1"), get("pl", "r2"),
                             get("pl", "p2"), get("pl", "t"), get("pl", "tau"),
                             get("pl", "weight"))
    rg = dataio.Range(get("rg", "r1"), get("rg", "p1"), get("rg", "r2"), get("rg", "p2"),
                      get("rg", "type1"), get("rg", "type2"), get("rg", "l"),
                      get("rg", "range"), get("rg", "precision"), get("rg", "weight"))
    meas = dataio.Measurements(d, pp, pl, rg)
    info = dict(kind=str(z["kind"]), n=int(z["n"]), l=int(z["l"]), b=int(z["b"]),
                has_gt=bool(z["has_gt"]), X_gt=z["X_gt"], d=d)
    return meas, info


def central_problem(name, r, use_precond=True):
    """Oracle problem + Graph of the centralised problem of a golden dataset."""
    meas, info = load_golden(name)
    ra = info["kind"] == "pyfg"
    g = graph.Graph(0, r, info["d"], ra=ra).set_measurements(meas)
    Q = g.quadratic_matrix()
    L = mf.Layout(r, info["d"], g.n, g.l, g.b, ra=ra)
    reg = graph.precond_regularization(Q, ra)
    P = solver.QuadraticProblem(L, Q, precond_reg=reg, use_precond=use_precond)
    return P, g, info


def gpu_problem(P, precond=True, device=0):
    """dcora_b200.QuadraticProblem from an oracle problem's Q (through the C-ABI)."""
    import dcora_b200 as dc
    L = P.L
    ag = dc.AgentData(n=L.n, l=L.l, b=L.b, Q=P.Q,
                      G=P.G if np.any(P.G) else None,
                      precond_reg=P.reg if precond else None)
    return dc.QuadraticProblem(ag, L.r, L.d, dc.ORDER_RA if L.ra else dc.ORDER_SE, device)


def split_contiguous(meas, n, num_robots):
    """The partition rule of examples/MultiRobotExample.cpp:56-118: contiguous
    ranges of n/num_robots poses per robot, the last robot takes the rest."""
    per = n // num_robots
    rob = np.minimum(np.arange(n) // per, num_robots - 1)
    start = np.arange(num_robots) * per
    pp = meas.pp
    r1, r2 = rob[pp.p1], rob[pp.p2]
    out = dataio.PosePose(meas.d, r1, pp.p1 - start[r1], r2, pp.p2 - start[r2],
                          pp.R, pp.t, pp.kappa, pp.tau, pp.weight)
    counts = np.bincount(rob, minlength=num_robots)
    return dataio.Measurements(meas.d, pp=out), counts, start


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def product_measurements(meas, num_robots=None):
    """oracle per-robot (robot id, local index) measurements -> the product's
    global-index Measurements + Partition (robot-major numbering)."""
    from dcora_b200 import graph as pg
    pp, pl, rg = meas.pp, meas.pl, meas.rg
    A = num_robots or int(max([a.max() for a in (pp.r1, pp.r2, pl.r1, pl.r2, rg.r1, rg.r2)
                               if len(a)]) + 1)
    n_a = np.zeros(A, np.int64)
    b_a = np.zeros(A, np.int64)

    def upd(cnt, r_, p_, sel):
        if sel.any():
            np.maximum.at(cnt, r_[sel], p_[sel] + 1)
    t = np.ones(len(pp), bool)
    upd(n_a, pp.r1, pp.p1, t); upd(n_a, pp.r2, pp.p2, t)
    t = np.ones(len(pl), bool)
    upd(n_a, pl.r1, pl.p1, t); upd(b_a, pl.r2, pl.p2, t)
    upd(n_a, rg.r1, rg.p1, rg.type1 == dataio.POSE); upd(n_a, rg.r2, rg.p2, rg.type2 == dataio.POSE)
    upd(b_a, rg.r1, rg.p1, rg.type1 == dataio.LANDMARK)
    upd(b_a, rg.r2, rg.p2, rg.type2 == dataio.LANDMARK)
    poff = np.concatenate([[0], np.cumsum(n_a)])
    boff = np.concatenate([[0], np.cumsum(b_a)])
    n, b = int(poff[-1]), int(boff[-1])

    def tv(r_, p_, ty):
        return np.where(ty == dataio.POSE, poff[r_] + p_, n + boff[r_] + p_)
    M = pg.Measurements(
        d=meas.d, n=n, b=b,
        pp_i=poff[pp.r1] + pp.p1, pp_j=poff[pp.r2] + pp.p2, pp_R=pp.R, pp_t=pp.t,
        pp_kappa=pp.kappa, pp_tau=pp.tau, pp_w=pp.weight,
        pl_i=poff[pl.r1] + pl.p1, pl_j=boff[pl.r2] + pl.p2, pl_t=pl.t, pl_tau=pl.tau,
        pl_w=pl.weight,
        rg_i=tv(rg.r1, rg.p1, rg.type1), rg_j=tv(rg.r2, rg.p2, rg.type2),
        rg_range=rg.range, rg_prec=rg.precision, rg_w=rg.weight)
    part = pg.Partition(A, np.repeat(np.arange(A), n_a), np.repeat(np.arange(A), b_a))
    return M, part


def write_g2o(path, meas, name=None):
    """Write the pose-pose measurements of a golden fixture as a g2o file whose
    information matrices reproduce kappa / tau under the reference reader's
    formulas (DCORA_utils.cpp:288-292,337-346): Itt = tau I, Irr = 2 kappa I (3-D),
    I33 = kappa (2-D).  No VERTEX lines (like data/kitti_00.g2o).  3-D rotations
    are written from the raw file quaternions (tests/golden/quats.npz) when the
    dataset `name` has them: the reference reader does not normalise quaternions."""
    from scipy.spatial.transform import Rotation
    pp = meas.pp
    d = meas.d
    with open(path, "w") as fh:
        if d == 3:
            quats = np.load(os.path.join(GOLDEN, "quats.npz"))
            q = quats[name] if name in quats.files else Rotation.from_matrix(pp.R).as_quat()
            for e in range(len(pp)):
                tau, kap = pp.tau[e], pp.kappa[e]
                info = np.zeros((6, 6))
                info[:3, :3] = tau * np.eye(3)
                info[3:, 3:] = 2.0 * kap * np.eye(3)
                up = [info[a, b] for a in range(6) for b in range(a, 6)]
                vals = list(pp.t[e]) + list(q[e]) + up
                fh.write(f"EDGE_SE3:QUAT {pp.p1[e]} {pp.p2[e]} " +
                         " ".join(f"{v:.17g}" for v in vals) + "\n")
        else:
            th = np.arctan2(pp.R[:, 1, 0], pp.R[:, 0, 0])
            for e in range(len(pp)):
                tau, kap = pp.tau[e], pp.kappa[e]
                vals = [pp.t[e][0], pp.t[e][1], th[e], tau, 0.0, 0.0, tau, 0.0, kap]
                fh.write(f"EDGE_SE2 {pp.p1[e]} {pp.p2[e]} " +
                         " ".join(f"{v:.17g}" for v in vals) + "\n")


def oracle_measurements(M, part):
    """The reverse of product_measurements: the product's global-index Measurements +
    Partition -> the oracle's per-robot (robot id, local index) measurement arrays.  Unit
    spheres are numbered per source robot in measurement order (DCORA_utils.cpp:1092-1094)."""
    rob_p = np.asarray(part.robot_of_pose, np.int64)
    rob_l = np.asarray(part.robot_of_lmk, np.int64)
    A = int(part.num_robots)

    def local(rob):
        loc = np.zeros(len(rob), np.int64)
        for a in range(A):
            sel = rob == a
            loc[sel] = np.arange(int(sel.sum()))
        return loc
    loc_p, loc_l = local(rob_p), local(rob_l)
    d, n = M.d, M.n
    pp = dataio.PosePose(d, rob_p[M.pp_i], loc_p[M.pp_i], rob_p[M.pp_j], loc_p[M.pp_j],
                         M.pp_R, M.pp_t, M.pp_kappa, M.pp_tau, M.pp_w)
    pl = dataio.PoseLandmark(d, rob_p[M.pl_i], loc_p[M.pl_i], rob_l[M.pl_j], loc_l[M.pl_j],
                             M.pl_t, M.pl_tau, M.pl_w)

    def endpoint(v):
        is_l = v >= n
        vp, vl = np.where(is_l, 0, v), np.where(is_l, v - n, 0)
        rob = np.where(is_l, rob_l[vl] if len(rob_l) else 0, rob_p[vp])
        idx = np.where(is_l, loc_l[vl] if len(loc_l) else 0, loc_p[vp])
        return rob, idx, np.where(is_l, dataio.LANDMARK, dataio.POSE)
    r1, p1, t1 = endpoint(M.rg_i)
    r2, p2, t2 = endpoint(M.rg_j)
    sph = np.zeros(len(r1), np.int64)
    cnt = np.zeros(A, np.int64)
    for e in range(len(r1)):
        sph[e] = cnt[r1[e]]
        cnt[r1[e]] += 1
    rg = dataio.Range(r1, p1, r2, p2, t1, t2, sph, M.rg_range, M.rg_prec, M.rg_w)
    return dataio.Measurements(d, pp, pl, rg)


def write_pyfg(path, M, part, gt=None):
    """Write the product's global-index Measurements + Partition as a PyFG file (the format of
    data/*.pyfg, read by src/DCORA_utils.cpp:437-1167): robot letters A, B, ..; pose i of robot
    R is "R<local index>", landmark j "LR<local index>".  Covariances are chosen so that the
    reader's formulas give back kappa / tau / precision: cov_t = I / tau, cov_R = I / (2 kappa)
    (3-D) or 1 / kappa (2-D), range cov = 1 / precision.  gt = (R (n, d, d), p (n, d)) fills
    the VERTEX lines (identity / zero otherwise); landmarks are written at the origin."""
    from scipy.spatial.transform import Rotation
    d, n = M.d, M.n
    rob_p = np.asarray(part.robot_of_pose, np.int64)
    rob_l = np.asarray(part.robot_of_lmk, np.int64) if M.b else np.zeros(0, np.int64)

    def local(rob):
        loc = np.zeros(len(rob), np.int64)
        for a in range(int(part.num_robots)):
            sel = rob == a
            loc[sel] = np.arange(int(sel.sum()))
        return loc
    loc_p, loc_l = local(rob_p), local(rob_l)
    psym = [chr(65 + int(rob_p[i])) + str(int(loc_p[i])) for i in range(n)]
    lsym = ["L" + chr(65 + int(rob_l[j])) + str(int(loc_l[j])) for j in range(M.b)]
    Rg = gt[0] if gt is not None else np.tile(np.eye(d), (n, 1, 1))
    pg = gt[1] if gt is not None else np.zeros((n, d))

    def pose_fields(R, t):
        if d == 2:
            return list(t) + [float(np.arctan2(R[1, 0], R[0, 0]))]
        return list(t) + list(Rotation.from_matrix(R).as_quat())

    def upper(diag):
        k = len(diag)
        return [diag[i] if i == j else 0.0 for i in range(k) for j in range(i, k)]
    f = lambda vals: " ".join(f"{v:.17g}" for v in vals)
    with open(path, "w") as fh:
        for i in range(n):
            fh.write(f"VERTEX_SE{d}{':QUAT' if d == 3 else ''} 0.0 {psym[i]} {f(pose_fields(Rg[i], pg[i]))}\n")
        for j in range(M.b):
            fh.write(f"VERTEX_XY{'Z' if d == 3 else ''} {lsym[j]} {f(np.zeros(d))}\n")
        for e in range(len(M.pp_i)):
            cov = [1.0 / M.pp_tau[e]] * d + ([1.0 / (2.0 * M.pp_kappa[e])] * 3 if d == 3
                                             else [1.0 / M.pp_kappa[e]])
            fh.write(f"EDGE_SE{d}{':QUAT' if d == 3 else ''} 0.0 {psym[M.pp_i[e]]} {psym[M.pp_j[e]]} "
                     f"{f(pose_fields(M.pp_R[e], M.pp_t[e]))} {f(upper(cov))}\n")
        for e in range(len(M.pl_i)):
            fh.write(f"EDGE_SE{d}_XY{'Z' if d == 3 else ''} 0.0 {psym[M.pl_i[e]]} {lsym[M.pl_j[e]]} "
                     f"{f(M.pl_t[e])} {f(upper([1.0 / M.pl_tau[e]] * d))}\n")
        for e in range(len(M.rg_i)):
            s = [psym[v] if v < n else lsym[v - n] for v in (int(M.rg_i[e]), int(M.rg_j[e]))]
            fh.write(f"EDGE_RANGE 0.0 {s[0]} {s[1]} {M.rg_range[e]:.17g} {1.0 / M.rg_prec[e]:.17g}\n")
