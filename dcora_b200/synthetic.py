"""Synthetic 3-D grid pose graphs with range edges (BASELINE.json config 5,
SURVEY.md 8d) and the spanning-tree initial guess used for them.

Grid of S x S x S poses with unit spacing; the pose numbering follows a
boustrophedon path through all cells (so that consecutive poses are grid
neighbours and "odometry" is part of the edge set); one relative-pose edge per
pair of 6-neighbour cells, oriented from the lower to the higher pose index;
noise: translation sigma_t (tau = 1/sigma_t^2), rotation a small-angle
Gaussian of std 1/sqrt(2 kappa) per axis (the isotropic Langevin concentration
kappa the reference's tests use, tests/testRobust.cpp:175-176); range edges
between random pose pairs within `range_radius`, unique (src, dst)
(src/Graph.cpp:174-181), precision 1/sigma_r^2.  All draws come from one
seeded numpy Generator.
"""
from __future__ import annotations

import numpy as np

from .graph import Measurements


def _rotvec_to_R(v):
    """Rodrigues formula, (m,3) -> (m,3,3)."""
    th = np.linalg.norm(v, axis=1)
    k = v / np.maximum(th, 1e-300)[:, None]
    Kx = np.zeros((len(v), 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)


def _random_rotations(rng, m):
    q = rng.standard_normal((m, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = np.empty((m, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def path_index(S):
    """idx[x, y, z] = position of cell (x, y, z) on the boustrophedon path."""
    x, y, z = np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij")
    # inside plane z the rows y are swept forwards (even z) or backwards (odd z);
    # inside a row x runs forwards / backwards alternately so the path is connected
    yy = np.where(z % 2 == 0, y, S - 1 - y)
    row = z * S + yy                      # global row counter along the path
    xx = np.where(row % 2 == 0, x, S - 1 - x)
    return (row * S + xx).astype(np.int64)


def grid3d(S, range_fraction=0.1, seed=0, sigma_t=0.1, kappa=10000.0, sigma_r=0.1,
           range_radius=3, noiseless=False):
    """Returns (Measurements, ground truth (R (n,3,3), p (n,3)), cell index map)."""
    rng = np.random.default_rng(seed)
    n = S ** 3
    idx = path_index(S)
    pos = np.empty((n, 3))
    gx, gy, gz = np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij")
    pos[idx.reshape(-1)] = np.stack([gx, gy, gz], -1).reshape(-1, 3).astype(np.float64)
    Rgt = _random_rotations(rng, n)
    # 6-neighbour edges
    ei, ej = [], []
    for ax in range(3):
        sl_a = [slice(None)] * 3
        sl_b = [slice(None)] * 3
        sl_a[ax], sl_b[ax] = slice(0, S - 1), slice(1, S)
        a, b = idx[tuple(sl_a)].reshape(-1), idx[tuple(sl_b)].reshape(-1)
        ei.append(np.minimum(a, b)); ej.append(np.maximum(a, b))
    ei, ej = np.concatenate(ei), np.concatenate(ej)
    order = np.lexsort((ej, ei))
    ei, ej = ei[order], ej[order]
    m = len(ei)
    tau = 1.0 / sigma_t ** 2
    ns = 0.0 if noiseless else 1.0
    Rn = _rotvec_to_R(ns * rng.standard_normal((m, 3)) / np.sqrt(2.0 * kappa))
    Rij = np.transpose(Rgt[ei], (0, 2, 1)) @ Rgt[ej] @ Rn
    tij = np.einsum("mji,mj->mi", Rgt[ei], pos[ej] - pos[ei]) + ns * sigma_t * rng.standard_normal((m, 3))
    kw = dict(d=3, n=n, pp_i=ei, pp_j=ej, pp_R=Rij, pp_t=tij,
              pp_kappa=np.full(m, kappa), pp_tau=np.full(m, tau))
    nr = int(round(range_fraction * m))
    if nr > 0:
        # offsets within the radius (excluding 0)
        o = np.arange(-range_radius, range_radius + 1)
        off = np.stack(np.meshgrid(o, o, o, indexing="ij"), -1).reshape(-1, 3)
        nn = (off ** 2).sum(1)
        off = off[(nn > 0) & (nn <= range_radius ** 2)]
        src_cells, dst_cells = [], []
        have = 0
        seen = np.zeros(0, np.int64)
        while have < nr:
            q = int((nr - have) * 1.6) + 16
            c = rng.integers(0, S, size=(q, 3))
            t = c + off[rng.integers(0, len(off), size=q)]
            ok = ((t >= 0) & (t < S)).all(1)
            c, t = c[ok], t[ok]
            s_i, d_i = idx[c[:, 0], c[:, 1], c[:, 2]], idx[t[:, 0], t[:, 1], t[:, 2]]
            key = s_i * n + d_i
            _, first = np.unique(key, return_index=True)
            first = np.sort(first)
            key, s_i, d_i = key[first], s_i[first], d_i[first]
            new = ~np.isin(key, seen)
            key, s_i, d_i = key[new][:nr - have], s_i[new][:nr - have], d_i[new][:nr - have]
            seen = np.concatenate([seen, key])
            src_cells.append(s_i); dst_cells.append(d_i)
            have += len(key)
        ri, rj = np.concatenate(src_cells), np.concatenate(dst_cells)
        rho = np.linalg.norm(pos[rj] - pos[ri], axis=1) + ns * sigma_r * rng.standard_normal(nr)
        kw.update(rg_i=ri, rg_j=rj, rg_range=np.abs(rho), rg_prec=np.full(nr, 1.0 / sigma_r ** 2))
    return Measurements(**kw), (Rgt, pos), idx


def spanning_tree_init(meas: Measurements, S, idx):
    """Initial guess by composing relative-pose measurements along a spanning
    tree of the grid rooted at cell (0,0,0): along x on the first row, along y
    in the first plane, then along z (tree depth <= 3(S-1), so drift stays
    small).  Stands in for the reference's centralised chordal initialisation
    (DCORA_solver.cpp:218-302, out of scope here).  Returns (R (n,3,3), t (n,3))
    with pose of cell (0,0,0) at the identity."""
    n = meas.n
    key = meas.pp_i * n + meas.pp_j
    order = np.argsort(key)
    skey = key[order]

    def edge_of(a, b):
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        e = order[np.searchsorted(skey, lo * n + hi)]
        return e, (a == lo)          # forward: measurement is a -> b

    R = np.zeros((n, 3, 3))
    t = np.zeros((n, 3))
    R[idx[0, 0, 0]] = np.eye(3)

    def step(a, b):
        e, fwd = edge_of(a, b)
        Rm, tm = meas.pp_R[e], meas.pp_t[e]
        # forward: T_b = T_a * T_m ; backward: T_b = T_a * T_m^-1
        Rinv = np.transpose(Rm, (0, 2, 1))
        tinv = -np.einsum("mij,mj->mi", Rinv, tm)
        Ruse = np.where(fwd[:, None, None], Rm, Rinv)
        tuse = np.where(fwd[:, None], tm, tinv)
        R[b] = R[a] @ Ruse
        t[b] = t[a] + np.einsum("mij,mj->mi", R[a], tuse)

    for x in range(1, S):
        step(idx[x - 1:x, 0, 0].reshape(-1), idx[x:x + 1, 0, 0].reshape(-1))
    for y in range(1, S):
        step(idx[:, y - 1, 0].reshape(-1), idx[:, y, 0].reshape(-1))
    for z in range(1, S):
        step(idx[:, :, z - 1].reshape(-1), idx[:, :, z].reshape(-1))
    return R, t


def lifted_initial_guess(meas: Measurements, idx, gp, S, r):
    """Initial iterate of the synthetic grid: the spanning-tree guess lifted to rank r with
    YLift = [I_d; 0] (Agent::initialize lifts with a shared YLift, Agent.cpp:502); unit
    spheres from the guessed translations (sign convention DCORA_utils.cpp:1136-1137).
    Returns X (r, gp.K) in the reference column ordering of `gp`."""
    R, t = spanning_tree_init(meas, S, idx)
    d, n = meas.d, meas.n
    X = np.zeros((r, gp.K))
    X[:d, gp.rot_cols(np.arange(n)).reshape(-1)] = np.transpose(R, (1, 0, 2)).reshape(d, -1)
    X[:d, gp.trans_col(np.arange(n))] = t.T
    if len(meas.rg_i):
        dv = t[meas.rg_i] - t[meas.rg_j]
        dv /= np.maximum(np.linalg.norm(dv, axis=1), 1e-12)[:, None]
        X[:d, gp.sphere_col(np.arange(len(meas.rg_i)))] = dv.T
    return X
