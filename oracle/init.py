"""Oracle: initialisation routines (TEST INFRASTRUCTURE), dense and edge by edge.

chordalInitialization / odometryInitialization (src/DCORA_solver.cpp:218-302) with the B
matrices of constructBMatrices (src/DCORA_utils.cpp:1543-1630) assembled densely, one
measurement at a time, and solved with numpy's least squares (the reference: SPQR)."""
import numpy as np

from . import manifold as mf


def odometry_initialization(pp, n):
    d = pp.d
    T = [None] * n
    T[0] = (np.eye(d), np.zeros(d))
    odo = {}
    for e in range(len(pp)):
        if pp.p2[e] == pp.p1[e] + 1 and int(pp.p1[e]) not in odo:
            odo[int(pp.p1[e])] = e
    for dst in range(1, n):
        e = odo[dst - 1]
        Rs, ts = T[dst - 1]
        T[dst] = (Rs @ pp.R[e], ts + Rs @ pp.t[e])
    return np.array([x[0] for x in T]), np.array([x[1] for x in T])


def chordal_initialization(pp, n):
    d, m = pp.d, len(pp)
    d2 = d * d
    # B3 (69c): for vec(R) column-major per pose: sqrt(kappa) (R_j - R_i R_ij)
    B3 = np.zeros((m * d2, n * d2))
    for e in range(m):
        i, j = int(pp.p1[e]), int(pp.p2[e])
        sk = np.sqrt(pp.weight[e] * pp.kappa[e])
        # vec(R_i R_ij) = (R_ij^T kron I) vec(R_i)
        B3[e * d2:(e + 1) * d2, i * d2:(i + 1) * d2] = -sk * np.kron(pp.R[e].T, np.eye(d))
        B3[e * d2:(e + 1) * d2, j * d2:(j + 1) * d2] = sk * np.eye(d2)
    cR = B3[:, :d2] @ np.eye(d).reshape(-1, order="F")
    rvec = -np.linalg.lstsq(B3[:, d2:], cR, rcond=None)[0]
    R = np.empty((n, d, d))
    R[0] = np.eye(d)
    for i in range(1, n):
        R[i] = mf.project_rotation(rvec[(i - 1) * d2:i * d2].reshape(d, d, order="F"))
    # B1, B2 (69a, 69b): sqrt(tau) (t_j - t_i - R_i t_ij)
    B1 = np.zeros((m * d, n * d))
    c = np.zeros(m * d)
    for e in range(m):
        i, j = int(pp.p1[e]), int(pp.p2[e])
        st = np.sqrt(pp.weight[e] * pp.tau[e])
        B1[e * d:(e + 1) * d, i * d:(i + 1) * d] = -st * np.eye(d)
        B1[e * d:(e + 1) * d, j * d:(j + 1) * d] = st * np.eye(d)
        c[e * d:(e + 1) * d] = -st * (R[i] @ pp.t[e])
    tred = -np.linalg.lstsq(B1[:, d:], c, rcond=None)[0]
    t = np.zeros((n, d))
    t[1:] = tred.reshape(n - 1, d)
    return R, t
