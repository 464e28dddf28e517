"""Oracle: lifted SE / RA product-manifold closed forms (TEST INFRASTRUCTURE).

Restates
  * src/manifold/LiftedManifold.cpp:18-110 (product structure:
    SE = (Stiefel(r,d) x R^r)^n ; RA = Stiefel(r,d)^n x Oblique(r,l) x R^{r x n}
    x R^{r x b}),
  * the closed forms the reference's tests compare ROPTLIB against:
    src/DCORA_utils.cpp:1677-1711 (projectToStiefelManifold /
    ObliqueManifold / symBlockDiagProduct), :2033-2051 (tangent projections),
    :2201-2220 (projectToSEMatrix / RAMatrix),
  * ROPTLIB (yuluntian/ROPTLIB@feature/cmake, not vendored) Stiefel
    ChooseStieParamsSet3 semantics as published: Euclidean metric, extrinsic
    representation, qf retraction (thin QR, diag(R)>0), EucGrad2Grad =
    projection, EucHvToHv = proj(exix - etax*sym(x^T egf)).  The polar
    retraction is provided as the alternative (variant unpinned).
All matrices are r x k numpy arrays in the reference column ordering.
"""
from __future__ import annotations

import numpy as np


class Layout:
    def __init__(self, r, d, n, l=0, b=0, ra=False):
        self.r, self.d, self.n, self.l, self.b, self.ra = r, d, n, l, b, ra
        if not ra:
            assert l == 0 and b == 0
            self.k = (d + 1) * n
            self.rot = (np.arange(n)[:, None] * (d + 1) + np.arange(d))
            self.sph = np.zeros(0, np.int64)
        else:
            self.k = (d + 1) * n + l + b
            self.rot = (np.arange(n)[:, None] * d + np.arange(d))
            self.sph = d * n + np.arange(l)

    def with_rank(self, r):
        return Layout(r, self.d, self.n, self.l, self.b, self.ra)


def _sym(A):
    return 0.5 * (A + np.swapaxes(A, -1, -2))


def _blocks(L, X):
    """(n, r, d) view-copy of the Stiefel blocks."""
    return np.transpose(X[:, L.rot], (1, 0, 2))


def _set_blocks(L, X, B):
    X[:, L.rot] = np.transpose(B, (1, 0, 2))


def tangent_project(L: Layout, X, V):
    """V - Y sym(Y^T V) per Stiefel block, v - y(y^T v) per sphere column,
    identity elsewhere (DCORA_utils.cpp:2033-2051; testManifold.cpp:354-440)."""
    out = np.array(V, dtype=np.float64, copy=True)
    if L.n:
        Y, W = _blocks(L, X), _blocks(L, V)
        S = _sym(np.einsum("nri,nrj->nij", Y, W))
        _set_blocks(L, out, W - np.einsum("nri,nij->nrj", Y, S))
    if L.l:
        y, v = X[:, L.sph], V[:, L.sph]
        out[:, L.sph] = v - y * np.sum(y * v, axis=0)
    return out


def ehess_to_rhess(L: Layout, X, EG, V, HV):
    """ROPTLIB Stiefel::EucHvToHv (Euclidean metric):
    proj_X(HV - V sym(Y^T EG)); sphere: proj(hv - v (y^T eg)); Euclidean: hv."""
    out = np.array(HV, dtype=np.float64, copy=True)
    if L.n:
        Y, G_, W, H = (_blocks(L, X), _blocks(L, EG), _blocks(L, V),
                       _blocks(L, HV))
        S = _sym(np.einsum("nri,nrj->nij", Y, G_))
        _set_blocks(L, out, H - np.einsum("nri,nij->nrj", W, S))
    if L.l:
        y, g, v = X[:, L.sph], EG[:, L.sph], V[:, L.sph]
        out[:, L.sph] = HV[:, L.sph] - v * np.sum(y * g, axis=0)
    return tangent_project(L, X, out)


def retract(L: Layout, X, V, kind="qf"):
    """ROPTLIB Retraction on the product manifold: Stiefel qf(Y+V) with
    diag(R)>0 (kind='qf') or polar factor (kind='polar'); sphere: normalise
    y+v; Euclidean: p+v.  (QuadraticProblem.cpp:236-259.)"""
    out = X + V
    if L.n:
        A = _blocks(L, out)
        if kind == "qf":
            Qf, Rf = np.linalg.qr(A)
            s = np.sign(np.einsum("nii->ni", Rf))
            s[s == 0] = 1.0
            _set_blocks(L, out, Qf * s[:, None, :])
        else:
            U, _, Vt = np.linalg.svd(A, full_matrices=False)
            _set_blocks(L, out, U @ Vt)
    if L.l:
        y = out[:, L.sph]
        out[:, L.sph] = y / np.linalg.norm(y, axis=0)
    return out


def project_manifold(L: Layout, M):
    """projectToSEMatrix / projectToRAMatrix (DCORA_utils.cpp:2201-2220):
    polar factor U V^T of each r x d block, normalised sphere columns."""
    out = np.array(M, dtype=np.float64, copy=True)
    if L.n:
        U, _, Vt = np.linalg.svd(_blocks(L, M), full_matrices=False)
        _set_blocks(L, out, U @ Vt)
    if L.l:
        y = out[:, L.sph]
        out[:, L.sph] = y / np.linalg.norm(y, axis=0)
    return out


def project_rotation(M):
    """projectToRotationGroup (DCORA_utils.cpp:1661-1675)."""
    U, _, Vt = np.linalg.svd(M)
    if np.linalg.det(U) * np.linalg.det(Vt) > 0:
        return U @ Vt
    U = U.copy()
    U[:, -1] *= -1
    return U @ Vt


def random_point(L: Layout, rng):
    return project_manifold(L, rng.standard_normal((L.r, L.k)))


def lift(L: Layout, X_d, rng=None, Ylift=None):
    """X = YLift * T (Agent.cpp:502): lift a rank-d solution to rank r."""
    if Ylift is None:
        A = rng.standard_normal((L.r, L.d))
        Ylift, _ = np.linalg.qr(A)
    return Ylift @ X_d
