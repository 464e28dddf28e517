"""Oracle: dual certificate + minimum-eigenvalue test (TEST INFRASTRUCTURE).

Restates src/DCORA_utils.cpp:
  * constructDualCertificateMatrixPGO      :1898-1931
  * constructDualCertificateMatrixRASLAM   :1933-1982
  * fastVerification                       :1713-1735
  * isSparseSymmetricMatrixPSD             :1737-1747  (CHOLMOD LLT succeeds)
  * computeMinimumEigenPair (spectrum shift) :1809-1896, (shift-invert) :1751-1805
  * projectSolutionRASLAM                  :1984-2031
Spectra v1.0.1 (implicitly restarted Lanczos) and CHOLMOD are third-party and
not vendored; ARPACK (scipy eigsh, the same IRL algorithm) and a dense/ARPACK
PSD decision stand in.  PARITY UNPINNED by the reference's tests (it has
none for certification); anchored on the noiseless fixtures (S >= 0, lambda_min
= 0 at ground truth) and on a dense eigendecomposition for small problems.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import manifold as mf


def lambda_blocks(L: mf.Layout, X, Q):
    """Returns (QXt^T as r x k, stiefel Lambda blocks (n,d,d), oblique (l,))."""
    XQ = (Q @ X.T).T
    Y = np.transpose(X[:, L.rot], (1, 0, 2))          # (n, r, d)
    QY = np.transpose(XQ[:, L.rot], (1, 0, 2))        # rows of QXt for rot i
    Pm = np.einsum("nri,nrj->nij", QY, Y)             # (QXt_i) * X_i
    lam = 0.5 * (Pm + np.swapaxes(Pm, 1, 2))
    ob = np.sum(X[:, L.sph] * XQ[:, L.sph], axis=0) if L.l else np.zeros(0)
    return XQ, lam, ob


def certificate_matrix(L: mf.Layout, X, Q):
    """S = Q - Lambda (sparse, reference ordering)."""
    _, lam, ob = lambda_blocks(L, X, Q)
    d, n = L.d, L.n
    rows = np.repeat(L.rot[:, :, None], d, axis=2).ravel()
    cols = np.repeat(L.rot[:, None, :], d, axis=1).ravel()
    Lam = sp.csr_matrix((lam.ravel(), (rows, cols)), shape=(L.k, L.k))
    if L.l:
        Lam = Lam + sp.csr_matrix((ob, (L.sph, L.sph)), shape=(L.k, L.k))
    return (sp.csr_matrix(Q) - Lam).tocsr()


def min_eig_pair(M, max_iterations=1000, min_eig_num_tol=1e-4, ncv=20, rng=None):
    """computeMinimumEigenPair, spectrum-shift variant (:1809-1896)."""
    k = M.shape[0]
    ncv = min(ncv, k)
    if k <= 64:
        w, V = np.linalg.eigh(M.toarray())
        return w[0], V[:, 0]
    lam_lm, v = spla.eigsh(M, k=1, which="LM", ncv=ncv,
                           maxiter=max_iterations * k, tol=1e-4)
    lam_lm = lam_lm[0]
    if lam_lm < 0:
        return lam_lm, v[:, 0]
    C = M - 2 * lam_lm * sp.identity(k)
    v0 = np.asarray(M.getrow(0).todense()).ravel()
    rng = rng or np.random.default_rng(0)
    pert = rng.standard_normal(k)
    pert /= np.linalg.norm(pert)
    xinit = v0 + 0.03 * np.linalg.norm(v0) * pert
    try:
        w, v = spla.eigsh(C, k=1, which="LM", ncv=ncv, v0=xinit,
                          maxiter=max_iterations * k,
                          tol=min_eig_num_tol / lam_lm)
        return w[0] + 2 * lam_lm, v[:, 0]
    except spla.ArpackNoConvergence:
        # shift-invert fallback (:1751-1805), sigma = -10 then halving
        sigma = -10.0
        for _ in range(10):
            try:
                w, v = spla.eigsh(M, k=1, sigma=sigma, which="LM", ncv=ncv)
                return w[0], v[:, 0]
            except Exception:
                sigma /= 2
        raise


def lambda_min(S, tol=1e-10, ncv=40):
    """Tight lambda_min(S) for the 1e-6-relative parity check (the reference
    never computes it when S+eta*I is PSD; SURVEY 7 hard-part 6)."""
    k = S.shape[0]
    if k <= 3000:
        return float(np.linalg.eigvalsh(S.toarray())[0])
    lm = spla.eigsh(S, k=1, which="LM", ncv=20, tol=1e-6,
                    return_eigenvectors=False)[0]
    C = S - abs(lm) * 1.01 * sp.identity(k)
    w = spla.eigsh(C, k=1, which="LM", ncv=ncv, tol=tol, maxiter=200000,
                   return_eigenvectors=False)[0]
    return float(w + abs(lm) * 1.01)


def is_psd(M):
    """Stands in for CHOLMOD supernodal LLT success (:1737-1747)."""
    k = M.shape[0]
    if k <= 3000:
        return bool(np.linalg.eigvalsh(M.toarray())[0] > 0)
    try:
        # symmetric-mode sparse LU with diagonal pivots only: the pivots are
        # those of LDL^T, so all pivots positive <=> SPD.
        lu = spla.splu(M.tocsc(), permc_spec="MMD_AT_PLUS_A",
                       diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        if not np.array_equal(lu.perm_r, lu.perm_c):
            return False   # an off-diagonal pivot was needed: a zero pivot
        return bool(np.all(lu.U.diagonal() > 0))
    except RuntimeError:
        return False


def fast_verification(S, eta, rng=None):
    """fastVerification (:1713-1735).  Returns (is_psd, theta, v)."""
    k = S.shape[0]
    M = (S + eta * sp.identity(k)).tocsr()
    if is_psd(M):
        return True, None, None
    lam, v = min_eig_pair(M, 1000, eta, rng=rng)
    theta = float(v @ (S @ v))
    return False, theta, v


def project_solution_ra(L: mf.Layout, X):
    """projectSolutionRASLAM (:1984-2031).  L is the rank-r layout of X."""
    d, n = L.d, L.n
    U, s, _ = np.linalg.svd(X.T, full_matrices=False)
    Xp = (U[:, :d] * s[:d]).T
    Ld = L.with_rank(d)
    B = np.transpose(Xp[:, Ld.rot], (1, 0, 2))
    if np.sum(np.linalg.det(B) > 0) < n // 2:
        Rf = np.eye(d)
        Rf[d - 1, d - 1] = -1
        Xp = Rf @ Xp
        B = np.transpose(Xp[:, Ld.rot], (1, 0, 2))
    for i in range(n):
        Xp[:, Ld.rot[i]] = mf.project_rotation(B[i])
    if L.l:
        y = Xp[:, Ld.sph]
        Xp[:, Ld.sph] = y / np.linalg.norm(y, axis=0)
    return Xp


def align_lifted_trajectory(X, Tw0, d, is_global=True):
    """alignLiftedTrajectoryToFrame (DCORA_utils.cpp:2262-2289).  X: (r, (d+1) n) pose-graph
    ordering, Tw0: (r, d+1) lifted frame.  Returns the (d, (d+1) n) trajectory."""
    r, k = X.shape
    n = k // (d + 1)
    R0T = Tw0[:, :d].T
    T = R0T @ X
    ta = Tw0[:, d] if is_global else T[:, d]
    t0 = R0T @ ta          # (local alignment needs r == d, as in the reference)
    for i in range(n):
        c = i * (d + 1)
        T[:, c:c + d] = mf.project_rotation(T[:, c:c + d])
        T[:, c + d] -= t0
    return T


def count_below(S, sigma):
    """Number of eigenvalues of the symmetric S below sigma (Sylvester's law of inertia on the
    pivots of a symmetric-mode sparse LU without row exchanges, like is_psd); None if the
    factorisation needed an off-diagonal pivot."""
    k = S.shape[0]
    M = (S - sigma * sp.identity(k)).tocsc()
    try:
        lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                       options=dict(SymmetricMode=True))
    except RuntimeError:
        return None
    if not np.array_equal(lu.perm_r, lu.perm_c):
        return None
    return int(np.sum(lu.U.diagonal() < 0))


def lambda_min_bisect(S, rel=1e-3, floor=1e-7):
    """lambda_min(S) for large S without trusting a Krylov iteration to find the right end of
    a clustered spectrum: bracket it by inertia counts, bisect to a relative width `rel`, then
    one shift-invert Lanczos step from the lower end of the bracket (nearest eigenvalue above
    it = lambda_min).  Returns (lambda_min, v)."""
    lo, hi = -1e-3, 1e-3
    while True:
        c = count_below(S, lo)
        if c == 0:
            break
        hi, lo = lo, lo * 4.0
    while (count_below(S, hi) or 0) == 0:
        lo, hi = hi, hi * 4.0
    while hi - lo > max(rel * max(abs(lo), abs(hi)), floor):
        mid = 0.5 * (lo + hi)
        c = count_below(S, mid)
        if c is None:
            mid = lo + 0.37 * (hi - lo)
            c = count_below(S, mid)
        if c == 0:
            lo = mid
        else:
            hi = mid
    w, v = spla.eigsh(S.tocsc(), k=1, sigma=lo - 0.1 * (hi - lo) - 1e-12, which="LM", tol=1e-12)
    return float(w[0]), v[:, 0]
