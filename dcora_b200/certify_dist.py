"""Certification of the PARTITIONED problem across ranks (last row).

fastVerification of the centralised dual certificate S = Q - Lambda(X)
(src/DCORA_utils.cpp:1713-1735,1898-1982) for an iterate that lives distributed over
the ranks of an RBCD run: rows/columns of S are partitioned like the agents, the
halo of every product is the same public-column exchange RBCD uses (one all-gather),
and the small projected problems are all-reduced.

The product S V is assembled from existing device entry points, for a block V of r
vectors at once (the C-ABI works on K x r matrices; the r rows of V^T are independent
vectors for a right-multiplication by the symmetric Q):
    (V Q_global)_local = V Q_diag + [V | halo(V)] C      update_G(V, halo) ; egrad(V)
    (V Lambda)_local   : Lambda = sym-block-diag(X^T (X Q_global)) on the rotation
                         blocks, <r_k, (X Q)_k> on the unit spheres -- built and applied by the
                         library (dcora_problem_certificate_prepare / _apply, csrc/certify.cu).
lambda_min is computed like dcora_certify does on one handle (csrc/certify.cu): a Krylov-type
iteration from a RANDOM start (here block Davidson / Rayleigh-Ritz with block size r and thick
restart, residuals preconditioned by the agents' own Cholesky factors = block Jacobi) until its smallest Ritz pair has converged, then a Rayleigh-Ritz step on span(rows of
X) -- at a critical point they span the null space of S, the cluster a random start resolves
last -- and the smaller of the two upper bounds is reported.  (The rows of X must NOT seed the
iteration: they are exact eigenvectors, their Ritz pair "converges" at once and would hide a
negative eigenvalue at a saddle.)  torch is used for the tall-skinny dense algebra of the iteration (Gram
matrices, basis rotations: library GEMMs) and NCCL; S V itself runs in the library's kernels.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class DistCertificate:
    is_psd: bool
    lambda_min: float
    residual: float
    iterations: int
    products: int          # block products S V (each = r matrix-vector products)
    converged: bool
    theta: float = 0.0     # v^T S v of the returned direction (fastVerification's theta)
    v: object = None       # local rows of the unit eigenvector estimate (device tensor, K_local)


def certify_distributed(S, eta, tol=1e-9, max_iters=400, max_basis=None, seed=0,
                        early_exit=False, precondition=True) -> DistCertificate:
    """`S` is a dcora_b200.rbcd.RBCD (any world size) holding the iterate; every rank
    calls this collectively and gets the same result."""
    t = S.torch
    dist = None
    if S.world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
    P, r, dev = S.P, S.r, S.device
    K = S.plan.K
    f64 = t.float64

    def allsum(x):
        if dist is not None:
            dist.all_reduce(x, group=S.group)
        return x

    # the certificate products overwrite the handle's linear term and pass many
    # short-lived buffers: no iterate tracking while we are here
    P.set_iterate_tracking(False)
    try:
        pack = t.from_numpy(np.asarray(S.plan.pack_idx, np.int64)).to(dev)

        def halo_of(V):
            if S.world == 1:
                return None
            send = V.index_select(0, pack).contiguous()
            H = t.empty((S.world * send.shape[0], r), dtype=f64, device=dev)
            dist.all_gather_into_tensor(H, send, group=S.group)
            return H

        def times_Q(V):
            """(V Q_global) restricted to the local columns; V is (K, r)."""
            P.update_G(V, halo_of(V), None)
            return P.EucGrad(V)

        # Lambda from the iterate (constructDualCertificateMatrix*, :1898-1982): kept on the
        # handle by dcora_problem_certificate_prepare, applied by ..._certificate_apply
        X = S.X
        XQ = times_Q(X)
        P.certificate_prepare(X, XQ)

        nprod = 0

        def times_S(V):
            nonlocal nprod
            nprod += 1
            return P.certificate_apply(V, times_Q(V))

        def gram(A, B):
            return allsum(A.T @ B)

        def orthonormalise(W, B):
            """W (K, r) made orthonormal to the columns of B (K, s) and within itself;
            returns the kept columns (possibly fewer than r)."""
            for _ in range(2):
                if B is not None and B.shape[1]:
                    W = W - B @ gram(B, W)
            G = gram(W, W)
            ev, U = t.linalg.eigh(G)
            keep = ev > 1e-20 * max(float(ev.max()), 1e-300)
            if not bool(keep.any()):
                return W[:, :0]
            W = W @ (U[:, keep] / ev[keep].sqrt())
            if B is not None and B.shape[1]:
                W = W - B @ gram(B, W)
                nrm = allsum((W * W).sum(0)).sqrt()
                W = W / nrm
            return W

        # random start block: rows [c0, c0 + K) of ONE global random matrix, so that the
        # iteration is the same for every partition of the problem over ranks
        g = t.Generator(device="cpu").manual_seed(seed)
        R0 = t.rand((S.gp.K, r), generator=g, dtype=f64)[S.plan.c0:S.plan.c0 + K].to(dev) - 0.5
        B = orthonormalise(R0, None)
        # S applied column block by column block (pad to r columns)
        def times_S_cols(M):
            out = []
            for c0 in range(0, M.shape[1], r):
                blk = M[:, c0:c0 + r]
                w = blk.shape[1]
                if w < r:
                    blk = t.cat([blk, t.zeros((K, r - w), dtype=f64, device=dev)], 1)
                out.append(times_S(blk.contiguous())[:, :w])
            return t.cat(out, 1)

        SB = times_S_cols(B)
        smax = max_basis or max(12 * r, 40)
        keep_n = 3 * r
        lam, resid, it, conv = 0.0, float("inf"), 0, False
        scale = 1.0
        for it in range(1, max_iters + 1):
            Hm = gram(B, SB)
            Hm = 0.5 * (Hm + Hm.T)
            th, Y = t.linalg.eigh(Hm)
            scale = max(scale, float(th.abs().max()))
            b = min(r, B.shape[1])
            Xr = B @ Y[:, :b]
            Rr = SB @ Y[:, :b] - Xr * th[:b]
            rn = allsum((Rr * Rr).sum(0)).sqrt()
            lam, resid = float(th[0]), float(rn[0])
            if resid <= tol * scale:
                conv = True
                break
            # early decision (optional): a Ritz value is an upper bound of lambda_min
            if early_exit and lam < -eta and resid <= 0.1 * (-eta - lam):
                conv = True
                break
            if B.shape[1] + r > smax:            # thick restart on the smallest Ritz vectors
                kk = min(keep_n, B.shape[1])
                B, SB = B @ Y[:, :kk], SB @ Y[:, :kk]
            # Davidson correction: the residuals through the agents' Cholesky factors
            # (block-Jacobi (Q_aa + reg I)^-1: the Laplacian part of S that makes a plain
            # Krylov iteration slow), then orthonormalised against the basis
            Wc = Rr
            if precondition and P.factor_nnz() > 0:
                if b < r:
                    Wc = t.cat([Rr, t.zeros((K, r - b), dtype=f64, device=dev)], 1)
                Wc = P.factor_solve(Wc.contiguous())[:, :b]
            W = orthonormalise(Wc, B)
            if W.shape[1] == 0:
                conv = True
                break
            B = t.cat([B, W], 1)
            SB = t.cat([SB, times_S_cols(W)], 1)
        # Rayleigh-Ritz on span(rows of X)
        Bx = orthonormalise(X.clone(), None)
        if Bx.shape[1]:
            Hx = gram(Bx, times_S_cols(Bx))
            mu = float(t.linalg.eigvalsh(0.5 * (Hx + Hx.T))[0])
            lam = min(lam, mu)
        # the escape direction: the smallest Ritz vector of the random-start iteration
        v = Xr[:, 0].contiguous()
        v = v / allsum((v * v).sum().reshape(1)).sqrt()
        Vp = t.zeros((K, r), dtype=f64, device=dev)
        Vp[:, 0] = v
        theta = float(allsum((times_S(Vp)[:, 0] * v).sum().reshape(1)))
        # Verdict: `lam` is an upper bound of lambda_min(S).  Below -eta it proves "not PSD"
        # whether or not the iteration converged; "PSD" is only reported from a CONVERGED
        # smallest Ritz pair whose residual leaves a margin (there is no global factorisation
        # of the partitioned S to run the reference's exact test on)
        psd = bool(conv and lam + eta > 0.0 and lam - resid + eta > 0.0)
        return DistCertificate(is_psd=psd, lambda_min=lam, residual=resid,
                               iterations=it, products=nprod, converged=conv, theta=theta, v=v)
    finally:
        # restore the linear term of the RBCD state and re-arm the tracking
        S.exchange()
        P.update_G(S.X, S.H, None)
        if not S.acc:
            P.set_iterate_tracking(True)
