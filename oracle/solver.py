"""Oracle: QuadraticProblem + QuadraticOptimizer (RTR / tCG) (TEST INFRASTRUCTURE).

Restates
  * src/QuadraticProblem.cpp:38-123 (f, EucGrad, EucHessianEta, PreConditioner,
    RieGrad), :125-136,236-259 (Retract), :138-234 (escapeSaddle),
  * src/QuadraticOptimizer.cpp:28-108 (optimize / trustRegion),
    :135-201 (one RGD step), :234-280 (solver configuration, the
    Max_Iteration==1 radius-shrink retry loop),
  * ROPTLIB RTRNewton / SolversTR::Run / tCG_TR as published
    (yuluntian/ROPTLIB@feature/cmake -- NOT vendored in the reference tree;
    restated from the upstream algorithm: theta=1, kappa=0.1,
    Acceptence_Rho=0.1, Shrinked_tau=0.25, Magnified_tau=2; iteration
    sequences are therefore PARITY UNPINNED, converged results are compared).
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import manifold as mf

# ROPTLIB tCGstatusSet
TR_NEGCURVTURE, TR_EXCREGION, TR_LCON, TR_SCON, TR_MAXITER = range(5)


class ROptParameters:
    """include/DCORA/DCORA_types.h:152-200 defaults."""

    def __init__(self, **kw):
        self.method = "RTR"
        self.verbose = False
        self.gradnorm_tol = 1e-2
        self.RGD_stepsize = 1e-3
        self.RGD_use_preconditioner = True
        self.RTR_iterations = 3
        self.RTR_tCG_iterations = 50
        self.RTR_initial_radius = 100.0
        # reference hard-codes Solver->TimeBound = 5.0 s
        # (QuadraticOptimizer.cpp:252); None disables it for parity runs.
        self.time_bound_s = None
        self.retraction = "qf"
        self.__dict__.update(kw)


class QuadraticProblem:
    def __init__(self, layout: mf.Layout, Q, G=None, precond_reg=0.1,
                 use_precond=True):
        self.L = layout
        self.Q = sp.csr_matrix(Q)
        assert self.Q.shape == (layout.k, layout.k)
        self.G = (np.zeros((layout.r, layout.k)) if G is None
                  else np.asarray(G, np.float64))
        self.reg = precond_reg
        self._lu = None
        self.use_precond = use_precond
        self.n_spmm = 0
        self.n_precond = 0

    # QuadraticProblem.cpp:38-68 (X*Q with symmetric Q == (Q*X^T)^T)
    def XQ(self, X):
        self.n_spmm += 1
        return (self.Q @ X.T).T

    def f(self, X):
        return 0.5 * np.sum(self.XQ(X) * X) + np.sum(X * self.G)

    def egrad(self, X):
        return self.XQ(X) + self.G

    def ehess(self, V):
        return self.XQ(V)

    def rgrad(self, X):
        return mf.tangent_project(self.L, X, self.egrad(X))

    def rgrad_norm(self, X):
        return np.linalg.norm(self.rgrad(X))

    def _factor(self):
        if self._lu is None:
            P = (self.Q + self.reg * sp.identity(self.L.k)).tocsc()
            self._lu = spla.splu(P)   # stands in for CHOLMOD LLT (same solve)
        return self._lu

    def precondition(self, X, V):
        """QuadraticProblem.cpp:70-84: proj_X( solve(V^T)^T )."""
        self.n_precond += 1
        if not self.use_precond:
            return mf.tangent_project(self.L, X, V)
        W = self._factor().solve(np.ascontiguousarray(V.T)).T
        return mf.tangent_project(self.L, X, W)

    def retract(self, X, V, kind="qf"):
        return mf.retract(self.L, X, V, kind)

    def rhess(self, X, EG, V):
        return mf.ehess_to_rhess(self.L, X, EG, V, self.ehess(V))


class ROPTResult:
    def __init__(self):
        self.success = False
        self.fInit = self.gradNormInit = self.fOpt = self.gradNormOpt = 0.0
        self.elapsedMs = 0.0
        self.tCGStatus = TR_MAXITER
        self.outer_iters = 0
        self.inner_iters = 0
        self.accepted = True


def tcg(P: QuadraticProblem, X, EG, g, Delta, max_inner, min_inner=0,
        theta=1.0, kappa=0.1):
    """ROPTLIB SolversTR::tCG_TR with eta0 = 0 (useRand=false)."""
    r = g.copy()
    e_Pe = 0.0
    norm_r0 = np.sqrt(np.sum(r * r))
    z = P.precondition(X, r)
    z_r = np.sum(z * r)
    d_Pd = z_r
    delta = -z
    e_Pd = 0.0
    eta = np.zeros_like(g)
    status = TR_MAXITER
    j = 0
    for j in range(max_inner):
        Hd = P.rhess(X, EG, delta)
        d_Hd = np.sum(delta * Hd)
        alpha = z_r / d_Hd if d_Hd != 0 else np.inf
        e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd
        if d_Hd <= 0 or e_Pe_new >= Delta * Delta:
            tau = (-e_Pd + np.sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd
            eta = eta + tau * delta
            status = TR_NEGCURVTURE if d_Hd <= 0 else TR_EXCREGION
            break
        e_Pe = e_Pe_new
        eta = eta + alpha * delta
        r = r + alpha * Hd
        norm_r = np.sqrt(np.sum(r * r))
        if j >= min_inner and norm_r <= norm_r0 * min(norm_r0 ** theta, kappa):
            status = TR_LCON if kappa < norm_r0 ** theta else TR_SCON
            break
        z = P.precondition(X, r)
        zold_rold = z_r
        z_r = np.sum(z * r)
        beta = z_r / zold_rold
        delta = -z + beta * delta
        e_Pd = beta * (e_Pd + alpha * d_Pd)
        d_Pd = z_r + beta * beta * d_Pd
    return eta, status, j + 1


def rtr_run(P: QuadraticProblem, X0, tol, initial_Delta, maximum_Delta,
            max_iter, max_inner, retraction="qf", time_bound_s=None,
            res: ROPTResult | None = None):
    """ROPTLIB SolversTR::Run (RTRNewton), GRAD_F stop criterion."""
    t0 = time.perf_counter()
    x1 = X0.copy()
    f1 = P.f(x1)
    EG = P.egrad(x1)
    gf1 = mf.tangent_project(P.L, x1, EG)
    ngf = np.linalg.norm(gf1)
    Delta = initial_Delta
    it = 0
    accepted = False
    status = TR_MAXITER
    inner_total = 0
    stop = False
    while (not stop) and it < max_iter:
        eta, status, nin = tcg(P, x1, EG, gf1, Delta, max_inner)
        inner_total += nin
        x2 = P.retract(x1, eta, retraction)
        f2 = P.f(x2)
        Heta = P.rhess(x1, EG, eta)
        rho = (f1 - f2) / (-np.sum(eta * (gf1 + 0.5 * Heta)))
        if rho > 0.75:
            if status in (TR_EXCREGION, TR_NEGCURVTURE):
                Delta = min(2.0 * Delta, maximum_Delta)
        elif rho < 0.25:
            Delta = 0.25 * Delta
        accepted = (rho > 0.1 or
                    (abs(f1 - f2) / (abs(f1) + 1) < np.sqrt(np.finfo(float).eps)
                     and f2 < f1))
        if accepted:
            x1, f1 = x2, f2
            EG = P.egrad(x1)
            gf1 = mf.tangent_project(P.L, x1, EG)
            ngf = np.linalg.norm(gf1)
            stop = ngf < tol
        it += 1
        if time_bound_s is not None and time.perf_counter() - t0 > time_bound_s:
            stop = True
    if res is not None:
        res.tCGStatus = status
        res.outer_iters += it
        res.inner_iters += inner_total
        res.accepted = accepted
    return x1, accepted


class QuadraticOptimizer:
    def __init__(self, problem: QuadraticProblem, params: ROptParameters | None = None):
        self.P = problem
        self.params = params or ROptParameters()
        self.result = ROPTResult()

    def optimize(self, Y):
        """QuadraticOptimizer.cpp:28-50."""
        P, res = self.P, self.result
        res.fInit = P.f(Y)
        res.gradNormInit = P.rgrad_norm(Y)
        t0 = time.perf_counter()
        if self.params.method == "RTR":
            Yopt = self.trust_region(Y)
        else:
            Yopt = self.gradient_descent(Y)
        res.elapsedMs = (time.perf_counter() - t0) * 1e3
        res.fOpt = P.f(Yopt)
        res.gradNormOpt = P.rgrad_norm(Yopt)
        res.success = True
        return Yopt

    def trust_region(self, Y):
        """QuadraticOptimizer.cpp:52-108,234-280."""
        p = self.params
        if self.P.rgrad_norm(Y) < p.gradnorm_tol:
            return Y
        if p.RTR_iterations == 1:
            radius = p.RTR_initial_radius
            total = 0
            while True:
                X, acc = rtr_run(self.P, Y, p.gradnorm_tol, radius, radius, 1,
                                 p.RTR_tCG_iterations, p.retraction,
                                 p.time_bound_s, self.result)
                if acc:
                    return X
                if total > 10:
                    return Y
                radius /= 4
                total += 1
        X, _ = rtr_run(self.P, Y, p.gradnorm_tol, p.RTR_initial_radius,
                       5 * p.RTR_initial_radius, p.RTR_iterations,
                       p.RTR_tCG_iterations, p.retraction, p.time_bound_s,
                       self.result)
        return X

    def gradient_descent(self, Y):
        """One RGD step (QuadraticOptimizer.cpp:124-201)."""
        p, P = self.params, self.P
        g = P.rgrad(Y)
        if p.RGD_use_preconditioner:
            g = P.precondition(Y, g)
        return P.retract(Y, -p.RGD_stepsize * g, p.retraction)


def escape_saddle(P_next: QuadraticProblem, Xopt, theta, v, gradient_tolerance,
                  preconditioned_gradient_tolerance, is_second_order=False,
                  retraction="qf"):
    """QuadraticProblem::escapeSaddle (QuadraticProblem.cpp:138-234).
    ``P_next`` is the rank-(r+1) problem.  Returns (X, success)."""
    r = P_next.L.r
    k = P_next.L.k
    assert Xopt.shape == (r - 1, k)
    X_plus = np.zeros((r, k))
    X_plus[:r - 1] = Xopt
    X_dot = np.zeros((r, k))
    X_dot[r - 1] = v
    alpha_min = 1e-6
    alpha = (max(16 * alpha_min, 100 * gradient_tolerance / abs(theta))
             if is_second_order else 1.0)
    alphas, fvals = [], []
    FX = P_next.f(X_plus)
    while alpha >= alpha_min:
        Xt = P_next.retract(X_plus, alpha * X_dot, retraction)
        Ft = P_next.f(Xt)
        g = P_next.rgrad(Xt)
        gn = np.linalg.norm(g)
        pgn = np.linalg.norm(P_next.precondition(Xt, g))
        alphas.append(alpha)
        fvals.append(Ft)
        if Ft < FX and gn > gradient_tolerance and pgn > preconditioned_gradient_tolerance:
            return Xt, True
        alpha /= 2
    i = int(np.argmin(fvals))
    if fvals[i] < FX:
        return P_next.retract(X_plus, alphas[i] * X_dot, retraction), True
    return None, False
