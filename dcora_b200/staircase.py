"""Riemannian staircase with certification on one (centralised) problem.

Host-side control flow of the reference drivers
(examples/SingleRobotExample_RASLAM.cpp:182-283, examples/MultiRobotExample.cpp:172-364 for a
single robot): solve at rank r, test the dual certificate S = Q - Lambda(X)
(fastVerification), on failure escapeSaddle to rank r + 1.  All arithmetic runs on the device
through dcora_b200.api; Q and its Cholesky factor are uploaded once and shared by every rank
(QuadraticProblem.with_rank).  The C++ twin is examples/staircase_pgo.cpp."""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from .api import QuadraticOptimizer, QuadraticProblem, ROptParameters


@dataclass
class StaircaseResult:
    X: np.ndarray
    rank: int
    certified: bool
    cost: float            # f(X) (the drivers print 2 f)
    gradnorm: float
    lambda_min: float
    theta: float
    levels: list = field(default_factory=list)   # per rank: dict of stats
    seconds: float = 0.0


def riemannian_staircase(problem: QuadraticProblem, X0, r_max=None, eta=1e-3,
                         params: ROptParameters | None = None, escape_tol=1e-6,
                         certify_tol=1e-9, max_matvecs=20000) -> StaircaseResult:
    """`problem` is the rank-r0 problem of the start X0 (r0 x K)."""
    t_start = time.perf_counter()
    params = params or ROptParameters(RTR_iterations=300, RTR_tCG_iterations=500,
                                      gradnorm_tol=1e-6, time_bound_s=0.0)
    P = problem
    X = np.asarray(X0, dtype=np.float64)
    r = P.r
    r_max = r_max or r + 6
    levels = []
    res = None
    while True:
        t0 = time.perf_counter()
        opt = QuadraticOptimizer(P, params)
        X = opt.optimize(X)
        st = opt.getOptResult()
        t1 = time.perf_counter()
        cert, v = P.certify(X, eta, max_matvecs=max_matvecs, tol=certify_tol)
        t2 = time.perf_counter()
        levels.append(dict(r=r, f=st.fOpt, gradnorm=st.gradNormOpt, outer=st.outer_iterations,
                           inner=st.inner_iterations, is_psd=bool(cert.is_psd),
                           lambda_min=cert.lambda_min, theta=cert.theta, matvecs=cert.matvecs,
                           solve_s=t1 - t0, certify_s=t2 - t1))
        res = StaircaseResult(X, r, bool(cert.is_psd), st.fOpt, st.gradNormOpt, cert.lambda_min,
                              cert.theta, levels)
        if cert.is_psd or r + 1 > r_max:
            break
        # MultiRobotExample.cpp:336-339: the escape direction must be resolved
        if cert.theta >= -eta / 2:
            break
        Pn = P.with_rank(r + 1)
        Xn, ok = Pn.escapeSaddle(X, cert.theta, v, escape_tol, escape_tol)
        if not ok:
            break
        P, X, r = Pn, Xn, r + 1
    res.seconds = time.perf_counter() - t_start
    return res
