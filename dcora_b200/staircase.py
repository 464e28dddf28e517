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
    state: object = None     # rbcd_staircase: the RBCD state of the last level
    rounded: object = None   # rbcd_staircase(round_solution=True): (d, K) SE(d) solution


def riemannian_staircase(problem: QuadraticProblem, X0, r_max=None, eta=1e-3,
                         params: ROptParameters | None = None, escape_tol=1e-6,
                         certify_tol=1e-9, max_matvecs=20000,
                         second_order_escape=False) -> StaircaseResult:
    """`problem` is the rank-r0 problem of the start X0 (r0 x K).  `second_order_escape`: the
    isSecondOrder flag of escapeSaddle (true in SingleRobotExample_RASLAM.cpp:272-277, the
    centralised problem; false in the multi-robot drivers)."""
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
        Xn, ok = Pn.escapeSaddle(X, cert.theta, v, escape_tol, escape_tol,
                                  isSecondOrder=second_order_escape)
        if not ok:
            break
        P, X, r = Pn, Xn, r + 1
    res.seconds = time.perf_counter() - t_start
    return res


# ---------------------------------------------------------------------------------------------
# the multi-agent staircase: RBCD(++) -> distributed certificate -> distributed escape -> r + 1
# (examples/MultiRobotExample.cpp:172-364, examples/MultiRobotExample_RASLAM.cpp:204-512)

def escape_distributed(S_next, X_dev, v_dev, theta, gradient_tolerance,
                       preconditioned_gradient_tolerance, is_second_order=False):
    """QuadraticProblem::escapeSaddle (src/QuadraticProblem.cpp:138-234) on the partitioned
    problem: `S_next` is the rank r+1 RBCD state, X_dev (K_local, r) the critical point,
    v_dev (K_local,) the local rows of the unit direction of negative curvature.  Cost and
    gradient norm are the global ones (monitor); the preconditioned gradient norm uses the
    agents' own factors (block Jacobi), the distributed stand-in of the centralised
    preconditioner.  Every rank calls this collectively.  Returns (X_next (K_local, r+1), ok)."""
    t = S_next.torch
    K, r = X_dev.shape
    Xp = t.zeros((K, r + 1), dtype=t.float64, device=X_dev.device)
    Xp[:, :r] = X_dev
    Xdot = t.zeros_like(Xp)
    Xdot[:, r] = v_dev
    alpha_min = 1e-6
    alpha = (max(16 * alpha_min, 100 * gradient_tolerance / abs(theta))
             if is_second_order else 1.0)

    def evaluate(X):
        S_next.set_device_X(X)
        f, gn = S_next.monitor()                 # leaves G(X) in the handle
        rg = S_next.P.RieGrad(X)
        pg = S_next.P.PreCondition(X, rg)
        p2 = (pg * pg).sum().reshape(1)
        if S_next.world > 1:
            import torch.distributed as dist
            dist.all_reduce(p2, group=S_next.group)
        return f, gn, float(p2.sqrt())
    f0, _, _ = evaluate(Xp)
    trials = []
    while alpha >= alpha_min:
        Xt = S_next.P.Retract(Xp, alpha * Xdot)
        f, gn, pgn = evaluate(Xt)
        trials.append((f, alpha))
        if f < f0 and gn > gradient_tolerance and pgn > preconditioned_gradient_tolerance:
            return Xt, True
        alpha /= 2
    f, a = min(trials)
    if f < f0:
        Xt = S_next.P.Retract(Xp, a * Xdot)
        S_next.set_device_X(Xt)
        return Xt, True
    return None, False


def rbcd_staircase(S, X0_global=None, r_max=None, eta=1e-3, gradnorm_tol=0.1, max_iters=1000,
                   check_every=5, certify_tol=1e-8, escape_tol=1e-4, round_solution=False,
                   log=None, certify_max_iters=400, stall_checks=0,
                   adaptive_restart=False) -> StaircaseResult:
    """Distributed Riemannian staircase on an RBCD state `S` (dcora_b200.rbcd.RBCD of any world
    size; every rank calls this collectively).  The drivers' defaults: RGradNormTol 0.1,
    min_eig_num_tol 1e-3, <= 1000 iterations per level (MultiRobotExample.cpp:121-133).

    Per level: parallel-coloured RBCD / RBCD++ sweeps until the global gradient norm is below
    `gradnorm_tol` -> certify_distributed -> (not PSD) escape_distributed to rank r + 1 on a
    handle that shares Q, the Cholesky factors and the coupling with the previous level.
    `adaptive_restart`: drop the Nesterov momentum whenever the monitored gradient norm went up
    between two checks (gradient-based adaptive restart) on top of the fixed restart interval.
    `stall_checks` > 0: when an accelerated level has not improved its best gradient norm by
    20 % over that many consecutive checks, it continues with plain (monotone) RBCD sweeps.
    Returns a StaircaseResult whose X is the LOCAL (r, K_local) iterate of this rank."""
    from .certify_dist import certify_distributed
    t_start = time.perf_counter()
    if X0_global is not None:
        S.set_X(X0_global)
    r = S.r
    r_max = r_max or r + 5
    levels, res = [], None
    while True:
        t0 = time.perf_counter()
        f = gn = None
        sweeps = 0
        best, since_best, switched_at = float("inf"), 0, None
        prev_gn, n_adaptive = None, 0
        for it in range(int(max_iters)):
            S.iterate()
            sweeps += 1
            if sweeps % check_every == 0 or it + 1 == max_iters:
                f, gn = S.monitor()
                if log:
                    log(f"rank {r} sweep {sweeps}: cost {2 * f:.8g} gradnorm {gn:.4g} "
                        f"({time.perf_counter() - t_start:.1f} s)")
                if gn < gradnorm_tol:
                    break
                if adaptive_restart and S.acc and prev_gn is not None and gn > prev_gn:
                    S.restart_acceleration()
                    n_adaptive += 1
                prev_gn = gn
                if gn < 0.8 * best:
                    best, since_best = gn, 0
                else:
                    since_best += 1
                if stall_checks and S.acc and since_best >= stall_checks:
                    S.disable_acceleration()
                    switched_at = sweeps
                    if log:
                        log(f"rank {r} sweep {sweeps}: RBCD++ stalled at gradnorm {best:.3g}, "
                            f"continuing with plain RBCD")
        t1 = time.perf_counter()
        cert = certify_distributed(S, eta, tol=certify_tol, max_iters=certify_max_iters)
        t2 = time.perf_counter()
        levels.append(dict(r=r, f=f, gradnorm=gn, sweeps=sweeps, plain_rbcd_from=switched_at,
                           adaptive_restarts=n_adaptive,
                           is_psd=bool(cert.is_psd),
                           lambda_min=cert.lambda_min, theta=cert.theta, residual=cert.residual,
                           products=cert.products, converged=bool(cert.converged),
                           solve_s=t1 - t0, certify_s=t2 - t1))
        if log:
            log(f"rank {r}: lambda_min {cert.lambda_min:.4g} psd {cert.is_psd} "
                f"({cert.products} block products, {t2 - t1:.2f} s)")
        res = StaircaseResult(None, r, bool(cert.is_psd), f, gn, cert.lambda_min, cert.theta, levels)
        if cert.is_psd or r + 1 > r_max:
            break
        if cert.theta >= -eta / 2:       # MultiRobotExample.cpp:333: direction not resolved
            break
        Sn = S.with_rank(r + 1)
        Xn, ok = escape_distributed(Sn, S.X, cert.v, cert.theta, escape_tol, escape_tol)
        if not ok:
            break
        Sn.set_device_X(Xn)
        S, r = Sn, r + 1
    res.X = S.get_X()
    res.state = S
    if round_solution and S.world == 1:
        res.rounded = S.P.round_solution(S.X).cpu().numpy().T
    res.seconds = time.perf_counter() - t_start
    return res
