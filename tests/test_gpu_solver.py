"""GPU parity tests of QuadraticOptimizer (RTR/tCG), certification and the
saddle escape against the oracle and the known answers."""
import numpy as np
import pytest

from oracle import certify, manifold as mf, solver
from tests.common import central_problem, gpu_problem, rel_err

pytestmark = pytest.mark.gpu


def _params(**kw):
    import dcora_b200 as dc
    return dc.ROptParameters(time_bound_s=0.0, **kw)


@pytest.mark.parametrize("name", ["pose_graph_optimization_test_2d",
                                  "pose_graph_optimization_test_3d",
                                  "range_aided_slam_test_2d",
                                  "range_aided_slam_test_3d"])
def test_ground_truth_is_a_fixed_point(name):
    # tests/testAgent.cpp:101-242: iterate() at the optimum keeps it (1e-9)
    import dcora_b200 as dc
    r = 2 if "2d" in name else 3
    P, g, info = central_problem(name, r)
    gp = gpu_problem(P)
    X = info["X_gt"]
    assert abs(gp.f(X)) < 1e-12
    opt = dc.QuadraticOptimizer(gp, _params())
    Xo = opt.optimize(X)
    assert np.allclose(Xo, X, atol=1e-9)
    res, v = gp.certify(X, 1e-4)
    assert res.is_psd == 1
    assert res.lambda_min > -1e-6


@pytest.mark.parametrize("name,r,tol,iters", [("smallGrid3D", 5, 1e-6, 300),
                                              ("smallGrid3D", 9, 1e-6, 300),
                                              ("tinyGrid3D", 4, 1e-7, 300),
                                              ("range_aided_slam_test_3d", 5, 1e-7, 3000)])
def test_rtr_converged_cost_matches_oracle(name, r, tol, iters):
    # north_star: final cost relative error <= 1e-8 at the same tolerance.
    # The range-aided fixture is noiseless (f* = 0, so the bound is taken
    # relative to fInit) and, with the reference's tiny RA regularisation
    # lambda_max/(1e6-1) (Graph.cpp:1919-1960), RTR needs ~600 outer iterations
    # from a random point on the CPU oracle as well.
    import dcora_b200 as dc
    P, g, info = central_problem(name, r)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(1))
    kw = dict(RTR_iterations=iters, RTR_tCG_iterations=500, gradnorm_tol=tol)
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters(**kw))
    Xo = oo.optimize(X0)
    go = dc.QuadraticOptimizer(gp, _params(**kw))
    Xg = go.optimize(X0)
    rg = go.getOptResult()
    assert rg.gradNormOpt < tol
    assert abs(rg.fInit - oo.result.fInit) <= 1e-12 * abs(oo.result.fInit)
    assert oo.result.gradNormOpt < tol
    assert abs(rg.fOpt - oo.result.fOpt) <= 1e-8 * abs(oo.result.fOpt) + 1e-13 * abs(rg.fInit)
    # the returned statistics are those of the returned point
    assert abs(P.f(Xg) - rg.fOpt) <= 1e-10 * abs(rg.fOpt) + 1e-13 * abs(rg.fInit)
    assert abs(P.rgrad_norm(Xg) - rg.gradNormOpt) <= 1e-6 * tol + 1e-12


def test_smallgrid_known_optimum_and_certificate():
    import dcora_b200 as dc
    P, g, info = central_problem("smallGrid3D", 5)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(1))
    go = dc.QuadraticOptimizer(gp, _params(RTR_iterations=300, RTR_tCG_iterations=500,
                                           gradnorm_tol=1e-6))
    X = go.optimize(X0)
    assert abs(2 * go.getOptResult().fOpt - 1025.3980521) < 1e-5
    res, v = gp.certify(X, 1e-3, tol=1e-10)
    S = certify.certificate_matrix(P.L, X, P.Q)
    lam_o = float(np.linalg.eigvalsh(S.toarray())[0])
    assert res.is_psd == 1 and certify.fast_verification(S, 1e-3)[0]
    assert abs(res.lambda_min - lam_o) <= 1e-6 * max(abs(lam_o), 1e-3)


def test_default_parameters_one_call_matches_oracle_trajectory():
    # default ROptParameters (3 outer x 50 tCG, tol 1e-2): same iteration
    # counts and cost as the oracle's restatement on a well-conditioned case
    import dcora_b200 as dc
    P, g, info = central_problem("smallGrid3D", 5)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(2))
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters())
    Xo = oo.optimize(X0)
    go = dc.QuadraticOptimizer(gp, _params())
    Xg = go.optimize(X0)
    rg = go.getOptResult()
    assert rg.outer_iterations == oo.result.outer_iters
    assert rg.inner_iterations == oo.result.inner_iters
    assert rel_err(Xg, Xo) < 1e-7
    assert abs(rg.fOpt - oo.result.fOpt) <= 1e-9 * abs(oo.result.fOpt)


def test_saddle_is_detected_and_escaped():
    # rank-3 critical point of smallGrid3D from a poor start is typically a
    # saddle; the certificate must say so on both sides and escapeSaddle must
    # decrease the cost at rank 4.
    import dcora_b200 as dc
    P3, g, info = central_problem("smallGrid3D", 3)
    gp3 = gpu_problem(P3)
    X0 = mf.random_point(P3.L, np.random.default_rng(7))
    kw = dict(RTR_iterations=300, RTR_tCG_iterations=500, gradnorm_tol=1e-6)
    go = dc.QuadraticOptimizer(gp3, _params(**kw))
    X = go.optimize(X0)
    S = certify.certificate_matrix(P3.L, X, P3.Q)
    lam_o = float(np.linalg.eigvalsh(S.toarray())[0])
    res, v = gp3.certify(X, 1e-3, tol=1e-10)
    assert (res.is_psd == 1) == (lam_o + 1e-3 > 0)
    assert abs(res.lambda_min - lam_o) <= 1e-6 * max(abs(lam_o), 1e-3)
    if res.is_psd:
        pytest.skip("start converged to the global optimum")
    assert abs(res.theta - lam_o) <= 1e-6 * abs(lam_o)
    assert abs(abs(v @ (S @ v)) - abs(lam_o)) <= 1e-5 * abs(lam_o)
    gp4 = gp3.with_rank(4)
    X4, ok = gp4.escapeSaddle(X, res.theta, v, 1e-6, 1e-6)
    assert ok
    f3 = go.getOptResult().fOpt
    assert gp4.f(X4) < f3
    # oracle agrees on the escape point given the same direction
    P4 = solver.QuadraticProblem(P3.L.with_rank(4), P3.Q, precond_reg=P3.reg)
    X4o, oko = solver.escape_saddle(P4, X, res.theta, v, 1e-6, 1e-6)
    assert oko and rel_err(X4, X4o) < 1e-9


def test_batched_agents_match_individual_solves():
    # three independent agent problems in one handle == three separate solves
    import dcora_b200 as dc
    names = ["tinyGrid3D", "smallGrid3D", "tinyGrid3D"]
    Ps = [central_problem(n, 5)[0] for n in names]
    rng = np.random.default_rng(4)
    X0s = [mf.random_point(P.L, rng) for P in Ps]
    ags = [dc.AgentData(n=P.L.n, Q=P.Q, precond_reg=P.reg) for P in Ps]
    gb = dc.QuadraticProblem(ags, 5, 3)
    kw = dict(RTR_iterations=50, RTR_tCG_iterations=100, gradnorm_tol=1e-5)
    ob = dc.QuadraticOptimizer(gb, _params(**kw))
    Xb = ob.optimize(np.concatenate(X0s, axis=1))
    off = 0
    for P, X0, rb in zip(Ps, X0s, ob.results):
        g1 = gpu_problem(P)
        o1 = dc.QuadraticOptimizer(g1, _params(**kw))
        X1 = o1.optimize(X0)
        r1 = o1.getOptResult()
        assert rb.outer_iterations == r1.outer_iterations
        assert rb.inner_iterations == r1.inner_iterations
        assert np.array_equal(Xb[:, off:off + P.L.k], X1)    # bit-identical
        off += P.L.k
    # masked solve leaves inactive agents untouched
    Xm = ob.optimize(np.concatenate(X0s, axis=1), active=[0, 1, 0])
    assert np.array_equal(Xm[:, :Ps[0].L.k], X0s[0])
    assert not np.array_equal(Xm[:, Ps[0].L.k:Ps[0].L.k + Ps[1].L.k], X0s[1])


def test_staircase_driver_escapes_the_saddle_and_certifies():
    # rank-3 start that converges to a saddle (see test_saddle_is_detected_and_escaped):
    # the staircase driver must climb and end at the certified optimum
    import dcora_b200 as dc
    from dcora_b200.staircase import riemannian_staircase
    P3, g, info = central_problem("smallGrid3D", 3)
    gp3 = gpu_problem(P3)
    X0 = mf.random_point(P3.L, np.random.default_rng(7))
    res = riemannian_staircase(gp3, X0, r_max=8, eta=1e-3)
    assert res.certified and res.lambda_min > -1e-3
    assert abs(2 * res.cost - 1025.3980521) < 1e-5
    assert res.levels[0]["r"] == 3
    if len(res.levels) > 1:
        assert not res.levels[0]["is_psd"] and res.levels[-1]["f"] < res.levels[0]["f"]


@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("range_aided_slam_test_3d", 5)])
@pytest.mark.parametrize("precond", [True, False])
def test_rgd_step_matches_oracle(name, r, precond):
    # ROptMethod::RGD: one Riemannian gradient step with constant step size
    # (QuadraticOptimizer.cpp:124-201), with and without the preconditioner
    import dcora_b200 as dc
    P, g, info = central_problem(name, r)
    gp = gpu_problem(P)
    P.reg = gp.precond_reg(0)
    P._lu = None
    X0 = mf.random_point(P.L, np.random.default_rng(3))
    kw = dict(method="RGD", RGD_stepsize=1e-4, RGD_use_preconditioner=precond)
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters(**kw))
    Xo = oo.optimize(X0)
    go = dc.QuadraticOptimizer(gp, _params(**kw))
    Xg = go.optimize(X0)
    rg = go.getOptResult()
    assert rel_err(Xg, Xo) < 1e-10
    assert abs(rg.fInit - oo.result.fInit) <= 1e-12 * abs(oo.result.fInit)
    assert abs(rg.fOpt - oo.result.fOpt) <= 1e-10 * abs(oo.result.fOpt)
    assert abs(rg.gradNormOpt - oo.result.gradNormOpt) <= 1e-8 * oo.result.gradNormOpt


def test_single_iteration_mode_shrinks_the_radius_until_accepted():
    # RTR_iterations == 1: QuadraticOptimizer.cpp:254-273 re-runs the solver with the
    # radius divided by 4 (at most 10 times) until the step is accepted
    import dcora_b200 as dc
    P, g, info = central_problem("smallGrid3D", 5)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(4))
    kw = dict(RTR_iterations=1, RTR_tCG_iterations=50, RTR_initial_radius=1e4)
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters(**kw))
    Xo = oo.optimize(X0)
    go = dc.QuadraticOptimizer(gp, _params(**kw))
    Xg = go.optimize(X0)
    rg = go.getOptResult()
    assert rg.fOpt < rg.fInit
    assert abs(rg.fOpt - oo.result.fOpt) <= 1e-8 * abs(oo.result.fOpt)
    assert rel_err(Xg, Xo) < 1e-6


def test_polar_retraction_in_the_solver():
    # the retraction variant is selectable (ROPTLIB's Stiefel default is qf; the
    # reference's metric projections are polar): both must agree with the oracle
    import dcora_b200 as dc
    P, g, info = central_problem("smallGrid3D", 5)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(5))
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters(retraction="polar"))
    Xo = oo.optimize(X0)
    go = dc.QuadraticOptimizer(gp, _params(retraction=dc.RETRACT_POLAR))
    Xg = go.optimize(X0)
    assert rel_err(Xg, Xo) < 1e-7
    assert abs(go.getOptResult().fOpt - oo.result.fOpt) <= 1e-9 * abs(oo.result.fOpt)


@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("range_aided_slam_test_3d", 5), ("sphere2500", 4)])
def test_stable_decrease_option_reaches_the_same_optimum(name, r):
    # ROptParameters.stable_decrease: the trust-region ratio from <egrad, d> + 0.5 <dQ, d>
    # (d = x+ - x) instead of f(x) - f(x+) -- the same number up to rounding, so the solve
    # follows the same trajectory on well-scaled problems and ends at the same cost
    import dcora_b200 as dc
    P, g, info = central_problem(name, r)
    gp = gpu_problem(P)
    X0 = mf.random_point(P.L, np.random.default_rng(4))
    out = []
    for stable in (False, True):
        go = dc.QuadraticOptimizer(gp, _params(RTR_iterations=60, RTR_tCG_iterations=200,
                                               gradnorm_tol=1e-6, stable_decrease=stable))
        go.optimize(X0)
        out.append(go.getOptResult())
    a, b = out
    assert abs(a.fOpt - b.fOpt) <= 1e-9 * abs(a.fOpt) + 1e-12
    assert b.gradNormOpt < 1e-6 or b.outer_iterations == 60
    assert abs(a.outer_iterations - b.outer_iterations) <= 2
