"""The native twin of the CPU oracle (oracle/native, C++ -O3: the CPU baseline of bench.py)
against the numpy/scipy oracle: same operations, same RTR/tCG iteration counts, same optimum."""
import numpy as np
import pytest

from oracle import manifold as mf, native, solver
from tests.common import central_problem, rel_err


@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("range_aided_slam_test_3d", 4),
                                    ("range_aided_slam_test_2d", 3), ("tinyGrid3D", 4)])
def test_operations_match_the_scipy_oracle(name, r):
    P, g, info = central_problem(name, r)
    rng = np.random.default_rng(3)
    X = mf.random_point(P.L, rng)
    V = rng.standard_normal((r, P.L.k))
    assert rel_err(native.apply("egrad", P.L, P.Q, X=X), P.egrad(X)) < 1e-13
    assert rel_err(native.apply("tangent", P.L, P.Q, X=X, V=V), mf.tangent_project(P.L, X, V)) < 1e-13
    T = mf.tangent_project(P.L, X, V)
    assert rel_err(native.apply("retract_qf", P.L, P.Q, X=X, V=0.1 * T), mf.retract(P.L, X, 0.1 * T, "qf")) < 1e-12
    assert rel_err(native.apply("retract_polar", P.L, P.Q, X=X, V=0.1 * T), mf.retract(P.L, X, 0.1 * T, "polar")) < 1e-11
    assert rel_err(native.apply("project", P.L, P.Q, X=V), mf.project_manifold(P.L, V)) < 1e-11
    ref = P._factor().solve(np.ascontiguousarray(V.T)).T
    assert rel_err(native.apply("solve", P.L, P.Q, V=V, reg=P.reg), ref) < 1e-8
    assert rel_err(native.apply("rhess", P.L, P.Q, X=X, V=T), P.rhess(X, P.egrad(X), T)) < 1e-12


@pytest.mark.parametrize("name,r", [("smallGrid3D", 5), ("range_aided_slam_test_3d", 4)])
def test_default_update_follows_the_scipy_oracle(name, r):
    # Agent::updateX with the default ROptParameters (3 RTR x 50 tCG, tol 1e-2)
    P, g, info = central_problem(name, r)
    X0 = mf.random_point(P.L, np.random.default_rng(2))
    oo = solver.QuadraticOptimizer(P, solver.ROptParameters())
    Xo = oo.optimize(X0)
    Xn, st = native.agent_update(P.L, P.Q, None, P.reg, X0)
    assert abs(st["fInit"] - oo.result.fInit) <= 1e-12 * abs(oo.result.fInit)
    assert int(st["outer"]) == oo.result.outer_iters
    assert int(st["inner"]) == oo.result.inner_iters
    assert rel_err(Xn, Xo) < 1e-7
    assert abs(st["fOpt"] - oo.result.fOpt) <= 1e-9 * abs(oo.result.fOpt) + 1e-13 * abs(st["fInit"])


def test_converged_optimum_of_smallgrid():
    P, g, info = central_problem("smallGrid3D", 5)
    X0 = mf.random_point(P.L, np.random.default_rng(1))
    Xn, st = native.agent_update(P.L, P.Q, None, P.reg, X0, max_outer=300, max_inner=500, tol=1e-6)
    assert st["gradNormOpt"] < 1e-6
    assert abs(2 * st["fOpt"] - 1025.3980521) < 1e-5     # SE-Sync optimum
    assert abs(P.f(Xn) - st["fOpt"]) <= 1e-10 * abs(st["fOpt"])
