"""CPU tests: the oracle against the reference's golden vectors / known answers
(SURVEY 8c).  These pin the restatement before it is trusted as the checker."""
import numpy as np
import pytest

from oracle import certify, graph, manifold as mf, solver
from tests.common import central_problem, load_golden, split_contiguous

NOISELESS = ["pose_graph_optimization_test_2d", "pose_graph_optimization_test_3d",
             "range_aided_slam_test_2d", "range_aided_slam_test_3d"]


@pytest.mark.parametrize("name", NOISELESS)
def test_ground_truth_is_optimum_of_noiseless_fixtures(name):
    # tests/testAgent.cpp:101-242 -- ground truth is the fixed point: f*=0, grad=0
    P, g, info = central_problem(name, info_rank(name))
    X = info["X_gt"]
    assert abs(P.f(X)) < 1e-12
    assert P.rgrad_norm(X) < 1e-7          # files carry 9 significant digits
    assert np.abs(P.XQ(X)).max() < 1e-8
    S = certify.certificate_matrix(P.L, X, P.Q)
    assert np.linalg.eigvalsh(S.toarray())[0] > -1e-8   # S >= 0 at the optimum


def info_rank(name):
    return 2 if "2d" in name else 3


@pytest.mark.parametrize("name", NOISELESS)
def test_iterate_keeps_ground_truth_fixed(name):
    # testAgentInitializeIterateOptimize{SE,RA}: optimize() from the ground
    # truth returns the ground truth (to 1e-9 in the reference)
    P, g, info = central_problem(name, info_rank(name))
    X = info["X_gt"]
    Xo = solver.QuadraticOptimizer(P, solver.ROptParameters()).optimize(X)
    assert np.allclose(Xo, X, atol=1e-9)


def test_projection_closed_form_se():
    # tests/testManifold.cpp:354-390
    rng = np.random.default_rng(0)
    r, d, n = 5, 3, 10
    L = mf.Layout(r, d, n)
    X = mf.random_point(L, rng)
    V = rng.standard_normal((r, L.k))
    W = mf.tangent_project(L, X, V)
    for i in range(n):
        Y = X[:, i * 4:i * 4 + 3]
        Vi = V[:, i * 4:i * 4 + 3]
        M = Y.T @ Vi
        assert np.allclose(W[:, i * 4:i * 4 + 3], Vi - Y @ (0.5 * (M + M.T)))
        assert np.allclose(W[:, i * 4 + 3], V[:, i * 4 + 3])
        # tangent: Y^T W skew
        A = Y.T @ W[:, i * 4:i * 4 + 3]
        assert np.allclose(A + A.T, 0, atol=1e-12)


def test_projection_closed_form_ra_and_feasibility():
    # tests/testManifold.cpp:392-440, tests/testUtils.cpp:38-48,82-92
    rng = np.random.default_rng(1)
    r, d, n, l, b = 4, 2, 7, 5, 3
    L = mf.Layout(r, d, n, l, b, ra=True)
    X = mf.random_point(L, rng)
    for i in range(n):
        Y = X[:, L.rot[i]]
        assert np.allclose(Y.T @ Y, np.eye(d), atol=1e-12)
    assert np.allclose(np.linalg.norm(X[:, L.sph], axis=0), 1.0)
    assert np.allclose(mf.project_manifold(L, X), X, atol=1e-12)   # fixed point
    V = rng.standard_normal((r, L.k))
    W = mf.tangent_project(L, X, V)
    assert np.allclose(np.sum(W[:, L.sph] * X[:, L.sph], axis=0), 0, atol=1e-12)
    tr = d * n + l
    assert np.allclose(W[:, tr:], V[:, tr:])
    for kind in ("qf", "polar"):
        Xn = mf.retract(L, X, 0.3 * W, kind)
        for i in range(n):
            Y = Xn[:, L.rot[i]]
            assert np.allclose(Y.T @ Y, np.eye(d), atol=1e-12)
        assert np.allclose(np.linalg.norm(Xn[:, L.sph], axis=0), 1.0)


def test_smallgrid_known_optimum():
    # literature optimum (SE-Sync): 2 f* = 1025.398 for smallGrid3D
    P, g, info = central_problem("smallGrid3D", 5)
    X0 = mf.random_point(P.L, np.random.default_rng(1))
    opt = solver.QuadraticOptimizer(P, solver.ROptParameters(
        RTR_iterations=100, RTR_tCG_iterations=500, gradnorm_tol=1e-6))
    X = opt.optimize(X0)
    assert abs(2 * opt.result.fOpt - 1025.3980521) < 1e-5
    S = certify.certificate_matrix(P.L, X, P.Q)
    ok, theta, v = certify.fast_verification(S, 1e-3)
    assert ok


def test_agent_matrices_are_slices_of_the_global_matrix():
    # Q_agent = Q_global[own, own], G = X_c * Q_global[c, own]  (Graph.cpp:579-822)
    meas, info = load_golden("smallGrid3D")
    d, n, r, R = info["d"], info["n"], 5, 5
    part, counts, start = split_contiguous(meas, n, R)
    gg = graph.Graph(0, r, d).set_measurements(meas)
    Qg = gg.quadratic_matrix().toarray()
    rng = np.random.default_rng(3)
    X = mf.random_point(mf.Layout(r, d, n), rng)
    dh = d + 1
    for a in range(R):
        g = graph.Graph(a, r, d).set_measurements(part)
        assert g.n == counts[a]
        own = np.arange(start[a] * dh, (start[a] + counts[a]) * dh)
        oth = np.setdiff1d(np.arange(n * dh), own)
        assert np.allclose(g.quadratic_matrix().toarray(), Qg[np.ix_(own, own)], atol=1e-9)
        nbr = {}
        for b in range(R):
            if b == a:
                continue
            for i in range(counts[b]):
                c0 = (start[b] + i) * dh
                nbr[("pose", b, i)] = X[:, c0:c0 + dh]
        G = g.linear_matrix(nbr)
        assert np.allclose(G, X[:, oth] @ Qg[np.ix_(oth, own)], atol=1e-8)


def test_ra_agent_matrices_are_slices_of_the_global_matrix():
    # same identity for range-aided graphs with two robots + map landmarks
    for name in ("range_aided_slam_test_2d", "range_aided_slam_test_3d"):
        gmeas, info = load_golden(name)
        lmeas, _ = load_golden(name, prefix="local_")
        d, r = info["d"], info["d"] + 2
        n, l, b = info["n"], info["l"], info["b"]
        gg = graph.Graph(0, r, d, ra=True).set_measurements(gmeas)
        Qg = gg.quadratic_matrix().toarray()
        Lg = mf.Layout(r, d, n, l, b, ra=True)
        X = mf.random_point(Lg, np.random.default_rng(5))
        # robots 0 (A) and 1 (B): 3 poses, 5 spheres, 1 landmark each
        # (std::map order of getLocalToGlobalStateMapping: robot-major)
        for a in (0, 1):
            g = graph.Graph(a, r, d, ra=True).set_measurements(lmeas)
            assert (g.n, g.l, g.b) == (3, 5, 1)
            rot = (np.arange(3)[:, None] * d + np.arange(d) + a * 3 * d).ravel()
            sph = d * n + a * 5 + np.arange(5)
            tr = d * n + l + a * 3 + np.arange(3)
            lm = d * n + l + n + a + np.arange(1)
            own = np.concatenate([rot, sph, tr, lm])
            oth = np.setdiff1d(np.arange(Lg.k), own)
            assert np.allclose(g.quadratic_matrix().toarray(), Qg[np.ix_(own, own)], atol=1e-9)
            o = 1 - a
            nbr = {}
            for i in range(3):
                P = np.zeros((r, d + 1))
                P[:, :d] = X[:, (o * 3 + i) * d:(o * 3 + i + 1) * d]
                P[:, d] = X[:, d * n + l + o * 3 + i]
                nbr[("pose", o, i)] = P
            for i in range(5):
                nbr[("sph", o, i)] = X[:, d * n + o * 5 + i]
            nbr[("lmk", o, 0)] = X[:, d * n + l + n + o]
            G = g.linear_matrix(nbr)
            assert np.allclose(G, X[:, oth] @ Qg[np.ix_(oth, own)], atol=1e-9)
