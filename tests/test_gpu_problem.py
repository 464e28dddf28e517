"""GPU parity tests (through the C-ABI) of the QuadraticProblem-level entry
points against the oracle: f, EucGrad, EucHessianEta, RieGrad, PreConditioner,
Retract, projections -- SE and RA orderings, 2-D and 3-D, several ranks."""
import numpy as np
import pytest

from oracle import manifold as mf
from tests.common import central_problem, gpu_problem, rel_err

pytestmark = pytest.mark.gpu

CASES = [("tinyGrid3D", 3), ("smallGrid3D", 5), ("sphere2500", 4), ("torus3D", 6),
         ("city10000", 3), ("kitti_00", 5), ("range_aided_slam_test_2d", 2),
         ("range_aided_slam_test_3d", 5), ("tiers", 3), ("tiers", 7),
         # high staircase ranks: every kernel is templated on r = 1..16
         ("smallGrid3D", 9), ("smallGrid3D", 16), ("range_aided_slam_test_2d", 8),
         ("range_aided_slam_test_3d", 12)]


@pytest.fixture(scope="module", params=CASES, ids=[f"{n}-r{r}" for n, r in CASES])
def pair(request):
    name, r = request.param
    P, g, info = central_problem(name, r)
    rng = np.random.default_rng(11)
    P.G = 0.1 * rng.standard_normal((r, P.L.k))       # exercise the linear term
    gp = gpu_problem(P)
    X = mf.random_point(P.L, rng)
    V = rng.standard_normal((r, P.L.k))
    return P, gp, X, V


def test_f(pair):
    P, gp, X, V = pair
    f_o = P.f(X)
    assert abs(gp.f(X) - f_o) <= 1e-12 * max(1.0, abs(f_o))


def test_egrad_hessvec(pair):
    P, gp, X, V = pair
    assert rel_err(gp.EucGrad(X), P.egrad(X)) < 1e-13
    assert rel_err(gp.EucHessianEta(V), P.ehess(V)) < 1e-13


def test_rgrad(pair):
    P, gp, X, V = pair
    assert rel_err(gp.RieGrad(X), P.rgrad(X)) < 1e-12
    n_o = P.rgrad_norm(X)
    assert abs(gp.RieGradNorm(X) - n_o) <= 1e-12 * n_o


def test_project_tangent(pair):
    P, gp, X, V = pair
    assert rel_err(gp.projectToTangentSpace(X, V), mf.tangent_project(P.L, X, V)) < 1e-13


def test_retract_qf_and_polar(pair):
    P, gp, X, V = pair
    W = 0.2 * mf.tangent_project(P.L, X, V)
    assert rel_err(gp.Retract(X, W, 0), mf.retract(P.L, X, W, "qf")) < 1e-12
    assert rel_err(gp.Retract(X, W, 1), mf.retract(P.L, X, W, "polar")) < 1e-12


def test_project_manifold(pair):
    P, gp, X, V = pair
    M = X + 0.3 * V
    assert rel_err(gp.project(M), mf.project_manifold(P.L, M)) < 1e-11


def test_precondition(pair):
    P, gp, X, V = pair
    ref = P.precondition(X, V)
    assert rel_err(gp.PreCondition(X, V), ref) < 1e-8


def test_factor_solve_without_projection(pair):
    # dcora_problem_precond with X = NULL: the Cholesky solve (Q + reg I)^-1 V alone
    # (graph_->preconditioner()->solve, QuadraticProblem.cpp:79)
    P, gp, X, V = pair
    ref = P._factor().solve(np.ascontiguousarray(V.T)).T
    assert rel_err(gp.factor_solve(V), ref) < 1e-8


def test_device_resident_arguments(pair):
    torch = pytest.importorskip("torch")
    P, gp, X, V = pair
    Xd = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    out = gp.EucGrad(Xd)
    assert out.is_cuda
    assert rel_err(out.cpu().numpy().T, P.egrad(X)) < 1e-13


@pytest.fixture(scope="module")
def big_agent():
    from dcora_b200 import graph as pg, synthetic
    from oracle import solver
    meas, gt, idx = synthetic.grid3d(15, 0.1, seed=8)
    gpm = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, 1))
    Q = gpm.Q()
    L = mf.Layout(4, 3, int(gpm.n_a[0]), int(gpm.l_a[0]), 0, ra=True)
    assert L.k >= 12000
    P = solver.QuadraticProblem(L, Q, precond_reg=0.5)
    rng = np.random.default_rng(3)
    X = mf.random_point(L, rng)
    V = rng.standard_normal((4, L.k))
    ref_solve = P._factor().solve(np.ascontiguousarray(V.T)).T
    return L, Q, X, V, ref_solve, P.precondition(X, V)


@pytest.mark.parametrize("warps", [64, 96, 128, 256])
def test_precondition_cluster_shapes(big_agent, warps):
    # K5 on a LARGE agent (>= 12000 columns) with every cluster shape the drivers choose from:
    # 64 warps (8 CTAs x 8), 96 (12 x 8), 128 (16 x 8), 256 (16 x 16) per agent
    # (dcora_set_trsv_max_warps); all must solve (Q + reg I) z = v like the oracle's LU
    import dcora_b200 as dc
    from dcora_b200 import _lib
    L, Q, X, V, ref_solve, ref_precond = big_agent
    lib = _lib.load()
    lib.dcora_set_trsv_max_warps(warps)
    try:
        ag = dc.AgentData(n=L.n, l=L.l, b=0, Q=Q, precond_reg=0.5)
        gp = dc.QuadraticProblem(ag, 4, 3, dc.ORDER_RA)
    finally:
        lib.dcora_set_trsv_max_warps(64)
    assert rel_err(gp.factor_solve(V), ref_solve) < 1e-9
    assert rel_err(gp.PreCondition(X, V), ref_precond) < 1e-9
