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
