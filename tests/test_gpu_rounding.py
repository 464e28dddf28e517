"""GPU parity of the rounding entry points (projectSolutionRASLAM,
alignLiftedTrajectoryToFrame; src/DCORA_utils.cpp:1984-2031, :2262-2289) against the oracle."""
import numpy as np
import pytest

from oracle import certify, manifold as mf
from tests.common import central_problem, gpu_problem, rel_err

pytestmark = pytest.mark.gpu

CASES = [("tinyGrid3D", 3), ("smallGrid3D", 5), ("city10000", 4), ("range_aided_slam_test_2d", 4),
         ("range_aided_slam_test_3d", 6), ("tiers", 5), ("smallGrid3D", 16)]


def _proper(Ld, Xd):
    """St(d, d) = O(d): make every rotation block of a rank-d point a proper rotation."""
    B = np.transpose(Xd[:, Ld.rot], (1, 0, 2))
    neg = np.linalg.det(B) < 0
    Xd[:, Ld.rot[neg, -1]] *= -1
    return Xd


def _noisy_lift(L, info, rng, noise):
    """A rank-r point close to a lifted rank-d solution (so that the truncation is well
    separated, as at the end of a staircase), in a random r-frame."""
    Ld = L.with_rank(L.d)
    Xd = _proper(Ld, mf.random_point(Ld, rng))
    Y, _ = np.linalg.qr(rng.standard_normal((L.r, L.d)))
    X = Y @ Xd + noise * rng.standard_normal((L.r, L.k))
    return mf.project_manifold(L, X)


@pytest.mark.parametrize("name,r", CASES, ids=[f"{n}-r{r}" for n, r in CASES])
def test_round_solution_matches_oracle(name, r):
    P, g, info = central_problem(name, r)
    L = P.L
    gp = gpu_problem(P, precond=False)
    rng = np.random.default_rng(5)
    X = _noisy_lift(L, info, rng, 0.02)
    got = gp.round_solution(X)
    want = certify.project_solution_ra(L, X)
    assert got.shape == want.shape == (L.d, L.k)
    # the rank-d truncation is unique up to the sign of each row; after the reflection vote
    # the two results differ by a diag(+-1) of determinant +1
    s = np.sign(np.sum(got * want, axis=1))
    assert np.prod(s) > 0
    assert rel_err(s[:, None] * got, want) < 1e-9
    # feasibility at rank d: rotations in SO(d), unit spheres of norm one
    Ld = L.with_rank(L.d)
    B = np.transpose(got[:, Ld.rot], (1, 0, 2))
    assert np.allclose(np.einsum("nij,nkj->nik", B, B), np.eye(L.d), atol=1e-12)
    assert np.allclose(np.linalg.det(B), 1.0, atol=1e-12)
    if L.l:
        assert np.allclose(np.linalg.norm(got[:, Ld.sph], axis=0), 1.0, atol=1e-13)
    # the cost of the rounded solution is a gauge-free scalar
    Pd = type(P)(Ld, P.Q, precond_reg=None, use_precond=False)
    assert abs(Pd.f(got) - Pd.f(want)) <= 1e-9 * max(1.0, abs(Pd.f(want)))


def test_round_solution_reflects_improper_solutions():
    P, g, info = central_problem("smallGrid3D", 5)
    L = P.L
    gp = gpu_problem(P, precond=False)
    rng = np.random.default_rng(9)
    Ld = L.with_rank(3)
    Xd = _proper(Ld, mf.random_point(Ld, rng))
    Y = np.zeros((5, 3)); Y[0, 0] = Y[1, 1] = 1.0; Y[2, 2] = -1.0     # improper embedding
    X = Y @ Xd
    got = gp.round_solution(X)
    B = np.transpose(got[:, Ld.rot], (1, 0, 2))
    assert np.allclose(np.linalg.det(B), 1.0, atol=1e-12)
    want = certify.project_solution_ra(L, X)
    s = np.sign(np.sum(got * want, axis=1))
    assert np.prod(s) > 0 and rel_err(s[:, None] * got, want) < 1e-10


@pytest.mark.parametrize("name,r,is_global", [("smallGrid3D", 5, True), ("city10000", 3, True),
                                              ("smallGrid3D", 3, False), ("city10000", 2, False)])
def test_align_trajectory_matches_oracle(name, r, is_global):
    P, g, info = central_problem(name, r)
    L = P.L
    gp = gpu_problem(P, precond=False)
    rng = np.random.default_rng(3)
    X = mf.random_point(L, rng)
    R0 = np.linalg.qr(rng.standard_normal((r, L.d)))[0]
    if r == L.d:
        # St(d, d) = O(d): keep R0^T R_i proper (the SO(d) projection of an exactly
        # orthogonal improper block is not unique)
        X = _proper(L, X)
        if np.linalg.det(R0) < 0:
            R0[:, -1] *= -1
    Tw0 = np.concatenate([R0, rng.standard_normal((r, 1))], axis=1)
    got = gp.align_trajectory(X, Tw0, is_global)
    want = certify.align_lifted_trajectory(X, Tw0, L.d, is_global)
    assert rel_err(got, want) < 1e-11
