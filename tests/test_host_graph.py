"""CPU tests of the product's host-side data plumbing (dcora_b200/graph.py,
rbcd.py planning) against the oracle's per-agent Graph restatement."""
import numpy as np
import pytest
import scipy.sparse as sp

from dcora_b200 import graph as pg, rbcd, synthetic
from oracle import graph as og, manifold as mf
from tests.common import load_golden, product_measurements, split_contiguous


def _csr(t, k):
    return sp.csr_matrix((t[2], t[1], t[0]), shape=(k, k))


def _multi(name, robots):
    if name.startswith("range_aided"):
        meas, info = load_golden(name, prefix="local_")
        return meas, True
    meas, info = load_golden(name)
    m, counts, start = split_contiguous(meas, info["n"], robots)
    return m, False


CASES = [("smallGrid3D", 5), ("tinyGrid3D", 3), ("range_aided_slam_test_2d", 2),
         ("range_aided_slam_test_3d", 2), ("kitti_00", 8)]


@pytest.mark.parametrize("name,robots", CASES)
def test_agent_Q_matches_oracle_graph(name, robots):
    meas, ra = _multi(name, robots)
    M, part = product_measurements(meas, robots)
    gp = pg.GlobalProblem(M, part, ra=ra)
    diag, cross = gp.split()
    for a in range(robots):
        g = og.Graph(a, 5, meas.d, ra=ra).set_measurements(meas)
        assert (g.n, g.l, g.b) == (gp.n_a[a], gp.l_a[a], gp.b_a[a])
        Qo = g.quadratic_matrix()
        Qp = _csr(diag[a], g.k)
        assert abs(Qo - Qp).max() <= 1e-12 * max(1.0, abs(Qo).max())


@pytest.mark.parametrize("name,robots", CASES)
def test_coupling_product_equals_oracle_linear_term(name, robots):
    meas, ra = _multi(name, robots)
    M, part = product_measurements(meas, robots)
    gp = pg.GlobalProblem(M, part, ra=ra)
    _, cross = gp.split(agents=[])
    r = 5
    rng = np.random.default_rng(3)
    X = rng.standard_normal((r, gp.K))
    G = (cross.T @ X.T).T          # G = X * C
    graphs = [og.Graph(a, r, meas.d, ra=ra).set_measurements(meas) for a in range(robots)]
    # neighbour dictionaries from X
    nbr = {}
    for a, g in enumerate(graphs):
        Xa = X[:, gp.coff[a]:gp.coff[a + 1]]
        for i in range(g.n):
            nbr[("pose", a, i)] = np.column_stack([Xa[:, g.rot_cols(i)], Xa[:, g.trans_col(i)]])
        for i in range(g.l):
            nbr[("sph", a, i)] = Xa[:, g.sphere_col(i)]
        for i in range(g.b):
            nbr[("lmk", a, i)] = Xa[:, g.landmark_col(i)]
    for a, g in enumerate(graphs):
        Go = g.linear_matrix(nbr)
        assert Go is not None
        Gp = G[:, gp.coff[a]:gp.coff[a + 1]]
        assert np.abs(Go - Gp).max() <= 1e-10 * max(1.0, np.abs(Go).max())


@pytest.mark.parametrize("world", [1, 2, 3])
def test_rank_plans_reproduce_G(world):
    meas, ra = _multi("range_aided_slam_test_3d", 2) if world == 2 else _multi("smallGrid3D", 5)
    robots = 2 if world == 2 else 5
    M, part = product_measurements(meas, robots)
    gp = pg.GlobalProblem(M, part, ra=ra)
    _, cross = gp.split(agents=[])
    plans = rbcd.make_plans(gp, world, cross)
    r = 4
    X = np.random.default_rng(0).standard_normal((r, gp.K))
    G = (cross.T @ X.T).T
    packs = [X[:, p.c0:p.c0 + p.K][:, p.pack_idx] for p in plans]
    H = np.concatenate(packs, axis=1) if world > 1 else np.zeros((r, 0))
    for p in plans:
        XE = np.concatenate([X[:, p.c0:p.c0 + p.K], H], axis=1)
        Gp = (p.coupling.T @ XE.T).T
        assert np.abs(Gp - G[:, p.c0:p.c0 + p.K]).max() < 1e-12 * max(1.0, np.abs(G).max())
    # colours: adjacent agents never share a colour
    adj = gp.agent_graph(cross)
    col = plans[0].colors
    for a in range(gp.A):
        assert all(col[a] != col[b] for b in adj[a])


def test_synthetic_grid_is_consistent():
    S = 6
    meas, (Rgt, pos), idx = synthetic.grid3d(S, range_fraction=0.1, seed=1, noiseless=True)
    n = S ** 3
    assert meas.n == n and len(meas.pp_i) == 3 * S * S * (S - 1)
    # consecutive poses on the path are grid neighbours and all carry an edge
    key = set((meas.pp_i * n + meas.pp_j).tolist())
    assert all((i * n + i + 1) in key for i in range(n - 1))
    assert len(np.unique(meas.rg_i * n + meas.rg_j)) == len(meas.rg_i)
    # noiseless data: the ground truth has zero cost
    part = pg.Partition.contiguous(n, 4)
    gp = pg.GlobalProblem(meas, part)
    Q = gp.Q()
    X = np.zeros((3, gp.K))
    X[:, gp.rot_cols(np.arange(n)).reshape(-1)] = np.transpose(Rgt, (1, 0, 2)).reshape(3, -1)
    X[:, gp.trans_col(np.arange(n))] = pos.T
    e = np.arange(len(meas.rg_i))
    dirs = (pos[meas.rg_i] - pos[meas.rg_j])
    X[:, gp.sphere_col(e)] = (dirs / np.linalg.norm(dirs, axis=1)[:, None]).T
    assert abs(np.sum((Q @ X.T).T * X)) < 1e-9
    # the spanning-tree guess reproduces the ground truth up to the gauge
    Ri, ti = synthetic.spanning_tree_init(meas, S, idx)
    R0, p0 = Rgt[idx[0, 0, 0]], pos[idx[0, 0, 0]]
    assert np.allclose(R0 @ Ri, Rgt, atol=1e-9)
    assert np.allclose((R0 @ ti.T).T + p0, pos, atol=1e-9)


def test_balanced_colouring_is_valid_and_even():
    # RBCD sweeps colour by colour; every class runs its agents concurrently, so classes must be
    # independent sets and of even size (one K5 wave per colour on the device)
    meas, gt, idx = synthetic.grid3d(12, range_fraction=0.1, seed=2)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, 48))
    _, cross = gp.split(agents=[])
    adj = gp.agent_graph(cross)
    col = pg.greedy_coloring(adj)
    assert all(col[a] != col[b] for a in range(gp.A) for b in adj[a])
    sizes = np.bincount(col)
    raw = -np.ones(gp.A, np.int64)            # the unbalanced greedy colouring, for reference
    for a in sorted(range(gp.A), key=lambda x: -len(adj[x])):
        used = {raw[b] for b in adj[a] if raw[b] >= 0}
        c = 0
        while c in used:
            c += 1
        raw[a] = c
    assert len(sizes) == raw.max() + 1                   # no colour added
    assert sizes.max() <= np.bincount(raw).max()         # never worse than greedy
    assert sizes.max() - sizes.min() <= max(2, np.bincount(raw).max() - np.bincount(raw).min())


@pytest.mark.parametrize("name", ["tinyGrid3D", "smallGrid3D", "pose_graph_optimization_test_2d",
                                  "pose_graph_optimization_test_3d"])
def test_initialisation_matches_oracle(name):
    # chordalInitialization / odometryInitialization (src/DCORA_solver.cpp:218-302): sparse
    # normal equations (product) vs dense edge-by-edge least squares (oracle)
    from dcora_b200 import init
    from oracle import init as oinit
    from tests.common import load_golden, product_measurements
    meas, info = load_golden(name)
    M, part = product_measurements(meas, 1)
    R, t = init.chordal_initialization(M)
    Ro, to = oinit.chordal_initialization(meas.pp, info["n"])
    assert np.abs(R - Ro).max() < 1e-9 and np.abs(t - to).max() < 1e-8 * max(1.0, np.abs(to).max())
    assert np.allclose(np.linalg.det(R), 1.0) and np.allclose(R[0], np.eye(M.d))
    if np.all(np.isin(np.arange(info["n"] - 1), M.pp_i[M.pp_j == M.pp_i + 1])):
        R2, t2 = init.odometry_initialization(M)
        Ro2, to2 = oinit.odometry_initialization(meas.pp, info["n"])
        assert np.abs(R2 - Ro2).max() < 1e-12 and np.abs(t2 - to2).max() < 1e-10 * max(1.0, np.abs(to2).max())
    if name.startswith("pose_graph_optimization_test"):
        # noiseless fixture: the chordal solution IS the ground truth in the frame of pose 0
        gp = pg.GlobalProblem(M, part)
        X = init.lift(gp, R, t, M.d + 1)
        assert 0.5 * np.sum((X @ gp.Q()) * X) < 1e-9


def test_chordal_initialisation_beats_the_spanning_tree_guess():
    from dcora_b200 import init
    meas, gt, idx = synthetic.grid3d(6, 0.1, seed=1)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, 4))
    Q = gp.Q()
    cost = lambda X: 0.5 * float(np.sum((X @ Q) * X))
    Xc = init.lift(gp, *init.chordal_initialization(meas), 5)
    Xs = synthetic.lifted_initial_guess(meas, idx, gp, 6, 5)
    Xo = init.lift(gp, *init.odometry_initialization(meas), 5)
    assert cost(Xc) < 0.2 * cost(Xs) < cost(Xo)
    # feasible: rotations orthonormal, unit spheres of norm one
    i = np.arange(meas.n)
    B = Xc[:, gp.rot_cols(i)].transpose(1, 0, 2)
    assert np.allclose(np.einsum("nri,nrj->nij", B, B), np.eye(3), atol=1e-12)
    assert np.allclose(np.linalg.norm(Xc[:, gp.sphere_col(np.arange(len(meas.rg_i)))], axis=0), 1.0)


@pytest.mark.parametrize("robots", [1, 4])
def test_synthetic_grid_agent_matrices_match_the_oracle_graph(robots):
    # configs[4]-shaped data (noisy 3-D grid + range edges, contiguous agents): the product's
    # vectorised assembly (graph.py) against the oracle's per-measurement Graph
    # (constructQuadraticCostTermRASLAM, src/Graph.cpp:824-1188) agent by agent
    import scipy.sparse as sp
    from oracle import graph as og
    from tests.common import oracle_measurements
    meas, gt, idx = synthetic.grid3d(6, 0.15, seed=3)
    part = pg.Partition.contiguous(meas.n, robots)
    gp = pg.GlobalProblem(meas, part)
    om = oracle_measurements(gp.meas, part)
    diag, cross = gp.split()
    for a in range(robots):
        g = og.Graph(a, 4, 3, ra=True).set_measurements(om)
        Qo = g.quadratic_matrix()
        rp, col, val = diag[a]
        Qp = sp.csr_matrix((val, col, rp), shape=Qo.shape)
        assert (g.n, g.l, g.b) == (int(gp.n_a[a]), int(gp.l_a[a]), int(gp.b_a[a]))
        assert abs(Qo - Qp).max() <= 1e-12 * abs(Qo).max()
    # and the blocks + the cross-agent part make up the global matrix
    B = sp.block_diag([sp.csr_matrix((v, c, r), shape=(len(r) - 1, len(r) - 1))
                       for (r, c, v) in (diag[a] for a in range(robots))]).tocsr()
    assert abs(B + cross - gp.Q()).max() == 0.0
