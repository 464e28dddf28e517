"""CPU tests of the host side of the product: the C-ABI library loads, exports
every symbol include/dcora_b200.h declares, refuses to compute without a GPU,
and its setup code (layout conversion, sparse Cholesky) is correct."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dcora_b200 import _lib
from dcora_b200._lib import AgentDesc
from tests.common import central_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "dcora_b200.h")).read()
    declared = set(re.findall(r"\b(dcora_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def _desc(P, keep, reg=0.1):
    Q = P.Q.tocsr()
    rp, ci, v = (Q.indptr.astype(np.int32), Q.indices.astype(np.int32),
                 Q.data.astype(np.float64))
    keep += [rp, ci, v]
    d = AgentDesc()
    d.n, d.l, d.b = P.L.n, P.L.l, P.L.b
    d.rowptr = rp.ctypes.data_as(_lib.c_int32_p)
    d.col = ci.ctypes.data_as(_lib.c_int32_p)
    d.val = v.ctypes.data_as(_lib.c_double_p)
    d.precond_reg = reg
    return d


def test_no_cpu_fallback_without_gpu(lib):
    if lib.dcora_device_count() > 0:
        pytest.skip("a GPU is present")
    P, g, info = central_problem("tinyGrid3D", 4)
    keep = []
    descs = (AgentDesc * 1)(_desc(P, keep))
    h = C.c_void_p()
    rc = lib.dcora_problem_create(descs, 1, 4, 3, 0, 0, C.byref(h))
    assert rc == -1 and not h.value
    assert b"no CPU fallback" in lib.dcora_last_error()


@pytest.mark.parametrize("name,r", [("tinyGrid3D", 3), ("smallGrid3D", 5),
                                    ("sphere2500", 4), ("city10000", 3),
                                    ("kitti_00", 5), ("range_aided_slam_test_2d", 4),
                                    ("range_aided_slam_test_3d", 5), ("tiers", 3)])
def test_layout_conversion_is_exact(lib, name, r):
    P, g, info = central_problem(name, r)
    keep = []
    descs = (AgentDesc * 1)(_desc(P, keep))
    err = C.c_double(-1)
    rc = lib.dcora_selfcheck_layout(descs, 1, r, info["d"], int(P.L.ra), C.byref(err))
    assert rc == 0
    assert err.value <= 1e-9 * max(1.0, abs(P.Q).max())


def test_layout_conversion_batch(lib):
    Ps = [central_problem(n, 5)[0] for n in ("tinyGrid3D", "smallGrid3D", "tinyGrid3D")]
    keep = []
    descs = (AgentDesc * 3)(*[_desc(P, keep) for P in Ps])
    err = C.c_double(-1)
    assert lib.dcora_selfcheck_layout(descs, 3, 5, 3, 0, C.byref(err)) == 0
    assert err.value <= 1e-9


@pytest.mark.parametrize("name", ["smallGrid3D", "sphere2500", "city10000", "tiers"])
def test_host_cholesky(lib, name):
    P, g, info = central_problem(name, 3)
    Q = P.Q.tocsr()
    rp, ci, v = Q.indptr.astype(np.int32), Q.indices.astype(np.int32), Q.data
    res, nnz = C.c_double(), C.c_int64()
    dh, npose = ((info["d"] + 1, P.L.n) if not P.L.ra else (1, 0))
    rc = lib.dcora_selfcheck_cholesky(Q.shape[0], rp.ctypes.data_as(_lib.c_int32_p),
                                      ci.ctypes.data_as(_lib.c_int32_p),
                                      v.ctypes.data_as(_lib.c_double_p), dh, npose,
                                      P.reg, C.byref(res), C.byref(nnz))
    assert rc == 0
    assert res.value < 1e-9
    assert nnz.value >= Q.shape[0]


def test_cholesky_rejects_indefinite(lib):
    rp = np.array([0, 2, 4], np.int32)
    ci = np.array([0, 1, 0, 1], np.int32)
    v = np.array([1.0, 2.0, 2.0, 1.0])
    res, nnz = C.c_double(), C.c_int64()
    rc = lib.dcora_selfcheck_cholesky(2, rp.ctypes.data_as(_lib.c_int32_p),
                                      ci.ctypes.data_as(_lib.c_int32_p),
                                      v.ctypes.data_as(_lib.c_double_p), 1, 0, 0.0,
                                      C.byref(res), C.byref(nnz))
    assert rc == -4


def _internal_order(P, info):
    """Q in the library's internal column order (pose blocks, spheres,
    landmarks) -- dcora_b200/csrc/layout.cpp."""
    import scipy.sparse as sp
    Q = P.Q.tocsr()
    L = P.L
    if not L.ra:
        return Q
    k, d, n, l = L.k, info["d"], L.n, L.l
    c = np.arange(k)
    m = np.empty(k, dtype=np.int64)
    dn = d * n
    rot = c < dn
    m[rot] = (c[rot] // d) * (d + 1) + c[rot] % d
    sph = (c >= dn) & (c < dn + l)
    m[sph] = (d + 1) * n + (c[sph] - dn)
    tr = (c >= dn + l) & (c < dn + l + n)
    m[tr] = (c[tr] - dn - l) * (d + 1) + d
    lm = c >= dn + l + n
    m[lm] = (d + 1) * n + l + (c[lm] - dn - l - n)
    Pm = sp.csr_matrix((np.ones(k), (m, c)), shape=(k, k))
    Q = (Pm @ Q @ Pm.T).tocsr()
    Q.sort_indices()
    return Q


@pytest.mark.parametrize("ordering", [0, 1])
@pytest.mark.parametrize("name", ["tinyGrid3D", "smallGrid3D", "sphere2500", "city10000",
                                  "kitti_00", "tiers", "range_aided_slam_test_3d",
                                  "range_aided_slam_test_2d"])
def test_block_trsv_schedule(lib, name, ordering):
    """The nested-dissection block schedule the preconditioner kernel walks
    solves (Q + reg I) x = b (replayed sequentially on the host)."""
    P, g, info = central_problem(name, 3)
    Q = _internal_order(P, info)
    rp, ci, v = Q.indptr.astype(np.int32), Q.indices.astype(np.int32), Q.data.astype(np.float64)
    res, nnz, stored = C.c_double(), C.c_int64(), C.c_int64()
    nb, lf, lb = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.dcora_selfcheck_trsv(Q.shape[0], rp.ctypes.data_as(_lib.c_int32_p),
                                  ci.ctypes.data_as(_lib.c_int32_p),
                                  v.ctypes.data_as(_lib.c_double_p), info["d"] + 1, P.L.n,
                                  P.reg, ordering, C.byref(res), C.byref(nnz), C.byref(stored),
                                  C.byref(nb), C.byref(lf), C.byref(lb))
    assert rc == 0
    assert res.value < 1e-9
    assert stored.value >= nnz.value - Q.shape[0] * 0 - 1 or stored.value > 0
    assert lf.value == lb.value >= 1
    if ordering == 1 and Q.shape[0] > 2000:
        # the block elimination tree of the ND ordering is shallow
        assert lf.value <= 64, (lf.value, nb.value)


def test_project_to_rotation_group_matches_oracle():
    """dcora_project_to_rotation_group (host entry of the rounding code) against the
    oracle's SVD projection, including reflections and near-singular inputs."""
    import dcora_b200 as dc
    from dcora_b200.api import project_to_rotation_group
    from oracle import manifold as mf
    rng = np.random.default_rng(0)
    for d in (2, 3):
        for trial in range(40):
            M = rng.standard_normal((d, d))
            if trial % 4 == 1:
                M[:, -1] = -M[:, -1]
            if trial % 4 == 2:      # close to a rotation
                M = mf.project_rotation(M) + 1e-3 * rng.standard_normal((d, d))
            if trial % 4 == 3:      # rank deficient + tiny noise
                M[:, 0] = M[:, 1] * 2 + 1e-9 * rng.standard_normal(d)
            R = project_to_rotation_group(M)
            want = mf.project_rotation(M)
            assert np.allclose(R @ R.T, np.eye(d), atol=1e-12)
            assert abs(np.linalg.det(R) - 1.0) < 1e-12
            if trial % 4 != 3:
                assert np.allclose(R, want, atol=1e-9), (d, trial)


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("S,robots", [(5, 6), (6, 8)])
def test_coupling_layout_is_exact_and_key_ordered(lib, S, robots, world):
    """build_coupling_layout with the global-column keys of rbcd.make_plans: the converted
    block / remainder structures reproduce [X | H] * C, and every output column's sources are
    stored in ascending key order (the rank-count-independent accumulation order)."""
    from dcora_b200 import graph as pg, rbcd, synthetic
    from dcora_b200._lib import HaloSegment
    meas, gt, idx = synthetic.grid3d(S, 0.2, seed=S)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, robots))
    for plan in rbcd.make_plans(gp, world):
        diag, _ = gp.split(agents=range(plan.a0, plan.a1))
        keep = []
        descs = (AgentDesc * (plan.a1 - plan.a0))()
        for i, a in enumerate(range(plan.a0, plan.a1)):
            rp, ci, v = diag[a]
            rp, ci, v = rp.astype(np.int32), ci.astype(np.int32), v.astype(np.float64)
            keep += [rp, ci, v]
            descs[i].n, descs[i].l, descs[i].b = int(gp.n_a[a]), int(gp.l_a[a]), int(gp.b_a[a])
            descs[i].rowptr = rp.ctypes.data_as(_lib.c_int32_p)
            descs[i].col = ci.ctypes.data_as(_lib.c_int32_p)
            descs[i].val = v.ctypes.data_as(_lib.c_double_p)
            descs[i].precond_reg = 0.1
        nseg = world if world > 1 else 0
        segs = (HaloSegment * max(1, nseg))()
        for s in range(nseg):
            segs[s].n, segs[s].l, segs[s].b = plan.seg
        Cm = plan.coupling.tocsr()
        rp, ci, v = Cm.indptr.astype(np.int32), Cm.indices.astype(np.int32), Cm.data.astype(np.float64)
        key = np.ascontiguousarray(plan.row_key, dtype=np.int64)
        err, bad = C.c_double(-1), C.c_int64(-1)
        rc = lib.dcora_selfcheck_coupling(descs, plan.a1 - plan.a0, 4, 3, 1, segs, nseg,
                                          rp.ctypes.data_as(_lib.c_int32_p),
                                          ci.ctypes.data_as(_lib.c_int32_p),
                                          v.ctypes.data_as(_lib.c_double_p),
                                          key.ctypes.data_as(C.POINTER(C.c_int64)),
                                          C.byref(err), C.byref(bad))
        assert rc == 0, lib.dcora_last_error()
        assert err.value <= 1e-9 * max(1.0, np.abs(v).max() if len(v) else 1.0)
        assert bad.value == 0


def test_subtree_parallel_cholesky_is_bit_identical(lib):
    """The numeric phase factors independent elimination-tree subtrees on several host threads
    (few large agents leave cores idle): same factor, bit for bit, whatever the thread count."""
    from dcora_b200 import graph as pg, synthetic
    import scipy.sparse as sp
    meas, gt, idx = synthetic.grid3d(14, 0.1, seed=2)
    gp = pg.GlobalProblem(meas, pg.Partition.contiguous(meas.n, 2))
    diag, _ = gp.split(agents=[0])
    rp, ci, v = diag[0]
    k = len(rp) - 1
    assert k >= 4000          # below that the serial path is taken
    out = []
    for threads in (1, 2, 5):
        cs, nnz = C.c_double(), C.c_int64()
        rc = lib.dcora_selfcheck_cholesky_threads(
            k, rp.ctypes.data_as(_lib.c_int32_p), ci.ctypes.data_as(_lib.c_int32_p),
            v.ctypes.data_as(_lib.c_double_p), 1, 0, 0.05, threads, C.byref(cs), C.byref(nnz))
        assert rc == 0
        out.append((cs.value, nnz.value))
    assert out[0] == out[1] == out[2], out


_CHOL_SNIPPET = r"""
import ctypes as C, sys, json
import numpy as np
sys.path.insert(0, {root!r})
from dcora_b200 import _lib
from tests.common import central_problem
from tests.test_host import _internal_order
lib = _lib.load()
out = {{}}
for name in ("city10000", "tiers", "sphere2500"):
    P, g, info = central_problem(name, 3)
    Q = _internal_order(P, info)
    rp, ci, v = Q.indptr.astype(np.int32), Q.indices.astype(np.int32), Q.data.astype(np.float64)
    cs, nnz = C.c_double(), C.c_int64()
    rc = lib.dcora_selfcheck_cholesky_threads(Q.shape[0], rp.ctypes.data_as(_lib.c_int32_p),
        ci.ctypes.data_as(_lib.c_int32_p), v.ctypes.data_as(_lib.c_double_p), info["d"] + 1, P.L.n,
        P.reg, 1, C.byref(cs), C.byref(nnz))
    out[name] = (rc, cs.value, nnz.value)
print(json.dumps(out))
"""


def test_supernodal_and_uplooking_factors_agree():
    """The supernodal numeric phase (default) and the scalar up-looking one (DCORA_CHOL=u) compute
    the same factor on the same ordering: nnz equal, position-weighted checksum to 1e-12."""
    import json
    import subprocess
    import sys
    res = {}
    for mode in ("s", "u"):
        env = dict(os.environ, DCORA_CHOL=mode)
        out = subprocess.run([sys.executable, "-c", _CHOL_SNIPPET.format(root=ROOT)], env=env,
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        res[mode] = json.loads(out.stdout.splitlines()[-1])
    for name, (rc, cs, nnz) in res["s"].items():
        rc_u, cs_u, nnz_u = res["u"][name]
        assert rc == rc_u == 0 and nnz == nnz_u
        assert abs(cs - cs_u) <= 1e-12 * abs(cs_u), (name, cs, cs_u)
