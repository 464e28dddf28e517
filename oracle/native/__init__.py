"""ctypes binding of the native CPU oracle (oracle/native/dcora_oracle.cpp): one
Agent::updateX = Cholesky of Q + reg I + QuadraticOptimizer::optimize, in C++ -O3.
TEST INFRASTRUCTURE: used by tests/ (cross-check of the scipy oracle) and by bench.py's
cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB, build

_lib = None
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.oracle_agent_update.restype = C.c_int
        lib.oracle_agent_update.argtypes = [C.c_int] * 6 + [_ip, _ip, _dp, _dp, C.c_double, _dp, _dp,
                                                            C.c_int, C.c_int, C.c_double, C.c_double,
                                                            C.c_int, _dp]
        lib.oracle_problem_create.restype = C.c_void_p
        lib.oracle_problem_create.argtypes = [C.c_int] * 6 + [_ip, _ip, _dp, C.c_double]
        lib.oracle_problem_destroy.restype = None
        lib.oracle_problem_destroy.argtypes = [C.c_void_p]
        lib.oracle_problem_update.restype = C.c_int
        lib.oracle_problem_update.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_int,
                                              C.c_double, C.c_double, C.c_int, _dp]
        lib.oracle_apply.restype = C.c_int
        lib.oracle_apply.argtypes = [C.c_int] * 7 + [_ip, _ip, _dp, _dp, C.c_double, _dp, _dp, _dp]
        _lib = lib
    return _lib


def _csr(Q):
    Q = Q.tocsr()
    Q.sort_indices()
    return (np.ascontiguousarray(Q.indptr, np.int32), np.ascontiguousarray(Q.indices, np.int32),
            np.ascontiguousarray(Q.data, np.float64))


def _f(M):
    """r x k numpy -> column-major buffer."""
    return np.asfortranarray(np.asarray(M, np.float64))


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def agent_update(L, Q, G, reg, X0, max_outer=3, max_inner=50, tol=1e-2, delta0=100.0,
                 retraction="qf"):
    """Returns (X, stats dict).  L: oracle.manifold.Layout; reg < 0 or None: no preconditioner."""
    lib = load()
    rp, ci, va = _csr(Q)
    Gf = _f(G) if G is not None and np.any(G) else None
    X = _f(X0)
    out = np.empty_like(X)
    st = np.zeros(12)
    rc = lib.oracle_agent_update(L.r, L.d, L.n, L.l, L.b, int(L.ra), _ptr(rp, _ip), _ptr(ci, _ip),
                                 _ptr(va, _dp), _ptr(Gf, _dp), -1.0 if reg is None else float(reg),
                                 _ptr(X, _dp), _ptr(out, _dp), int(max_outer), int(max_inner),
                                 float(tol), float(delta0), 0 if retraction == "qf" else 1,
                                 _ptr(st, _dp))
    if rc != 0:
        raise RuntimeError("native oracle: Cholesky factorisation failed")
    return out, dict(zip(STAT_KEYS, st))


STAT_KEYS = ("fInit", "gradNormInit", "fOpt", "gradNormOpt", "outer", "inner", "factor_s",
             "solve_s", "nnzL", "tcg_status", "precond", "spmm")


class AgentProblem:
    """One agent's Q with its Cholesky factor built once (the reference caches the factor
    until Q changes, Graph.cpp:530-533); `update` = one Agent::updateX.  Distinct objects
    may be driven from distinct threads (the call releases the GIL)."""

    def __init__(self, L, Q, reg):
        self.lib = load()
        self.L = L
        rp, ci, va = _csr(Q) if not isinstance(Q, tuple) else (
            np.ascontiguousarray(Q[0], np.int32), np.ascontiguousarray(Q[1], np.int32),
            np.ascontiguousarray(Q[2], np.float64))
        self.h = self.lib.oracle_problem_create(L.r, L.d, L.n, L.l, L.b, int(L.ra), _ptr(rp, _ip),
                                                _ptr(ci, _ip), _ptr(va, _dp),
                                                -1.0 if reg is None else float(reg))
        if not self.h:
            raise RuntimeError("native oracle: Cholesky factorisation failed")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_problem_destroy(self.h)
            self.h = None

    def update(self, G, X0, max_outer=3, max_inner=50, tol=1e-2, delta0=100.0, retraction="qf"):
        Gf = _f(G) if G is not None else None
        X = _f(X0)
        out = np.empty_like(X)
        st = np.zeros(12)
        rc = self.lib.oracle_problem_update(self.h, _ptr(Gf, _dp), _ptr(X, _dp), _ptr(out, _dp),
                                            int(max_outer), int(max_inner), float(tol),
                                            float(delta0), 0 if retraction == "qf" else 1,
                                            _ptr(st, _dp))
        assert rc == 0
        return out, dict(zip(STAT_KEYS, st))


OPS = {"egrad": 0, "tangent": 1, "retract_qf": 2, "retract_polar": 3, "project": 4, "solve": 5,
       "rhess": 6}


def apply(op, L, Q, X=None, V=None, G=None, reg=0.1):
    lib = load()
    rp, ci, va = _csr(Q)
    Gf = _f(G) if G is not None and np.any(G) else None
    Xf = _f(X) if X is not None else None
    Vf = _f(V) if V is not None else None
    out = np.empty((L.r, L.k), order="F")
    rc = lib.oracle_apply(OPS[op], L.r, L.d, L.n, L.l, L.b, int(L.ra), _ptr(rp, _ip), _ptr(ci, _ip),
                          _ptr(va, _dp), _ptr(Gf, _dp), float(reg), _ptr(Xf, _dp), _ptr(Vf, _dp),
                          _ptr(out, _dp))
    if rc != 0:
        raise RuntimeError(f"native oracle: op {op} failed ({rc})")
    return out
