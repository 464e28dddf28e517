"""Host-side mirror of the reference interface for the hot path, over the C-ABI.

Names and argument meaning follow DCORA::QuadraticProblem
(include/DCORA/QuadraticProblem.h:44-185), DCORA::QuadraticOptimizer
(include/DCORA/QuadraticOptimizer.h:30-120), ROptParameters / ROPTResult
(include/DCORA/DCORA_types.h:152-233) and the certification free functions
(include/DCORA/DCORA_utils.h:224-292).  Nothing here computes: every method is
one call into libdcora_b200.so, which runs on the GPU or fails.

Matrices are numpy arrays of shape (r, K) (any memory order; copied to the
column-major layout the reference uses) or CUDA torch tensors of shape (K, r),
contiguous -- the same bytes, already resident in HBM.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import AgentDesc, CertifyResult, HaloSegment, ROptParams, ROptResult, check

C_int64_p = C.POINTER(C.c_int64)

ORDER_SE, ORDER_RA = 0, 1
RETRACT_QF, RETRACT_POLAR = 0, 1
TCG_STATUS = ["TR_NEGCURVTURE", "TR_EXCREGION", "TR_LCON", "TR_SCON", "TR_MAXITER"]


@dataclass
class AgentData:
    """What Graph hands to QuadraticProblem for one agent: Q (scipy CSR or
    (rowptr, col, val)), optional G (r x k), n / l / b and the preconditioner
    regularisation (None disables the Cholesky preconditioner)."""
    n: int
    Q: object
    l: int = 0
    b: int = 0
    G: np.ndarray | None = None
    precond_reg: float | str | None = 0.1   # None: no factor; "auto_ra": RA rule


@dataclass
class ROptParameters:
    method: str = "RTR"
    verbose: bool = False
    gradnorm_tol: float = 1e-2
    RGD_stepsize: float = 1e-3
    RGD_use_preconditioner: bool = True
    RTR_iterations: int = 3
    RTR_tCG_iterations: int = 50
    RTR_initial_radius: float = 100.0
    time_bound_s: float = 5.0      # QuadraticOptimizer.cpp:252; <=0 disables
    retraction: int = RETRACT_QF
    # rho of RTR from <egrad, d> + 0.5 <dQ, d> instead of f(x) - f(x+): see dcora_ropt_params
    stable_decrease: bool = False

    def to_c(self):
        p = ROptParams()
        p.method = 0 if self.method == "RTR" else 1
        p.verbose = int(self.verbose)
        p.gradnorm_tol = self.gradnorm_tol
        p.RGD_stepsize = self.RGD_stepsize
        p.RGD_use_preconditioner = int(self.RGD_use_preconditioner)
        p.RTR_iterations = self.RTR_iterations
        p.RTR_tCG_iterations = self.RTR_tCG_iterations
        p.RTR_initial_radius = self.RTR_initial_radius
        p.time_bound_s = self.time_bound_s
        p.retraction = self.retraction
        p.stable_decrease = int(self.stable_decrease)
        return p


@dataclass
class ROPTResult:
    success: bool = False
    fInit: float = 0.0
    gradNormInit: float = 0.0
    fOpt: float = 0.0
    gradNormOpt: float = 0.0
    elapsedMs: float = 0.0
    tCGStatus: int = 4
    outer_iterations: int = 0
    inner_iterations: int = 0
    accepted: bool = False


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Arg:
    """Pointer view of a matrix argument (keeps the backing buffer alive)."""

    def __init__(self, x, r, K, out=False):
        self.torch = _is_torch(x)
        if self.torch:
            assert x.is_cuda and x.is_contiguous() and tuple(x.shape) == (K, r), \
                "CUDA tensors must be contiguous with shape (K, r)"
            if not out:
                # the handle launches on its own stream: the producer of a
                # device-resident input must have finished
                import torch
                torch.cuda.current_stream(x.device).synchronize()
            self.buf = x
            self.ptr = C.c_void_p(x.data_ptr())
        else:
            a = np.asarray(x, dtype=np.float64)
            assert a.shape == (r, K), f"expected shape {(r, K)}, got {a.shape}"
            if out:
                # the library writes column-major r x K doubles straight into this buffer
                assert a is x and a.flags.f_contiguous and a.flags.writeable, \
                    "output arrays must be float64, Fortran-ordered (column-major) and writeable"
                self.buf = a
            else:
                self.buf = np.asfortranarray(a)
            self.ptr = self.buf.ctypes.data_as(C.c_void_p)


def _new_out(like, r, K):
    if _is_torch(like):
        import torch
        return torch.empty((K, r), dtype=torch.float64, device=like.device)
    return np.empty((r, K), dtype=np.float64, order="F")


def project_to_rotation_group(M):
    """projectToRotationGroup (src/DCORA_utils.cpp:1661-1675), d = 2 or 3."""
    a = np.asfortranarray(np.asarray(M, dtype=np.float64))
    d = a.shape[0]
    assert a.shape == (d, d)
    out = np.empty((d, d), order="F")
    check(_lib.load().dcora_project_to_rotation_group(
        a.ctypes.data_as(C.POINTER(C.c_double)), d, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


class QuadraticProblem:
    """Batch of `len(agents)` agent problems (1 = the reference's single
    QuadraticProblem).  src/QuadraticProblem.cpp:19-34."""

    def __init__(self, agents, r, d, order=ORDER_SE, device=0, _handle=None,
                 _parent=None):
        self.lib = _lib.load()
        self.r, self.d, self.order, self.device = r, d, order, device
        if _handle is not None:
            self.h = _handle
            self.agents = _parent.agents
            self.A = _parent.A
            self.K = _parent.K
            self.koff = _parent.koff
            self.KH = getattr(_parent, "KH", 0)
            return
        if isinstance(agents, AgentData):
            agents = [agents]
        self.agents = agents
        self.A = len(agents)
        descs = (AgentDesc * self.A)()
        keep = []
        koff = [0]
        for i, ag in enumerate(agents):
            if isinstance(ag.Q, tuple):
                rowptr, col, val = ag.Q
            else:
                Q = ag.Q.tocsr()
                rowptr, col, val = Q.indptr, Q.indices, Q.data
            rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
            col = np.ascontiguousarray(col, dtype=np.int32)
            val = np.ascontiguousarray(val, dtype=np.float64)
            k = (d + 1) * ag.n + ag.l + ag.b
            assert len(rowptr) == k + 1, "Q has the wrong dimension"
            keep += [rowptr, col, val]
            descs[i].n, descs[i].l, descs[i].b = ag.n, ag.l, ag.b
            descs[i].rowptr = rowptr.ctypes.data_as(_lib.c_int32_p)
            descs[i].col = col.ctypes.data_as(_lib.c_int32_p)
            descs[i].val = val.ctypes.data_as(_lib.c_double_p)
            if ag.G is not None:
                G = np.asfortranarray(np.asarray(ag.G, np.float64))
                assert G.shape == (r, k)
                keep.append(G)
                descs[i].G = G.ctypes.data_as(_lib.c_double_p)
            descs[i].precond_reg = (-1.0 if ag.precond_reg is None else
                                    -2.0 if ag.precond_reg == "auto_ra" else ag.precond_reg)
            koff.append(koff[-1] + k)
        self.koff = koff
        self.K = koff[-1]
        self.KH = 0
        h = C.c_void_p()
        check(self.lib.dcora_problem_create(descs, self.A, r, d, order, device,
                                            C.byref(h)))
        self.h = h

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and h.value:
            self.lib.dcora_problem_destroy(h)
            self.h = None

    # -- dims (QuadraticProblem.h:58-94)
    def relaxation_rank(self):
        return self.r

    def dimension(self):
        return self.d

    def problem_dimension(self):
        return self.K

    def with_rank(self, r):
        """The same graph at another relaxation rank (the staircase re-creates
        Graph/QuadraticProblem per rank; Q is rank independent)."""
        h = C.c_void_p()
        check(self.lib.dcora_problem_clone_rank(self.h, r, C.byref(h)))
        return QuadraticProblem(None, r, self.d, self.order, self.device,
                                _handle=h, _parent=self)

    def nnz(self):
        return self.lib.dcora_problem_nnz(self.h)

    def factor_nnz(self):
        return self.lib.dcora_problem_factor_nnz(self.h)

    def precond_reg(self, agent=0):
        return self.lib.dcora_problem_precond_reg(self.h, agent)

    def launch_count(self):
        return self.lib.dcora_problem_launch_count(self.h)

    def stream(self):
        return self.lib.dcora_problem_stream(self.h)

    # -- data
    def set_G(self, G):
        if G is None:
            check(self.lib.dcora_problem_set_G(self.h, None))
        else:
            a = _Arg(G, self.r, self.K)
            check(self.lib.dcora_problem_set_G(self.h, a.ptr))

    # -- QuadraticProblem.cpp:38-123
    def _scalars(self, fn, X):
        a = _Arg(X, self.r, self.K)
        out = (C.c_double * self.A)()
        check(fn(self.h, a.ptr, out))
        v = np.array(out[:])
        return float(v[0]) if self.A == 1 else v

    def f(self, X):
        return self._scalars(self.lib.dcora_problem_f, X)

    def RieGradNorm(self, X):
        return self._scalars(self.lib.dcora_problem_rgradnorm, X)

    def _unary(self, fn, X):
        a = _Arg(X, self.r, self.K)
        out = _new_out(X, self.r, self.K)
        o = _Arg(out, self.r, self.K, out=True)
        check(fn(self.h, a.ptr, o.ptr))
        return out

    def _binary(self, fn, X, V, *extra):
        a, b = _Arg(X, self.r, self.K), _Arg(V, self.r, self.K)
        out = _new_out(X, self.r, self.K)
        o = _Arg(out, self.r, self.K, out=True)
        check(fn(self.h, a.ptr, b.ptr, o.ptr, *extra))
        return out

    def EucGrad(self, X):
        return self._unary(self.lib.dcora_problem_egrad, X)

    def EucHessianEta(self, V):
        return self._unary(self.lib.dcora_problem_hessvec, V)

    def RieGrad(self, X):
        return self._unary(self.lib.dcora_problem_rgrad, X)

    def PreCondition(self, X, V):
        return self._binary(self.lib.dcora_problem_precond, X, V)

    def factor_solve(self, V):
        """(Q + reg I)^-1 V per agent (the preconditioner's Cholesky solve without the
        tangent projection: dcora_problem_precond with X = NULL)."""
        v = _Arg(V, self.r, self.K)
        out = _new_out(V, self.r, self.K)
        o = _Arg(out, self.r, self.K, out=True)
        check(self.lib.dcora_problem_precond(self.h, None, v.ptr, o.ptr))
        return out

    def Retract(self, X, V, retraction=RETRACT_QF):
        return self._binary(self.lib.dcora_problem_retract, X, V, retraction)

    def projectToTangentSpace(self, X, V):
        return self._binary(self.lib.dcora_problem_project_tangent, X, V)

    def project(self, M):
        """LiftedSEManifold::project / LiftedRAManifold::project."""
        return self._unary(self.lib.dcora_problem_project_manifold, M)

    # -- QuadraticProblem.cpp:138-234
    def escapeSaddle(self, Xopt, theta, v, gradient_tolerance,
                     preconditioned_gradient_tolerance, isSecondOrder=False,
                     retraction=RETRACT_QF):
        """`self` is the rank-r problem; Xopt is (r-1) x K.  Returns (X, ok)."""
        a = _Arg(Xopt, self.r - 1, self.K)
        vv = np.ascontiguousarray(v, dtype=np.float64)
        assert vv.shape == (self.K,)
        out = np.empty((self.r, self.K), dtype=np.float64, order="F")
        rc = check(self.lib.dcora_problem_escape_saddle(
            self.h, a.ptr, float(theta), vv.ctypes.data_as(C.c_void_p),
            float(gradient_tolerance), float(preconditioned_gradient_tolerance),
            int(isSecondOrder), retraction, out.ctypes.data_as(C.c_void_p)))
        return out, rc == 0

    # -- certification (DCORA_utils.cpp:1713-1982)
    def certify(self, X, eta, max_matvecs=20000, tol=1e-9, want_vector=True):
        a = _Arg(X, self.r, self.K)
        res = CertifyResult()
        v = np.zeros(self.K) if want_vector else None
        check(self.lib.dcora_certify(
            self.h, a.ptr, float(eta), int(max_matvecs), float(tol), C.byref(res),
            v.ctypes.data_as(C.c_void_p) if v is not None else None))
        return res, v

    # -- multi-agent RBCD (Agent.cpp:113-152,844-906,1202-1278; Graph.cpp:685-822)
    def set_coupling(self, C, halo_segments=(), row_key=None):
        """Register the cross-agent coupling: C is a scipy CSR (or (rowptr, col,
        val)) of shape (K + KH, K); halo_segments = [(n, l, b), ...]; row_key (K + KH int64,
        optional): accumulation order of the sources (dcora_problem_set_coupling_keyed)."""
        if isinstance(C, tuple):
            rowptr, col, val = C
        else:
            C = C.tocsr()
            rowptr, col, val = C.indptr, C.indices, C.data
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        segs = (HaloSegment * max(1, len(halo_segments)))()
        KH = 0
        for i, (n, l, b) in enumerate(halo_segments):
            segs[i].n, segs[i].l, segs[i].b = n, l, b
            KH += (self.d + 1) * n + l + b
        assert len(rowptr) == self.K + KH + 1, "coupling has the wrong number of rows"
        key = None
        if row_key is not None:
            key = np.ascontiguousarray(row_key, dtype=np.int64)
            assert key.shape == (self.K + KH,)
        check(self.lib.dcora_problem_set_coupling_keyed(
            self.h, segs, len(halo_segments), rowptr.ctypes.data_as(_lib.c_int32_p),
            col.ctypes.data_as(_lib.c_int32_p), val.ctypes.data_as(_lib.c_double_p),
            key.ctypes.data_as(C_int64_p) if key is not None else None))
        self.KH = KH

    def _mask(self, active):
        if active is None:
            return None, None
        arr = np.ascontiguousarray(active, dtype=np.uint8)
        assert arr.shape == (self.A,)
        return arr, arr.ctypes.data_as(_lib.c_uint8_p)

    def update_G(self, X, H=None, active=None):
        a = _Arg(X, self.r, self.K)
        h = _Arg(H, self.r, self.KH) if (H is not None and self.KH > 0) else None
        keep, act = self._mask(active)
        check(self.lib.dcora_problem_update_G(self.h, a.ptr, h.ptr if h else None, act))

    def pack_columns(self, X, idx, out=None):
        a = _Arg(X, self.r, self.K)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        if out is None:
            out = _new_out(X, self.r, len(idx))
        o = _Arg(out, self.r, len(idx), out=True)
        check(self.lib.dcora_problem_pack_columns(
            self.h, a.ptr, idx.ctypes.data_as(_lib.c_int32_p), len(idx), o.ptr))
        return out

    def set_iterate_tracking(self, on=True):
        """See dcora_problem_set_iterate_tracking: call again after writing to the tracked
        device buffer yourself."""
        check(self.lib.dcora_problem_set_iterate_tracking(self.h, int(bool(on))))

    def load_iterate(self, X_host, X_device):
        """numpy (r, K) column-major (pinned or pageable) -> CUDA tensor (K, r)."""
        a = np.asarray(X_host)
        assert a.dtype == np.float64 and a.shape == (self.r, self.K) and a.flags.f_contiguous
        d = _Arg(X_device, self.r, self.K, out=True)
        check(self.lib.dcora_problem_load_iterate(self.h, a.ctypes.data_as(C.c_void_p), d.ptr))

    def store_iterate(self, X_device, X_host):
        a = X_host
        assert a.dtype == np.float64 and a.shape == (self.r, self.K) and a.flags.f_contiguous
        d = _Arg(X_device, self.r, self.K)
        check(self.lib.dcora_problem_store_iterate(self.h, d.ptr, a.ctypes.data_as(C.c_void_p)))

    def round_solution(self, X):
        """projectSolutionRASLAM (src/DCORA_utils.cpp:1984-2031) of the
        whole batch as one solution: (r, K) -> (d, K)."""
        a = _Arg(X, self.r, self.K)
        out = _new_out(X, self.d, self.K)
        o = _Arg(out, self.d, self.K, out=True)
        check(self.lib.dcora_problem_round_solution(self.h, a.ptr, o.ptr))
        return out

    def align_trajectory(self, X, Tw0, is_global=True):
        """alignLiftedTrajectoryToFrame (src/DCORA_utils.cpp:2262-2289); Tw0: (r, d+1)."""
        a = _Arg(X, self.r, self.K)
        T = np.asfortranarray(np.asarray(Tw0, dtype=np.float64))
        assert T.shape == (self.r, self.d + 1)
        out = _new_out(X, self.d, self.K)
        o = _Arg(out, self.d, self.K, out=True)
        check(self.lib.dcora_problem_align_trajectory(
            self.h, a.ptr, T.ctypes.data_as(C.POINTER(C.c_double)), int(bool(is_global)), o.ptr))
        return out

    def certificate_prepare(self, X, XQ):
        """Lambda(X) of the handle's agents from X and (X Q_global) on its columns."""
        a, b = _Arg(X, self.r, self.K), _Arg(XQ, self.r, self.K)
        check(self.lib.dcora_problem_certificate_prepare(self.h, a.ptr, b.ptr))

    def certificate_apply(self, V, QV, out=None):
        """S V = QV - V Lambda for a block of r vectors (after certificate_prepare)."""
        a, b = _Arg(V, self.r, self.K), _Arg(QV, self.r, self.K)
        if out is None:
            out = _new_out(V, self.r, self.K)
        o = _Arg(out, self.r, self.K, out=True)
        check(self.lib.dcora_problem_certificate_apply(self.h, a.ptr, b.ptr, o.ptr))
        return out

    def cost_terms(self, X):
        a = _Arg(X, self.r, self.K)
        q = (C.c_double * self.A)()
        l = (C.c_double * self.A)()
        check(self.lib.dcora_problem_cost_terms(self.h, a.ptr, q, l))
        return np.array(q[:]), np.array(l[:])

    def combine_project(self, a, A, b, B, active=None):
        x, y = _Arg(A, self.r, self.K), _Arg(B, self.r, self.K)
        out = _new_out(A, self.r, self.K)
        o = _Arg(out, self.r, self.K, out=True)
        keep, act = self._mask(active)
        check(self.lib.dcora_problem_combine_project(self.h, float(a), x.ptr, float(b),
                                                     y.ptr, act, o.ptr))
        return out

    def time_kernel(self, kind, iters=20):
        """(mean ms, algorithmic bytes) of one application of a kernel class."""
        ms, by = C.c_double(), C.c_double()
        check(self.lib.dcora_problem_time_kernel(self.h, KERNELS[kind], iters,
                                                 C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def time_spmm(self, iters=20):
        ms, by = C.c_double(), C.c_double()
        check(self.lib.dcora_problem_time_spmm(self.h, iters, C.byref(ms), C.byref(by)))
        return ms.value, by.value


KERNELS = {"spmm": 0, "precond": 1, "project": 2, "retract": 3, "dot": 4, "axpby": 5,
           "rhess": 6}


class QuadraticOptimizer:
    """src/QuadraticOptimizer.cpp:20-50."""

    def __init__(self, problem: QuadraticProblem, params: ROptParameters | None = None):
        self.problem = problem
        self.params = params or ROptParameters()
        self._raw = ()
        self._results: list[ROPTResult] | None = []

    # setters (QuadraticOptimizer.h:52-96)
    def setVerbose(self, v): self.params.verbose = v
    def setAlgorithm(self, alg): self.params.method = alg
    def setRGDStepsize(self, s): self.params.RGD_stepsize = s
    def setRTRIterations(self, n): self.params.RTR_iterations = n
    def setGradientNormTolerance(self, t): self.params.gradnorm_tol = t
    def setRTRInitialRadius(self, r): self.params.RTR_initial_radius = r
    def setRTRtCGIterations(self, n): self.params.RTR_tCG_iterations = n

    def optimize(self, Y, active=None, out=None):
        """`out` may be Y itself (the solve is then in place)."""
        P = self.problem
        a = _Arg(Y, P.r, P.K)
        if out is None:
            out = _new_out(Y, P.r, P.K)
        o = _Arg(out, P.r, P.K, out=True)
        res = (ROptResult * P.A)()
        prm = self.params.to_c()
        act = None
        if active is not None:
            act_arr = np.ascontiguousarray(active, dtype=np.uint8)
            assert act_arr.shape == (P.A,)
            act = act_arr.ctypes.data_as(_lib.c_uint8_p)
        check(P.lib.dcora_rtr_solve(P.h, a.ptr, o.ptr, C.byref(prm), act, res))
        self._raw = res          # converted lazily: a batch holds hundreds of agents
        self._results = None
        return out

    @property
    def results(self):
        if self._results is None:
            self._results = [ROPTResult(bool(x.success), x.fInit, x.gradNormInit, x.fOpt,
                                        x.gradNormOpt, x.elapsedMs, x.tCGStatus,
                                        x.outer_iterations, x.inner_iterations,
                                        bool(x.accepted)) for x in self._raw]
        return self._results

    def getOptResult(self):
        return self.results[0] if len(self.results) == 1 else self.results
