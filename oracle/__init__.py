"""CPU oracle for the DCORA Riemannian-staircase hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference algorithm (adthoms/dcora, files cited per function as
``path:line`` relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  Nothing under ``dcora_b200/`` imports it and
the product path never falls back to it: the product fails loudly when the
CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * data matrices / manifold closed forms / fixed-point behaviour: PINNED on
    the reference's four noiseless fixtures (f*=0, grad=0 at the VERTEX ground
    truth, tests/testAgent.cpp:101-456) and on the closed forms that
    tests/testManifold.cpp:354-440 compares ROPTLIB against.
  * RTR/tCG iteration sequence, retraction variant, preconditioner numerics,
    certification (fastVerification / computeMinimumEigenPair /
    constructDualCertificateMatrix*) and escapeSaddle: PARITY UNPINNED by the
    reference's own tests (it has none for them) and the third-party
    arithmetic (ROPTLIB yuluntian/ROPTLIB@feature/cmake, Spectra v1.0.1,
    CHOLMOD) is not vendored in /root/reference and cannot be built here.
    These are restated from the published algorithms and anchored on the
    reference's call sites; agreement is checked on converged results.
"""
