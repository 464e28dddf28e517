/*
 * dcora_b200 -- C-ABI of the B200-native DCORA Riemannian-staircase hot path.
 *
 * Every entry point replaces one piece of the reference's C++ class surface
 * (adthoms/dcora; citations are path:line in the reference checkout).  The
 * library is a plain shared object (extern "C", pointers + sizes, no torch /
 * Eigen types); INTEGRATION.md shows the reference-side shim that binds it.
 *
 * Conventions
 *  - All dense matrices are fp64, COLUMN-MAJOR r x k (Eigen default,
 *    include/DCORA/DCORA_types.h:34), in the reference column ordering:
 *      DCORA_ORDER_SE: [Y1 p1 | Y2 p2 | ...]               k=(d+1)n
 *      DCORA_ORDER_RA: [Y1..Yn | r1..rl | p1..pn | L1..Lb] k=(d+1)n+l+b
 *    (include/DCORA/manifold/LiftedVariable.h:27-30,148-152).
 *  - A handle holds a BATCH of `num_agents` independent agent problems
 *    (block-diagonal Q); dense arguments of a batch are the per-agent
 *    matrices concatenated in agent order (agent a occupies r*k_a doubles).
 *    num_agents==1 is the reference's single QuadraticProblem.
 *  - Dense pointer arguments may be host OR device pointers (detected with
 *    cudaPointerGetAttributes).  Host pointers are staged through pinned
 *    memory inside the call; device pointers are used in place (inputs
 *    already resident in HBM).
 *  - Return value: 0 ok; <0 hard error (CUDA / allocation / bad argument,
 *    text in dcora_last_error()); >0 soft failure mirroring the reference's
 *    bool / std::optional returns.  There is NO CPU fallback: without a CUDA
 *    device every compute entry point returns DCORA_ERR_CUDA.
 *  - A handle is not re-entrant; distinct handles may be driven from
 *    distinct host threads (one CUDA stream per handle).
 */
#ifndef DCORA_B200_H_
#define DCORA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCORA_OK 0
#define DCORA_ERR_CUDA (-1)
#define DCORA_ERR_ARG (-2)
#define DCORA_ERR_ALLOC (-3)
#define DCORA_ERR_FACTOR (-4) /* Q + reg*I not positive definite */
#define DCORA_SOFT_FAIL 1

#define DCORA_ORDER_SE 0
#define DCORA_ORDER_RA 1

#define DCORA_REG_NONE (-1.0)
#define DCORA_REG_AUTO_RA (-2.0)

#define DCORA_RETRACT_QF 0    /* ROPTLIB Stiefel qf retraction, diag(R)>0 */
#define DCORA_RETRACT_POLAR 1 /* polar factor (projectToStiefelManifold)  */

/* ROPTLIB::tCGstatusSet as stored in ROPTResult.tCGStatus
 * (include/DCORA/DCORA_types.h:226). */
#define DCORA_TCG_NEGCURVTURE 0
#define DCORA_TCG_EXCREGION 1
#define DCORA_TCG_LCON 2
#define DCORA_TCG_SCON 3
#define DCORA_TCG_MAXITER 4

typedef struct dcora_problem dcora_problem;

/* One agent's data matrices, as Graph hands them to QuadraticProblem:
 * Graph::quadraticMatrix() (CSR, int32 indices, include/DCORA/DCORA_types.h:36,
 * src/Graph.cpp:523-528), Graph::linearMatrix() (src/Graph.cpp:535-540) and
 * the regularisation of Graph::constructPreconditioner (src/Graph.cpp:1901-1917). */
typedef struct {
  int32_t n, l, b;        /* poses, unit spheres, landmarks (Graph::n/l/b)     */
  const int32_t *rowptr;  /* k+1                                                */
  const int32_t *col;     /* nnz                                                */
  const double *val;      /* nnz                                                */
  const double *G;        /* r x k column-major, or NULL for G = 0              */
  double precond_reg;     /* reg in chol(Q + reg*I); DCORA_REG_NONE disables the
                             factor; DCORA_REG_AUTO_RA applies the range-aided
                             rule lambda_max(Q)/(1e6-1), fallback 0.1
                             (src/Graph.cpp:1919-1960)                          */
} dcora_agent_desc;

/* ROptParameters (include/DCORA/DCORA_types.h:152-200) + the hard-coded
 * Solver->TimeBound (src/QuadraticOptimizer.cpp:252) made explicit. */
typedef struct {
  int32_t method;             /* 0 = RTR, 1 = RGD                               */
  int32_t verbose;
  double gradnorm_tol;        /* 1e-2                                           */
  double RGD_stepsize;        /* 1e-3                                           */
  int32_t RGD_use_preconditioner;
  int32_t RTR_iterations;     /* 3                                              */
  int32_t RTR_tCG_iterations; /* 50                                             */
  double RTR_initial_radius;  /* 100                                            */
  double time_bound_s;        /* 5.0 in the reference; <=0 disables             */
  int32_t retraction;         /* DCORA_RETRACT_QF / _POLAR                      */
  int32_t stable_decrease;    /* 0 (reference arithmetic): rho from f(x) - f(x+);
                                 1: from <egrad(x), d> + 0.5 <d Q, d>, d = x+ - x -- the
                                 same number without subtracting two large costs (an
                                 agent's local cost can be 1e9 while the decrease RTR
                                 tests is 1e-6)                                   */
} dcora_ropt_params;

/* ROPTResult (include/DCORA/DCORA_types.h:203-233), one per agent. */
typedef struct {
  int32_t success;
  int32_t tCGStatus;
  double fInit, gradNormInit, fOpt, gradNormOpt, elapsedMs;
  int32_t outer_iterations; /* RTR iterations taken                             */
  int32_t inner_iterations; /* total tCG iterations                             */
  int32_t accepted;         /* Solver->latestStepAccepted()                     */
  int32_t reserved;
} dcora_ropt_result;

void dcora_ropt_params_default(dcora_ropt_params *p);
const char *dcora_last_error(void);
/* Number of visible CUDA devices (0 when none / no driver). */
int dcora_device_count(void);

/* QuadraticProblem::QuadraticProblem(graph) (src/QuadraticProblem.cpp:19-34)
 * for `num_agents` graphs of common rank r and dimension d.  Uploads Q (as
 * BSR of (d+1)x(d+1) pose blocks + a scalar CSR remainder), G, and -- when
 * precond_reg >= 0 -- builds chol(Q_a + reg*I) on the host (own sparse
 * Cholesky; replaces Eigen::CholmodDecomposition, src/Graph.cpp:1910-1911)
 * and uploads the level-scheduled factor. */
int dcora_problem_create(const dcora_agent_desc *agents, int32_t num_agents,
                         int32_t r, int32_t d, int32_t order, int32_t device,
                         dcora_problem **out);
void dcora_problem_destroy(dcora_problem *p);

/* Same Q / factor at another relaxation rank (the staircase re-creates the
 * Graph per rank but Q does not depend on r: SURVEY 8(a17)).  Shares the
 * device matrices of `src`. */
int dcora_problem_clone_rank(const dcora_problem *src, int32_t r,
                             dcora_problem **out);

int64_t dcora_problem_dim(const dcora_problem *p);      /* K = sum_a k_a     */
int32_t dcora_problem_rank(const dcora_problem *p);
int32_t dcora_problem_num_agents(const dcora_problem *p);
int64_t dcora_problem_nnz(const dcora_problem *p);      /* nnz(Q) as given   */
int64_t dcora_problem_factor_nnz(const dcora_problem *p);
/* Regularisation actually used for agent a's factor (resolves
 * DCORA_REG_AUTO_RA); < 0 when the agent has no factor. */
double dcora_problem_precond_reg(const dcora_problem *p, int32_t agent);
/* cudaStream_t the handle launches on (for CUDA-event timing by callers). */
void *dcora_problem_stream(const dcora_problem *p);
/* Kernel launches issued by this handle since creation. */
int64_t dcora_problem_launch_count(const dcora_problem *p);

/* Graph::setNeighborStates -> constructG (src/Graph.cpp:333-340,565-577):
 * replace the linear term.  G: r x K, or NULL for zero. */
int dcora_problem_set_G(dcora_problem *p, const double *G);

/* QuadraticProblem::f (src/QuadraticProblem.cpp:38-51): f_out[num_agents]. */
int dcora_problem_f(dcora_problem *p, const double *X, double *f_out);
/* QuadraticProblem::EucGrad (:53-59): EG = X*Q + G. */
int dcora_problem_egrad(dcora_problem *p, const double *X, double *EG);
/* QuadraticProblem::EucHessianEta (:61-68): HV = V*Q. */
int dcora_problem_hessvec(dcora_problem *p, const double *V, double *HV);
/* QuadraticProblem::PreConditioner (:70-84): proj_X((Q+reg I)^-1 V^T)^T.
 * X == NULL: the plain solve (Q+reg I)^-1 V without the tangent projection
 * (the CHOLMOD solve of :79 alone; used as a block-Jacobi preconditioner by the
 * distributed certificate). */
int dcora_problem_precond(dcora_problem *p, const double *X, const double *V,
                          double *OUT);
/* QuadraticProblem::RieGrad / RieGradNorm (:86-123). */
int dcora_problem_rgrad(dcora_problem *p, const double *X, double *RG);
int dcora_problem_rgradnorm(dcora_problem *p, const double *X, double *norm_out);
/* QuadraticProblem::Retract (:125-136,236-259). */
int dcora_problem_retract(dcora_problem *p, const double *X, const double *V,
                          double *Xout, int32_t retraction);
/* LiftedSEManifold/LiftedRAManifold::projectToTangentSpace
 * (src/manifold/LiftedManifold.cpp:37-41,104-108). */
int dcora_problem_project_tangent(dcora_problem *p, const double *X,
                                  const double *V, double *OUT);
/* LiftedSEManifold/LiftedRAManifold::project = projectToSEMatrix /
 * projectToRAMatrix (src/DCORA_utils.cpp:2201-2220). */
int dcora_problem_project_manifold(dcora_problem *p, const double *M,
                                   double *OUT);

/* QuadraticOptimizer::optimize (src/QuadraticOptimizer.cpp:28-50) for every
 * agent of the batch at once; active[a]==0 leaves agent a untouched
 * (active==NULL: all).  results[num_agents]. */
int dcora_rtr_solve(dcora_problem *p, const double *X0, double *Xopt,
                    const dcora_ropt_params *params, const uint8_t *active,
                    dcora_ropt_result *results);

/* Certification of a CENTRALISED problem (num_agents == 1):
 * constructDualCertificateMatrix{PGO,RASLAM} + fastVerification
 * (src/DCORA_utils.cpp:1713-1735,1898-1982).  S = Q - Lambda is applied
 * implicitly; lambda_min(S) comes from a device thick-restart Lanczos
 * (replaces CHOLMOD's PSD test and Spectra, :1737-1896).
 *   is_psd     : 1 iff lambda_min(S + eta I) > 0 (the reference's verdict)
 *   lambda_min : smallest eigenvalue of S
 *   theta      : v^T S v   (fastVerification's *theta)
 *   v          : k doubles, reference ordering (may be NULL) */
typedef struct {
  int32_t is_psd;
  int32_t converged;
  int32_t matvecs;
  int32_t restarts;
  double lambda_min;
  double theta;
  double residual;
} dcora_certify_result;
int dcora_certify(dcora_problem *p, const double *X, double eta,
                  int32_t max_matvecs, double tol, dcora_certify_result *res,
                  double *v);

/* QuadraticProblem::escapeSaddle (src/QuadraticProblem.cpp:138-234).  `p` is
 * the rank-r problem, Xopt is (r-1) x k, X is r x k.  Returns DCORA_SOFT_FAIL
 * when no trial point decreases the cost (the reference returns false). */
int dcora_problem_escape_saddle(dcora_problem *p, const double *Xopt,
                                double theta, const double *v,
                                double gradient_tolerance,
                                double preconditioned_gradient_tolerance,
                                int32_t is_second_order, int32_t retraction,
                                double *X);

/* ---- multi-agent RBCD (Agent::iterate / updateX, src/Agent.cpp:535-596,
 * 1216-1278; driver loop examples/MultiRobotExample.cpp:223-307) ---------- */

/* One block of neighbour public states received from another process: a
 * pseudo-agent of n poses, l unit spheres and b landmarks, columns in the same
 * ordering as the agents (PoseDict / UnitSphereDict / LandmarkDict of
 * Agent::getSharedStateDicts, src/Agent.cpp:113-152, as a packed matrix). */
typedef struct {
  int32_t n, l, b;
} dcora_halo_segment;

/* Register the cross-agent coupling C, a (K + KH) x K CSR matrix holding the
 * blocks of the global data matrix that couple variables of different agents:
 * rows are [the K columns of this handle's agents | the halo segments],
 * columns are the K local columns.  G = [X | H] * C then equals what
 * Graph::constructLinearCostTerm{PGO,RASLAM} (src/Graph.cpp:685-822,
 * 1190-1772) accumulates edge by edge from the neighbour dictionaries.
 * nseg may be 0 (all agents of the problem live in this handle). */
int dcora_problem_set_coupling(dcora_problem *p, const dcora_halo_segment *segs,
                               int32_t nseg, const int32_t *rowptr,
                               const int32_t *col, const double *val);
/* The same with an accumulation order: row_key[K + KH] gives every row of C (every source
 * column of [X | H]) a key, and the contributions to one column of G are summed in ascending
 * key order (NULL: ascending row).  With the GLOBAL column number of each source as key the
 * sum does not depend on how the agents are dealt to processes: the iterates of an N-rank
 * run are bit-identical to the 1-rank run's. */
int dcora_problem_set_coupling_keyed(dcora_problem *p, const dcora_halo_segment *segs,
                                     int32_t nseg, const int32_t *rowptr,
                                     const int32_t *col, const double *val,
                                     const int64_t *row_key);
int64_t dcora_problem_halo_dim(const dcora_problem *p); /* KH */

/* Graph::setNeighborStates + constructG (src/Graph.cpp:333-340,565-577) for
 * the agents with active[a] != 0 (NULL: all): G_a = ([X | H] * C)_a.
 * X: r x K current iterates of the local agents; H: r x KH halo (may be NULL
 * when KH == 0). */
int dcora_problem_update_G(dcora_problem *p, const double *X, const double *H,
                           const uint8_t *active);

/* Agent::getSharedStateDicts as a gather: out[:, j] = X[:, idx[j]], idx are
 * reference column indices into the r x K batch (host array, cached on the
 * device on first use / when it changes). out: r x count. */
int dcora_problem_pack_columns(dcora_problem *p, const double *X,
                               const int32_t *idx, int64_t count, double *out);

/* Iterate tracking for drivers that keep ONE device-resident iterate buffer and
 * change it only through this library (the RBCD loop: Agent::X,
 * include/DCORA/Agent.h:698).  While tracking is on, the handle remembers the
 * device pointer whose content equals its internal copy of the iterate: calls
 * that receive that same pointer (update_G, rtr_solve, cost_terms, rgradnorm)
 * skip the layout conversion of their input, and an in-place rtr_solve writes
 * back only the columns of the agents it updated.  The caller must call this
 * function again (on = 1 re-arms, forgetting the pointer) whenever it has
 * written to the buffer itself.  Needs a registered coupling; off by default. */
int dcora_problem_set_iterate_tracking(dcora_problem *p, int32_t on);

/* Moving the device-resident iterate in and out: a driver whose Agent::X lives in host
 * memory between RBCD iterations (setX / getX, include/DCORA/Agent.h:583,596) uploads it
 * once per iteration into its device buffer and reads the result back once.  Both are
 * r x K column-major doubles in the reference ordering; the copies run on the handle's
 * stream and are complete on return.  load also forgets any tracked pointer (the buffer's
 * content changed). */
int dcora_problem_load_iterate(dcora_problem *p, const double *X_host, double *X_device);
int dcora_problem_store_iterate(dcora_problem *p, const double *X_device, double *X_host);

/* Rounding (src/DCORA_utils.cpp).
 *
 * dcora_project_to_rotation_group: projectToRotationGroup (:1661-1675), host, d x d
 * column-major in and out (d = 2 or 3).
 *
 * dcora_problem_round_solution: projectSolutionRASLAM (:1984-2031) for
 * the whole batch treated as one solution: X is r x K (reference ordering, host or device),
 * Xd is d x K in the same ordering.  Rank-d truncation of X through the r x r Gram matrix,
 * reflection vote over the rotation blocks, rotations to SO(d), unit spheres normalised.
 * The truncation is unique up to the sign of each of its d rows; the vote removes the
 * improper part, so two implementations agree up to diag(+-1) of determinant +1.
 *
 * dcora_problem_align_trajectory: alignLiftedTrajectoryToFrame (:2262-2289).  Tw0 is the
 * r x (d+1) lifted frame [R0 t0] column-major; Td (d x K) = R0^T X with every rotation block
 * projected to SO(d) and R0^T ta subtracted from every translation (and landmark), ta = t0
 * (is_global) or the first translation of R0^T X (local; needs r == d as in the reference). */
int dcora_project_to_rotation_group(const double *M, int32_t d, double *R);
int dcora_problem_round_solution(dcora_problem *p, const double *X, double *Xd);
int dcora_problem_align_trajectory(dcora_problem *p, const double *X, const double *Tw0,
                                   int32_t is_global, double *Td);

/* Certificate of the PARTITIONED problem (constructDualCertificateMatrix{PGO,RASLAM},
 * src/DCORA_utils.cpp:1898-1982, applied implicitly and blockwise).
 * dcora_problem_certificate_prepare: Lambda of this handle's agents from the iterate X and
 * XQ = (X * Q_global) restricted to the handle's columns (dcora_problem_update_G with the halo
 * + dcora_problem_egrad give it): sym(Y_i^T (XQ)_i) per rotation block, <r_k, (XQ)_k> per unit
 * sphere; kept on the handle until the next prepare / certify / RTR call.
 * dcora_problem_certificate_apply: OUT = QV - V * Lambda for a block of r vectors (r x K like an
 * iterate), i.e. S V on the handle's columns given QV = V * Q_global. */
int dcora_problem_certificate_prepare(dcora_problem *p, const double *X, const double *XQ);
int dcora_problem_certificate_apply(dcora_problem *p, const double *V, const double *QV,
                                    double *OUT);

/* Per-agent 0.5 <X Q, X> and <X, G>: the global cost of the partitioned
 * problem is sum_a quad_a + 0.5 lin_a (cross terms appear in two agents). */
int dcora_problem_cost_terms(dcora_problem *p, const double *X, double *quad_out,
                             double *lin_out);

/* Nesterov-accelerated RBCD++ (Agent::updateY / updateV, src/Agent.cpp:
 * 1202-1214): OUT = projectToManifold(a * A + b * B) for the active agents
 * (others: OUT = A). */
int dcora_problem_combine_project(dcora_problem *p, double a, const double *A,
                                  double b, const double *B, const uint8_t *active,
                                  double *OUT);
/* OUT = a * A + b * B (no projection; the X - Y term of Agent::updateV). */
int dcora_problem_combine(dcora_problem *p, double a, const double *A, double b,
                          const double *B, double *OUT);

/* K1 in isolation, for roofline measurement: runs the Q-SpMM `iters` times on
 * device-resident data and returns the mean kernel time (CUDA events on the
 * handle's stream) and the algorithmic bytes of one launch
 * (8*nnz(Q) + 16*r*K). */
int dcora_problem_time_spmm(dcora_problem *p, int32_t iters, double *ms_mean,
                            double *algorithmic_bytes);
/* The same for every hot-path kernel class: mean duration of one application
 * on internally generated device-resident data (all agents active) and its
 * algorithmic bytes as defined in SURVEY 8d / DESIGN.md:
 *   SPMM     8 nnz(Q) + 16 rK            PRECOND  2 (8 nnz(L) + 16 rK) + 24 rK
 *   PROJECT  24 rK   RETRACT 24 rK   DOT 16 rK   AXPBY 24 rK   RHESS 40 rK */
#define DCORA_KERNEL_SPMM 0
#define DCORA_KERNEL_PRECOND 1
#define DCORA_KERNEL_PROJECT 2
#define DCORA_KERNEL_RETRACT 3
#define DCORA_KERNEL_DOT 4
#define DCORA_KERNEL_AXPBY 5
#define DCORA_KERNEL_RHESS 6
int dcora_problem_time_kernel(dcora_problem *p, int32_t kind, int32_t iters,
                              double *ms_mean, double *algorithmic_bytes);

/* Profiling aid: runs the preconditioner kernel (K5) with a slot trace for one
 * agent and returns the %globaltimer stamps (ns) of its first CTA: stamps[0] =
 * start, stamps[1] = right-hand side loaded, then one stamp per finished barrier
 * slot (forward slots first), one after the write-back, terminated by 0.
 * `active` (num_agents bytes, NULL: every agent) selects the agents that run
 * concurrently with the traced one (which is always active). */
int dcora_debug_trsv_trace(dcora_problem *p, int32_t agent, const uint8_t *active,
                           uint64_t *stamps, int32_t max_stamps);

/* Warps the preconditioner solve (K5) of a large agent (>= 12000 columns) is dealt to by
 * handles created AFTER this call: 64 (default: clusters of 4 CTAs), 96 (6 CTAs), 128 (8 CTAs)
 * or 256 (16 CTAs, non-portable cluster size); more than 64 serves ranks <= 8 only.  The solve
 * is a latency chain per agent, so when few agents are solved concurrently (small colour
 * classes: several GPUs) more CTAs per agent shorten it while the clusters still fit the
 * machine in one wave (<= 24 agents at 96, <= 16 at 128, <= 8 at 256). */
void dcora_set_trsv_max_warps(int32_t warps);

/* Host-only self checks of the setup code (no GPU needed, no usable result):
 * max |Y_layout - Y_csr| of the CSR -> BSR/remainder conversion on internally
 * generated X, and the relative residual of the host Cholesky factor of
 * Q + reg*I on an internally generated right-hand side. */
int dcora_selfcheck_layout(const dcora_agent_desc *agents, int32_t num_agents,
                           int32_t r, int32_t d, int32_t order,
                           double *max_abs_err);
int dcora_selfcheck_cholesky_threads(int32_t k, const int32_t *rowptr, const int32_t *col,
                                     const double *val, int32_t dh, int32_t npose, double reg,
                                     int32_t threads, double *checksum, int64_t *factor_nnz);
int dcora_selfcheck_coupling(const dcora_agent_desc *agents, int32_t num_agents,
                             int32_t r, int32_t d, int32_t order,
                             const dcora_halo_segment *segs, int32_t nseg,
                             const int32_t *rowptr, const int32_t *col, const double *val,
                             const int64_t *row_key, double *max_abs_err,
                             int64_t *order_violations);
int dcora_selfcheck_cholesky(int32_t k, const int32_t *rowptr, const int32_t *col,
                             const double *val, int32_t dh, int32_t npose,
                             double reg, double *rel_residual,
                             int64_t *factor_nnz);

/* Host self check of the block triangular-solve schedule the preconditioner
 * kernel walks (nested-dissection factor regrouped into inverted diagonal
 * blocks + off-diagonal pose blocks, replayed sequentially on the host):
 * relative residual of (Q + reg I) x = b, nnz(L), values stored by the
 * schedule, number of diagonal blocks and block-tree depth per direction.
 * ordering: 0 = minimum degree (one block per supervariable), 1 = nested
 * dissection (what dcora_problem_create uses). */
int dcora_selfcheck_trsv(int32_t k, const int32_t *rowptr, const int32_t *col,
                         const double *val, int32_t dh, int32_t npose, double reg,
                         int32_t ordering, double *rel_residual, int64_t *factor_nnz,
                         int64_t *stored_values, int32_t *nblocks,
                         int32_t *levels_fwd, int32_t *levels_bwd);

#ifdef __cplusplus
}
#endif
#endif /* DCORA_B200_H_ */
