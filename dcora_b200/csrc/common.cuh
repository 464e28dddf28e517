// dcora_b200 internal definitions shared by the .cu / .cpp translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dcora_b200.h"

namespace dcora {

constexpr int MAX_RANK = 16;      // templated ranks 1..MAX_RANK
constexpr int POSE_CHUNK = 64;    // poses per chunk (pose region)
constexpr int EXTRA_CHUNK = 512;  // columns per chunk (sphere / landmark region)
constexpr int NUM_SMS = 148;      // B200
constexpr int DCORA_COUNT_RING = 8;  // slots of the solver's lagged "agents still iterating" counts

void set_error(const std::string &msg);
// Host threads of the set-up code (layout conversion, factorisation, tapes): the machine's
// cores divided by the ranks on this node (LOCAL_WORLD_SIZE), NOT OMP_NUM_THREADS -- torchrun
// exports OMP_NUM_THREADS=1, which would serialise the per-agent factorisations.  Override
// with DCORA_SETUP_THREADS.
int setup_threads();

#define DCORA_CUDA(call)                                                        \
  do {                                                                          \
    cudaError_t err__ = (call);                                                 \
    if (err__ != cudaSuccess) {                                                 \
      ::dcora::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + \
                         " at " + __FILE__ + ":" + std::to_string(__LINE__));   \
      return DCORA_ERR_CUDA;                                                    \
    }                                                                           \
  } while (0)

#define DCORA_TRY(call)            \
  do {                             \
    int rc__ = (call);             \
    if (rc__ < 0) return rc__;     \
  } while (0)

// A chunk is a run of internal columns that belongs to one agent and one
// manifold factor kind; every vector kernel is launched one CTA per chunk so
// that per-agent reductions are deterministic and per-agent masks are cheap.
enum ChunkKind : int { KIND_POSE = 0, KIND_SPHERE = 1, KIND_EUCLID = 2 };
struct Chunk {
  int col0;   // first internal column
  int ncols;  // number of columns (pose chunks: multiple of d+1)
  int kind;
  int agent;
};

// Host-side result of converting the agents' CSR matrices (reference
// ordering) to the internal device layout.  See layout.cpp.
struct HostLayout {
  int r = 0, d = 0, order = 0, A = 0;
  int64_t K = 0;           // total columns
  int64_t nnz_in = 0;      // nnz(Q) as given
  std::vector<int64_t> coff;  // A+1 column offsets (internal == reference)
  std::vector<int> n, l, b;   // per agent
  std::vector<int> poff;      // A+1 pose offsets
  bool identity_perm = true;
  std::vector<int> perm;      // internal col -> reference col (global), size K
  // BSR of Q^T over pose blocks (row = output pose block)
  int np = 0;
  std::vector<int> browptr, bcol;  // bcol = internal column start of block col
  std::vector<double> bval;        // [a][b] row-major, a = output col in block
  std::vector<int> prowcol;        // np: internal col start of pose block row
  // scalar remainder (entries touching a sphere / landmark row or column)
  std::vector<int> rem_row, rem_rowptr, rem_col;
  std::vector<uint8_t> rem_assign;  // 1: Y[row] = sum, 0: Y[row] += sum
  std::vector<double> rem_val;
  std::vector<Chunk> chunks;
  std::vector<int> agent_chunk0;  // A+1
};

int build_layout(const dcora_agent_desc *agents, int A, int r, int d, int order,
                 HostLayout *out);

// Host-side converted coupling matrix (layout.cpp: build_coupling_layout).
struct HostCoupling {
  int64_t KH = 0;              // halo columns
  bool identity_perm = true;
  std::vector<int> perm_h;     // internal halo col -> reference halo col
  std::vector<int> browptr, bcol, prowcol;
  std::vector<double> bval;
  std::vector<int> rem_row, rem_rowptr, rem_col;
  std::vector<uint8_t> rem_assign;
  std::vector<double> rem_val;
};
int build_coupling_layout(int d, int order, const std::vector<int64_t> &coff,
                          const std::vector<int> &n, const std::vector<int> &l,
                          const std::vector<int> &poff, const dcora_halo_segment *segs,
                          int nseg, const int64_t *rowptr, const int32_t *col,
                          const double *val, const int64_t *row_key, HostCoupling *out);

// Device copy of the coupling (rbcd.cu).
struct DeviceCoupling {
  int device = 0;
  int64_t KH = 0, nnzb = 0, nrem_rows = 0, nnz_in = 0;
  std::vector<int> rem_agent_h;
  bool identity_perm = true;
  int *perm_h = nullptr;
  int *browptr = nullptr, *bcol = nullptr;
  double *bval = nullptr;
  int *rem_row = nullptr, *rem_rowptr = nullptr, *rem_col = nullptr, *rem_agent = nullptr;
  uint8_t *rem_assign = nullptr;
  double *rem_val = nullptr;
  ~DeviceCoupling();
};

// Host sparse Cholesky factor of one agent (chol.cpp), internal ordering.
constexpr int DCORA_ORDERING_MD = 0;  // minimum degree, one block per supervariable
constexpr int DCORA_ORDERING_ND = 1;  // nested dissection, blocks = subdomains / separators
constexpr int TRSV_LEAF_ROWS = 32;    // ND recursion stops at subdomains of <= 32 rows
struct HostFactor {
  int k = 0;
  std::vector<int> perm;            // new -> old (internal col within agent)
  std::vector<int> Lp, Li;          // CSC of L (k+1, nnz), rows sorted
  std::vector<double> Lx;
  int64_t nnz = 0;                  // nnz(L) (kept after Li / Lx are released)
  std::vector<int> blk_ptr;         // factor-row boundaries of the diagonal blocks
  std::vector<int> pose_start;      // k: first factor row of the row's pose block, -1 if scalar
};
// A given as full symmetric CSR (internal ordering, agent-local); `dh` and
// `npose` describe the supervariable structure used by the ordering.
int sparse_cholesky(int k, const std::vector<int> &Ap, const std::vector<int> &Ai,
                    const std::vector<double> &Ax, int dh, int npose,
                    HostFactor *out, int ordering = DCORA_ORDERING_ND,
                    int leaf_rows = TRSV_LEAF_ROWS, int threads = 1);

double lambda_max_lanczos(int k, const int *Ap, const int *Ai, const double *Ax,
                          double tol, int max_steps);

// Device-side block triangular-solve tapes for the whole batch (sptrsv.cu; the
// tape format is documented in trsv.h).
struct DeviceFactor {
  bool present = false;
  int64_t nnzL = 0;           // nnz of the scalar Cholesky factors
  int64_t stored_values = 0;  // factor values one direction of the tapes stores
  int64_t tape_bytes = 0;     // both directions, with headers and padding
  int A = 0;
  int nwarps = 8;             // warps per agent (8, 16, 32 or 64): cluster shape at launch
  int max_levels = 0;
  unsigned long long *tape = nullptr;  // all tapes, 8-byte words
  long long *tape_off = nullptr;       // A * nwarps: first word of each warp's tape
  int *tape_chunks = nullptr;          // A * nwarps: length in records
  int *nslots = nullptr;               // A: barrier slots of every agent's tapes
  long long *arow0 = nullptr;          // A + 1: first panel row of every agent
  int *frow_col = nullptr;             // K: internal column of every factor row (global)
  int *col_frow = nullptr;             // K: factor row of every internal column (global)
  int max_slots = 0;
  // cost model of the schedule (worst agent): busiest-warp iterations / rounds summed over
  // the slots, and the perfectly balanced iteration count
  int64_t crit_iters = 0, crit_rounds = 0, ideal_iters = 0;
  void release();
};

// Compacted launch lists of a subset of the batch's agents (one colour of an
// RBCD sweep): masked kernels are launched over these lists instead of over
// the whole batch, so a colour step costs what its agents cost.
struct ActiveSet {
  int n_agents = 0, n_chunks = 0, n_tiles = 0, n_rem = 0, n_crem = 0;
  int *agents = nullptr;      // agent ids
  int *chunks = nullptr;      // chunk ids
  int2 *tiles = nullptr;      // SpMM tiles: [first, end) pose block rows
  int *rem_rows = nullptr;    // remainder rows of Q
  int *crem_rows = nullptr;   // remainder rows of the coupling matrix
  uint8_t *mask = nullptr;    // A bytes: the subset itself
  const void *coupling_of = nullptr;  // coupling the crem list was built for
  ~ActiveSet();
};

struct DeviceMatrices {
  int device = 0;
  int d = 0, order = 0, A = 0, np = 0;
  int64_t K = 0, nnz_in = 0, nnzb = 0, nrem_rows = 0, nrem = 0;
  int nchunks = 0;
  std::vector<int64_t> coff;
  std::vector<int> n, l, b, poff, agent_chunk0_h, rem_agent_h;
  std::vector<double> reg;  // per agent: regularisation of the factor (<0: none)
  // single-agent handles: Q as full symmetric CSR in the internal ordering (exact PSD test)
  std::vector<int> hostQ_ptr, hostQ_col;
  std::vector<double> hostQ_val;
  bool identity_perm = true;
  int *perm = nullptr;
  int *browptr = nullptr, *bcol = nullptr, *prowcol = nullptr;
  int *pose_agent = nullptr;  // np: agent of each pose block row
  int *rem_agent = nullptr;   // nrem_rows: agent of each remainder row
  double *bval = nullptr;
  int *rem_row = nullptr, *rem_rowptr = nullptr, *rem_col = nullptr;
  uint8_t *rem_assign = nullptr;
  double *rem_val = nullptr;
  Chunk *chunks = nullptr;
  int *agent_chunk0 = nullptr;
  int *chunk_pose0 = nullptr;   // nchunks: first pose block row of a pose chunk (-1 otherwise)
  int2 *chunk_rem = nullptr;    // nchunks: [first, end) remainder rows with their row in the chunk
  DeviceFactor F;
  ~DeviceMatrices();
};

}  // namespace dcora

// The opaque handle.
struct dcora_problem {
  int r = 0;
  std::shared_ptr<dcora::DeviceMatrices> M;
  std::shared_ptr<dcora::DeviceCoupling> C;  // cross-agent coupling (RBCD), optional
  double *xe = nullptr;            // [X | H] extended vector, r*(K+KH)
  // iterate tracking (dcora_problem_set_iterate_tracking): `tracked` is the
  // caller's device buffer whose content equals the first K columns of xe
  bool tracking = false;
  const double *tracked = nullptr;
  double *hstage = nullptr;        // halo staging (reference layout)
  // launch lists of the caller's current agent subset (null: whole batch)
  std::shared_ptr<dcora::ActiveSet> act;
  std::map<std::string, std::shared_ptr<dcora::ActiveSet>> act_cache;
  int *pack_idx = nullptr;         // cached public-column index list (device)
  std::vector<int32_t> pack_idx_host;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  int64_t launches = 0;
  size_t vec_len = 0;              // r*K doubles
  double *G = nullptr;             // internal layout (always allocated, zero if none)
  bool has_G = false;
  bool timing_ready = false;       // WS_IO0 holds a feasible point for time_kernel
  std::vector<double *> ws;        // workspace vectors (internal layout)
  double *stage = nullptr;         // staging vector (reference layout)
  double *partials = nullptr;      // nchunks * NPART doubles
  double *scal = nullptr;          // per-agent scalar block (device)
  double *h_scal = nullptr;        // pinned mirror
  int *flags = nullptr;            // per-agent int flags (device)
  int *h_flags = nullptr;          // pinned mirror
  cudaEvent_t count_ev[dcora::DCORA_COUNT_RING] = {};  // read-back events of the count ring (solver.cu)
  // certificate workspace, kept across dcora_certify calls (certify.cu): one arena for the
  // Krylov bases and the rank-1 twin of this handle that applies Q to single vectors
  double *cert_ws = nullptr;
  size_t cert_ws_len = 0;
  dcora_problem *cert_q1 = nullptr;
  double *vec(int i);              // lazily allocated workspace vector i
  ~dcora_problem();
};
