// Handle life-cycle, host<->device staging, and the QuadraticProblem-level
// entry points of the C-ABI (include/dcora_b200.h).
#include <omp.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <cstring>
#include <mutex>

#include "kernels.cuh"
#include "solver.cuh"
#include "trsv.h"

namespace dcora {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

ActiveSet::~ActiveSet() {
  cudaFree(agents); cudaFree(chunks); cudaFree(tiles); cudaFree(rem_rows); cudaFree(crem_rows);
  cudaFree(mask);
}

DeviceMatrices::~DeviceMatrices() {
  cudaSetDevice(device);
  cudaFree(perm); cudaFree(browptr); cudaFree(bcol); cudaFree(prowcol);
  cudaFree(pose_agent); cudaFree(rem_agent); cudaFree(bval);
  cudaFree(rem_row); cudaFree(rem_rowptr); cudaFree(rem_col);
  cudaFree(rem_assign); cudaFree(rem_val); cudaFree(chunks); cudaFree(agent_chunk0);
  cudaFree(chunk_pose0); cudaFree(chunk_rem);
  F.release();
}

template <typename T>
static cudaError_t upload(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src.empty()) return cudaSuccess;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static bool is_device_ptr(const void *ptr) {
  if (!ptr) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// Bring a reference-layout r x K matrix (host or device) into the internal
// layout vector `dst` (device).
int import_matrix(dcora_problem *p, const double *src, double *dst, int r) {
  DeviceMatrices &M = *p->M;
  const size_t bytes = (size_t)M.K * r * sizeof(double);
  const bool dev = is_device_ptr(src);
  if (M.identity_perm) {
    if (src != dst)
      DCORA_CUDA(cudaMemcpyAsync(dst, src, bytes,
                                 dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 p->stream));
    return DCORA_OK;
  }
  const double *s = src;
  if (!dev) {
    DCORA_CUDA(cudaMemcpyAsync(p->stage, src, bytes, cudaMemcpyHostToDevice, p->stream));
    s = p->stage;
  }
  return launch_permute(p, s, dst, r, true);
}

int export_matrix(dcora_problem *p, const double *src, double *dst, int r) {
  DeviceMatrices &M = *p->M;
  // any write into the tracked buffer other than the solver's own write-back
  // (which re-establishes it afterwards) ends the equality with p->xe
  if (dst == p->tracked) p->tracked = nullptr;
  const size_t bytes = (size_t)M.K * r * sizeof(double);
  const bool dev = is_device_ptr(dst);
  if (M.identity_perm) {
    if (src != dst)
      DCORA_CUDA(cudaMemcpyAsync(dst, src, bytes,
                                 dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                 p->stream));
  } else if (dev) {
    DCORA_TRY(launch_permute(p, src, dst, r, false));
  } else {
    DCORA_TRY(launch_permute(p, src, p->stage, r, false));
    DCORA_CUDA(cudaMemcpyAsync(dst, p->stage, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  // The C-ABI is synchronous: results are complete when a call returns, also
  // for device destinations (the caller's stream is not ours).
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

bool pointer_on_device(const void *ptr) { return is_device_ptr(ptr); }

int acquire_iterate(dcora_problem *p, const double *X, double *buf, double **out) {
  if (p->tracking && p->xe && is_device_ptr(X)) {
    if (X != p->tracked) {
      DCORA_TRY(import_matrix(p, X, p->xe, p->r));
      p->tracked = X;
    }
    *out = p->xe;
    return DCORA_OK;
  }
  if (!buf) return DCORA_ERR_ALLOC;
  // p->xe is about to hold something that is not the tracked buffer's content
  if (buf == p->xe) p->tracked = nullptr;
  DCORA_TRY(import_matrix(p, X, buf, p->r));
  *out = buf;
  return DCORA_OK;
}

int read_scalars(dcora_problem *p) {
  DCORA_CUDA(cudaMemcpyAsync(p->h_scal, p->scal, sizeof(double) * NS * p->M->A,
                             cudaMemcpyDeviceToHost, p->stream));
  DCORA_CUDA(cudaMemcpyAsync(p->h_flags, p->flags, sizeof(int) * NF * p->M->A,
                             cudaMemcpyDeviceToHost, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

static int alloc_handle_buffers(dcora_problem *p) {
  DeviceMatrices &M = *p->M;
  p->vec_len = (size_t)M.K * p->r;
  p->ws.assign(WS_COUNT, nullptr);
  DCORA_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  DCORA_CUDA(cudaMalloc(&p->G, std::max<size_t>(1, p->vec_len) * sizeof(double)));
  DCORA_CUDA(cudaMemsetAsync(p->G, 0, p->vec_len * sizeof(double), p->stream));
  if (!M.identity_perm)
    DCORA_CUDA(cudaMalloc(&p->stage, std::max<size_t>(1, p->vec_len) * sizeof(double)));
  DCORA_CUDA(cudaMalloc(&p->partials, std::max(1, M.nchunks) * NPART * sizeof(double)));
  DCORA_CUDA(cudaMalloc(&p->scal, sizeof(double) * NS * M.A));
  DCORA_CUDA(cudaMalloc(&p->flags, sizeof(int) * (NF * M.A + 4 + DCORA_COUNT_RING)));  // + counters
  DCORA_CUDA(cudaMallocHost(&p->h_scal, sizeof(double) * NS * M.A));
  DCORA_CUDA(cudaMallocHost(&p->h_flags, sizeof(int) * (NF * M.A + 4 + DCORA_COUNT_RING)));
  DCORA_CUDA(cudaMemsetAsync(p->scal, 0, sizeof(double) * NS * M.A, p->stream));
  std::vector<int> fl((size_t)NF * M.A + 4 + DCORA_COUNT_RING, 0);
  for (int a = 0; a < M.A; ++a) { fl[a * NF + F_ALL] = 1; fl[a * NF + F_USER] = 1; }
  DCORA_CUDA(cudaMemcpyAsync(p->flags, fl.data(), fl.size() * sizeof(int),
                             cudaMemcpyHostToDevice, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

}  // namespace dcora

using namespace dcora;

double *dcora_problem::vec(int i) {
  if (i < 0 || i >= (int)ws.size()) return nullptr;
  if (!ws[i]) {
    // zero-filled: the triangular-solve panels are read through (d+1)-row
    // windows whose unused rows meet zero coefficients and must be finite
    if (cudaMalloc(&ws[i], std::max<size_t>(1, vec_len) * sizeof(double)) != cudaSuccess ||
        cudaMemsetAsync(ws[i], 0, std::max<size_t>(1, vec_len) * sizeof(double), stream) != cudaSuccess) {
      set_error("workspace allocation failed");
      ws[i] = nullptr;
    }
  }
  return ws[i];
}

dcora_problem::~dcora_problem() {
  if (M) cudaSetDevice(M->device);
  for (double *w : ws) cudaFree(w);
  cudaFree(G); cudaFree(stage); cudaFree(xe); cudaFree(hstage); cudaFree(pack_idx); cudaFree(partials); cudaFree(scal); cudaFree(flags);
  cudaFreeHost(h_scal); cudaFreeHost(h_flags);
  for (cudaEvent_t e : count_ev) if (e) cudaEventDestroy(e);
  cudaFree(cert_ws);
  delete cert_q1;
  if (stream && owns_stream) cudaStreamDestroy(stream);
}

extern "C" {

const char *dcora_last_error(void) { return g_error.c_str(); }

int dcora_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void dcora_ropt_params_default(dcora_ropt_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->method = 0;
  p->verbose = 0;
  p->gradnorm_tol = 1e-2;
  p->RGD_stepsize = 1e-3;
  p->RGD_use_preconditioner = 1;
  p->RTR_iterations = 3;
  p->RTR_tCG_iterations = 50;
  p->RTR_initial_radius = 100;
  p->time_bound_s = 5.0;
  p->retraction = DCORA_RETRACT_QF;
}

int dcora_problem_create(const dcora_agent_desc *agents, int32_t num_agents,
                         int32_t r, int32_t d, int32_t order, int32_t device,
                         dcora_problem **out) {
  if (!out) return DCORA_ERR_ARG;
  *out = nullptr;
  if (dcora_device_count() <= device || device < 0) {
    set_error("no CUDA device " + std::to_string(device) +
              " (the dcora_b200 hot path has no CPU fallback)");
    return DCORA_ERR_CUDA;
  }
  DCORA_CUDA(cudaSetDevice(device));
  // DCORA_SETUP_TIMING=1: where the set-up time goes (stderr)
  const bool timing = std::getenv("DCORA_SETUP_TIMING") != nullptr;
  auto tprev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[dcora setup] %-28s %8.3f s\n", what,
                 std::chrono::duration<double>(now - tprev).count());
    tprev = now;
  };
  HostLayout H;
  DCORA_TRY(build_layout(agents, num_agents, r, d, order, &H));
  lap("layout (CSR -> BSR)");
  auto M = std::make_shared<DeviceMatrices>();
  M->device = device;
  M->d = d; M->order = order; M->A = H.A; M->np = H.np; M->K = H.K;
  M->nnz_in = H.nnz_in;
  M->nnzb = (int64_t)H.bcol.size();
  M->nrem_rows = (int64_t)H.rem_row.size();
  M->nrem = (int64_t)H.rem_col.size();
  M->nchunks = (int)H.chunks.size();
  M->coff = H.coff; M->n = H.n; M->l = H.l; M->b = H.b; M->poff = H.poff;
  M->agent_chunk0_h = H.agent_chunk0;
  M->identity_perm = H.identity_perm;
  std::vector<int> pose_agent(H.np), rem_agent(H.rem_row.size());
  for (int a = 0; a < H.A; ++a)
    for (int i = H.poff[a]; i < H.poff[a + 1]; ++i) pose_agent[i] = a;
  {
    int a = 0;
    for (size_t s = 0; s < H.rem_row.size(); ++s) {
      while (H.rem_row[s] >= H.coff[a + 1]) ++a;
      rem_agent[s] = a;
    }
  }
  DCORA_CUDA(upload(&M->perm, H.perm));
  DCORA_CUDA(upload(&M->browptr, H.browptr));
  DCORA_CUDA(upload(&M->bcol, H.bcol));
  DCORA_CUDA(upload(&M->bval, H.bval));
  DCORA_CUDA(upload(&M->prowcol, H.prowcol));
  DCORA_CUDA(upload(&M->pose_agent, pose_agent));
  DCORA_CUDA(upload(&M->rem_agent, rem_agent));
  M->rem_agent_h = rem_agent;
  DCORA_CUDA(upload(&M->rem_row, H.rem_row));
  DCORA_CUDA(upload(&M->rem_rowptr, H.rem_rowptr));
  DCORA_CUDA(upload(&M->rem_col, H.rem_col));
  DCORA_CUDA(upload(&M->rem_val, H.rem_val));
  DCORA_CUDA(upload(&M->rem_assign, H.rem_assign));
  DCORA_CUDA(upload(&M->chunks, H.chunks));
  DCORA_CUDA(upload(&M->agent_chunk0, H.agent_chunk0));
  {
    // per chunk: its pose block rows and its remainder rows (both are sorted by column)
    std::vector<int> pose0(H.chunks.size(), -1);
    std::vector<int2> remr(H.chunks.size());
    size_t s = 0;
    for (size_t c = 0; c < H.chunks.size(); ++c) {
      const Chunk &ch = H.chunks[c];
      if (ch.kind == KIND_POSE)
        pose0[c] = H.poff[ch.agent] + (ch.col0 - (int)H.coff[ch.agent]) / (d + 1);
      while (s < H.rem_row.size() && H.rem_row[s] < ch.col0) ++s;
      remr[c].x = (int)s;
      while (s < H.rem_row.size() && H.rem_row[s] < ch.col0 + ch.ncols) ++s;
      remr[c].y = (int)s;
    }
    DCORA_CUDA(upload(&M->chunk_pose0, pose0));
    DCORA_CUDA(upload(&M->chunk_rem, remr));
  }

  // ---- centralised problems keep Q on the host (internal ordering): the exact PSD test
  // of the certificate (dcora_certify) factors S + eta I = Q - Lambda + eta I
  if (num_agents == 1) {
    const dcora_agent_desc &g = agents[0];
    const int k = (d + 1) * g.n + g.l + g.b;
    std::vector<int> ref2int(k);
    for (int c = 0; c < k; ++c) ref2int[H.perm[c]] = c;
    M->hostQ_ptr.assign(k + 1, 0);
    for (int R = 0; R < k; ++R) M->hostQ_ptr[ref2int[R] + 1] = g.rowptr[R + 1] - g.rowptr[R];
    for (int i = 0; i < k; ++i) M->hostQ_ptr[i + 1] += M->hostQ_ptr[i];
    M->hostQ_col.resize(M->hostQ_ptr[k]);
    M->hostQ_val.resize(M->hostQ_ptr[k]);
    for (int R = 0; R < k; ++R) {
      int w = M->hostQ_ptr[ref2int[R]];
      for (int q = g.rowptr[R]; q < g.rowptr[R + 1]; ++q) {
        M->hostQ_col[w] = ref2int[g.col[q]];
        M->hostQ_val[w] = g.val[q];
        ++w;
      }
    }
  }

  lap("upload of Q");
  // ---- preconditioner factors (host) -> device
  bool any = false;
  for (int a = 0; a < num_agents; ++a)
    any = any || agents[a].precond_reg >= 0.0 || agents[a].precond_reg == DCORA_REG_AUTO_RA;
  M->reg.assign(num_agents, -1.0);
  if (any) {
    std::vector<HostFactor> HF(num_agents);
    std::vector<AgentTape> AT(num_agents);
    // launch cluster = the largest agent's; every agent's own split depends
    // only on its size
    int tape_warps = 8;
    for (int a = 0; a < num_agents; ++a)
      tape_warps = std::max(tape_warps, trsv_warps_for((double)(H.coff[a + 1] - H.coff[a])));
    std::vector<int> rcs(num_agents, DCORA_OK);
    const int dh = d + 1;
    // few large agents: the spare cores factor independent elimination-tree subtrees of an agent
    const int chol_threads = std::max(1, setup_threads() / std::max(1, num_agents));
#pragma omp parallel for schedule(dynamic, 1) num_threads(setup_threads())
    for (int a = 0; a < num_agents; ++a) {
      const dcora_agent_desc &g = agents[a];
      const int k = dh * g.n + g.l + g.b;
      // P = Q + reg I in the agent-local internal ordering (full symmetric CSR)
      const int c0 = (int)H.coff[a];
      std::vector<int> ref2int(k);
      for (int c = 0; c < k; ++c) ref2int[H.perm[c0 + c] - c0] = c;
      // H.perm: internal -> reference; invert
      std::vector<int> Ap(k + 1, 0);
      for (int R = 0; R < k; ++R) Ap[ref2int[R] + 1] = g.rowptr[R + 1] - g.rowptr[R] + 1;
      for (int i = 0; i < k; ++i) Ap[i + 1] += Ap[i];
      std::vector<int> Ai(Ap[k]);
      std::vector<double> Ax(Ap[k]);
      double reg = g.precond_reg >= 0.0 ? g.precond_reg : 0.0;
      if (g.precond_reg == DCORA_REG_AUTO_RA) {
        const double lm = lambda_max_lanczos(k, g.rowptr, g.col, g.val, 1e-3, 300);
        reg = lm > 0.0 ? lm / (1e6 - 1.0) : 0.1;
      }
      M->reg[a] = reg;
      for (int R = 0; R < k; ++R) {
        const int ri = ref2int[R];
        int w = Ap[ri];
        bool diag = false;
        for (int q = g.rowptr[R]; q < g.rowptr[R + 1]; ++q) {
          const int ci = ref2int[g.col[q]];
          double v = g.val[q];
          if (ci == ri) { v += reg; diag = true; }
          Ai[w] = ci; Ax[w] = v; ++w;
        }
        if (!diag) { Ai[w] = ri; Ax[w] = reg; ++w; }
        // unused slot (diag present): mark with a harmless zero on the diagonal
        if (w < Ap[ri + 1]) { Ai[w] = ri; Ax[w] = 0.0; }
      }
      // symmetrise numerically: P <- (P + P^T)/2 is implied by using the upper
      // triangle only inside sparse_cholesky (as CHOLMOD does for symmetric
      // input).
      rcs[a] = sparse_cholesky(k, Ap, Ai, Ax, dh, g.n, &HF[a], DCORA_ORDERING_ND, TRSV_LEAF_ROWS,
                               chol_threads);
      if (rcs[a] == DCORA_OK) {
        rcs[a] = build_agent_tape(HF[a], dh, trsv_warps_for((double)k), tape_warps, c0, &AT[a]);
        // only nnz(L) and the permutation are needed from here on
        const size_t nnz = HF[a].Lx.size();
        std::vector<int>().swap(HF[a].Li);
        std::vector<double>().swap(HF[a].Lx);
        HF[a].Lx.resize(0);
        HF[a].nnz = (int64_t)nnz;
      }
    }
    for (int a = 0; a < num_agents; ++a)
      if (rcs[a] != DCORA_OK) {
        set_error("Cholesky factorisation of Q + reg*I failed for agent " + std::to_string(a));
        return DCORA_ERR_FACTOR;
      }
    lap("factorisation + tapes");
    DCORA_TRY(upload_factor(M.get(), HF, AT, tape_warps, H));
    lap("upload of the tapes");
  }

  dcora_problem *p = new dcora_problem();
  p->r = r;
  p->M = M;
  int rc = alloc_handle_buffers(p);
  if (rc != DCORA_OK) { delete p; return rc; }
  // linear term
  bool anyG = false;
  for (int a = 0; a < num_agents; ++a) anyG = anyG || agents[a].G != nullptr;
  if (anyG) {
    std::vector<double> Gh((size_t)H.K * r, 0.0);
    for (int a = 0; a < num_agents; ++a)
      if (agents[a].G)
        std::memcpy(&Gh[(size_t)H.coff[a] * r], agents[a].G,
                    sizeof(double) * (size_t)(H.coff[a + 1] - H.coff[a]) * r);
    rc = dcora_problem_set_G(p, Gh.data());
    if (rc != DCORA_OK) { delete p; return rc; }
  }
  *out = p;
  return DCORA_OK;
}

int dcora_problem_clone_rank(const dcora_problem *src, int32_t r, dcora_problem **out) {
  if (!src || !out || r < 1 || r > MAX_RANK) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(src->M->device));
  dcora_problem *p = new dcora_problem();
  p->r = r;
  p->M = src->M;
  p->C = src->C;
  int rc = alloc_handle_buffers(p);
  if (rc != DCORA_OK) { delete p; return rc; }
  *out = p;
  return DCORA_OK;
}

void dcora_problem_destroy(dcora_problem *p) { delete p; }

int64_t dcora_problem_dim(const dcora_problem *p) { return p ? p->M->K : 0; }
int32_t dcora_problem_rank(const dcora_problem *p) { return p ? p->r : 0; }
int32_t dcora_problem_num_agents(const dcora_problem *p) { return p ? p->M->A : 0; }
int64_t dcora_problem_nnz(const dcora_problem *p) { return p ? p->M->nnz_in : 0; }
int64_t dcora_problem_factor_nnz(const dcora_problem *p) { return p ? p->M->F.nnzL : 0; }
double dcora_problem_precond_reg(const dcora_problem *p, int32_t agent) {
  if (!p || agent < 0 || agent >= (int)p->M->reg.size()) return -1.0;
  return p->M->reg[agent];
}
void *dcora_problem_stream(const dcora_problem *p) { return p ? (void *)p->stream : nullptr; }
int64_t dcora_problem_launch_count(const dcora_problem *p) { return p ? p->launches : 0; }

#define DCORA_ENTER(p)                         \
  if (!(p)) return DCORA_ERR_ARG;              \
  DCORA_CUDA(cudaSetDevice((p)->M->device));

int dcora_problem_set_G(dcora_problem *p, const double *G) {
  DCORA_ENTER(p);
  if (!G) {
    DCORA_CUDA(cudaMemsetAsync(p->G, 0, p->vec_len * sizeof(double), p->stream));
    p->has_G = false;
    return DCORA_OK;
  }
  DCORA_TRY(import_matrix(p, G, p->G, p->r));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  p->has_G = true;
  return DCORA_OK;
}

// f_a = 0.5 <X Q, X> + <X, G> per agent -> scal slot `slot`
static int compute_f(dcora_problem *p, const double *X, double *tmp, int slot, int mask) {
  DCORA_TRY(launch_spmm(p, X, nullptr, tmp, mask));
  const double *xs[2] = {tmp, X};
  const double *ys[2] = {X, p->G};
  DCORA_TRY(launch_dots(p, p->has_G ? 2 : 1, xs, ys, mask));
  return launch_combine_f(p, slot, p->has_G ? 1 : 0, mask);
}

int dcora_problem_f(dcora_problem *p, const double *X, double *f_out) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1);
  if (!x || !t) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(compute_f(p, x, t, S_TMP0, F_ALL));
  DCORA_TRY(read_scalars(p));
  for (int a = 0; a < p->M->A; ++a) f_out[a] = p->h_scal[a * NS + S_TMP0];
  return DCORA_OK;
}

int dcora_problem_egrad(dcora_problem *p, const double *X, double *EG) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1);
  if (!x || !t) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(launch_spmm(p, x, p->has_G ? p->G : nullptr, t, -1));
  return export_matrix(p, t, EG, p->r);
}

int dcora_problem_hessvec(dcora_problem *p, const double *V, double *HV) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1);
  if (!x || !t) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, V, x, p->r));
  DCORA_TRY(launch_spmm(p, x, nullptr, t, -1));
  return export_matrix(p, t, HV, p->r);
}

int dcora_problem_precond(dcora_problem *p, const double *X, const double *V,
                          double *OUT) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *v = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !v || !o) return DCORA_ERR_ALLOC;
  if (X) DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(import_matrix(p, V, v, p->r));
  DCORA_TRY(launch_precond(p, X ? x : nullptr, v, o, -1));
  return export_matrix(p, o, OUT, p->r);
}

int dcora_problem_rgrad(dcora_problem *p, const double *X, double *RG) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !t || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(launch_spmm(p, x, p->has_G ? p->G : nullptr, t, -1));
  DCORA_TRY(launch_project_tangent(p, x, t, o, -1));
  return export_matrix(p, o, RG, p->r);
}

int dcora_problem_set_iterate_tracking(dcora_problem *p, int32_t on) {
  if (!p) return DCORA_ERR_ARG;
  if (on && !p->C) {
    set_error("dcora_problem_set_iterate_tracking: register the coupling first "
              "(dcora_problem_set_coupling)");
    return DCORA_ERR_ARG;
  }
  p->tracking = on != 0;
  p->tracked = nullptr;
  return DCORA_OK;
}

int dcora_problem_rgradnorm(dcora_problem *p, const double *X, double *norm_out) {
  DCORA_ENTER(p);
  double *x = nullptr, *t = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!t || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(acquire_iterate(p, X, p->vec(WS_IO0), &x));
  DCORA_TRY(launch_spmm(p, x, p->has_G ? p->G : nullptr, t, -1));
  DCORA_TRY(launch_project_tangent(p, x, t, o, -1));
  const double *xs[1] = {o};
  DCORA_TRY(launch_dots(p, 1, xs, xs, -1));
  const int dst[1] = {S_TMP0};
  DCORA_TRY(launch_reduce_scalars(p, 1, dst));
  DCORA_TRY(read_scalars(p));
  for (int a = 0; a < p->M->A; ++a) norm_out[a] = std::sqrt(p->h_scal[a * NS + S_TMP0]);
  return DCORA_OK;
}

int dcora_problem_retract(dcora_problem *p, const double *X, const double *V,
                          double *Xout, int32_t retraction) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *v = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !v || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(import_matrix(p, V, v, p->r));
  DCORA_TRY(launch_retract(p, x, v, o, retraction, -1));
  return export_matrix(p, o, Xout, p->r);
}

int dcora_problem_project_tangent(dcora_problem *p, const double *X, const double *V,
                                  double *OUT) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *v = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !v || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(import_matrix(p, V, v, p->r));
  DCORA_TRY(launch_project_tangent(p, x, v, o, -1));
  return export_matrix(p, o, OUT, p->r);
}

int dcora_problem_project_manifold(dcora_problem *p, const double *Mx, double *OUT) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *o = p->vec(WS_IO2);
  if (!x || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, Mx, x, p->r));
  DCORA_TRY(launch_project_manifold(p, x, o, -1));
  return export_matrix(p, o, OUT, p->r);
}

int dcora_problem_time_kernel(dcora_problem *p, int32_t kind, int32_t iters,
                              double *ms_mean, double *algorithmic_bytes) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1), *o = p->vec(WS_IO2), *s = p->vec(WS_S);
  if (!x || !t || !o || !s) return DCORA_ERR_ALLOC;
  const double rk = (double)p->r * (double)p->M->K;
  double bytes = 0.0;
  auto run = [&]() -> int {
    switch (kind) {
      case DCORA_KERNEL_SPMM:
        bytes = 8.0 * (double)p->M->nnz_in + 16.0 * rk;
        return launch_spmm(p, x, nullptr, t, -1);
      case DCORA_KERNEL_PRECOND:
        bytes = 2.0 * (8.0 * (double)p->M->F.nnzL + 16.0 * rk) + 24.0 * rk;
        return launch_precond(p, x, t, o, -1);
      case DCORA_KERNEL_PROJECT:
        bytes = 24.0 * rk;
        return launch_project_tangent(p, x, t, o, -1);
      case DCORA_KERNEL_RETRACT:
        bytes = 24.0 * rk;
        return launch_retract(p, x, t, o, DCORA_RETRACT_QF, -1);
      case DCORA_KERNEL_DOT: {
        bytes = 16.0 * rk;
        const double *xs[1] = {x}, *ys[1] = {t};
        return launch_dots(p, 1, xs, ys, -1);
      }
      case DCORA_KERNEL_AXPBY:
        bytes = 24.0 * rk;
        return launch_axpby(p, o, -1, 0.5, x, -1, 0.25, t, -1);
      case DCORA_KERNEL_RHESS:
        bytes = 40.0 * rk;
        return launch_rhess_correct(p, x, s, t, t, o, -1);
      default:
        set_error("dcora_problem_time_kernel: unknown kind");
        return DCORA_ERR_ARG;
    }
  };
  // a feasible point in x, something finite elsewhere
  DCORA_CUDA(cudaMemsetAsync(t, 0, p->vec_len * sizeof(double), p->stream));
  DCORA_CUDA(cudaMemsetAsync(s, 0, p->vec_len * sizeof(double), p->stream));
  DCORA_TRY(launch_axpby(p, t, -1, 0.0, t, -1, 0.0, nullptr, -1));
  if (!p->timing_ready) {
    // deterministic pseudo-random start, projected to the manifold
    std::vector<double> h(p->vec_len);
    uint64_t st = 0x9E3779B97F4A7C15ull;
    for (auto &v : h) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      v = (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
    DCORA_CUDA(cudaMemcpyAsync(t, h.data(), sizeof(double) * p->vec_len, cudaMemcpyHostToDevice,
                               p->stream));
    DCORA_CUDA(cudaStreamSynchronize(p->stream));
    DCORA_TRY(launch_project_manifold(p, t, x, -1));
    p->timing_ready = true;
  }
  DCORA_TRY(launch_copy(p, t, x, -1));
  cudaEvent_t e0, e1;
  DCORA_CUDA(cudaEventCreate(&e0));
  DCORA_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) DCORA_TRY(run());
  DCORA_CUDA(cudaEventRecord(e0, p->stream));
  for (int i = 0; i < iters; ++i) DCORA_TRY(run());
  DCORA_CUDA(cudaEventRecord(e1, p->stream));
  DCORA_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  DCORA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_mean) *ms_mean = (double)ms / std::max(1, iters);
  if (algorithmic_bytes) *algorithmic_bytes = bytes;
  return DCORA_OK;
}

int dcora_debug_trsv_trace(dcora_problem *p, int32_t agent, const uint8_t *active,
                           uint64_t *stamps, int32_t max_stamps) {
  DCORA_ENTER(p);
  double *x = p->vec(WS_IO0), *t = p->vec(WS_IO1), *o = p->vec(WS_IO2);
  if (!x || !t || !o || !stamps || max_stamps < 4 || agent < 0 || agent >= p->M->A)
    return DCORA_ERR_ARG;
  if (!p->timing_ready) {
    double ms, by;
    DCORA_TRY(dcora_problem_time_kernel(p, DCORA_KERNEL_PROJECT, 1, &ms, &by));
  }
  unsigned long long *d = nullptr;
  DCORA_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * max_stamps));
  DCORA_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * max_stamps, p->stream));
  int mask = -1;
  if (active) {
    std::vector<uint8_t> act(active, active + p->M->A);
    act[agent] = 1;
    DCORA_TRY(set_user_mask(p, act.data()));
    mask = F_USER;
  }
  for (int it = 0; it < 3; ++it) DCORA_TRY(launch_precond(p, x, t, o, mask, d, agent));
  DCORA_CUDA(cudaMemcpyAsync(stamps, d, sizeof(unsigned long long) * max_stamps,
                             cudaMemcpyDeviceToHost, p->stream));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  cudaFree(d);
  if (active) DCORA_TRY(set_user_mask(p, nullptr));
  return DCORA_OK;
}

int dcora_problem_time_spmm(dcora_problem *p, int32_t iters, double *ms_mean,
                            double *algorithmic_bytes) {
  return dcora_problem_time_kernel(p, DCORA_KERNEL_SPMM, iters, ms_mean, algorithmic_bytes);
}

}  // extern "C"
