// Host-only self checks of the setup code (layout conversion, sparse
// Cholesky).  They return an error norm on internally generated data; they do
// not produce results a caller could use instead of the GPU path.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "common.cuh"

using namespace dcora;

static inline double prand(uint64_t &st) {
  st ^= st << 13; st ^= st >> 7; st ^= st << 17;
  return (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

extern "C" int dcora_selfcheck_layout(const dcora_agent_desc *agents, int32_t num_agents,
                                      int32_t r, int32_t d, int32_t order,
                                      double *max_abs_err) {
  HostLayout H;
  DCORA_TRY(build_layout(agents, num_agents, r, d, order, &H));
  const int dh = d + 1;
  uint64_t st = 88172645463325252ull;
  std::vector<double> Xref((size_t)H.K * r), Xint((size_t)H.K * r);
  for (auto &v : Xref) v = prand(st);
  for (int64_t c = 0; c < H.K; ++c)
    for (int q = 0; q < r; ++q) Xint[c * r + q] = Xref[(size_t)H.perm[c] * r + q];
  // reference: Y = X * Q per agent from the CSR as given
  std::vector<double> Yref((size_t)H.K * r, 0.0), Yint((size_t)H.K * r, 0.0);
  for (int a = 0; a < num_agents; ++a) {
    const dcora_agent_desc &g = agents[a];
    const int64_t c0 = H.coff[a];
    const int k = (int)(H.coff[a + 1] - c0);
    for (int R = 0; R < k; ++R)
      for (int p = g.rowptr[R]; p < g.rowptr[R + 1]; ++p)
        for (int q = 0; q < r; ++q)
          Yref[(c0 + g.col[p]) * r + q] += Xref[(c0 + R) * r + q] * g.val[p];
  }
  // converted structures, same arithmetic the kernels perform
  for (int i = 0; i < H.np; ++i)
    for (int p = H.browptr[i]; p < H.browptr[i + 1]; ++p)
      for (int a = 0; a < dh; ++a)
        for (int b = 0; b < dh; ++b)
          for (int q = 0; q < r; ++q)
            Yint[(size_t)(H.prowcol[i] + a) * r + q] +=
                H.bval[(size_t)p * dh * dh + a * dh + b] * Xint[(size_t)(H.bcol[p] + b) * r + q];
  for (size_t s = 0; s < H.rem_row.size(); ++s)
    for (int p = H.rem_rowptr[s]; p < H.rem_rowptr[s + 1]; ++p)
      for (int q = 0; q < r; ++q)
        Yint[(size_t)H.rem_row[s] * r + q] += H.rem_val[p] * Xint[(size_t)H.rem_col[p] * r + q];
  double err = 0.0;
  for (int64_t c = 0; c < H.K; ++c)
    for (int q = 0; q < r; ++q)
      err = std::fmax(err, std::fabs(Yint[c * r + q] - Yref[(size_t)H.perm[c] * r + q]));
  // chunk table must tile [0, K) exactly
  int64_t next = 0;
  for (const Chunk &ch : H.chunks) {
    if (ch.col0 != next) err = INFINITY;
    next += ch.ncols;
  }
  if (next != H.K) err = INFINITY;
  if (max_abs_err) *max_abs_err = err;
  return DCORA_OK;
}

// Coupling layout (build_coupling_layout): G = [X | H] * C through the converted block /
// remainder structures against the CSR as given, on generated data; with keys, also counts
// the output columns whose sources are not stored in ascending key order.
extern "C" int dcora_selfcheck_coupling(const dcora_agent_desc *agents, int32_t num_agents,
                                        int32_t r, int32_t d, int32_t order,
                                        const dcora_halo_segment *segs, int32_t nseg,
                                        const int32_t *rowptr, const int32_t *col,
                                        const double *val, const int64_t *row_key,
                                        double *max_abs_err, int64_t *order_violations) {
  HostLayout H;
  DCORA_TRY(build_layout(agents, num_agents, r, d, order, &H));
  int64_t KH = 0;
  for (int s = 0; s < nseg; ++s) KH += (int64_t)(d + 1) * segs[s].n + segs[s].l + segs[s].b;
  const int64_t K = H.K, KE = K + KH;
  std::vector<int64_t> rp(rowptr, rowptr + KE + 1);
  HostCoupling C;
  DCORA_TRY(build_coupling_layout(d, order, H.coff, H.n, H.l, H.poff, segs, nseg, rp.data(), col,
                                  val, row_key, &C));
  const int dh = d + 1;
  uint64_t st = 1234567891234567ull;
  std::vector<double> Xref((size_t)KE * r), Xint((size_t)KE * r);
  for (auto &v : Xref) v = prand(st);
  std::vector<int64_t> i2r(KE);   // extended internal column -> extended reference column
  for (int64_t c = 0; c < K; ++c) i2r[c] = H.perm[c];
  for (int64_t c = 0; c < KH; ++c) i2r[K + c] = K + C.perm_h[c];
  for (int64_t c = 0; c < KE; ++c)
    for (int q = 0; q < r; ++q) Xint[c * r + q] = Xref[(size_t)i2r[c] * r + q];
  std::vector<double> Gref((size_t)K * r, 0.0), Gint((size_t)K * r, 0.0);
  for (int64_t R = 0; R < KE; ++R)
    for (int64_t p = rp[R]; p < rp[R + 1]; ++p)
      for (int q = 0; q < r; ++q) Gref[(size_t)col[p] * r + q] += Xref[R * r + q] * val[p];
  int64_t bad = 0;
  auto key_of = [&](int ci) { return row_key ? row_key[i2r[ci]] : (int64_t)ci; };
  for (int i = 0; i < H.np; ++i) {
    for (int p = C.browptr[i]; p < C.browptr[i + 1]; ++p) {
      if (p > C.browptr[i] && key_of(C.bcol[p - 1]) > key_of(C.bcol[p])) ++bad;
      for (int a = 0; a < dh; ++a)
        for (int b = 0; b < dh; ++b)
          for (int q = 0; q < r; ++q)
            Gint[(size_t)(C.prowcol[i] + a) * r + q] +=
                C.bval[(size_t)p * dh * dh + a * dh + b] * Xint[(size_t)(C.bcol[p] + b) * r + q];
    }
  }
  for (size_t s = 0; s < C.rem_row.size(); ++s)
    for (int p = C.rem_rowptr[s]; p < C.rem_rowptr[s + 1]; ++p) {
      if (p > C.rem_rowptr[s] && key_of(C.rem_col[p - 1]) > key_of(C.rem_col[p])) ++bad;
      for (int q = 0; q < r; ++q)
        Gint[(size_t)C.rem_row[s] * r + q] += C.rem_val[p] * Xint[(size_t)C.rem_col[p] * r + q];
    }
  double err = 0.0;
  for (int64_t c = 0; c < K; ++c)
    for (int q = 0; q < r; ++q)
      err = std::fmax(err, std::fabs(Gint[c * r + q] - Gref[(size_t)H.perm[c] * r + q]));
  if (max_abs_err) *max_abs_err = err;
  if (order_violations) *order_violations = bad;
  return DCORA_OK;
}

extern "C" int dcora_selfcheck_cholesky(int32_t k, const int32_t *rowptr, const int32_t *col,
                                        const double *val, int32_t dh, int32_t npose,
                                        double reg, double *rel_residual,
                                        int64_t *factor_nnz) {
  std::vector<int> Ap(rowptr, rowptr + k + 1), Ai;
  std::vector<double> Ax;
  // append reg on the diagonal (an explicit extra entry per row)
  std::vector<int> Bp(k + 1, 0);
  for (int i = 0; i < k; ++i) Bp[i + 1] = Bp[i] + (Ap[i + 1] - Ap[i]) + 1;
  Ai.resize(Bp[k]); Ax.resize(Bp[k]);
  for (int i = 0; i < k; ++i) {
    int w = Bp[i];
    for (int p = Ap[i]; p < Ap[i + 1]; ++p) { Ai[w] = col[p]; Ax[w] = val[p]; ++w; }
    Ai[w] = i; Ax[w] = reg;
  }
  HostFactor F;
  DCORA_TRY(sparse_cholesky(k, Bp, Ai, Ax, dh, npose, &F));
  if (factor_nnz) *factor_nnz = (int64_t)F.Lx.size();
  uint64_t st = 1234567ull;
  std::vector<double> b(k), y(k), x(k);
  for (auto &v : b) v = prand(st);
  for (int i = 0; i < k; ++i) y[i] = b[F.perm[i]];
  for (int j = 0; j < k; ++j) {           // L y = Pb (column oriented)
    y[j] /= F.Lx[F.Lp[j]];
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) y[F.Li[p]] -= F.Lx[p] * y[j];
  }
  for (int j = k - 1; j >= 0; --j) {      // L^T z = y
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) y[j] -= F.Lx[p] * y[F.Li[p]];
    y[j] /= F.Lx[F.Lp[j]];
  }
  for (int i = 0; i < k; ++i) x[F.perm[i]] = y[i];
  double rn = 0.0, bn = 0.0;
  for (int i = 0; i < k; ++i) {
    double s = reg * x[i];
    for (int p = Ap[i]; p < Ap[i + 1]; ++p) s += val[p] * x[col[p]];
    rn += (s - b[i]) * (s - b[i]);
    bn += b[i] * b[i];
  }
  if (rel_residual) *rel_residual = std::sqrt(rn / bn);
  return DCORA_OK;
}

// Factor of Q + reg*I with `threads` host threads (subtree-parallel numeric phase): position
// weighted checksum of L and nnz(L); the factor must not depend on the thread count.
extern "C" int dcora_selfcheck_cholesky_threads(int32_t k, const int32_t *rowptr, const int32_t *col,
                                                const double *val, int32_t dh, int32_t npose,
                                                double reg, int32_t threads, double *checksum,
                                                int64_t *factor_nnz) {
  std::vector<int> Bp(k + 1, 0), Ai;
  std::vector<double> Ax;
  for (int i = 0; i < k; ++i) Bp[i + 1] = Bp[i] + (rowptr[i + 1] - rowptr[i]) + 1;
  Ai.resize(Bp[k]); Ax.resize(Bp[k]);
  for (int i = 0; i < k; ++i) {
    int w = Bp[i];
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) { Ai[w] = col[p]; Ax[w] = val[p]; ++w; }
    Ai[w] = i; Ax[w] = reg;
  }
  HostFactor F;
  DCORA_TRY(sparse_cholesky(k, Bp, Ai, Ax, dh, npose, &F, DCORA_ORDERING_ND, TRSV_LEAF_ROWS, threads));
  double cs = 0.0;
  for (size_t i = 0; i < F.Lx.size(); ++i) cs += F.Lx[i] * std::cos(0.001 * (double)i) + 1e-3 * F.Li[i];
  if (checksum) *checksum = cs;
  if (factor_nnz) *factor_nnz = (int64_t)F.Lx.size();
  return DCORA_OK;
}

#include "trsv.h"

static int factor_for_check(int32_t k, const int32_t *rowptr, const int32_t *col,
                            const double *val, int32_t dh, int32_t npose, double reg,
                            int32_t ordering, HostFactor *F) {
  std::vector<int> Bp(k + 1, 0), Ai;
  std::vector<double> Ax;
  for (int i = 0; i < k; ++i) Bp[i + 1] = Bp[i] + (rowptr[i + 1] - rowptr[i]) + 1;
  Ai.resize(Bp[k]); Ax.resize(Bp[k]);
  for (int i = 0; i < k; ++i) {
    int w = Bp[i];
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) { Ai[w] = col[p]; Ax[w] = val[p]; ++w; }
    Ai[w] = i; Ax[w] = reg;
  }
  return sparse_cholesky(k, Bp, Ai, Ax, dh, npose, F, ordering, TRSV_LEAF_ROWS);
}

// Host self check of the block triangular-solve tapes (trsv_host.cpp): the
// factor of Q + reg*I is built with the requested ordering, compiled into the
// per-warp tapes the kernel streams, and the tapes are replayed sequentially on
// the host for an internally generated right-hand side.  Reports the relative
// residual ||(Q + reg I) x - b|| / ||b||, nnz(L), the number of factor values
// on the tapes per direction (zero-padded pose blocks + inverted diagonal
// blocks), the number of diagonal blocks and the depth of the block
// elimination tree per direction.
extern "C" int dcora_selfcheck_trsv(int32_t k, const int32_t *rowptr, const int32_t *col,
                                    const double *val, int32_t dh, int32_t npose,
                                    double reg, int32_t ordering, double *rel_residual,
                                    int64_t *factor_nnz, int64_t *stored_values,
                                    int32_t *nblocks, int32_t *levels_fwd,
                                    int32_t *levels_bwd) {
  HostFactor F;
  DCORA_TRY(factor_for_check(k, rowptr, col, val, dh, npose, reg, ordering, &F));
  AgentTape S;
  const char *nwe = getenv("DCORA_SELFCHECK_WARPS");
  const int nw = nwe ? atoi(nwe) : 48;
  DCORA_TRY(build_agent_tape(F, dh, nw, std::max(nw, 64), 0, &S));
  if (factor_nnz) *factor_nnz = (int64_t)F.Lx.size();
  if (stored_values) *stored_values = S.stored_values / 2;
  if (nblocks) *nblocks = S.nblocks;
  if (levels_fwd) *levels_fwd = S.nlev[0];
  if (levels_bwd) *levels_bwd = S.nlev[1];
  uint64_t st = 7654321ull;
  std::vector<double> b(k), x(k, 0.0), T(k + 8, 0.0), X(k + 8, 0.0);
  for (auto &v : b) v = prand(st);
  agent_tape_apply_host(S, F, 0, b.data(), x.data(), T.data(), X.data());
  double rn = 0.0, bn = 0.0;
  for (int i = 0; i < k; ++i) {
    double s = reg * x[i];
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) s += val[p] * x[col[p]];
    rn += (s - b[i]) * (s - b[i]);
    bn += b[i] * b[i];
  }
  if (rel_residual) *rel_residual = std::sqrt(rn / bn);
  return DCORA_OK;
}
