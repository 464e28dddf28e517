// K6/K7: dual certificate S = Q - Lambda applied implicitly, and lambda_min(S)
// by a device thick-restart Lanczos over the K1 SpMV (r = 1).
//
// Replaces constructDualCertificateMatrix{PGO,RASLAM} (DCORA_utils.cpp:1898-1982),
// fastVerification / isSparseSymmetricMatrixPSD (:1713-1747, CHOLMOD) and
// computeMinimumEigenPair (:1751-1896, Spectra v1.0.1).  Spectra's implicitly
// restarted Lanczos and the thick-restart Lanczos used here span the same
// Krylov spaces; S is never materialised.  Also hosts escapeSaddle
// (QuadraticProblem.cpp:138-234).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "solver.cuh"

namespace dcora {

namespace {

// out = qv - Lambda * v  (Lambda blocks live in a rank-R shaped buffer, see
// launch_sym_ytg: d*d per pose block at pb = col0*R, one scalar per sphere)
template <int DH>
__global__ void __launch_bounds__(128)
lambda_apply_kernel(const Chunk *__restrict__ chunks, int R,
                    const double *__restrict__ Lam, const double *__restrict__ v,
                    const double *__restrict__ qv, double *__restrict__ out) {
  constexpr int D = DH - 1;
  const Chunk ch = chunks[blockIdx.x];
  if (ch.kind == KIND_POSE) {
    for (int t = threadIdx.x; t < ch.ncols; t += 128) {
      const int pi = t / DH, a = t - pi * DH;
      const int col = ch.col0 + t;
      double s = qv[col];
      if (a < D) {
        const size_t pb = (size_t)(ch.col0 + pi * DH) * R;
#pragma unroll
        for (int b = 0; b < D; ++b) s -= Lam[pb + a * D + b] * v[ch.col0 + pi * DH + b];
      }
      out[col] = s;
    }
  } else if (ch.kind == KIND_SPHERE) {
    for (int t = threadIdx.x; t < ch.ncols; t += 128) {
      const int col = ch.col0 + t;
      out[col] = qv[col] - Lam[(size_t)col * R] * v[col];
    }
  } else {
    for (int t = threadIdx.x; t < ch.ncols; t += 128) out[ch.col0 + t] = qv[ch.col0 + t];
  }
}

// The same for a block of R vectors at once (the rows of V^T, K x R column-major R x K like
// every iterate): OUT = QV - V * Lambda.  The distributed certificate applies S to R vectors
// per product (dcora_problem_certificate_apply).
template <int DH>
__global__ void __launch_bounds__(128)
lambda_apply_block_kernel(const Chunk *__restrict__ chunks, int R,
                          const double *__restrict__ Lam, const double *__restrict__ V,
                          const double *__restrict__ QV, double *__restrict__ OUT) {
  constexpr int D = DH - 1;
  const Chunk ch = chunks[blockIdx.x];
  const int n = ch.ncols * R;
  if (ch.kind == KIND_POSE) {
    for (int t = threadIdx.x; t < n; t += 128) {
      const int c = t / R, q = t - c * R;
      const int pi = c / DH, a = c - pi * DH;
      const size_t o = (size_t)(ch.col0 + c) * R + q;
      double s = QV[o];
      if (a < D) {
        const size_t pb = (size_t)(ch.col0 + pi * DH) * R;
#pragma unroll
        for (int b = 0; b < D; ++b) s -= Lam[pb + a * D + b] * V[pb + (size_t)b * R + q];
      }
      OUT[o] = s;
    }
  } else if (ch.kind == KIND_SPHERE) {
    for (int t = threadIdx.x; t < n; t += 128) {
      const int c = t / R;
      const size_t o = (size_t)ch.col0 * R + t;
      OUT[o] = QV[o] - Lam[(size_t)(ch.col0 + c) * R] * V[o];
    }
  } else {
    for (int t = threadIdx.x; t < n; t += 128) OUT[(size_t)ch.col0 * R + t] = QV[(size_t)ch.col0 * R + t];
  }
}

constexpr int LZ_BLOCK = 256;
constexpr int LZ_ROWS = 2048;   // rows per CTA in the tall-skinny kernels
constexpr int LZ_MAXM = 64;

// partial[blk][c] = sum_{i in blk} V[c][i] * w[i],  c < j
__global__ void __launch_bounds__(LZ_BLOCK)
basis_dots_kernel(const double *__restrict__ V, int64_t ld, int j,
                  const double *__restrict__ w, int64_t K, double *__restrict__ partial) {
  __shared__ double sm[LZ_BLOCK / 32];
  const int64_t i0 = (int64_t)blockIdx.x * LZ_ROWS;
  const int64_t i1 = min(K, i0 + LZ_ROWS);
  for (int c = 0; c < j; ++c) {
    const double *vc = V + (size_t)c * ld;
    double acc = 0.0;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += LZ_BLOCK) acc = fma(vc[i], w[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < LZ_BLOCK / 32; ++q) t += sm[q];
      partial[(size_t)blockIdx.x * LZ_MAXM + c] = t;
    }
    __syncthreads();
  }
}

// h[c] = sum_blk partial[blk][c]  (fixed order)
__global__ void __launch_bounds__(LZ_BLOCK)
basis_dots_reduce_kernel(const double *__restrict__ partial, int nblk, int j,
                         double *__restrict__ h) {
  __shared__ double sm[LZ_BLOCK / 32];
  const int c = blockIdx.x;
  if (c >= j) return;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nblk; b += LZ_BLOCK) acc += partial[(size_t)b * LZ_MAXM + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < LZ_BLOCK / 32; ++q) t += sm[q];
    h[c] = t;
  }
}

// w[i] -= sum_c V[c][i] h[c]
__global__ void __launch_bounds__(LZ_BLOCK)
basis_axpy_kernel(const double *__restrict__ V, int64_t ld, int j,
                  const double *__restrict__ h, double *__restrict__ w, int64_t K) {
  __shared__ double sh[LZ_MAXM];
  if (threadIdx.x < j) sh[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * LZ_BLOCK + threadIdx.x;
  if (i >= K) return;
  double acc = w[i];
  for (int c = 0; c < j; ++c) acc = fma(-sh[c], V[(size_t)c * ld + i], acc);
  w[i] = acc;
}

// out[c][i] = sum_q V[q][i] * Y[q*ldy + c],  c < kk  (Ritz vectors)
__global__ void __launch_bounds__(LZ_BLOCK)
basis_rotate_kernel(const double *__restrict__ V, int64_t ld, int m, int kk,
                    const double *__restrict__ Y, double *__restrict__ out, int64_t K) {
  extern __shared__ double sy[];  // m * kk
  for (int t = threadIdx.x; t < m * kk; t += LZ_BLOCK) sy[t] = Y[t];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * LZ_BLOCK + threadIdx.x;
  if (i >= K) return;
  for (int c0 = 0; c0 < kk; c0 += 8) {
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
    for (int q = 0; q < m; ++q) {
      const double v = V[(size_t)q * ld + i];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < kk) acc[u] = fma(v, sy[q * kk + c0 + u], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (c0 + u < kk) out[(size_t)(c0 + u) * ld + i] = acc[u];
  }
}

__global__ void __launch_bounds__(LZ_BLOCK)
scale_copy_kernel(const double *__restrict__ x, double s, double *__restrict__ y, int64_t K) {
  const int64_t i = (int64_t)blockIdx.x * LZ_BLOCK + threadIdx.x;
  if (i < K) y[i] = s * x[i];
}

// out[i] = x[i * r + q]: row q of an r x K column-major matrix as a vector
__global__ void __launch_bounds__(LZ_BLOCK)
extract_row_kernel(const double *__restrict__ x, int r, int q, double *__restrict__ out,
                   int64_t K) {
  const int64_t i = (int64_t)blockIdx.x * LZ_BLOCK + threadIdx.x;
  if (i < K) out[i] = x[i * r + q];
}

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (host, tiny)
void jacobi_eig(int n, std::vector<double> &Amat, std::vector<double> &evec,
                std::vector<double> &eval) {
  evec.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) evec[(size_t)i * n + i] = 1.0;
  auto Aat = [&](int i, int j) -> double & { return Amat[(size_t)i * n + j]; };
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? dg : off) += Aat(i, j) * Aat(i, j);
    if (off <= 1e-32 * (dg + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = Aat(p, q);
        if (std::fabs(apq) < 1e-300) continue;
        const double tau = (Aat(q, q) - Aat(p, p)) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = Aat(k, p), akq = Aat(k, q);
          Aat(k, p) = c * akp - s * akq;
          Aat(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = Aat(p, k), aqk = Aat(q, k);
          Aat(p, k) = c * apk - s * aqk;
          Aat(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = evec[(size_t)k * n + p], vkq = evec[(size_t)k * n + q];
          evec[(size_t)k * n + p] = c * vkp - s * vkq;
          evec[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  eval.resize(n);
  for (int i = 0; i < n; ++i) eval[i] = Aat(i, i);
}

}  // namespace

// sub-handle of rank 1 sharing matrices and stream with `p`
static int make_rank1(dcora_problem *p, dcora_problem **q) {
  dcora_problem *s = nullptr;
  DCORA_TRY(dcora_problem_clone_rank(p, 1, &s));
  cudaStreamDestroy(s->stream);
  s->stream = p->stream;
  s->owns_stream = false;
  *q = s;
  return DCORA_OK;
}

static int apply_S(dcora_problem *p, dcora_problem *q1, const double *Lam,
                   const double *v, double *tmp, double *out) {
  DeviceMatrices &M = *p->M;
  DCORA_TRY(launch_spmm(q1, v, nullptr, tmp, -1));
  if (M.d == 3)
    lambda_apply_kernel<4><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, Lam, v, tmp, out);
  else
    lambda_apply_kernel<3><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, Lam, v, tmp, out);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

}  // namespace dcora

using namespace dcora;

extern "C" int dcora_certify(dcora_problem *p, const double *X, double eta,
                             int32_t max_matvecs, double tol,
                             dcora_certify_result *res, double *v_out) {
  if (!p || !X || !res) return DCORA_ERR_ARG;
  if (p->M->A != 1) {
    set_error("dcora_certify needs a centralised problem (num_agents == 1)");
    return DCORA_ERR_ARG;
  }
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  const int64_t K = M.K;
  double *x = p->vec(WS_IO0), *xq = p->vec(WS_IO1), *Lam = p->vec(WS_S);
  if (!x || !xq || !Lam) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  // Lambda = sym-block-diag(X^T (X Q)) -- G is ignored exactly as the
  // reference does (valid for the centralised problem only)
  DCORA_TRY(launch_spmm(p, x, nullptr, xq, -1));
  DCORA_TRY(launch_sym_ytg(p, x, xq, Lam, -1));

  if (!p->cert_q1) DCORA_TRY(make_rank1(p, &p->cert_q1));   // kept for the next call
  dcora_problem *q1 = p->cert_q1;

  const int m = (int)std::min<int64_t>(48, K);        // Krylov dimension
  const int keep = std::max(1, std::min(m / 2, 16));  // Ritz vectors kept
  double *basis[2] = {nullptr, nullptr};
  double *w = nullptr, *tmp = nullptr, *d_h = nullptr, *d_part = nullptr, *d_Y = nullptr;
  const int nblk = (int)((K + LZ_ROWS - 1) / LZ_ROWS);
  const int gridK = (int)((K + LZ_BLOCK - 1) / LZ_BLOCK);
  auto cleanup = [&]() {};   // the workspace belongs to the handle
#define CERT_CUDA(call)                                    \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) {                              \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
      cleanup();                                           \
      return DCORA_ERR_CUDA;                               \
    }                                                      \
  } while (0)
  {
    // one arena, allocated on the first call and reused (sizes depend on K only)
    const size_t n_basis = (size_t)(m + 1) * K;
    const size_t need = 2 * n_basis + 2 * (size_t)K + LZ_MAXM + (size_t)nblk * LZ_MAXM +
                        (size_t)LZ_MAXM * LZ_MAXM;
    if (p->cert_ws_len < need) {
      cudaFree(p->cert_ws);
      p->cert_ws = nullptr;
      p->cert_ws_len = 0;
      CERT_CUDA(cudaMalloc(&p->cert_ws, sizeof(double) * need));
      p->cert_ws_len = need;
    }
    double *a = p->cert_ws;
    basis[0] = a; a += n_basis;
    basis[1] = a; a += n_basis;
    w = a; a += K;
    tmp = a; a += K;
    d_h = a; a += LZ_MAXM;
    d_part = a; a += (size_t)nblk * LZ_MAXM;
    d_Y = a;
  }

  // deterministic start vector (the reference perturbs S.row(0) randomly,
  // DCORA_utils.cpp:1862-1866; any vector with a component along the minimum
  // eigenvector will do)
  {
    std::vector<double> v0(K);
    uint64_t st = 0x9E3779B97F4A7C15ull;
    double nrm = 0.0;
    for (int64_t i = 0; i < K; ++i) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      v0[i] = (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
      nrm += v0[i] * v0[i];
    }
    nrm = 1.0 / std::sqrt(nrm);
    for (int64_t i = 0; i < K; ++i) v0[i] *= nrm;
    CERT_CUDA(cudaMemcpyAsync(basis[0], v0.data(), sizeof(double) * K, cudaMemcpyHostToDevice,
                              p->stream));
    CERT_CUDA(cudaStreamSynchronize(p->stream));
  }

  std::vector<double> T((size_t)m * m, 0.0), evec, eval, h(LZ_MAXM), Yk;
  int cur = 0, k = 0, matvecs = 0, restarts = 0, converged = 0;
  double lam = 0.0, resid = 0.0, scale = 0.0;
  std::vector<double> best_y(m, 0.0);
  int rc = DCORA_OK;
  while (true) {
    double *V = basis[cur];
    double beta_last = 0.0;
    for (int j = k; j < m; ++j) {
      rc = apply_S(p, q1, Lam, V + (size_t)j * K, tmp, w);
      if (rc < 0) { cleanup(); return rc; }
      ++matvecs;
      // two passes of classical Gram-Schmidt against V[0..j]; the first pass
      // yields the projected matrix column (alpha_j and, after a restart, the
      // coupling row)
      for (int pass = 0; pass < 2; ++pass) {
        basis_dots_kernel<<<nblk, LZ_BLOCK, 0, p->stream>>>(V, K, j + 1, w, K, d_part);
        basis_dots_reduce_kernel<<<j + 1, LZ_BLOCK, 0, p->stream>>>(d_part, nblk, j + 1, d_h);
        basis_axpy_kernel<<<gridK, LZ_BLOCK, 0, p->stream>>>(V, K, j + 1, d_h, w, K);
        p->launches += 3;
        if (pass == 0) {
          CERT_CUDA(cudaMemcpyAsync(h.data(), d_h, sizeof(double) * (j + 1),
                                    cudaMemcpyDeviceToHost, p->stream));
          CERT_CUDA(cudaStreamSynchronize(p->stream));
          T[(size_t)j * m + j] = h[j];
        }
      }
      // beta_j = ||w||
      basis_dots_kernel<<<nblk, LZ_BLOCK, 0, p->stream>>>(w, K, 1, w, K, d_part);
      basis_dots_reduce_kernel<<<1, LZ_BLOCK, 0, p->stream>>>(d_part, nblk, 1, d_h);
      p->launches += 2;
      double bb = 0.0;
      CERT_CUDA(cudaMemcpyAsync(&bb, d_h, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
      CERT_CUDA(cudaStreamSynchronize(p->stream));
      const double beta = std::sqrt(std::max(0.0, bb));
      beta_last = beta;
      if (j + 1 < m) { T[(size_t)(j + 1) * m + j] = beta; T[(size_t)j * m + j + 1] = beta; }
      const double inv = beta > 1e-300 ? 1.0 / beta : 0.0;
      scale_copy_kernel<<<gridK, LZ_BLOCK, 0, p->stream>>>(w, inv, V + (size_t)(j + 1) * K, K);
      p->launches += 1;
    }
    // Ritz pairs of the projected matrix
    std::vector<double> Tc(T);
    jacobi_eig(m, Tc, evec, eval);
    std::vector<int> idx(m);
    for (int i = 0; i < m; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return eval[a] < eval[b]; });
    scale = 0.0;
    for (int i = 0; i < m; ++i) scale = std::max(scale, std::fabs(eval[i]));
    lam = eval[idx[0]];
    resid = std::fabs(beta_last * evec[(size_t)(m - 1) * m + idx[0]]);
    for (int q = 0; q < m; ++q) best_y[q] = evec[(size_t)q * m + idx[0]];
    ++restarts;
    converged = resid <= tol * std::max(scale, 1e-300);
    // early verdict: lambda_min <= lam always (Ritz values bound from above),
    // so a Ritz value below -eta already decides "not PSD" once its residual
    // is small relative to the margin
    if (converged || matvecs >= max_matvecs || beta_last <= 1e-300) break;
    // thick restart: keep the `keep` smallest Ritz vectors
    const int kk = std::min(keep, m - 1);
    Yk.assign((size_t)m * kk, 0.0);
    for (int q = 0; q < m; ++q)
      for (int c = 0; c < kk; ++c) Yk[(size_t)q * kk + c] = evec[(size_t)q * m + idx[c]];
    CERT_CUDA(cudaMemcpyAsync(d_Y, Yk.data(), sizeof(double) * m * kk, cudaMemcpyHostToDevice,
                              p->stream));
    double *Vn = basis[1 - cur];
    basis_rotate_kernel<<<gridK, LZ_BLOCK, sizeof(double) * m * kk, p->stream>>>(V, K, m, kk,
                                                                                 d_Y, Vn, K);
    CERT_CUDA(cudaMemcpyAsync(Vn + (size_t)kk * K, V + (size_t)m * K, sizeof(double) * K,
                              cudaMemcpyDeviceToDevice, p->stream));
    p->launches += 1;
    std::fill(T.begin(), T.end(), 0.0);
    for (int c = 0; c < kk; ++c) {
      T[(size_t)c * m + c] = eval[idx[c]];
      const double s = beta_last * evec[(size_t)(m - 1) * m + idx[c]];
      T[(size_t)kk * m + c] = s;
      T[(size_t)c * m + kk] = s;
    }
    CERT_CUDA(cudaStreamSynchronize(p->stream));
    cur = 1 - cur;
    k = kk;
  }
  // minimum Ritz vector v = V y
  {
    double *V = basis[cur];
    CERT_CUDA(cudaMemcpyAsync(d_Y, best_y.data(), sizeof(double) * m, cudaMemcpyHostToDevice,
                              p->stream));
    basis_rotate_kernel<<<gridK, LZ_BLOCK, sizeof(double) * m, p->stream>>>(V, K, m, 1, d_Y, w, K);
    p->launches += 1;
    CERT_CUDA(cudaStreamSynchronize(p->stream));
  }
  // Rayleigh-Ritz on span(rows of X).  At a critical point S X^T = 0, so the
  // rows of X span (numerically) null vectors of S; the pencil's smallest
  // eigenvalues are clustered there at a relative gap of ~1e-9 of the spectrum
  // and a Krylov method from a random start resolves that cluster last.  Both
  // numbers are Rayleigh quotients, i.e. upper bounds of lambda_min(S): report
  // the smaller one (with its vector).
  const int r = p->r;
  if (2 * r <= m + 1) {
    double *B = basis[1 - cur];            // r rows of X, then S * rows
    const int gridr = gridK;
    for (int q = 0; q < r; ++q) {
      extract_row_kernel<<<gridr, LZ_BLOCK, 0, p->stream>>>(x, r, q, B + (size_t)q * K, K);
      p->launches += 1;
    }
    std::vector<double> Mm((size_t)r * r), Nm((size_t)r * r), hh(LZ_MAXM);
    for (int q = 0; q < r; ++q) {
      double *SBq = B + (size_t)(r + q) * K;
      rc = apply_S(p, q1, Lam, B + (size_t)q * K, tmp, SBq);
      if (rc < 0) { cleanup(); return rc; }
      ++matvecs;
      for (int which = 0; which < 2; ++which) {
        const double *rhs = which == 0 ? SBq : B + (size_t)q * K;
        basis_dots_kernel<<<nblk, LZ_BLOCK, 0, p->stream>>>(B, K, r, rhs, K, d_part);
        basis_dots_reduce_kernel<<<r, LZ_BLOCK, 0, p->stream>>>(d_part, nblk, r, d_h);
        p->launches += 2;
        CERT_CUDA(cudaMemcpyAsync(hh.data(), d_h, sizeof(double) * r, cudaMemcpyDeviceToHost,
                                  p->stream));
        CERT_CUDA(cudaStreamSynchronize(p->stream));
        for (int c = 0; c < r; ++c) (which == 0 ? Mm : Nm)[(size_t)c * r + q] = hh[c];
      }
    }
    // generalised symmetric problem M y = mu N y.  N = X X^T is singular when
    // the solution is rank deficient (the usual case above the optimal rank), so
    // it is reduced with its own eigen-decomposition: N = U D U^T, basis
    // W = U_k D_k^-1/2 over the numerically non-zero part, C = W^T M W.
    std::vector<double> Ns((size_t)r * r), Un, Dn;
    for (int i = 0; i < r; ++i)
      for (int j2 = 0; j2 < r; ++j2)
        Ns[(size_t)i * r + j2] = 0.5 * (Nm[(size_t)i * r + j2] + Nm[(size_t)j2 * r + i]);
    jacobi_eig(r, Ns, Un, Dn);
    double dmax = 0.0;
    for (int i = 0; i < r; ++i) dmax = std::max(dmax, Dn[i]);
    std::vector<int> keepc;
    for (int i = 0; i < r; ++i)
      if (Dn[i] > 1e-10 * dmax) keepc.push_back(i);
    const int rk = (int)keepc.size();
    if (rk > 0 && dmax > 0.0) {
      // W[q][c] = U[q][keepc[c]] / sqrt(D[keepc[c]])
      std::vector<double> Wn((size_t)r * rk);
      for (int q = 0; q < r; ++q)
        for (int c = 0; c < rk; ++c)
          Wn[(size_t)q * rk + c] = Un[(size_t)q * r + keepc[c]] / std::sqrt(Dn[keepc[c]]);
      std::vector<double> MW((size_t)r * rk, 0.0), Cm((size_t)rk * rk, 0.0);
      for (int i = 0; i < r; ++i)
        for (int c = 0; c < rk; ++c) {
          double sacc = 0.0;
          for (int t = 0; t < r; ++t)
            sacc += 0.5 * (Mm[(size_t)i * r + t] + Mm[(size_t)t * r + i]) * Wn[(size_t)t * rk + c];
          MW[(size_t)i * rk + c] = sacc;
        }
      for (int a2 = 0; a2 < rk; ++a2)
        for (int c = 0; c < rk; ++c) {
          double sacc = 0.0;
          for (int t = 0; t < r; ++t) sacc += Wn[(size_t)t * rk + a2] * MW[(size_t)t * rk + c];
          Cm[(size_t)a2 * rk + c] = sacc;
        }
      for (int a2 = 0; a2 < rk; ++a2)
        for (int c = a2 + 1; c < rk; ++c) {
          const double sym = 0.5 * (Cm[(size_t)a2 * rk + c] + Cm[(size_t)c * rk + a2]);
          Cm[(size_t)a2 * rk + c] = sym;
          Cm[(size_t)c * rk + a2] = sym;
        }
      std::vector<double> ev2, el2;
      jacobi_eig(rk, Cm, ev2, el2);
      int best = 0;
      for (int i = 1; i < rk; ++i) if (el2[i] < el2[best]) best = i;
      if (el2[best] < lam) {
        lam = el2[best];
        // coefficients on the rows of X: z = W y
        std::vector<double> z(r, 0.0);
        for (int q = 0; q < r; ++q)
          for (int c = 0; c < rk; ++c) z[q] += Wn[(size_t)q * rk + c] * ev2[(size_t)c * rk + best];
        CERT_CUDA(cudaMemcpyAsync(d_Y, z.data(), sizeof(double) * r, cudaMemcpyHostToDevice,
                                  p->stream));
        basis_rotate_kernel<<<gridK, LZ_BLOCK, sizeof(double) * r, p->stream>>>(B, K, r, 1, d_Y, w, K);
        p->launches += 1;
        CERT_CUDA(cudaStreamSynchronize(p->stream));
      }
    }
  }
  {
    // theta = v^T S v (fastVerification :1727) and the residual ||S v - theta v||
    double *Sw = basis[1 - cur];
    rc = apply_S(p, q1, Lam, w, tmp, Sw);
    if (rc < 0) { cleanup(); return rc; }
    double vals[3] = {0.0, 1.0, 0.0};
    const double *lhs[3] = {Sw, w, Sw}, *rhs[3] = {w, w, Sw};
    for (int t = 0; t < 3; ++t) {
      basis_dots_kernel<<<nblk, LZ_BLOCK, 0, p->stream>>>(lhs[t], K, 1, rhs[t], K, d_part);
      basis_dots_reduce_kernel<<<1, LZ_BLOCK, 0, p->stream>>>(d_part, nblk, 1, d_h);
      p->launches += 2;
      CERT_CUDA(cudaMemcpyAsync(&vals[t], d_h, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
      CERT_CUDA(cudaStreamSynchronize(p->stream));
    }
    const double vSv = vals[0], vv = vals[1], ss = vals[2];
    res->theta = vSv / vv;
    resid = std::sqrt(std::max(0.0, ss / vv - res->theta * res->theta));
    if (v_out) {
      scale_copy_kernel<<<gridK, LZ_BLOCK, 0, p->stream>>>(w, 1.0 / std::sqrt(vv), tmp, K);
      rc = export_matrix(q1, tmp, v_out, 1);
      if (rc < 0) { cleanup(); return rc; }
      CERT_CUDA(cudaStreamSynchronize(p->stream));
    }
  }
  res->lambda_min = lam;
  res->residual = resid;
  res->converged = converged;
  res->matvecs = matvecs;
  res->restarts = restarts;
  // Verdict.  `lam` is a Rayleigh quotient, i.e. an UPPER bound of lambda_min(S): a value
  // below -eta proves "not PSD" whether or not the iteration converged.  "PSD" is decided
  // like the reference does (fastVerification -> isSparseSymmetricMatrixPSD,
  // DCORA_utils.cpp:1713-1747): by an exact sparse Cholesky of S + eta I on the host, unless
  // the extreme Ritz pair has converged with a margin that makes the test redundant.
  // Centralised handles also CHECK the eigenvalue with the same factorisation, and fall back
  // to shift-invert iterations on the host factor when Lanczos missed the end of a clustered
  // spectrum (the role of the reference's shift-invert variant of computeMinimumEigenPair,
  // :1751-1805): ill-conditioned range-aided problems (tiers.pyfg) hide eigenvalues of size
  // 1e-3 next to a cluster at zero of a matrix of norm 1e5.
  int proven_psd = 0;
  const bool margin_ok = converged && lam - resid + eta > 0.0 && lam + eta > 0.0 &&
                         !getenv("DCORA_CERTIFY_ALWAYS_CHOLESKY");
  if (margin_ok) {
    proven_psd = 1;
  } else if (!M.hostQ_ptr.empty()) {
    // S = Q - Lambda in the internal ordering (duplicate entries add up) with one explicit
    // diagonal slot per row for the shift
    std::vector<double> hl((size_t)p->r * K);
    CERT_CUDA(cudaMemcpyAsync(hl.data(), Lam, sizeof(double) * hl.size(), cudaMemcpyDeviceToHost,
                              p->stream));
    CERT_CUDA(cudaStreamSynchronize(p->stream));
    const int k = (int)K, dd = M.d, dhh = M.d + 1, np = M.n[0], nl = M.l[0];
    std::vector<int> Ap(k + 1, 0), shift_pos(k);
    for (int i = 0; i < k; ++i) {
      int extra = 1;                                   // the shift on the diagonal
      if (i < dhh * np) { if (i % dhh < dd) extra += dd; }
      else if (i < dhh * np + nl) extra += 1;
      Ap[i + 1] = Ap[i] + (M.hostQ_ptr[i + 1] - M.hostQ_ptr[i]) + extra;
    }
    std::vector<int> Ai(Ap[k]);
    std::vector<double> Ax(Ap[k]);
    for (int i = 0; i < k; ++i) {
      int w2 = Ap[i];
      for (int q = M.hostQ_ptr[i]; q < M.hostQ_ptr[i + 1]; ++q) {
        Ai[w2] = M.hostQ_col[q]; Ax[w2] = M.hostQ_val[q]; ++w2;
      }
      shift_pos[i] = w2;
      Ai[w2] = i; Ax[w2] = 0.0; ++w2;
      if (i < dhh * np) {
        const int pi = i / dhh, a = i % dhh;
        if (a < dd) {
          // Lambda block of pose pi lives at pb = (pi * dh) * r: entry [a * d + b] = S[b][a]
          const size_t pb = (size_t)(pi * dhh) * p->r;
          for (int b = 0; b < dd; ++b) { Ai[w2] = pi * dhh + b; Ax[w2] = -hl[pb + a * dd + b]; ++w2; }
        }
      } else if (i < dhh * np + nl) {
        Ai[w2] = i; Ax[w2] = -hl[(size_t)i * p->r]; ++w2;
      }
    }
    HostFactor Fs;
    auto factor_shifted = [&](double sigma) {     // is S + sigma I positive definite?
      for (int i = 0; i < k; ++i) Ax[shift_pos[i]] = sigma;
      Fs = HostFactor();
      return sparse_cholesky(k, Ap, Ai, Ax, dhh, np, &Fs, DCORA_ORDERING_ND, TRSV_LEAF_ROWS,
                             setup_threads()) == DCORA_OK;
    };
    auto times_S = [&](const std::vector<double> &x, std::vector<double> &y) {
      for (int i = 0; i < k; ++i) {
        double acc = 0.0;
        for (int q = Ap[i]; q < Ap[i + 1]; ++q)
          if (q != shift_pos[i]) acc += Ax[q] * x[Ai[q]];
        y[i] = acc;
      }
    };
    bool need_host_pair = false;
    if (lam + eta > 0.0) {
      proven_psd = factor_shifted(eta) ? 1 : 0;
      need_host_pair = !proven_psd;     // the factorisation disproves Lanczos' estimate
    } else {
      // not PSD; is `lam` the end of the spectrum?  S - (lam - slack) I must be definite
      const double slack = std::max(1e-6 * std::fabs(lam), 1e-9);
      need_host_pair = !factor_shifted(-lam + slack);
    }
    if (need_host_pair) {
      // smallest sigma = eta * 2^j (at least 2 |lam|) with S + sigma I definite, then inverse
      // iterations; the shift follows the Rayleigh quotient down while it stays definite
      double sigma = std::max(2.0 * std::fabs(lam), 2.0 * eta);
      bool ok = false;
      for (int t = 0; t < 80 && !(ok = factor_shifted(sigma)); ++t) sigma *= 2.0;
      if (ok) {
        std::vector<double> x(k), y(k), z(k);
        uint64_t st = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < k; ++i) {
          st ^= st << 13; st ^= st >> 7; st ^= st << 17;
          x[i] = (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
        auto solve = [&](std::vector<double> &v2) {     // v2 <- (S + sigma I)^-1 v2
          for (int i = 0; i < k; ++i) y[i] = v2[Fs.perm[i]];
          for (int c = 0; c < k; ++c) {
            y[c] /= Fs.Lx[Fs.Lp[c]];
            for (int q = Fs.Lp[c] + 1; q < Fs.Lp[c + 1]; ++q) y[Fs.Li[q]] -= Fs.Lx[q] * y[c];
          }
          for (int c = k - 1; c >= 0; --c) {
            for (int q = Fs.Lp[c] + 1; q < Fs.Lp[c + 1]; ++q) y[c] -= Fs.Lx[q] * y[Fs.Li[q]];
            y[c] /= Fs.Lx[Fs.Lp[c]];
          }
          for (int i = 0; i < k; ++i) v2[Fs.perm[i]] = y[i];
        };
        double rho = 0.0, rho_prev = 1e300, rn = 0.0;
        for (int round = 0; round < 12; ++round) {
          for (int it = 0; it < 12; ++it) {
            solve(x);
            double nx = 0.0;
            for (double v2 : x) nx += v2 * v2;
            nx = 1.0 / std::sqrt(nx);
            for (double &v2 : x) v2 *= nx;
          }
          times_S(x, z);
          rho = 0.0;
          for (int i = 0; i < k; ++i) rho += x[i] * z[i];
          rn = 0.0;
          for (int i = 0; i < k; ++i) rn += (z[i] - rho * x[i]) * (z[i] - rho * x[i]);
          rn = std::sqrt(rn);
          if (std::fabs(rho - rho_prev) <= 1e-13 * std::max(std::fabs(rho), 1e-3) ||
              rn <= 1e-12 * std::max(std::fabs(rho), 1.0))
            break;
          rho_prev = rho;
          // move the shift 90 % of the way to -rho if S stays definite there (faster contraction)
          const double trial = -rho + 0.1 * (sigma + rho);
          if (trial > 0.0 && trial < sigma) {
            HostFactor keepF = Fs;
            if (factor_shifted(trial)) sigma = trial; else { Fs = keepF; }
          }
        }
        lam = rho;
        res->lambda_min = rho;
        res->theta = rho;
        res->residual = rn;
        res->converged = 1;
        if (v_out) {
          CERT_CUDA(cudaMemcpyAsync(tmp, x.data(), sizeof(double) * k, cudaMemcpyHostToDevice, p->stream));
          int rc2 = export_matrix(q1, tmp, v_out, 1);
          if (rc2 < 0) { cleanup(); return rc2; }
          CERT_CUDA(cudaStreamSynchronize(p->stream));
        }
        proven_psd = 0;   // either the exact test said so or lam < -eta
      }
    }
  }
  if (!proven_psd && lam + eta > 0.0 && M.hostQ_ptr.empty()) res->converged = 0;
  res->is_psd = proven_psd;
  cleanup();
#undef CERT_CUDA
  return DCORA_OK;
}

// Distributed certificate (dcora_b200/certify_dist.py): the multipliers of the batch from the
// iterate and X * Q_global, kept on the handle ...
extern "C" int dcora_problem_certificate_prepare(dcora_problem *p, const double *X,
                                                 const double *XQ) {
  if (!p || !X || !XQ) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  double *x = p->vec(WS_IO0), *xq = p->vec(WS_IO1), *Lam = p->vec(WS_S);
  if (!x || !xq || !Lam) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, p->r));
  DCORA_TRY(import_matrix(p, XQ, xq, p->r));
  DCORA_TRY(launch_sym_ytg(p, x, xq, Lam, -1));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));
  return DCORA_OK;
}

// ... and OUT = QV - V * Lambda for a block of r vectors (S V restricted to this handle's
// columns, QV = V * Q_global from dcora_problem_update_G + dcora_problem_egrad)
extern "C" int dcora_problem_certificate_apply(dcora_problem *p, const double *V,
                                               const double *QV, double *OUT) {
  if (!p || !V || !QV || !OUT) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  double *v = p->vec(WS_IO0), *qv = p->vec(WS_IO1), *o = p->vec(WS_IO2), *Lam = p->vec(WS_S);
  if (!v || !qv || !o || !Lam) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, V, v, p->r));
  DCORA_TRY(import_matrix(p, QV, qv, p->r));
  if (M.d == 3)
    lambda_apply_block_kernel<4><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, Lam, v, qv, o);
  else
    lambda_apply_block_kernel<3><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, Lam, v, qv, o);
  DCORA_LAUNCH_CHECK(p);
  return export_matrix(p, o, OUT, p->r);
}

extern "C" int dcora_problem_escape_saddle(dcora_problem *p, const double *Xopt,
                                           double theta, const double *v,
                                           double gradient_tolerance,
                                           double preconditioned_gradient_tolerance,
                                           int32_t is_second_order, int32_t retraction,
                                           double *Xout) {
  if (!p || !Xopt || !v || !Xout) return DCORA_ERR_ARG;
  if (p->M->A != 1 || p->r < 2) {
    set_error("escape_saddle needs a single-agent problem of rank >= 2");
    return DCORA_ERR_ARG;
  }
  DCORA_CUDA(cudaSetDevice(p->M->device));
  const int r = p->r;
  const int64_t K = p->M->K;
  // Host assembly of X_plus = [Xopt; 0] and X_dot = [0; v^T] in the reference
  // layout (QuadraticProblem.cpp:148-154); everything else runs on the device.
  std::vector<double> hx((size_t)r * K, 0.0), hd((size_t)r * K, 0.0), hin;
  const double *src = Xopt;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, Xopt) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
    hin.resize((size_t)(r - 1) * K);
    DCORA_CUDA(cudaMemcpy(hin.data(), Xopt, sizeof(double) * hin.size(), cudaMemcpyDeviceToHost));
    src = hin.data();
  } else {
    cudaGetLastError();
  }
  for (int64_t c = 0; c < K; ++c) {
    for (int q = 0; q < r - 1; ++q) hx[(size_t)c * r + q] = src[(size_t)c * (r - 1) + q];
    hd[(size_t)c * r + (r - 1)] = v[c];
  }
  double *xp = p->vec(WS_X1), *xd = p->vec(WS_ETA), *step = p->vec(WS_DELTA);
  double *xt = p->vec(WS_X2), *t0 = p->vec(WS_T0), *eg = p->vec(WS_EG), *g = p->vec(WS_GF);
  double *pg = p->vec(WS_Z);
  if (!xp || !xd || !step || !xt || !t0 || !eg || !g || !pg) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, hx.data(), xp, r));
  DCORA_TRY(import_matrix(p, hd.data(), xd, r));
  DCORA_CUDA(cudaStreamSynchronize(p->stream));

  const double alpha_min = 1e-6;
  double alpha = is_second_order
                     ? std::max(16 * alpha_min, 100 * gradient_tolerance / std::fabs(theta))
                     : 1.0;
  DCORA_TRY(eval_f(p, xp, t0, S_TMP0, -1));
  DCORA_TRY(read_scalars(p));
  const double FX = p->h_scal[S_TMP0];
  std::vector<double> alphas, fvals;
  auto trial = [&](double a, double *f, double *gn, double *pgn) -> int {
    DCORA_TRY(launch_axpby(p, step, -1, a, xd, -1, 0.0, nullptr, -1));
    DCORA_TRY(launch_retract(p, xp, step, xt, retraction, -1));
    DCORA_TRY(eval_f(p, xt, t0, S_TMP0, -1));
    if (gn) {
      DCORA_TRY(launch_spmm(p, xt, p->has_G ? p->G : nullptr, eg, -1));
      DCORA_TRY(launch_project_tangent(p, xt, eg, g, -1));
      DCORA_TRY(launch_precond(p, xt, g, pg, -1));
      const double *xs[2] = {g, pg};
      DCORA_TRY(launch_dots(p, 2, xs, xs, -1));
      const int dst[2] = {S_TMP1, S_TMP2};
      DCORA_TRY(launch_reduce_scalars(p, 2, dst));
    }
    DCORA_TRY(read_scalars(p));
    *f = p->h_scal[S_TMP0];
    if (gn) { *gn = std::sqrt(p->h_scal[S_TMP1]); *pgn = std::sqrt(p->h_scal[S_TMP2]); }
    return DCORA_OK;
  };
  while (alpha >= alpha_min) {
    double ft, gn, pgn;
    DCORA_TRY(trial(alpha, &ft, &gn, &pgn));
    alphas.push_back(alpha);
    fvals.push_back(ft);
    if (ft < FX && gn > gradient_tolerance && pgn > preconditioned_gradient_tolerance)
      return export_matrix(p, xt, Xout, r);
    alpha /= 2;
  }
  const size_t imin = std::min_element(fvals.begin(), fvals.end()) - fvals.begin();
  if (!fvals.empty() && fvals[imin] < FX) {
    double ft;
    DCORA_TRY(trial(alphas[imin], &ft, nullptr, nullptr));
    return export_matrix(p, xt, Xout, r);
  }
  return DCORA_SOFT_FAIL;
}
