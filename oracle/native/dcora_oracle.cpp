// CPU oracle, native twin (TEST INFRASTRUCTURE -- never linked into the product).
//
// A plain C++17 restatement of ONE Agent::updateX of the reference, compiled -O3
// -march=native, for two purposes:
//   * the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): the reference itself
//     cannot be built here (Eigen, SuiteSparse/CHOLMOD, ROPTLIB, Spectra are absent), and the
//     numpy/scipy oracle pays Python control flow and a SuperLU LU of an SPD matrix, which
//     flatters the GPU arm;
//   * a second, independent implementation the scipy oracle is cross-checked against
//     (tests/test_oracle_native.py).
// What it follows:
//   QuadraticProblem::f / EucGrad / EucHessianEta / PreConditioner / RieGrad
//       src/QuadraticProblem.cpp:38-123   (X*Q as CSR SpMM, Q symmetric)
//   Graph::constructPreconditioner           src/Graph.cpp:1901-1917 (Cholesky of Q + reg I;
//       CHOLMOD there, here a minimum-degree ordering on the pose/sphere/landmark graph and
//       an up-looking sparse Cholesky, both written for this file)
//   LiftedSE/RAManifold projection, retraction  src/manifold/LiftedManifold.cpp:18-110,
//       src/DCORA_utils.cpp:2033-2051; ROPTLIB Stiefel qf retraction / EucHvToHv as published
//   QuadraticOptimizer::optimize / trustRegion  src/QuadraticOptimizer.cpp:28-108,234-280 and
//       ROPTLIB RTRNewton::Run / tCG_TR as published (theta 1, kappa 0.1, rho' 0.1, x0.25 / x2)
// Matrices are column-major r x k in the REFERENCE column ordering
// (include/DCORA/manifold/LiftedVariable.h:27-30,148-152).  Single-threaded per call, like
// the reference (ENABLE_OPENMP OFF, CMakeLists.txt:53); callers parallelise over agents.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <set>
#include <vector>

namespace {

struct Layout {
  int r, d, n, l, b, ra, k;
  int rot(int i, int a) const { return ra ? d * i + a : (d + 1) * i + a; }
  int sph(int j) const { return d * n + j; }
  // supervariable of a column: pose i (rotation + translation columns), sphere, landmark
  int sv_of(int c) const {
    if (!ra) return c / (d + 1);
    if (c < d * n) return c / d;
    if (c < d * n + l) return n + (c - d * n);
    if (c < d * n + l + n) return c - d * n - l;
    return n + l + (c - d * n - l - n);
  }
};

struct Csr {
  int k;
  const int *ptr, *col;
  const double *val;
};

using Mat = std::vector<double>;  // r x k column-major

void spmm(const Csr &Q, int r, const double *X, double *Y) {
  for (int c = 0; c < Q.k; ++c) {
    double acc[16] = {0};
    for (int p = Q.ptr[c]; p < Q.ptr[c + 1]; ++p) {
      const double v = Q.val[p];
      const double *x = X + (size_t)Q.col[p] * r;
      for (int q = 0; q < r; ++q) acc[q] += v * x[q];
    }
    for (int q = 0; q < r; ++q) Y[(size_t)c * r + q] = acc[q];
  }
}

double dot(const Mat &a, const Mat &b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}

// ---- manifold ------------------------------------------------------------------------
// out = V - Y sym(Y^T V) per Stiefel block, v - y (y^T v) per sphere, identity elsewhere
void tangent_project(const Layout &L, const double *X, const double *V, double *out) {
  const int r = L.r, d = L.d;
  if (out != V) std::memcpy(out, V, sizeof(double) * (size_t)r * L.k);
  for (int i = 0; i < L.n; ++i) {
    double S[9];
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < d; ++c) {
        double s1 = 0, s2 = 0;
        const double *ya = X + (size_t)L.rot(i, a) * r, *yc = X + (size_t)L.rot(i, c) * r;
        const double *va = V + (size_t)L.rot(i, a) * r, *vc = V + (size_t)L.rot(i, c) * r;
        for (int q = 0; q < r; ++q) { s1 += ya[q] * vc[q]; s2 += yc[q] * va[q]; }
        S[a * d + c] = 0.5 * (s1 + s2);
      }
    double tmp[3 * 16];
    for (int c = 0; c < d; ++c)
      for (int q = 0; q < r; ++q) {
        double s = V[(size_t)L.rot(i, c) * r + q];
        for (int a = 0; a < d; ++a) s -= X[(size_t)L.rot(i, a) * r + q] * S[a * d + c];
        tmp[c * r + q] = s;
      }
    for (int c = 0; c < d; ++c)
      for (int q = 0; q < r; ++q) out[(size_t)L.rot(i, c) * r + q] = tmp[c * r + q];
  }
  for (int j = 0; j < L.l; ++j) {
    const double *y = X + (size_t)L.sph(j) * r, *v = V + (size_t)L.sph(j) * r;
    double t = 0;
    for (int q = 0; q < r; ++q) t += y[q] * v[q];
    double *o = out + (size_t)L.sph(j) * r;
    for (int q = 0; q < r; ++q) o[q] = v[q] - t * y[q];
  }
}

// ROPTLIB Stiefel::EucHvToHv: proj_X(HV - V sym(Y^T EG)); sphere: proj(hv - v (y^T eg))
void ehess_to_rhess(const Layout &L, const double *X, const double *EG, const double *V,
                    const double *HV, double *out) {
  const int r = L.r, d = L.d;
  std::vector<double> W(HV, HV + (size_t)r * L.k);
  for (int i = 0; i < L.n; ++i) {
    double S[9];
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < d; ++c) {
        double s1 = 0, s2 = 0;
        const double *ya = X + (size_t)L.rot(i, a) * r, *yc = X + (size_t)L.rot(i, c) * r;
        const double *ga = EG + (size_t)L.rot(i, a) * r, *gc = EG + (size_t)L.rot(i, c) * r;
        for (int q = 0; q < r; ++q) { s1 += ya[q] * gc[q]; s2 += yc[q] * ga[q]; }
        S[a * d + c] = 0.5 * (s1 + s2);
      }
    for (int c = 0; c < d; ++c)
      for (int q = 0; q < r; ++q) {
        double s = 0;
        for (int a = 0; a < d; ++a) s += V[(size_t)L.rot(i, a) * r + q] * S[a * d + c];
        W[(size_t)L.rot(i, c) * r + q] -= s;
      }
  }
  for (int j = 0; j < L.l; ++j) {
    const double *y = X + (size_t)L.sph(j) * r, *g = EG + (size_t)L.sph(j) * r;
    double t = 0;
    for (int q = 0; q < r; ++q) t += y[q] * g[q];
    for (int q = 0; q < r; ++q) W[(size_t)L.sph(j) * r + q] -= t * V[(size_t)L.sph(j) * r + q];
  }
  tangent_project(L, X, W.data(), out);
}

// thin QR of the r x d block (columns at `cols`), Q with diag(R) > 0, by Householder
void qf_block(int r, int d, double *A /* r x d col-major, in place -> Q */) {
  double Rm[9] = {0};
  std::vector<double> Qm((size_t)r * d, 0.0);
  // modified Gram-Schmidt with one re-orthogonalisation (r x 3: as accurate as Householder)
  for (int c = 0; c < d; ++c) {
    double *ac = A + (size_t)c * r;
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j < c; ++j) {
        const double *qj = A + (size_t)j * r;
        double h = 0;
        for (int q = 0; q < r; ++q) h += qj[q] * ac[q];
        for (int q = 0; q < r; ++q) ac[q] -= h * qj[q];
        Rm[j * d + c] += h;
      }
    double nrm = 0;
    for (int q = 0; q < r; ++q) nrm += ac[q] * ac[q];
    nrm = std::sqrt(nrm);
    for (int q = 0; q < r; ++q) ac[q] /= nrm;   // diag(R) = nrm > 0
  }
}

// polar factor U V^T = A (A^T A)^-1/2 of an r x d block (Jacobi eigen-decomposition of A^T A)
void polar_block(int r, int d, double *A) {
  double M[9], V[9] = {0};
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < d; ++c) {
      double s = 0;
      for (int q = 0; q < r; ++q) s += A[(size_t)a * r + q] * A[(size_t)c * r + q];
      M[a * d + c] = s;
    }
  for (int a = 0; a < d; ++a) V[a * d + a] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < d; ++c)
        if (a != c) off += M[a * d + c] * M[a * d + c];
    if (off < 1e-300) break;
    for (int p = 0; p < d - 1; ++p)
      for (int q = p + 1; q < d; ++q) {
        const double apq = M[p * d + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double tau = (M[q * d + q] - M[p * d + p]) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < d; ++k) {
          const double akp = M[k * d + p], akq = M[k * d + q];
          M[k * d + p] = c * akp - s * akq; M[k * d + q] = s * akp + c * akq;
        }
        for (int k = 0; k < d; ++k) {
          const double apk = M[p * d + k], aqk = M[q * d + k];
          M[p * d + k] = c * apk - s * aqk; M[q * d + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < d; ++k) {
          const double vkp = V[k * d + p], vkq = V[k * d + q];
          V[k * d + p] = c * vkp - s * vkq; V[k * d + q] = s * vkp + c * vkq;
        }
      }
  }
  double W[9];  // V diag(1/sqrt(lambda)) V^T
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < d; ++c) {
      double s = 0;
      for (int e = 0; e < d; ++e) s += V[a * d + e] * V[c * d + e] / std::sqrt(M[e * d + e]);
      W[a * d + c] = s;
    }
  std::vector<double> B((size_t)r * d);
  for (int c = 0; c < d; ++c)
    for (int q = 0; q < r; ++q) {
      double s = 0;
      for (int a = 0; a < d; ++a) s += A[(size_t)a * r + q] * W[a * d + c];
      B[(size_t)c * r + q] = s;
    }
  std::memcpy(A, B.data(), sizeof(double) * (size_t)r * d);
}

// kind 0: qf retraction of X + V, 1: polar retraction, 2: metric projection of X (V unused)
void retract(const Layout &L, const double *X, const double *V, int kind, double *out) {
  const int r = L.r, d = L.d;
  for (size_t i = 0; i < (size_t)r * L.k; ++i) out[i] = kind == 2 ? X[i] : X[i] + V[i];
  std::vector<double> A((size_t)r * d);
  for (int i = 0; i < L.n; ++i) {
    for (int a = 0; a < d; ++a)
      std::memcpy(&A[(size_t)a * r], out + (size_t)L.rot(i, a) * r, sizeof(double) * r);
    if (kind == 0) qf_block(r, d, A.data()); else polar_block(r, d, A.data());
    for (int a = 0; a < d; ++a)
      std::memcpy(out + (size_t)L.rot(i, a) * r, &A[(size_t)a * r], sizeof(double) * r);
  }
  for (int j = 0; j < L.l; ++j) {
    double *y = out + (size_t)L.sph(j) * r;
    double s = 0;
    for (int q = 0; q < r; ++q) s += y[q] * y[q];
    s = 1.0 / std::sqrt(s);
    for (int q = 0; q < r; ++q) y[q] *= s;
  }
}

// ---- sparse Cholesky of P = Q + reg I -------------------------------------------------
struct Factor {
  int k = 0;
  std::vector<int> perm, Lp, Li;   // perm: new -> old; CSC of L, diagonal first
  std::vector<double> Lx;
};

// minimum degree on the supervariable graph (explicit elimination graph)
std::vector<int> min_degree_order(const Layout &L, const Csr &Q) {
  const int nsv = L.n + L.l + L.b;
  std::vector<int> weight(nsv, 1);
  for (int i = 0; i < L.n; ++i) weight[i] = L.d + 1;
  std::vector<std::set<int>> adj(nsv);
  for (int c = 0; c < Q.k; ++c) {
    const int s = L.sv_of(c);
    for (int p = Q.ptr[c]; p < Q.ptr[c + 1]; ++p) {
      const int t = L.sv_of(Q.col[p]);
      if (t != s) { adj[s].insert(t); adj[t].insert(s); }
    }
  }
  auto degree = [&](int v) { int dg = 0; for (int u : adj[v]) dg += weight[u]; return dg; };
  std::set<std::pair<int, int>> heap;
  std::vector<int> deg(nsv);
  for (int v = 0; v < nsv; ++v) { deg[v] = degree(v); heap.insert({deg[v], v}); }
  std::vector<int> order;
  order.reserve(nsv);
  while (!heap.empty()) {
    const int p = heap.begin()->second;
    heap.erase(heap.begin());
    order.push_back(p);
    std::vector<int> nb(adj[p].begin(), adj[p].end());
    for (int u : nb) { heap.erase({deg[u], u}); adj[u].erase(p); }
    for (size_t i = 0; i < nb.size(); ++i)
      for (size_t j = i + 1; j < nb.size(); ++j) { adj[nb[i]].insert(nb[j]); adj[nb[j]].insert(nb[i]); }
    for (int u : nb) { deg[u] = degree(u); heap.insert({deg[u], u}); }
    adj[p].clear();
  }
  return order;
}

bool cholesky(const Layout &L, const Csr &Q, double reg, Factor *F) {
  const int k = Q.k;
  // column permutation from the supervariable order (columns of one supervariable adjacent)
  std::vector<std::vector<int>> cols_of(L.n + L.l + L.b);
  for (int c = 0; c < k; ++c) cols_of[L.sv_of(c)].push_back(c);
  F->perm.clear();
  for (int s : min_degree_order(L, Q))
    for (int c : cols_of[s]) F->perm.push_back(c);
  std::vector<int> pinv(k);
  for (int i = 0; i < k; ++i) pinv[F->perm[i]] = i;
  // upper triangle of P A P^T in CSC
  std::vector<int> Cp(k + 1, 0);
  for (int c = 0; c < k; ++c) {
    bool diag = false;
    for (int p = Q.ptr[c]; p < Q.ptr[c + 1]; ++p) {
      const int i = pinv[c], j = pinv[Q.col[p]];
      if (i <= j) Cp[j + 1]++;
      diag = diag || Q.col[p] == c;
    }
    if (!diag) Cp[pinv[c] + 1]++;
  }
  for (int j = 0; j < k; ++j) Cp[j + 1] += Cp[j];
  std::vector<int> Ci(Cp[k]), fill(Cp.begin(), Cp.end() - 1);
  std::vector<double> Cx(Cp[k]);
  for (int c = 0; c < k; ++c) {
    bool diag = false;
    for (int p = Q.ptr[c]; p < Q.ptr[c + 1]; ++p) {
      const int i = pinv[c], j = pinv[Q.col[p]];
      if (i <= j) { Ci[fill[j]] = i; Cx[fill[j]++] = Q.val[p] + (i == j ? reg : 0.0); }
      diag = diag || Q.col[p] == c;
    }
    if (!diag) { const int j = pinv[c]; Ci[fill[j]] = j; Cx[fill[j]++] = reg; }
  }
  // elimination tree and column counts
  std::vector<int> parent(k, -1), anc(k, -1), cnt(k, 1), flag(k, -1);
  for (int j = 0; j < k; ++j)
    for (int p = Cp[j]; p < Cp[j + 1]; ++p)
      for (int i = Ci[p]; i != -1 && i < j;) {
        const int nx = anc[i];
        anc[i] = j;
        if (nx == -1) parent[i] = j;
        i = nx;
      }
  for (int j = 0; j < k; ++j) {
    flag[j] = j;
    for (int p = Cp[j]; p < Cp[j + 1]; ++p)
      for (int i = Ci[p]; i != -1 && i < j && flag[i] != j; i = parent[i]) { cnt[i]++; flag[i] = j; }
  }
  F->k = k;
  F->Lp.assign(k + 1, 0);
  for (int j = 0; j < k; ++j) F->Lp[j + 1] = F->Lp[j] + cnt[j];
  F->Li.assign(F->Lp[k], 0);
  F->Lx.assign(F->Lp[k], 0.0);
  std::vector<int> cf(F->Lp.begin(), F->Lp.end() - 1), stack(k), pat(k);
  std::vector<double> x(k, 0.0);
  std::fill(flag.begin(), flag.end(), -1);
  for (int j = 0; j < k; ++j) {   // up-looking: row j of L
    int top = k;
    flag[j] = j;
    double dj = 0.0;
    for (int p = Cp[j]; p < Cp[j + 1]; ++p) {
      int i = Ci[p];
      if (i == j) { dj += Cx[p]; continue; }
      x[i] += Cx[p];
      int len = 0;
      for (; flag[i] != j; i = parent[i]) { stack[len++] = i; flag[i] = j; }
      while (len > 0) pat[--top] = stack[--len];
    }
    for (; top < k; ++top) {
      const int i = pat[top];
      const double lji = x[i] / F->Lx[F->Lp[i]];
      x[i] = 0.0;
      for (int p = F->Lp[i] + 1; p < cf[i]; ++p) x[F->Li[p]] -= F->Lx[p] * lji;
      dj -= lji * lji;
      const int q = cf[i]++;
      F->Li[q] = j;
      F->Lx[q] = lji;
    }
    if (!(dj > 0.0) || !std::isfinite(dj)) return false;
    const int q = cf[j]++;
    F->Li[q] = j;
    F->Lx[q] = std::sqrt(dj);
  }
  return true;
}

// W = (L L^T)^-1 V for the r right-hand sides at once (column-major r x k)
void chol_solve(const Factor &F, int r, const double *V, double *W) {
  const int k = F.k;
  std::vector<double> y((size_t)k * r);
  for (int i = 0; i < k; ++i) std::memcpy(&y[(size_t)i * r], V + (size_t)F.perm[i] * r, sizeof(double) * r);
  for (int j = 0; j < k; ++j) {
    double *yj = &y[(size_t)j * r];
    const double dinv = 1.0 / F.Lx[F.Lp[j]];
    for (int q = 0; q < r; ++q) yj[q] *= dinv;
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) {
      double *yi = &y[(size_t)F.Li[p] * r];
      const double lv = F.Lx[p];
      for (int q = 0; q < r; ++q) yi[q] -= lv * yj[q];
    }
  }
  for (int j = k - 1; j >= 0; --j) {
    double *yj = &y[(size_t)j * r];
    for (int p = F.Lp[j] + 1; p < F.Lp[j + 1]; ++p) {
      const double *yi = &y[(size_t)F.Li[p] * r];
      const double lv = F.Lx[p];
      for (int q = 0; q < r; ++q) yj[q] -= lv * yi[q];
    }
    const double dinv = 1.0 / F.Lx[F.Lp[j]];
    for (int q = 0; q < r; ++q) yj[q] *= dinv;
  }
  for (int i = 0; i < k; ++i) std::memcpy(W + (size_t)F.perm[i] * r, &y[(size_t)i * r], sizeof(double) * r);
}

// ---- problem + RTR --------------------------------------------------------------------
struct Problem {
  Layout L;
  Csr Q;
  const double *G;   // may be null
  const Factor *F;   // may be null (no preconditioner: projection only)
  long n_spmm = 0, n_precond = 0;
  size_t len() const { return (size_t)L.r * L.k; }
  void XQ(const double *X, double *Y) { ++n_spmm; spmm(Q, L.r, X, Y); }
  double f(const Mat &X) {
    Mat t(len());
    XQ(X.data(), t.data());
    double s = 0.5 * dot(t, X);
    if (G) for (size_t i = 0; i < len(); ++i) s += X[i] * G[i];
    return s;
  }
  void egrad(const Mat &X, Mat &EG) {
    XQ(X.data(), EG.data());
    if (G) for (size_t i = 0; i < len(); ++i) EG[i] += G[i];
  }
  void precondition(const Mat &X, const Mat &V, Mat &out) {
    ++n_precond;
    if (!F) { tangent_project(L, X.data(), V.data(), out.data()); return; }
    Mat w(len());
    chol_solve(*F, L.r, V.data(), w.data());
    tangent_project(L, X.data(), w.data(), out.data());
  }
  void rhess(const Mat &X, const Mat &EG, const Mat &V, Mat &out) {
    Mat hv(len());
    XQ(V.data(), hv.data());
    ehess_to_rhess(L, X.data(), EG.data(), V.data(), hv.data(), out.data());
  }
};

enum { TR_NEGCURVTURE, TR_EXCREGION, TR_LCON, TR_SCON, TR_MAXITER };

int tcg(Problem &P, const Mat &X, const Mat &EG, const Mat &g, double Delta, int max_inner,
        Mat &eta, int *status) {
  const size_t N = P.len();
  Mat r = g, z(N), delta(N), Hd(N);
  double e_Pe = 0.0, e_Pd = 0.0;
  const double nr0 = std::sqrt(dot(r, r));
  P.precondition(X, r, z);
  double z_r = dot(z, r), d_Pd = z_r;
  for (size_t i = 0; i < N; ++i) { delta[i] = -z[i]; eta[i] = 0.0; }
  *status = TR_MAXITER;
  int j = 0;
  for (; j < max_inner; ++j) {
    P.rhess(X, EG, delta, Hd);
    const double d_Hd = dot(delta, Hd);
    const double alpha = z_r / d_Hd;
    const double e_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
    if (d_Hd <= 0 || e_new >= Delta * Delta) {
      const double tau = (-e_Pd + std::sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
      for (size_t i = 0; i < N; ++i) eta[i] += tau * delta[i];
      *status = d_Hd <= 0 ? TR_NEGCURVTURE : TR_EXCREGION;
      return j + 1;
    }
    e_Pe = e_new;
    for (size_t i = 0; i < N; ++i) { eta[i] += alpha * delta[i]; r[i] += alpha * Hd[i]; }
    const double nr = std::sqrt(dot(r, r));
    if (nr <= nr0 * std::min(nr0, 0.1)) {
      *status = 0.1 < nr0 ? TR_LCON : TR_SCON;
      return j + 1;
    }
    P.precondition(X, r, z);
    const double zold = z_r;
    z_r = dot(z, r);
    const double beta = z_r / zold;
    for (size_t i = 0; i < N; ++i) delta[i] = -z[i] + beta * delta[i];
    e_Pd = beta * (e_Pd + alpha * d_Pd);
    d_Pd = z_r + beta * beta * d_Pd;
  }
  return j;
}

struct RunStats { int outer = 0, inner = 0, status = TR_MAXITER; bool accepted = false; };

void rtr_run(Problem &P, Mat &x1, double tol, double Delta0, double DeltaMax, int max_iter,
             int max_inner, int retraction, RunStats *st) {
  const size_t N = P.len();
  Mat EG(N), gf(N), eta(N), x2(N), Heta(N);
  double f1 = P.f(x1);
  P.egrad(x1, EG);
  tangent_project(P.L, x1.data(), EG.data(), gf.data());
  double Delta = Delta0;
  bool stop = false;
  for (int it = 0; !stop && it < max_iter; ++it) {
    int status;
    st->inner += tcg(P, x1, EG, gf, Delta, max_inner, eta, &status);
    st->status = status;
    retract(P.L, x1.data(), eta.data(), retraction, x2.data());
    const double f2 = P.f(x2);
    P.rhess(x1, EG, eta, Heta);
    double m = 0;
    for (size_t i = 0; i < N; ++i) m += eta[i] * (gf[i] + 0.5 * Heta[i]);
    const double rho = (f1 - f2) / (-m);
    if (rho > 0.75) {
      if (status == TR_EXCREGION || status == TR_NEGCURVTURE) Delta = std::min(2.0 * Delta, DeltaMax);
    } else if (rho < 0.25) {
      Delta *= 0.25;
    }
    st->accepted = rho > 0.1 ||
                   (std::fabs(f1 - f2) / (std::fabs(f1) + 1) < 1.4901161193847656e-08 && f2 < f1);
    if (st->accepted) {
      x1 = x2;
      f1 = f2;
      P.egrad(x1, EG);
      tangent_project(P.L, x1.data(), EG.data(), gf.data());
      stop = std::sqrt(dot(gf, gf)) < tol;
    }
    st->outer++;
  }
}

// An agent's problem with its data matrix copied and its preconditioner factored ONCE (the
// reference caches the CHOLMOD factor until Q changes, Graph.cpp:530-533,1894-1917).
struct OracleProblem {
  Layout L;
  std::vector<int> rowptr, col;
  std::vector<double> val;
  Factor F;
  bool has_factor = false;
  double factor_s = 0.0;
};

static int update_impl(const Layout &L, const Csr &Q, const Factor *F, const double *G,
                       const double *X0, double *Xout, int max_outer, int max_inner, double tol,
                       double delta0, int retraction, double *stats) {
  const auto t1 = std::chrono::steady_clock::now();
  Problem P{L, Q, G, F};
  Mat X(X0, X0 + P.len()), g(P.len()), EG(P.len());
  auto gradnorm = [&](const Mat &x) {
    P.egrad(x, EG);
    tangent_project(L, x.data(), EG.data(), g.data());
    return std::sqrt(dot(g, g));
  };
  stats[0] = P.f(X);
  stats[1] = gradnorm(X);
  RunStats st;
  if (!(stats[1] < tol)) {   // QuadraticOptimizer::trustRegion early return
    if (max_outer == 1) {    // QuadraticOptimizer.cpp:254-273
      double radius = delta0;
      for (int total = 0;; ++total) {
        Mat Xt = X;
        rtr_run(P, Xt, tol, radius, radius, 1, max_inner, retraction, &st);
        if (st.accepted) { X = Xt; break; }
        if (total > 10) break;
        radius /= 4;
      }
    } else {
      rtr_run(P, X, tol, delta0, 5 * delta0, max_outer, max_inner, retraction, &st);
    }
  }
  const auto t2 = std::chrono::steady_clock::now();
  stats[2] = P.f(X);
  stats[3] = gradnorm(X);
  stats[4] = st.outer; stats[5] = st.inner;
  stats[7] = std::chrono::duration<double>(t2 - t1).count();
  stats[8] = F ? (double)F->Lx.size() : 0.0;
  stats[9] = st.status;
  stats[10] = (double)P.n_precond;
  stats[11] = (double)P.n_spmm;
  std::memcpy(Xout, X.data(), sizeof(double) * P.len());
  return 0;
}

}  // namespace

extern "C" {

// stats[0..11] = fInit, gradNormInit, fOpt, gradNormOpt, outer, inner, factor seconds,
// solve seconds, nnz(L), tCG status, preconditioner applications, SpMMs.
void *oracle_problem_create(int r, int d, int n, int l, int b, int ra, const int *rowptr,
                            const int *col, const double *val, double reg) {
  OracleProblem *p = new OracleProblem();
  p->L = Layout{r, d, n, l, b, ra, (d + 1) * n + l + b};
  p->rowptr.assign(rowptr, rowptr + p->L.k + 1);
  p->col.assign(col, col + rowptr[p->L.k]);
  p->val.assign(val, val + rowptr[p->L.k]);
  if (reg >= 0.0) {
    const auto t0 = std::chrono::steady_clock::now();
    Csr Q{p->L.k, p->rowptr.data(), p->col.data(), p->val.data()};
    if (!cholesky(p->L, Q, reg, &p->F)) { delete p; return nullptr; }
    p->has_factor = true;
    p->factor_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return p;
}
void oracle_problem_destroy(void *h) { delete static_cast<OracleProblem *>(h); }

// One Agent::updateX on a created problem (thread safe for distinct handles)
int oracle_problem_update(void *h, const double *G, const double *X0, double *Xout, int max_outer,
                          int max_inner, double tol, double delta0, int retraction,
                          double *stats) {
  OracleProblem *p = static_cast<OracleProblem *>(h);
  Csr Q{p->L.k, p->rowptr.data(), p->col.data(), p->val.data()};
  stats[6] = p->factor_s;
  return update_impl(p->L, Q, p->has_factor ? &p->F : nullptr, G, X0, Xout, max_outer, max_inner,
                     tol, delta0, retraction, stats);
}

// One Agent::updateX: factor (Q + reg I) unless reg < 0, then QuadraticOptimizer::optimize.
// Returns 0, or -1 when the factorisation fails.
int oracle_agent_update(int r, int d, int n, int l, int b, int ra, const int *rowptr,
                        const int *col, const double *val, const double *G, double reg,
                        const double *X0, double *Xout, int max_outer, int max_inner, double tol,
                        double delta0, int retraction, double *stats) {
  Layout L{r, d, n, l, b, ra, (d + 1) * n + l + b};
  Csr Q{L.k, rowptr, col, val};
  Factor F;
  const auto t0 = std::chrono::steady_clock::now();
  if (reg >= 0.0 && !cholesky(L, Q, reg, &F)) return -1;
  stats[6] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return update_impl(L, Q, reg >= 0.0 ? &F : nullptr, G, X0, Xout, max_outer, max_inner, tol,
                     delta0, retraction, stats);
}

// single operations, for the cross-check against the scipy oracle.
// op 0: out = X Q (+ G)   1: tangent projection of V at X   2: qf retraction of X + V
// 3: polar retraction     4: metric projection of X          5: (Q + reg I)^-1 V (no projection)
// 6: Riemannian Hessian-vector product of V at X
int oracle_apply(int op, int r, int d, int n, int l, int b, int ra, const int *rowptr,
                 const int *col, const double *val, const double *G, double reg,
                 const double *X, const double *V, double *out) {
  Layout L{r, d, n, l, b, ra, (d + 1) * n + l + b};
  Csr Q{L.k, rowptr, col, val};
  Problem P{L, Q, G, nullptr};
  const size_t N = P.len();
  switch (op) {
    case 0: { Mat x(X, X + N), eg(N); P.egrad(x, eg); std::memcpy(out, eg.data(), 8 * N); return 0; }
    case 1: tangent_project(L, X, V, out); return 0;
    case 2: retract(L, X, V, 0, out); return 0;
    case 3: retract(L, X, V, 1, out); return 0;
    case 4: retract(L, X, nullptr, 2, out); return 0;
    case 5: {
      Factor F;
      if (!cholesky(L, Q, reg, &F)) return -1;
      chol_solve(F, r, V, out);
      return 0;
    }
    case 6: {
      Mat x(X, X + N), v(V, V + N), eg(N), hv(N);
      P.egrad(x, eg);
      P.rhess(x, eg, v, hv);
      std::memcpy(out, hv.data(), 8 * N);
      return 0;
    }
  }
  return -2;
}

}  // extern "C"
