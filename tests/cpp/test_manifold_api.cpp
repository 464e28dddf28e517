// C++ checks of the host mirror (include/dcora_b200/DCORA.hpp) written like the reference's
// own gtests, against closed forms computed on the host in this file:
//   * testLiftedSEManifoldProjectToTangentSpace / ...RA...  (tests/testManifold.cpp:354-440):
//     projectToTangentSpace == V - Y sym(Y^T V) on Stiefel blocks, identity on Euclidean
//     columns, v - y (y^T v) on unit-sphere columns (the reference's oblique assertion at
//     :433 is vacuous; here it is a real check);
//   * test*RandomManifold (:77-107): project() output is feasible and a fixed point;
//   * QuadraticProblem::f / RieGrad consistency on a small pose graph: f from the device
//     equals 0.5 <X Q, X> computed on the host from Graph::quadraticMatrix(), and
//     RieGrad == projectToTangentSpace(X, X Q).
// Needs a CUDA device.  Prints "ok <name>" per check, exits non-zero on the first failure.
#include <random>

#include "dcora_b200/DCORA.hpp"

using DCORA::Matrix;

static int failures = 0;
#define EXPECT(cond, name)                                             \
  do {                                                                 \
    if (cond) std::printf("ok %s\n", name);                            \
    else { std::printf("FAILED %s (%s:%d)\n", name, __FILE__, __LINE__); ++failures; } \
  } while (0)

static Matrix random_matrix(size_t r, size_t k, unsigned seed) {
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> N(0.0, 1.0);
  Matrix M(r, k);
  for (size_t j = 0; j < k; ++j)
    for (size_t i = 0; i < r; ++i) M(i, j) = N(rng);
  return M;
}

// closed form for one Stiefel block at columns [c0, c0+d): V - Y sym(Y^T V)
static double stiefel_block_error(const Matrix &X, const Matrix &V, const Matrix &P, size_t c0,
                                  size_t d) {
  const size_t r = X.rows();
  double err = 0.0;
  std::vector<double> S(d * d);
  for (size_t a = 0; a < d; ++a)
    for (size_t b = 0; b < d; ++b) {
      double yv = 0.0, vy = 0.0;
      for (size_t q = 0; q < r; ++q) {
        yv += X(q, c0 + a) * V(q, c0 + b);
        vy += V(q, c0 + a) * X(q, c0 + b);
      }
      S[a * d + b] = 0.5 * (yv + vy);
    }
  for (size_t b = 0; b < d; ++b)
    for (size_t q = 0; q < r; ++q) {
      double ref = V(q, c0 + b);
      for (size_t a = 0; a < d; ++a) ref -= X(q, c0 + a) * S[a * d + b];
      err = std::max(err, std::fabs(ref - P(q, c0 + b)));
    }
  return err;
}

static double stiefel_feasibility(const Matrix &X, size_t c0, size_t d) {
  double err = 0.0;
  for (size_t a = 0; a < d; ++a)
    for (size_t b = 0; b < d; ++b) {
      double s = 0.0;
      for (size_t q = 0; q < X.rows(); ++q) s += X(q, c0 + a) * X(q, c0 + b);
      err = std::max(err, std::fabs(s - (a == b ? 1.0 : 0.0)));
    }
  return err;
}

int main() {
  if (dcora_device_count() < 1) {
    std::printf("no CUDA device\n");
    return 77;
  }
  // ---------------- SE manifold (tests/testManifold.cpp:354-390)
  for (unsigned d = 2; d <= 3; ++d) {
    const unsigned r = 5, n = 37;
    DCORA::LiftedSEManifold M(r, d, n);
    const Matrix X = M.project(random_matrix(r, (d + 1) * n, 1));
    const Matrix V = random_matrix(r, (d + 1) * n, 2);
    const Matrix P = M.projectToTangentSpace(X, V);
    double e_rot = 0.0, e_tr = 0.0, e_feas = 0.0;
    for (unsigned i = 0; i < n; ++i) {
      e_rot = std::max(e_rot, stiefel_block_error(X, V, P, i * (d + 1), d));
      e_feas = std::max(e_feas, stiefel_feasibility(X, i * (d + 1), d));
      for (unsigned q = 0; q < r; ++q)
        e_tr = std::max(e_tr, std::fabs(P(q, i * (d + 1) + d) - V(q, i * (d + 1) + d)));
    }
    EXPECT(e_rot < 1e-12, d == 2 ? "SE(2) tangent projection, Stiefel blocks" : "SE(3) tangent projection, Stiefel blocks");
    EXPECT(e_tr == 0.0, "SE tangent projection, Euclidean columns");
    EXPECT(e_feas < 1e-12, "SE project(): Y^T Y = I");
    const Matrix X2 = M.project(X);
    double e_fix = 0.0;
    for (size_t j = 0; j < X.cols(); ++j)
      for (size_t i = 0; i < X.rows(); ++i) e_fix = std::max(e_fix, std::fabs(X2(i, j) - X(i, j)));
    EXPECT(e_fix < 1e-12, "SE project() is idempotent");
  }
  // ---------------- RA manifold (tests/testManifold.cpp:392-440), ordering [Y.. | r.. | p.. | L..]
  {
    const unsigned r = 6, d = 3, n = 11, l = 7, b = 3;
    const unsigned k = (d + 1) * n + l + b;
    DCORA::LiftedRAManifold M(r, d, n, l, b);
    const Matrix X = M.project(random_matrix(r, k, 3));
    const Matrix V = random_matrix(r, k, 4);
    const Matrix P = M.projectToTangentSpace(X, V);
    double e_rot = 0.0, e_ob = 0.0, e_eu = 0.0, e_norm = 0.0;
    for (unsigned i = 0; i < n; ++i) e_rot = std::max(e_rot, stiefel_block_error(X, V, P, i * d, d));
    for (unsigned i = 0; i < l; ++i) {
      const unsigned c = d * n + i;
      double yv = 0.0, yy = 0.0;
      for (unsigned q = 0; q < r; ++q) { yv += X(q, c) * V(q, c); yy += X(q, c) * X(q, c); }
      e_norm = std::max(e_norm, std::fabs(yy - 1.0));
      for (unsigned q = 0; q < r; ++q)
        e_ob = std::max(e_ob, std::fabs(P(q, c) - (V(q, c) - X(q, c) * yv)));
    }
    for (unsigned c = d * n + l; c < k; ++c)
      for (unsigned q = 0; q < r; ++q) e_eu = std::max(e_eu, std::fabs(P(q, c) - V(q, c)));
    EXPECT(e_rot < 1e-12, "RA tangent projection, Stiefel blocks");
    EXPECT(e_ob < 1e-12, "RA tangent projection, oblique columns");
    EXPECT(e_eu == 0.0, "RA tangent projection, Euclidean columns");
    EXPECT(e_norm < 1e-12, "RA project(): unit spheres");
  }
  // ---------------- QuadraticProblem on a small ring pose graph
  {
    const unsigned d = 3, n = 12, r = 4;
    std::vector<DCORA::RelativePosePoseMeasurement> meas;
    std::mt19937_64 rng(7);
    std::normal_distribution<double> N(0.0, 1.0);
    for (unsigned i = 0; i < n; ++i)
      for (unsigned hop = 1; hop <= 2; ++hop) {
        DCORA::RelativePosePoseMeasurement m;
        m.p1 = i; m.p2 = (i + hop) % n; m.d = d;
        if (m.p1 == m.p2) continue;
        double q[4] = {N(rng), N(rng), N(rng), N(rng)};
        const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        DCORA::detail::quat_to_rot(q[0] / nq, q[1] / nq, q[2] / nq, q[3] / nq, m.R);
        for (double &t : m.t) t = N(rng);
        m.kappa = 50.0 + i; m.tau = 10.0 + hop;
        meas.push_back(m);
      }
    auto graph = std::make_shared<DCORA::Graph>(0, r, d);
    graph->setMeasurements(meas);
    DCORA::QuadraticProblem problem(graph);
    DCORA::LiftedSEManifold M(r, d, n);
    const Matrix X = M.project(random_matrix(r, (d + 1) * n, 5));
    // host reference: XQ and 0.5 <XQ, X> from the CSR matrix
    const DCORA::SparseMatrix &Q = graph->quadraticMatrix();
    Matrix XQ(r, Q.cols);
    for (int32_t row = 0; row < Q.rows; ++row)
      for (int32_t p = Q.rowptr[row]; p < Q.rowptr[row + 1]; ++p)
        for (unsigned q = 0; q < r; ++q) XQ(q, Q.col[p]) += X(q, row) * Q.val[p];
    double f_ref = 0.0, scale = 0.0;
    for (size_t j = 0; j < X.cols(); ++j)
      for (unsigned q = 0; q < r; ++q) { f_ref += 0.5 * XQ(q, j) * X(q, j); scale = std::max(scale, std::fabs(XQ(q, j))); }
    const double f_dev = problem.f(X);
    EXPECT(std::fabs(f_dev - f_ref) <= 1e-12 * std::fabs(f_ref), "QuadraticProblem::f == 0.5 <XQ, X>");
    const Matrix EG = problem.EucGrad(X);
    double e_eg = 0.0;
    for (size_t j = 0; j < X.cols(); ++j)
      for (unsigned q = 0; q < r; ++q) e_eg = std::max(e_eg, std::fabs(EG(q, j) - XQ(q, j)));
    EXPECT(e_eg <= 1e-12 * scale, "QuadraticProblem::EucGrad == X Q");
    const Matrix RG = problem.RieGrad(X), PR = M.projectToTangentSpace(X, XQ);
    double e_rg = 0.0;
    for (size_t j = 0; j < X.cols(); ++j)
      for (unsigned q = 0; q < r; ++q) e_rg = std::max(e_rg, std::fabs(RG(q, j) - PR(q, j)));
    EXPECT(e_rg <= 1e-12 * scale, "QuadraticProblem::RieGrad == proj(X Q)");
    EXPECT(std::fabs(problem.RieGradNorm(X) - RG.norm()) <= 1e-12 * RG.norm(), "RieGradNorm");
    // a converged solve is a critical point and lowers the cost (testPrior-style RTR check)
    DCORA::ROptParameters prm;
    prm.RTR_iterations = 100; prm.RTR_tCG_iterations = 200; prm.gradnorm_tol = 1e-8; prm.time_bound_s = 0;
    DCORA::QuadraticOptimizer opt(&problem, prm);
    const Matrix Xs = opt.optimize(X);
    const DCORA::ROPTResult res = opt.getOptResult();
    EXPECT(res.success && res.gradNormOpt < 1e-8 && res.fOpt < res.fInit, "RTR converges");
    EXPECT(std::fabs(problem.f(Xs) - res.fOpt) <= 1e-10 * std::fabs(res.fOpt) + 1e-14, "fOpt is f(Xopt)");
    EXPECT(std::fabs(res.fInit - f_ref) <= 1e-12 * std::fabs(f_ref), "fInit is f(X0)");
  }
  // ---------------- range-aided graph: noiseless measurements, ground truth is optimal
  // (what tests/testAgent.cpp:101-242 asserts on data/range_aided_slam_test_3d.pyfg)
  {
    const unsigned d = 3, n = 9, r = 5;
    std::mt19937_64 rng(11);
    std::normal_distribution<double> N(0.0, 1.0);
    std::vector<std::vector<double>> Rg(n, std::vector<double>(9)), tg(n, std::vector<double>(3));
    for (unsigned i = 0; i < n; ++i) {
      double q[4] = {N(rng), N(rng), N(rng), N(rng)};
      const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      DCORA::detail::quat_to_rot(q[0] / nq, q[1] / nq, q[2] / nq, q[3] / nq, Rg[i].data());
      for (double &t : tg[i]) t = 3.0 * N(rng);
    }
    std::vector<DCORA::RelativePosePoseMeasurement> pp;
    for (unsigned i = 0; i < n; ++i)
      for (unsigned hop = 1; hop <= 2; ++hop) {
        const unsigned j = (i + hop) % n;
        DCORA::RelativePosePoseMeasurement m;
        m.p1 = i; m.p2 = j; m.d = d; m.kappa = 40.0; m.tau = 7.0;
        for (unsigned a = 0; a < 3; ++a) {            // R_ij = R_i^T R_j, t_ij = R_i^T (t_j - t_i)
          for (unsigned c = 0; c < 3; ++c) {
            double sacc = 0.0;
            for (unsigned e = 0; e < 3; ++e) sacc += Rg[i][e * 3 + a] * Rg[j][e * 3 + c];
            m.R[a * 3 + c] = sacc;
          }
          double st = 0.0;
          for (unsigned e = 0; e < 3; ++e) st += Rg[i][e * 3 + a] * (tg[j][e] - tg[i][e]);
          m.t[a] = st;
        }
        pp.push_back(m);
      }
    std::vector<DCORA::RangeMeasurement> rgs;
    for (unsigned i = 0; i < n; i += 2) {
      DCORA::RangeMeasurement g2;
      g2.p1 = i; g2.p2 = (i + 4) % n; g2.precision = 25.0;
      double s2 = 0.0;
      for (unsigned e = 0; e < 3; ++e) s2 += (tg[g2.p1][e] - tg[g2.p2][e]) * (tg[g2.p1][e] - tg[g2.p2][e]);
      g2.range = std::sqrt(s2);
      rgs.push_back(g2);
    }
    auto graph = std::make_shared<DCORA::Graph>(0, r, d, DCORA::GraphType::RangeAidedSLAMGraph);
    graph->setMeasurements(pp, rgs);
    DCORA::QuadraticProblem problem(graph);
    Matrix X(r, graph->k());
    for (unsigned i = 0; i < n; ++i) {
      for (unsigned c = 0; c < 3; ++c)
        for (unsigned e = 0; e < 3; ++e) X(e, graph->rotationColumn(i, c)) = Rg[i][e * 3 + c];
      for (unsigned e = 0; e < 3; ++e) X(e, graph->translationColumn(i)) = tg[i][e];
    }
    for (size_t k2 = 0; k2 < rgs.size(); ++k2)
      for (unsigned e = 0; e < 3; ++e)
        X(e, graph->unitSphereColumn((unsigned)k2)) =
            (tg[rgs[k2].p1][e] - tg[rgs[k2].p2][e]) / rgs[k2].range;
    EXPECT(graph->k() == 4 * n + rgs.size() && !problem.useSEManifold(), "RA graph dimensions");
    EXPECT(std::fabs(problem.f(X)) < 1e-9, "RA graph: f(ground truth) = 0");
    EXPECT(problem.RieGradNorm(X) < 1e-7, "RA graph: ground truth is a critical point");
    DCORA::ROptParameters prm;
    prm.time_bound_s = 0;
    DCORA::QuadraticOptimizer opt(&problem, prm);
    const Matrix Xs = opt.optimize(X);
    double e_fix = 0.0;
    for (size_t j = 0; j < X.cols(); ++j)
      for (unsigned q = 0; q < r; ++q) e_fix = std::max(e_fix, std::fabs(Xs(q, j) - X(q, j)));
    EXPECT(e_fix < 1e-9, "RA graph: optimize() keeps the ground truth fixed");
    double theta = 0.0;
    DCORA::Vector v;
    DCORA::CertificateInfo info;
    const bool ok = DCORA::fastVerification(problem, X, 1e-4, &theta, &v, &info);
    EXPECT(ok && info.lambda_min > -1e-6, "RA graph: ground truth is certified");
  }
  std::printf(failures ? "FAILED %d checks\n" : "all checks passed\n", failures);
  return failures ? 1 : 0;
}
