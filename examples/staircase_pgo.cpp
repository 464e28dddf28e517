// Single-agent Riemannian staircase with certification on a g2o pose graph
// (BASELINE.json configs[2]: sphere2500 / torus3D / city10000), written against
// the C++ host mirror of the reference classes (include/dcora_b200/DCORA.hpp).
//
// Control flow of examples/MultiRobotExample.cpp:121-364 for one robot (where
// RBCD degenerates to repeated QuadraticOptimizer::optimize calls): random
// feasible start at r_min, solve at rank r, build the dual certificate and test
// it (fastVerification), on failure escapeSaddle to rank r+1.
//
//   staircase_pgo <file.g2o> [--rmin 5] [--rmax 10] [--eta 1e-3] [--tol 1e-6]
//                 [--rtr 300] [--tcg 500] [--seed 1] [--check-q]
// Prints one JSON line per staircase level and a final summary line.
// --check-q only parses the file and assembles Q on the host (no GPU needed).
#include <chrono>
#include <cstring>
#include <random>

#include "dcora_b200/DCORA.hpp"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s file.g2o [--rmin R] [--rmax R] [--eta E] [--tol T] [--rtr N] "
                         "[--tcg N] [--seed S] [--check-q]\n", argv[0]);
    return 2;
  }
  unsigned r_min = 5, r_max = 10, seed = 1;
  int rtr = 300, tcg = 500;
  double eta = 1e-3, tol = 1e-6;
  bool check_q = false;
  for (int i = 2; i < argc; ++i) {
    auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
    if (!std::strcmp(argv[i], "--rmin")) r_min = std::atoi(next());
    else if (!std::strcmp(argv[i], "--rmax")) r_max = std::atoi(next());
    else if (!std::strcmp(argv[i], "--eta")) eta = std::atof(next());
    else if (!std::strcmp(argv[i], "--tol")) tol = std::atof(next());
    else if (!std::strcmp(argv[i], "--rtr")) rtr = std::atoi(next());
    else if (!std::strcmp(argv[i], "--tcg")) tcg = std::atoi(next());
    else if (!std::strcmp(argv[i], "--seed")) seed = std::atoi(next());
    else if (!std::strcmp(argv[i], "--check-q")) check_q = true;
  }
  const double t_start = now_s();
  size_t n = 0;
  const auto measurements = DCORA::read_g2o_file(argv[1], &n);
  DCORA_B200_CHECK(!measurements.empty(), "no edges");
  const unsigned d = measurements[0].d;
  const double t_read = now_s();

  if (check_q) {
    auto g = std::make_shared<DCORA::Graph>(0, r_min, d);
    g->setMeasurements(measurements);
    const DCORA::SparseMatrix &Q = g->quadraticMatrix();
    double fro = 0.0, chk = 0.0;
    for (int32_t i = 0; i < Q.rows; ++i)
      for (int32_t p = Q.rowptr[i]; p < Q.rowptr[i + 1]; ++p) {
        fro += Q.val[p] * Q.val[p];
        chk += Q.val[p] * std::cos(0.001 * i + 0.002 * Q.col[p]);
      }
    std::printf("{\"n\": %zu, \"d\": %u, \"m\": %zu, \"k\": %d, \"nnz\": %lld, \"fro\": %.17g, "
                "\"checksum\": %.17g}\n", n, d, measurements.size(), Q.rows,
                (long long)Q.nonZeros(), std::sqrt(fro), chk);
    return 0;
  }

  // random feasible start (MultiRobotExample.cpp:161-164: Matrix::Random + projectToSEMatrix)
  const unsigned k = (d + 1) * (unsigned)n;
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  DCORA::Matrix M(r_min, k);
  for (size_t j = 0; j < k; ++j)
    for (size_t i = 0; i < r_min; ++i) M(i, j) = U(rng);

  auto graph = std::make_shared<DCORA::Graph>(0, r_min, d);
  graph->setMeasurements(measurements);
  std::unique_ptr<DCORA::QuadraticProblem> problem(new DCORA::QuadraticProblem(graph));
  const double t_setup = now_s();
  DCORA::Matrix X = DCORA::LiftedSEManifold(r_min, d, (unsigned)n).project(M);

  DCORA::ROptParameters params;
  params.RTR_iterations = rtr;
  params.RTR_tCG_iterations = tcg;
  params.gradnorm_tol = tol;
  params.time_bound_s = 0.0;
  bool certified = false;
  double cost = 0.0, gradnorm = 0.0, lambda_min = 0.0;
  unsigned r = r_min;
  for (; r < r_max; ++r) {
    const double t0 = now_s();
    DCORA::QuadraticOptimizer optimizer(problem.get(), params);
    DCORA::Matrix Xopt = optimizer.optimize(X);
    const DCORA::ROPTResult res = optimizer.getOptResult();
    const double t1 = now_s();
    double theta = 0.0;
    DCORA::Vector v;
    DCORA::CertificateInfo info;
    const bool global_opt = DCORA::fastVerification(*problem, Xopt, eta, &theta, &v, &info);
    const double t2 = now_s();
    cost = 2.0 * res.fOpt;
    gradnorm = res.gradNormOpt;
    lambda_min = info.lambda_min;
    std::printf("{\"r\": %u, \"cost_2f\": %.12f, \"gradnorm\": %.3e, \"outer\": %d, \"inner\": %d, "
                "\"is_psd\": %s, \"theta\": %.6e, \"lambda_min\": %.6e, \"matvecs\": %d, "
                "\"solve_s\": %.4f, \"certify_s\": %.4f}\n",
                r, cost, gradnorm, res.outerIterations, res.innerIterations,
                global_opt ? "true" : "false", theta, lambda_min, info.matvecs, t1 - t0, t2 - t1);
    std::fflush(stdout);
    X = Xopt;
    if (global_opt) {
      certified = true;
      break;
    }
    // MultiRobotExample.cpp:336-339: the escape direction must have converged
    DCORA_B200_CHECK(theta < -eta / 2, "escape direction did not converge to the desired precision");
    std::unique_ptr<DCORA::QuadraticProblem> next(new DCORA::QuadraticProblem(*problem, r + 1));
    DCORA::Matrix Xnext;
    if (!next->escapeSaddle(Xopt, theta, v, tol, tol, &Xnext)) break;
    X = Xnext;
    problem = std::move(next);
  }
  const double t_end = now_s();
  std::printf("{\"summary\": true, \"file\": \"%s\", \"n\": %zu, \"d\": %u, \"certified\": %s, "
              "\"rank\": %u, \"cost_2f\": %.12f, \"gradnorm\": %.3e, \"lambda_min\": %.6e, "
              "\"read_s\": %.4f, \"setup_s\": %.4f, \"time_to_certified_s\": %.4f}\n",
              argv[1], n, d, certified ? "true" : "false", r, cost, gradnorm, lambda_min,
              t_read - t_start, t_setup - t_read, t_end - t_start);
  return certified ? 0 : 1;
}
