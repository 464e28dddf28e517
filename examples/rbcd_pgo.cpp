// Multi-agent RBCD on a g2o pose graph, C++ host side: the loop of
// examples/MultiRobotExample.cpp:223-307 with the agents of the process batched on one GPU
// (DCORA::AgentBatch, include/dcora_b200/DCORA.hpp).  The reference updates one greedily
// selected agent per iteration; here every iteration sweeps the colours of the agent graph and
// updates all agents of a colour at once (non-adjacent agents: each colour step is an exact
// block-coordinate descent step).
//
//   rbcd_pgo <file.g2o> <num_robots> [--rank 5] [--iters 200] [--tol 0.1] [--seed 1]
// Prints one JSON line per monitored iteration and a final summary line.
#include <chrono>
#include <cstring>
#include <random>

#include "dcora_b200/DCORA.hpp"

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s file.g2o num_robots [--rank R] [--iters N] [--tol T] [--seed S]\n", argv[0]);
    return 2;
  }
  const unsigned num_robots = std::atoi(argv[2]);
  unsigned r = 5, iters = 200, seed = 1;
  double tol = 0.1;  // RGradNormTol of the reference driver (MultiRobotExample.cpp:133)
  for (int i = 3; i < argc; ++i) {
    auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
    if (!std::strcmp(argv[i], "--rank")) r = std::atoi(next());
    else if (!std::strcmp(argv[i], "--iters")) iters = std::atoi(next());
    else if (!std::strcmp(argv[i], "--tol")) tol = std::atof(next());
    else if (!std::strcmp(argv[i], "--seed")) seed = std::atoi(next());
  }
  size_t n = 0;
  const auto measurements = DCORA::read_g2o_file(argv[1], &n);
  DCORA_B200_CHECK(!measurements.empty() && num_robots >= 1 && num_robots <= n, "bad input");
  const unsigned d = measurements[0].d;
  // contiguous partition: n / num_robots poses each, the last robot takes the rest
  // (MultiRobotExample.cpp:56-83)
  const unsigned per = (unsigned)(n / num_robots);
  std::vector<unsigned> poses_per_agent(num_robots, per);
  poses_per_agent.back() = (unsigned)n - per * (num_robots - 1);

  const auto t0 = std::chrono::steady_clock::now();
  auto graph = std::make_shared<DCORA::Graph>(0, r, d);
  graph->setMeasurements(measurements);
  DCORA::AgentBatch agents(graph->quadraticMatrix(), poses_per_agent, r, d);
  const std::vector<int> color = agents.coloring();
  int ncolors = 0;
  for (int c : color) ncolors = std::max(ncolors, c + 1);

  // random feasible start (MultiRobotExample.cpp:161-164)
  const unsigned k = (d + 1) * (unsigned)n;
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  DCORA::Matrix M(r, k);
  for (size_t j = 0; j < k; ++j)
    for (size_t i = 0; i < r; ++i) M(i, j) = U(rng);
  DCORA::Matrix X = DCORA::LiftedSEManifold(r, d, (unsigned)n).project(M);
  const double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  DCORA::ROptParameters params;   // reference defaults: 3 RTR x 50 tCG, tol 1e-2
  params.time_bound_s = 0.0;
  double cost = 0.0, gradnorm = 0.0;
  unsigned it = 0;
  const auto t1 = std::chrono::steady_clock::now();
  for (; it < iters; ++it) {
    for (int c = 0; c < ncolors; ++c) {
      std::vector<uint8_t> active(num_robots, 0);
      for (unsigned a = 0; a < num_robots; ++a) active[a] = color[a] == c;
      agents.updateLinearTerms(X, active);
      agents.optimize(&X, active, params);
    }
    agents.updateLinearTerms(X, {});
    agents.evaluate(X, &cost, &gradnorm);
    if (it % 10 == 0 || gradnorm < tol)
      std::printf("{\"iter\": %u, \"cost_2f\": %.9f, \"gradnorm\": %.6e}\n", it, 2 * cost, gradnorm);
    if (gradnorm < tol) break;   // MultiRobotExample.cpp:283-286
  }
  const double solve_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
  std::printf("{\"summary\": true, \"robots\": %u, \"colors\": %d, \"iterations\": %u, \"cost_2f\": %.9f, "
              "\"gradnorm\": %.6e, \"setup_s\": %.4f, \"solve_s\": %.4f}\n",
              num_robots, ncolors, it + 1, 2 * cost, gradnorm, setup_s, solve_s);
  return gradnorm < tol ? 0 : 1;
}
