// Multi-process RBCD / RBCD++ on a PyFG dataset, C++ host side, one process per GPU without any
// Python: the loop of examples/MultiRobotExample_RASLAM.cpp:290-452 with the agents of a process
// batched on its GPU (DCORA::DistributedAgentBatch) and the public states exchanged by
// ncclAllGather.  The processes are forked here (no mpirun / torchrun needed); rank 0 creates
// the NCCL id and hands it to the others through shared memory.
//
//   rbcd_ra <file.pyfg> [--ranks N] [--rank R] [--iters I] [--tol T] [--acc]
//           [--nesterov-blocks B] [--restart I] [--x0 x0.bin] [--out x.bin]
// --nesterov-blocks -1: the number of colours as block count of RBCD++ (a colour class moves per
// iteration); 0 (default): the number of agents as in the reference.
// x0.bin / x.bin: r x K doubles, column-major, agent-major reference ordering (default start:
// the file's ground truth lifted with [I; 0]).  Prints one JSON line per iteration (rank 0).
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstring>

#include "dcora_b200/DCORA_dist.hpp"

struct Shared {
  ncclUniqueId id;
  std::atomic<int> ready;
};

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s file.pyfg [--ranks N] [--rank R] [--iters I] [--tol T] [--acc] "
                         "[--x0 file] [--out file]\n", argv[0]);
    return 2;
  }
  int world = 1;
  unsigned r = 5, iters = 10;
  double tol = 0.1;   // RGradNormTol (MultiRobotExample_RASLAM.cpp:101)
  bool acc = false;
  int nesterov_blocks = 0;          // 0: agents (reference), -1: colours, > 0: given
  unsigned restart_interval = 30;   // AgentParameters::restartInterval
  std::string x0_file, out_file;
  for (int i = 2; i < argc; ++i) {
    auto next = [&]() { return i + 1 < argc ? argv[++i] : "0"; };
    if (!std::strcmp(argv[i], "--ranks")) world = std::atoi(next());
    else if (!std::strcmp(argv[i], "--rank")) r = std::atoi(next());
    else if (!std::strcmp(argv[i], "--iters")) iters = std::atoi(next());
    else if (!std::strcmp(argv[i], "--tol")) tol = std::atof(next());
    else if (!std::strcmp(argv[i], "--acc")) acc = true;
    else if (!std::strcmp(argv[i], "--nesterov-blocks")) nesterov_blocks = std::atoi(next());
    else if (!std::strcmp(argv[i], "--restart")) restart_interval = std::atoi(next());
    else if (!std::strcmp(argv[i], "--x0")) x0_file = next();
    else if (!std::strcmp(argv[i], "--out")) out_file = next();
  }
  // host-only work before the fork (no CUDA context yet)
  const DCORA::GlobalMeasurements meas = DCORA::read_pyfg_file(argv[1]);
  DCORA::PartitionedProblem problem(meas);
  const unsigned d = meas.d;
  const int64_t K = problem.K();
  DCORA_B200_CHECK(world >= 1 && world <= problem.numAgents(), "more processes than agents");
  DCORA::Matrix X0(r, (size_t)K);
  if (!x0_file.empty()) {
    std::ifstream in(x0_file, std::ios::binary);
    in.read(reinterpret_cast<char *>(X0.data()), sizeof(double) * X0.size());
    DCORA_B200_CHECK((size_t)in.gcount() == sizeof(double) * X0.size(), "x0 has the wrong size");
  } else {
    for (unsigned i = 0; i < meas.n; ++i) {
      for (unsigned c = 0; c < d; ++c)
        for (unsigned a = 0; a < d; ++a) X0(a, problem.rotCol(i, c)) = meas.gt_rotation[i * d * d + a * d + c];
      for (unsigned a = 0; a < d; ++a) X0(a, problem.transCol(i)) = meas.gt_translation[i * d + a];
    }
    for (unsigned j = 0; j < meas.b; ++j)
      for (unsigned a = 0; a < d; ++a) X0(a, problem.lmkCol(j)) = meas.gt_landmark[j * d + a];
    for (size_t e = 0; e < meas.ranges.size(); ++e) {
      const auto &rg = meas.ranges[e];
      const double *p1 = rg.landmark1 ? &meas.gt_landmark[rg.p1 * d] : &meas.gt_translation[rg.p1 * d];
      const double *p2 = rg.landmark2 ? &meas.gt_landmark[rg.p2 * d] : &meas.gt_translation[rg.p2 * d];
      double nrm = 0.0;
      for (unsigned a = 0; a < d; ++a) nrm += (p1[a] - p2[a]) * (p1[a] - p2[a]);
      nrm = std::sqrt(nrm);
      for (unsigned a = 0; a < d; ++a) X0(a, problem.sphereCol(e)) = (p1[a] - p2[a]) / nrm;
    }
  }
  if (!out_file.empty()) {   // sized once; every rank writes its own columns
    FILE *f = std::fopen(out_file.c_str(), "wb");
    DCORA_B200_CHECK(f != nullptr, "cannot create " + out_file);
    std::fclose(f);
    DCORA_B200_CHECK(truncate(out_file.c_str(), (off_t)(sizeof(double) * X0.size())) == 0, "truncate");
  }
  Shared *sh = static_cast<Shared *>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE,
                                          MAP_SHARED | MAP_ANONYMOUS, -1, 0));
  DCORA_B200_CHECK(sh != MAP_FAILED, "mmap");
  new (&sh->ready) std::atomic<int>(0);
  int rank = 0;
  std::vector<pid_t> children;
  for (int k = 1; k < world; ++k) {
    const pid_t pid = fork();
    DCORA_B200_CHECK(pid >= 0, "fork");
    if (pid == 0) { rank = k; children.clear(); break; }
    children.push_back(pid);
  }
  // ---- one process per GPU from here on
  DCORA_B200_CUDA(cudaSetDevice(rank));
  ncclComm_t comm = nullptr;
  if (world > 1) {
    if (rank == 0) {
      DCORA_B200_NCCL(ncclGetUniqueId(&sh->id));
      sh->ready.store(1, std::memory_order_release);
    } else {
      while (sh->ready.load(std::memory_order_acquire) == 0) usleep(1000);
    }
    DCORA_B200_NCCL(ncclCommInitRank(&comm, world, sh->id, rank));
  }
  int rc = 0;
  {
    const auto t0 = std::chrono::steady_clock::now();
    DCORA::DistributedAgentBatch agents(problem, r, rank, world, rank, comm, acc, restart_interval,
                                        nesterov_blocks);
    agents.setX(X0);
    const double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double cost = 0.0, gradnorm = 0.0;
    unsigned it = 0;
    const auto t1 = std::chrono::steady_clock::now();
    for (; it < iters; ++it) {
      agents.iterate();
      agents.evaluate(&cost, &gradnorm);
      if (rank == 0)
        std::printf("{\"iter\": %u, \"cost\": %.17g, \"gradnorm\": %.17g}\n", it, cost, gradnorm);
      if (gradnorm < tol) break;
    }
    const double solve_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    if (rank == 0)
      std::printf("{\"summary\": true, \"ranks\": %d, \"agents\": %d, \"colors\": %d, \"K\": %lld, "
                  "\"iterations\": %u, \"cost\": %.17g, \"gradnorm\": %.17g, \"setup_s\": %.4f, "
                  "\"solve_s\": %.4f}\n",
                  world, problem.numAgents(), agents.plan().num_colors, (long long)K,
                  std::min(it + 1, iters), cost, gradnorm, setup_s, solve_s);
    if (!out_file.empty()) {
      const DCORA::Matrix Xl = agents.getX();
      FILE *f = std::fopen(out_file.c_str(), "r+b");
      DCORA_B200_CHECK(f != nullptr, "cannot open " + out_file);
      std::fseek(f, (long)(sizeof(double) * (size_t)agents.plan().c0 * r), SEEK_SET);
      std::fwrite(Xl.data(), sizeof(double), Xl.size(), f);
      std::fclose(f);
    }
  }
  if (comm) ncclCommDestroy(comm);
  std::fflush(stdout);
  if (rank != 0) _exit(0);
  for (pid_t pid : children) {
    int st = 0;
    waitpid(pid, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 3;
  }
  return rc;
}
