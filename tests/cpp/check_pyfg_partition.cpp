// Host-only check of the C++ PyFG reader, the partitioned data matrices and the exchange plans
// (include/dcora_b200/DCORA_dist.hpp): prints one JSON line that tests/test_cpp_dist.py compares
// with the Python twin (dcora_b200/graph.py, rbcd.py) built from the same file.
//   check_pyfg_partition <file.pyfg> <world>
#include "dcora_b200/DCORA_dist.hpp"

static void stats(const DCORA::SparseMatrix &Q, double *fro, double *chk) {
  double f = 0.0, c = 0.0;
  for (int32_t i = 0; i < Q.rows; ++i)
    for (int32_t p = Q.rowptr[i]; p < Q.rowptr[i + 1]; ++p) {
      f += Q.val[p] * Q.val[p];
      c += Q.val[p] * std::cos(0.001 * i + 0.002 * Q.col[p]);
    }
  *fro = std::sqrt(f);
  *chk = c;
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const int world = std::atoi(argv[2]);
  const DCORA::GlobalMeasurements M = DCORA::read_pyfg_file(argv[1]);
  DCORA::PartitionedProblem P(M);
  double fro, chk;
  stats(P.Q(), &fro, &chk);
  std::printf("{\"d\": %u, \"n\": %u, \"b\": %u, \"l\": %zu, \"robots\": %u, \"K\": %lld, \"nnz\": %lld, "
              "\"fro\": %.17g, \"checksum\": %.17g, \"plans\": [",
              M.d, M.n, M.b, M.ranges.size(), M.num_robots, (long long)P.K(),
              (long long)P.Q().nonZeros(), fro, chk);
  for (int r = 0; r < world; ++r) {
    const DCORA::RankPlan plan = DCORA::makePlan(P, world, r);
    double cf, cc, df = 0.0;
    stats(plan.coupling, &cf, &cc);
    for (const auto &D : plan.diag) { double a, b2; stats(D, &a, &b2); df += a * a; }
    long long pack_sum = 0;
    for (int32_t v : plan.pack_idx) pack_sum += v;
    long long key_sum = 0;
    for (int64_t v : plan.row_key) key_sum += v;
    std::printf("%s{\"a0\": %d, \"a1\": %d, \"c0\": %lld, \"K\": %lld, \"seg\": [%d, %d, %d], "
                "\"coupling_rows\": %d, \"coupling_nnz\": %lld, \"coupling_fro\": %.17g, "
                "\"coupling_checksum\": %.17g, \"diag_fro\": %.17g, \"pack_sum\": %lld, "
                "\"key_sum\": %lld, \"num_colors\": %d, \"colors\": [",
                r ? ", " : "", plan.a0, plan.a1, (long long)plan.c0, (long long)plan.K, plan.seg.n,
                plan.seg.l, plan.seg.b, plan.coupling.rows, (long long)plan.coupling.nonZeros(), cf, cc,
                std::sqrt(df), pack_sum, key_sum, plan.num_colors);
    for (size_t a = 0; a < plan.colors.size(); ++a) std::printf("%s%d", a ? ", " : "", plan.colors[a]);
    std::printf("]}");
  }
  std::printf("]}\n");
  return 0;
}
