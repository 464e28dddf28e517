// Host-only check of DCORA::Graph's data matrix (pose-graph and range-aided): reads a plain
// text measurement list written by the test-suite from a golden fixture
//   D <d>
//   PP <i> <j> <R row-major d*d> <t d> <kappa> <tau> <weight>
//   RG <src> <dst> <src_is_landmark> <dst_is_landmark> <range> <precision> <weight>   (in unit-sphere order)
//   LM <num_landmarks>
// builds Q with DCORA::Graph and prints {k, nnz, fro, checksum} for comparison with the oracle.
#include <fstream>
#include <sstream>

#include "dcora_b200/DCORA.hpp"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  DCORA_B200_CHECK(in.good(), "cannot open input");
  unsigned d = 3, nlm = 0;
  std::vector<DCORA::RelativePosePoseMeasurement> pp;
  std::vector<DCORA::RangeMeasurement> rg;
  std::string line, tok;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    if (!(ss >> tok)) continue;
    if (tok == "D") ss >> d;
    else if (tok == "LM") ss >> nlm;
    else if (tok == "PP") {
      DCORA::RelativePosePoseMeasurement m;
      m.d = d;
      ss >> m.p1 >> m.p2;
      for (unsigned i = 0; i < d * d; ++i) ss >> m.R[i];
      for (unsigned i = 0; i < d; ++i) ss >> m.t[i];
      ss >> m.kappa >> m.tau >> m.weight;
      pp.push_back(m);
    } else if (tok == "RG") {
      DCORA::RangeMeasurement g;
      int l1, l2;
      ss >> g.p1 >> g.p2 >> l1 >> l2 >> g.range >> g.precision >> g.weight;
      g.landmark1 = l1 != 0; g.landmark2 = l2 != 0;
      rg.push_back(g);
    }
    DCORA_B200_CHECK(!ss.fail(), "malformed line: " + line);
  }
  const bool ra = !rg.empty() || nlm > 0;
  DCORA::Graph graph(0, d + 1, d, ra ? DCORA::GraphType::RangeAidedSLAMGraph : DCORA::GraphType::PoseGraph);
  if (ra) graph.setMeasurements(pp, rg, nlm); else graph.setMeasurements(pp);
  const DCORA::SparseMatrix &Q = graph.quadraticMatrix();
  double fro = 0.0, chk = 0.0;
  for (int32_t i = 0; i < Q.rows; ++i)
    for (int32_t p = Q.rowptr[i]; p < Q.rowptr[i + 1]; ++p) {
      fro += Q.val[p] * Q.val[p];
      chk += Q.val[p] * std::cos(0.001 * i + 0.002 * Q.col[p]);
    }
  std::printf("{\"n\": %u, \"l\": %u, \"b\": %u, \"k\": %d, \"nnz\": %lld, \"fro\": %.17g, \"checksum\": %.17g}\n",
              graph.n(), graph.l(), graph.b(), Q.rows, (long long)Q.nonZeros(), std::sqrt(fro), chk);
  return 0;
}
