// dcora_b200 -- C++ host side of the MULTI-PROCESS path (one process per GPU).
//
// What a maintainer of the reference needs next to DCORA.hpp to run the drivers
// (examples/MultiRobotExample_RASLAM.cpp:204-452) on several GPUs without Python:
//   * read_pyfg_file            the PyFG reader (src/DCORA_utils.cpp:437-1167) and the
//                               robot-major global numbering of getGlobalMeasurements /
//                               getLocalToGlobalStateMapping (:1169-1352)
//   * PartitionedProblem        the agents' data matrices and the cross-agent coupling of a
//                               partitioned pose-graph / range-aided problem
//                               (Graph::constructQuadraticCostTerm*, src/Graph.cpp:579-1188)
//   * makePlans                 which public states every process sends and where the received
//                               ones land (Graph::myPublic*IDs, Agent::getSharedStateDicts /
//                               updateNeighborStates, src/Agent.cpp:113-152,844-906)
//   * DistributedAgentBatch     the agents of one process as one batched device handle; public
//                               states travel by ncclAllGather on the handle's stream; RBCD and
//                               RBCD++ (Agent::iterate, updateX/Y/V, restart: Agent.cpp:535-596,
//                               1158-1278) over the colours of the agent graph.
// NCCL is linked directly (no torch).  The Python twin is dcora_b200/rbcd.py; both produce the
// same iterates (tests/test_cpp_dist.py).
#ifndef DCORA_B200_DCORA_DIST_HPP_
#define DCORA_B200_DCORA_DIST_HPP_

#include <cuda_runtime.h>
#include <nccl.h>

#include <map>
#include <numeric>
#include <set>

#include "dcora_b200/DCORA.hpp"

namespace DCORA {

#define DCORA_B200_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      std::fprintf(stderr, "dcora_b200: %s: %s\n", #expr, cudaGetErrorString(e__));         \
      std::abort();                                                                        \
    }                                                                                      \
  } while (0)
#define DCORA_B200_NCCL(expr)                                                              \
  do {                                                                                     \
    ncclResult_t e__ = (expr);                                                             \
    if (e__ != ncclSuccess) {                                                              \
      std::fprintf(stderr, "dcora_b200: %s: %s\n", #expr, ncclGetErrorString(e__));         \
      std::abort();                                                                        \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------ measurements
struct RelativePoseLandmarkMeasurement {  // include/DCORA/Measurements.h (pose -> landmark)
  unsigned p1 = 0, p2 = 0;                // pose, landmark
  double t[3] = {0, 0, 0};
  double tau = 1, weight = 1;
};

// All measurements of a dataset in GLOBAL numbering (robot-major: the poses of robot A first,
// in frame order, then robot B, ...; landmarks likewise) + the owner of every variable.
struct GlobalMeasurements {
  unsigned d = 0, n = 0, b = 0, num_robots = 0;
  std::vector<RelativePosePoseMeasurement> pose_pose;
  std::vector<RelativePoseLandmarkMeasurement> pose_landmark;
  std::vector<RangeMeasurement> ranges;   // endpoints: pose or landmark numbers (flags)
  std::vector<unsigned> robot_of_pose, robot_of_landmark;
  // ground truth of the VERTEX lines: rotations row-major d*d per pose, translations, landmarks
  std::vector<double> gt_rotation, gt_translation, gt_landmark;
  bool rangeAided() const { return !ranges.empty() || !pose_landmark.empty() || b > 0; }
};

namespace detail {
struct PyfgSymbol {
  unsigned robot = 0, frame = 0;
  bool landmark = false;
};
// getRobotAndStateIDFromSymbol / getStateTypeFromSymbol (DCORA_utils.cpp:584-628):
// "A12" pose 12 of robot A; "LB3" landmark 3 of robot B; "L7" landmark 7 of the map.
inline PyfgSymbol parse_symbol(const std::string &s) {
  PyfgSymbol o;
  DCORA_B200_CHECK(s.size() >= 2, "bad PyFG symbol " + s);
  size_t pos = 1;
  if (s[0] == 'L') {
    o.landmark = true;
    if (std::isupper((unsigned char)s[1])) { o.robot = (unsigned)(s[1] - 'A'); pos = 2; }
    else o.robot = (unsigned)('M' - 'A');   // MAP_ID (DCORA_types.h:43-44)
  } else {
    DCORA_B200_CHECK(std::isupper((unsigned char)s[0]), "bad PyFG symbol " + s);
    o.robot = (unsigned)(s[0] - 'A');
  }
  o.frame = (unsigned)std::stoul(s.substr(pos));
  return o;
}
// upper triangle (row by row) of a symmetric dim x dim matrix -> trace of the leading /
// trailing diagonal block
inline void read_upper(std::istringstream &ss, int dim, std::vector<double> *M) {
  M->assign((size_t)dim * dim, 0.0);
  for (int i = 0; i < dim; ++i)
    for (int j = i; j < dim; ++j) {
      double v;
      ss >> v;
      (*M)[i * dim + j] = (*M)[j * dim + i] = v;
    }
}
}  // namespace detail

// read_pyfg_file (DCORA_utils.cpp:437-1167) + getGlobalMeasurements (:1230-1352): precisions
// tau = d / tr(cov_t), kappa = 1 / cov_theta (2-D) or 3 / (2 tr(cov_R)) (3-D), range precision
// = 1 / cov; duplicate ranges are skipped (:1078-1086); priors are parsed and dropped (:1244).
inline GlobalMeasurements read_pyfg_file(const std::string &filename) {
  std::ifstream in(filename);
  DCORA_B200_CHECK(in.good(), "cannot open " + filename);
  struct PoseRec { std::vector<double> R, t; };
  std::map<std::pair<unsigned, unsigned>, PoseRec> poses;
  std::map<std::pair<unsigned, unsigned>, std::vector<double>> lmks;
  struct PP { detail::PyfgSymbol a, b; RelativePosePoseMeasurement m; };
  struct PL { detail::PyfgSymbol a, b; RelativePoseLandmarkMeasurement m; };
  struct RG { detail::PyfgSymbol a, b; double range, prec; };
  std::vector<PP> pps;
  std::vector<PL> pls;
  std::vector<RG> rgs;
  std::set<std::tuple<bool, unsigned, unsigned, bool, unsigned, unsigned>> seen;
  unsigned d = 0;
  std::string line, ty, ts, s1, s2;
  std::vector<double> cov;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    if (!(ss >> ty)) continue;
    if (d == 0) {  // getDimFromPyfgFirstLine (:376-435)
      DCORA_B200_CHECK(ty != "EDGE_RANGE", "first PyFG line carries no dimension");
      d = (ty.find("SE3") != std::string::npos || ty.find("XYZ") != std::string::npos) ? 3 : 2;
    }
    auto read_pose = [&](double *R, double *t) {
      if (d == 2) {
        double th;
        ss >> t[0] >> t[1] >> th;
        R[0] = std::cos(th); R[1] = -std::sin(th); R[2] = std::sin(th); R[3] = std::cos(th);
      } else {
        double q[4];
        ss >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3];
        detail::quat_to_rot(q[0], q[1], q[2], q[3], R);
      }
    };
    auto precisions = [&](double *kappa, double *tau) {
      const int dim = d == 2 ? 3 : 6;
      detail::read_upper(ss, dim, &cov);
      double trt = 0.0, trr = 0.0;
      for (unsigned i = 0; i < d; ++i) trt += cov[i * dim + i];
      for (int i = (int)d; i < dim; ++i) trr += cov[i * dim + i];
      *tau = d / trt;
      *kappa = d == 2 ? 1.0 / trr : 3.0 / (2.0 * trr);
    };
    if (ty == "VERTEX_SE2" || ty == "VERTEX_SE3:QUAT") {
      ss >> ts >> s1;
      const auto a = detail::parse_symbol(s1);
      PoseRec rec;
      rec.R.assign(d * d, 0.0); rec.t.assign(d, 0.0);
      read_pose(rec.R.data(), rec.t.data());
      DCORA_B200_CHECK(poses.emplace(std::make_pair(a.robot, a.frame), rec).second,
                       "duplicate pose " + s1);
    } else if (ty == "VERTEX_XY" || ty == "VERTEX_XYZ") {
      ss >> s1;
      const auto a = detail::parse_symbol(s1);
      std::vector<double> p(d);
      for (double &v : p) ss >> v;
      DCORA_B200_CHECK(lmks.emplace(std::make_pair(a.robot, a.frame), p).second,
                       "duplicate landmark " + s1);
    } else if (ty.find(":PRIOR") != std::string::npos) {
      continue;  // parsed by the reference, dropped by getGlobalMeasurements
    } else if (ty == "EDGE_SE2" || ty == "EDGE_SE3:QUAT") {
      PP e;
      ss >> ts >> s1 >> s2;
      e.a = detail::parse_symbol(s1); e.b = detail::parse_symbol(s2);
      e.m.d = d;
      read_pose(e.m.R, e.m.t);
      precisions(&e.m.kappa, &e.m.tau);
      pps.push_back(e);
    } else if (ty == "EDGE_SE2_XY" || ty == "EDGE_SE3_XYZ") {
      PL e;
      ss >> ts >> s1 >> s2;
      e.a = detail::parse_symbol(s1); e.b = detail::parse_symbol(s2);
      for (unsigned i = 0; i < d; ++i) ss >> e.m.t[i];
      detail::read_upper(ss, (int)d, &cov);
      double tr = 0.0;
      for (unsigned i = 0; i < d; ++i) tr += cov[i * d + i];
      e.m.tau = d / tr;
      pls.push_back(e);
    } else if (ty == "EDGE_RANGE") {
      RG e;
      double c;
      ss >> ts >> s1 >> s2 >> e.range >> c;
      e.a = detail::parse_symbol(s1); e.b = detail::parse_symbol(s2);
      e.prec = 1.0 / c;
      DCORA_B200_CHECK(e.range > 0.0, "range must be positive");
      if (!seen.emplace(e.a.landmark, e.a.robot, e.a.frame, e.b.landmark, e.b.robot, e.b.frame).second)
        continue;
      rgs.push_back(e);
    } else {
      DCORA_B200_CHECK(false, "unknown PyFG item " + ty);
    }
    DCORA_B200_CHECK(!ss.fail(), "malformed PyFG line: " + line);
  }
  // robot-major global numbering, robots compacted in id order
  GlobalMeasurements G;
  G.d = d;
  std::map<unsigned, unsigned> robot_index;
  for (const auto &kv : poses) robot_index.emplace(kv.first.first, 0);
  for (const auto &kv : lmks) robot_index.emplace(kv.first.first, 0);
  unsigned next = 0;
  for (auto &kv : robot_index) kv.second = next++;
  G.num_robots = next;
  std::map<std::pair<unsigned, unsigned>, unsigned> pose_id, lmk_id;
  for (const auto &kv : poses) {
    pose_id[kv.first] = G.n++;
    G.robot_of_pose.push_back(robot_index[kv.first.first]);
    G.gt_rotation.insert(G.gt_rotation.end(), kv.second.R.begin(), kv.second.R.end());
    G.gt_translation.insert(G.gt_translation.end(), kv.second.t.begin(), kv.second.t.end());
  }
  for (const auto &kv : lmks) {
    lmk_id[kv.first] = G.b++;
    G.robot_of_landmark.push_back(robot_index[kv.first.first]);
    G.gt_landmark.insert(G.gt_landmark.end(), kv.second.begin(), kv.second.end());
  }
  auto pose_of = [&](const detail::PyfgSymbol &s) {
    auto it = pose_id.find({s.robot, s.frame});
    DCORA_B200_CHECK(it != pose_id.end(), "edge refers to an unknown pose");
    return it->second;
  };
  auto lmk_of = [&](const detail::PyfgSymbol &s) {
    auto it = lmk_id.find({s.robot, s.frame});
    DCORA_B200_CHECK(it != lmk_id.end(), "edge refers to an unknown landmark");
    return it->second;
  };
  for (auto &e : pps) {
    e.m.p1 = pose_of(e.a); e.m.p2 = pose_of(e.b);
    e.m.r1 = G.robot_of_pose[e.m.p1]; e.m.r2 = G.robot_of_pose[e.m.p2];
    G.pose_pose.push_back(e.m);
  }
  for (auto &e : pls) {
    e.m.p1 = pose_of(e.a); e.m.p2 = lmk_of(e.b);
    G.pose_landmark.push_back(e.m);
  }
  for (auto &e : rgs) {
    RangeMeasurement m;
    m.landmark1 = e.a.landmark; m.landmark2 = e.b.landmark;
    m.p1 = e.a.landmark ? lmk_of(e.a) : pose_of(e.a);
    m.p2 = e.b.landmark ? lmk_of(e.b) : pose_of(e.b);
    m.range = e.range; m.precision = e.prec;
    G.ranges.push_back(m);
  }
  return G;
}

// ------------------------------------------------------------- PartitionedProblem
// Column bookkeeping + data matrices of a problem whose variables are owned by A agents.
// Agent-major reference ordering: agent a holds columns [coff[a], coff[a+1]) laid out like the
// reference's own agent-local variable -- pose graphs [Y1 p1 .. Yn pn], range-aided graphs
// [Y1..Yn | r1..rl | p1..pn | L1..Lb] (LiftedVariable.h:148-152).  A unit sphere belongs to the
// agent of its range's source endpoint, numbered per agent in measurement order
// (DCORA_utils.cpp:1092-1094).
class PartitionedProblem {
 public:
  explicit PartitionedProblem(const GlobalMeasurements &M) : M_(M), d_(M.d), ra_(M.rangeAided()) {
    A_ = (int)M.num_robots;
    DCORA_B200_CHECK(M.robot_of_pose.size() == M.n && M.robot_of_landmark.size() == M.b,
                     "one owner per variable");
    n_a.assign(A_, 0); l_a.assign(A_, 0); b_a.assign(A_, 0);
    local_pose_.resize(M.n); local_lmk_.resize(M.b);
    for (unsigned i = 0; i < M.n; ++i) local_pose_[i] = n_a[M.robot_of_pose[i]]++;
    for (unsigned j = 0; j < M.b; ++j) local_lmk_[j] = b_a[M.robot_of_landmark[j]]++;
    sph_robot_.resize(M.ranges.size()); local_sph_.resize(M.ranges.size());
    for (size_t e = 0; e < M.ranges.size(); ++e) {
      const RangeMeasurement &rg = M.ranges[e];
      sph_robot_[e] = rg.landmark1 ? M.robot_of_landmark[rg.p1] : M.robot_of_pose[rg.p1];
      local_sph_[e] = l_a[sph_robot_[e]]++;
    }
    coff.assign(A_ + 1, 0);
    for (int a = 0; a < A_; ++a) coff[a + 1] = coff[a] + (int64_t)(d_ + 1) * n_a[a] + l_a[a] + b_a[a];
    DCORA_B200_CHECK(coff[A_] < ((int64_t)1 << 31), "problem dimension exceeds int32");
  }
  unsigned d() const { return d_; }
  int numAgents() const { return A_; }
  bool rangeAided() const { return ra_; }
  int64_t K() const { return coff[A_]; }
  const GlobalMeasurements &measurements() const { return M_; }
  std::vector<unsigned> n_a, l_a, b_a;
  std::vector<int64_t> coff;

  int32_t rotCol(unsigned i, unsigned c) const {
    const unsigned a = M_.robot_of_pose[i], li = local_pose_[i];
    return (int32_t)(coff[a] + (ra_ ? li * d_ + c : li * (d_ + 1) + c));
  }
  int32_t transCol(unsigned i) const {
    const unsigned a = M_.robot_of_pose[i], li = local_pose_[i];
    return (int32_t)(coff[a] + (ra_ ? d_ * n_a[a] + l_a[a] + li : li * (d_ + 1) + d_));
  }
  int32_t sphereCol(size_t e) const {
    const unsigned a = sph_robot_[e];
    return (int32_t)(coff[a] + d_ * n_a[a] + local_sph_[e]);
  }
  int32_t lmkCol(unsigned j) const {
    const unsigned a = M_.robot_of_landmark[j];
    return (int32_t)(coff[a] + d_ * n_a[a] + l_a[a] + n_a[a] + local_lmk_[j]);
  }
  int agentOfCol(int64_t c) const {
    return (int)(std::upper_bound(coff.begin(), coff.end(), c) - coff.begin()) - 1;
  }

  // Global data matrix Q = A^T Omega A (Graph.cpp:579-683 for pose-pose edges; pose-landmark
  // and range terms of Graph.cpp:824-1188: residuals p_j - p_i - Y_i t and p_j - p_i - rho r).
  const SparseMatrix &Q() {
    if (Q_.rows) return Q_;
    const unsigned dh = d_ + 1;
    std::vector<SparseMatrix::Triplet> trip;
    trip.reserve(M_.pose_pose.size() * 4 * dh * dh + M_.ranges.size() * 9);
    std::vector<double> T(dh * dh), TO(dh * dh), om(dh);
    auto col = [&](unsigned pose, unsigned a) { return a < d_ ? rotCol(pose, a) : transCol(pose); };
    for (const auto &m : M_.pose_pose) {
      std::fill(T.begin(), T.end(), 0.0);
      for (unsigned a = 0; a < d_; ++a) {
        for (unsigned c = 0; c < d_; ++c) T[a * dh + c] = m.R[a * d_ + c];
        T[a * dh + d_] = m.t[a];
        om[a] = m.weight * m.kappa;
      }
      T[d_ * dh + d_] = 1.0;
      om[d_] = m.weight * m.tau;
      for (unsigned a = 0; a < dh; ++a)
        for (unsigned c = 0; c < dh; ++c) TO[a * dh + c] = T[a * dh + c] * om[c];
      for (unsigned a = 0; a < dh; ++a)
        for (unsigned c = 0; c < dh; ++c) {
          double s = 0.0;
          for (unsigned e = 0; e < dh; ++e) s += TO[a * dh + e] * T[c * dh + e];
          if (s != 0.0) trip.push_back({col(m.p1, a), col(m.p1, c), s});
          if (TO[a * dh + c] != 0.0) {
            trip.push_back({col(m.p1, a), col(m.p2, c), -TO[a * dh + c]});
            trip.push_back({col(m.p2, c), col(m.p1, a), -TO[a * dh + c]});
          }
        }
      for (unsigned a = 0; a < dh; ++a) trip.push_back({col(m.p2, a), col(m.p2, a), om[a]});
    }
    for (const auto &m : M_.pose_landmark) {
      // z = [Y_i (d columns), p_i, L_j], coefficients c = [-t, -1, +1], weight w tau
      std::vector<int32_t> idx(d_ + 2);
      std::vector<double> c(d_ + 2);
      for (unsigned a = 0; a < d_; ++a) { idx[a] = rotCol(m.p1, a); c[a] = -m.t[a]; }
      idx[d_] = transCol(m.p1); c[d_] = -1.0;
      idx[d_ + 1] = lmkCol(m.p2); c[d_ + 1] = 1.0;
      const double w = m.weight * m.tau;
      for (unsigned a = 0; a < d_ + 2; ++a)
        for (unsigned b2 = 0; b2 < d_ + 2; ++b2) trip.push_back({idx[a], idx[b2], w * c[a] * c[b2]});
    }
    for (size_t e = 0; e < M_.ranges.size(); ++e) {
      const RangeMeasurement &rg = M_.ranges[e];
      const int32_t idx[3] = {rg.landmark1 ? lmkCol(rg.p1) : transCol(rg.p1),
                              rg.landmark2 ? lmkCol(rg.p2) : transCol(rg.p2), sphereCol(e)};
      const double c[3] = {-1.0, 1.0, rg.range};
      const double w = rg.weight * rg.precision;
      for (int a = 0; a < 3; ++a)
        for (int b2 = 0; b2 < 3; ++b2) trip.push_back({idx[a], idx[b2], w * c[a] * c[b2]});
    }
    Q_ = SparseMatrix::fromTriplets((int32_t)K(), (int32_t)K(), std::move(trip));
    return Q_;
  }
  // diagonal blocks of agents [a0, a1) in agent-local columns + the cross-agent remainder
  // (all rows, global columns)
  void split(int a0, int a1, std::vector<SparseMatrix> *diag, SparseMatrix *cross) {
    const SparseMatrix &Qg = Q();
    diag->assign(std::max(0, a1 - a0), SparseMatrix());
    for (int a = a0; a < a1; ++a) {
      SparseMatrix &D = (*diag)[a - a0];
      D.rows = D.cols = (int32_t)(coff[a + 1] - coff[a]);
      D.rowptr.assign(D.rows + 1, 0);
    }
    cross->rows = cross->cols = Qg.rows;
    cross->rowptr.assign(Qg.rows + 1, 0);
    cross->col.clear(); cross->val.clear();
    int a = 0;
    for (int32_t row = 0; row < Qg.rows; ++row) {
      while (row >= coff[a + 1]) ++a;
      const bool mine = a >= a0 && a < a1;
      for (int32_t p = Qg.rowptr[row]; p < Qg.rowptr[row + 1]; ++p) {
        const int32_t c = Qg.col[p];
        if (c >= coff[a] && c < coff[a + 1]) {
          if (mine) {
            (*diag)[a - a0].col.push_back((int32_t)(c - coff[a]));
            (*diag)[a - a0].val.push_back(Qg.val[p]);
          }
        } else {
          cross->col.push_back(c);
          cross->val.push_back(Qg.val[p]);
        }
      }
      if (mine) (*diag)[a - a0].rowptr[row - coff[a] + 1] = (int32_t)(*diag)[a - a0].col.size();
      cross->rowptr[row + 1] = (int32_t)cross->col.size();
    }
  }
  // agents sharing a measurement
  std::vector<std::set<int>> agentGraph(const SparseMatrix &cross) const {
    std::vector<std::set<int>> adj(A_);
    int a = 0;
    for (int32_t row = 0; row < cross.rows; ++row) {
      while (row >= coff[a + 1]) ++a;
      for (int32_t p = cross.rowptr[row]; p < cross.rowptr[row + 1]; ++p)
        adj[a].insert(agentOfCol(cross.col[p]));
    }
    return adj;
  }
  // greedy colouring (largest degree first) followed by the class balancing of
  // dcora_b200/graph.py: agents of one colour share no measurement
  static std::vector<int> coloring(const std::vector<std::set<int>> &adj) {
    const int A = (int)adj.size();
    std::vector<int> order(A), color(A, -1);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int x, int y) { return adj[x].size() > adj[y].size(); });
    for (int a : order) {
      int c = 0;
      for (;; ++c) {
        bool clash = false;
        for (int b : adj[a]) clash = clash || color[b] == c;
        if (!clash) break;
      }
      color[a] = c;
    }
    if (A == 0) return color;
    const int ncol = *std::max_element(color.begin(), color.end()) + 1;
    for (int round = 0; round < 4 * A; ++round) {
      std::vector<int> size(ncol, 0), by_size(ncol);
      for (int c : color) size[c]++;
      const int big = (int)(std::max_element(size.begin(), size.end()) - size.begin());
      std::iota(by_size.begin(), by_size.end(), 0);
      std::stable_sort(by_size.begin(), by_size.end(), [&](int x, int y) { return size[x] < size[y]; });
      bool moved = false;
      for (int tgt : by_size) {
        if (size[tgt] + 1 >= size[big]) break;
        for (int a = 0; a < A && !moved; ++a) {
          if (color[a] != big) continue;
          bool ok = true;
          for (int b : adj[a]) ok = ok && color[b] != tgt;
          if (ok) { color[a] = tgt; moved = true; }
        }
        if (moved) break;
      }
      if (!moved) break;
    }
    return color;
  }

 private:
  friend struct RankPlan;
  GlobalMeasurements M_;
  unsigned d_;
  bool ra_;
  int A_ = 0;
  std::vector<unsigned> local_pose_, local_lmk_, sph_robot_, local_sph_;
  SparseMatrix Q_;
};

// ------------------------------------------------------------------- exchange plan
// Contiguous block of agents per process; every process publishes the states other processes
// need as one fixed-size segment [rotations | spheres | translations | landmarks] (padded to
// the largest process), all-gathered into the halo H = [segment of rank 0 | rank 1 | ...].
struct RankPlan {
  int rank = 0, world = 1, a0 = 0, a1 = 0;
  int64_t c0 = 0, K = 0;                 // first global column / number of local columns
  dcora_halo_segment seg = {0, 0, 0};
  int64_t seg_cols = 0;
  std::vector<int32_t> pack_idx;         // local reference column feeding every segment column
  SparseMatrix coupling;                 // (K + world * seg_cols) x K
  std::vector<int64_t> row_key;          // global column behind every coupling row
  std::vector<int> colors;               // colour of every GLOBAL agent
  int num_colors = 0;
  std::vector<SparseMatrix> diag;        // data matrices of the local agents
};

inline std::pair<int, int> rankAgents(int A, int world, int rank) {
  const int base = A / world, rem = A % world;
  const int a0 = rank * base + std::min(rank, rem);
  return {a0, a0 + base + (rank < rem ? 1 : 0)};
}

inline RankPlan makePlan(PartitionedProblem &P, int world, int rank) {
  const int A = P.numAgents();
  const unsigned d = P.d(), dh = d + 1;
  const GlobalMeasurements &M = P.measurements();
  RankPlan plan;
  plan.rank = rank; plan.world = world;
  std::tie(plan.a0, plan.a1) = rankAgents(A, world, rank);
  plan.c0 = P.coff[plan.a0];
  plan.K = P.coff[plan.a1] - plan.c0;
  SparseMatrix cross;
  P.split(plan.a0, plan.a1, &plan.diag, &cross);
  plan.colors = PartitionedProblem::coloring(P.agentGraph(cross));
  plan.num_colors = A ? *std::max_element(plan.colors.begin(), plan.colors.end()) + 1 : 0;
  const int64_t Kg = P.K();
  std::vector<int> rank_of_agent(A);
  for (int r = 0; r < world; ++r) {
    const auto blk = rankAgents(A, world, r);
    for (int a = blk.first; a < blk.second; ++a) rank_of_agent[a] = r;
  }
  auto rank_of_col = [&](int64_t c) { return rank_of_agent[P.agentOfCol(c)]; };
  if (world == 1) {
    plan.coupling = cross;
    plan.row_key.resize(Kg);
    std::iota(plan.row_key.begin(), plan.row_key.end(), (int64_t)0);
    return plan;
  }
  // kind / variable of every global column
  std::vector<int8_t> kind(Kg, 0);
  std::vector<unsigned> vid(Kg, 0);
  for (unsigned i = 0; i < M.n; ++i) {
    for (unsigned c = 0; c < d; ++c) vid[P.rotCol(i, c)] = i;
    vid[P.transCol(i)] = i;
  }
  for (size_t e = 0; e < M.ranges.size(); ++e) { kind[P.sphereCol(e)] = 1; vid[P.sphereCol(e)] = (unsigned)e; }
  for (unsigned j = 0; j < M.b; ++j) { kind[P.lmkCol(j)] = 2; vid[P.lmkCol(j)] = j; }
  // public variables per rank: rows of `cross` with an entry in another rank's column
  std::vector<std::set<unsigned>> pub_pose(world), pub_sph(world), pub_lmk(world);
  for (int32_t row = 0; row < cross.rows; ++row) {
    const int rr = rank_of_col(row);
    bool is_public = false;
    for (int32_t p = cross.rowptr[row]; p < cross.rowptr[row + 1] && !is_public; ++p)
      is_public = rank_of_col(cross.col[p]) != rr;
    if (!is_public) continue;
    (kind[row] == 0 ? pub_pose : kind[row] == 1 ? pub_sph : pub_lmk)[rr].insert(vid[row]);
  }
  size_t n_s = 0, l_s = 0, b_s = 0;
  for (int r = 0; r < world; ++r) {
    n_s = std::max(n_s, pub_pose[r].size());
    l_s = std::max(l_s, pub_sph[r].size());
    b_s = std::max(b_s, pub_lmk[r].size());
  }
  plan.seg = {(int32_t)n_s, (int32_t)l_s, (int32_t)b_s};
  plan.seg_cols = (int64_t)dh * n_s + l_s + b_s;
  const bool ra = P.rangeAided();
  std::vector<int64_t> halo_of(Kg, -1), halo_key((size_t)world * plan.seg_cols, -1);
  for (int r = 0; r < world; ++r) {
    const int64_t rc0 = P.coff[rankAgents(A, world, r).first];
    std::vector<int64_t> seg_ref(plan.seg_cols, -1);
    size_t j = 0;
    for (unsigned i : pub_pose[r]) {
      for (unsigned c = 0; c < d; ++c)
        seg_ref[ra ? j * d + c : j * dh + c] = P.rotCol(i, c);
      seg_ref[ra ? d * n_s + l_s + j : j * dh + d] = P.transCol(i);
      ++j;
    }
    j = 0;
    for (unsigned e : pub_sph[r]) seg_ref[d * n_s + j++] = P.sphereCol(e);
    j = 0;
    for (unsigned q : pub_lmk[r]) seg_ref[d * n_s + l_s + n_s + j++] = P.lmkCol(q);
    for (int64_t s = 0; s < plan.seg_cols; ++s) {
      if (seg_ref[s] < 0) continue;
      halo_of[seg_ref[s]] = r * plan.seg_cols + s;
      halo_key[r * plan.seg_cols + s] = seg_ref[s];
    }
    if (r == rank) {
      plan.pack_idx.assign(plan.seg_cols, 0);
      for (int64_t s = 0; s < plan.seg_cols; ++s)
        if (seg_ref[s] >= 0) plan.pack_idx[s] = (int32_t)(seg_ref[s] - rc0);
    }
  }
  // coupling rows [local columns | halo], columns = local columns: the transpose view of the
  // cross entries whose column is local (cross is symmetric: entry (row, col) <-> (col, row))
  std::vector<SparseMatrix::Triplet> trip;
  const int64_t c1 = plan.c0 + plan.K;
  for (int32_t row = 0; row < cross.rows; ++row)
    for (int32_t p = cross.rowptr[row]; p < cross.rowptr[row + 1]; ++p) {
      const int32_t c = cross.col[p];
      if (c < plan.c0 || c >= c1) continue;
      int64_t src;
      if (row >= plan.c0 && row < c1) src = row - plan.c0;
      else {
        DCORA_B200_CHECK(halo_of[row] >= 0, "coupling source is not public");
        src = plan.K + halo_of[row];
      }
      trip.push_back({(int32_t)src, (int32_t)(c - plan.c0), cross.val[p]});
    }
  plan.coupling = SparseMatrix::fromTriplets((int32_t)(plan.K + world * plan.seg_cols),
                                             (int32_t)plan.K, std::move(trip));
  plan.row_key.resize(plan.K + world * plan.seg_cols);
  for (int64_t c = 0; c < plan.K; ++c) plan.row_key[c] = plan.c0 + c;
  std::copy(halo_key.begin(), halo_key.end(), plan.row_key.begin() + plan.K);
  return plan;
}

// ------------------------------------------------------------ DistributedAgentBatch
// The agents of ONE process on one GPU + the exchange with the other processes.
// X, V, Y live in device memory in the agent-major reference ordering of the local columns
// (r x K_local, column-major).
class DistributedAgentBatch {
 public:
  // comm may be nullptr when world == 1
  DistributedAgentBatch(PartitionedProblem &P, unsigned r, int rank, int world, int device,
                        ncclComm_t comm, bool acceleration = false, unsigned restart_interval = 30,
                        int nesterov_blocks = 0)
      : r_(r), d_(P.d()), A_global_(P.numAgents()), comm_(comm), acc_(acceleration),
        restart_interval_(restart_interval), nesterov_blocks_(nesterov_blocks) {
    DCORA_B200_CUDA(cudaSetDevice(device));
    plan_ = makePlan(P, world, rank);
    const int A = plan_.a1 - plan_.a0;
    std::vector<dcora_agent_desc> desc(A);
    for (int a = 0; a < A; ++a) {
      const int g = plan_.a0 + a;
      desc[a].n = (int32_t)P.n_a[g]; desc[a].l = (int32_t)P.l_a[g]; desc[a].b = (int32_t)P.b_a[g];
      desc[a].rowptr = plan_.diag[a].rowptr.data();
      desc[a].col = plan_.diag[a].col.data();
      desc[a].val = plan_.diag[a].val.data();
      desc[a].G = nullptr;
      desc[a].precond_reg = P.rangeAided() ? DCORA_REG_AUTO_RA : 0.1;
    }
    // CTAs per agent of the preconditioner solve: by the largest colour class on any process
    // (see dcora_set_trsv_max_warps and dcora_b200/rbcd.py)
    int biggest = 0;
    for (int rk = 0; rk < world; ++rk) {
      const auto blk = rankAgents(A_global_, world, rk);
      std::vector<int> cnt(plan_.num_colors + 1, 0);
      for (int a = blk.first; a < blk.second; ++a) biggest = std::max(biggest, ++cnt[plan_.colors[a]]);
    }
    const int warps = (r > 8 || biggest == 0) ? 64 : biggest <= 8 ? 256 : biggest <= 16 ? 128
                      : biggest <= 24 ? 96 : 64;
    dcora_set_trsv_max_warps(warps);
    h_ = std::make_shared<detail::Handle>();
    DCORA_B200_CALL(dcora_problem_create(desc.data(), A, (int32_t)r, (int32_t)d_,
                                         P.rangeAided() ? DCORA_ORDER_RA : DCORA_ORDER_SE, device,
                                         &h_->p));
    dcora_set_trsv_max_warps(64);
    std::vector<dcora_halo_segment> segs(world > 1 ? world : 0, plan_.seg);
    DCORA_B200_CALL(dcora_problem_set_coupling_keyed(
        h_->p, segs.data(), (int32_t)segs.size(), plan_.coupling.rowptr.data(),
        plan_.coupling.col.data(), plan_.coupling.val.data(), plan_.row_key.data()));
    DCORA_B200_CALL(dcora_problem_set_iterate_tracking(h_->p, acc_ ? 0 : 1));
    const size_t bytes = sizeof(double) * (size_t)plan_.K * r_;
    for (double **buf : {&X_, &V_, &Y_, &T_, &W_}) DCORA_B200_CUDA(cudaMalloc(buf, std::max<size_t>(bytes, 8)));
    if (world > 1) {
      DCORA_B200_CUDA(cudaMalloc(&send_, sizeof(double) * (size_t)plan_.seg_cols * r_ + 8));
      DCORA_B200_CUDA(cudaMalloc(&H_, sizeof(double) * (size_t)world * plan_.seg_cols * r_ + 8));
      DCORA_B200_CUDA(cudaMemset(H_, 0, sizeof(double) * (size_t)world * plan_.seg_cols * r_ + 8));
    }
    dcora_ropt_params_default(&params_);
    params_.time_bound_s = 0.0;
  }
  ~DistributedAgentBatch() {
    for (double *buf : {X_, V_, Y_, T_, W_, send_, H_}) cudaFree(buf);
  }
  DistributedAgentBatch(const DistributedAgentBatch &) = delete;
  DistributedAgentBatch &operator=(const DistributedAgentBatch &) = delete;

  const RankPlan &plan() const { return plan_; }
  dcora_ropt_params &params() { return params_; }
  int64_t agentUpdates() const { return updates_; }

  // Xglobal: r x K_global (host, agent-major reference ordering)
  void setX(const Matrix &Xglobal) {
    DCORA_B200_CHECK(Xglobal.rows() == r_, "X must have r rows");
    DCORA_B200_CUDA(cudaMemcpy(X_, Xglobal.data() + (size_t)plan_.c0 * r_,
                               sizeof(double) * (size_t)plan_.K * r_, cudaMemcpyHostToDevice));
    if (acc_) {  // initializeAcceleration (Agent.cpp:1178-1187)
      copy(V_, X_); copy(Y_, X_);
      gamma_ = alpha_ = 0.0;
    } else {
      DCORA_B200_CALL(dcora_problem_set_iterate_tracking(h_->p, 1));
    }
  }
  Matrix getX() const {  // the LOCAL columns
    Matrix X(r_, (size_t)plan_.K);
    DCORA_B200_CUDA(cudaMemcpy(X.data(), X_, sizeof(double) * (size_t)plan_.K * r_, cudaMemcpyDeviceToHost));
    return X;
  }
  // Agent::getSharedStateDicts -> updateNeighborStates for every process at once
  void exchange(const double *src) {
    if (plan_.world == 1) return;
    DCORA_B200_CALL(dcora_problem_pack_columns(h_->p, src, plan_.pack_idx.data(),
                                               (int64_t)plan_.pack_idx.size(), send_));
    cudaStream_t st = (cudaStream_t)dcora_problem_stream(h_->p);
    DCORA_B200_NCCL(ncclAllGather(send_, H_, (size_t)plan_.seg_cols * r_, ncclDouble, comm_, st));
    DCORA_B200_CUDA(cudaStreamSynchronize(st));
  }
  // one RBCD iteration: every colour of the agent graph once
  void iterate() {
    for (int c = 0; c < plan_.num_colors; ++c) {
      std::vector<uint8_t> active(plan_.a1 - plan_.a0, 0);
      for (int a = plan_.a0; a < plan_.a1; ++a) active[a - plan_.a0] = plan_.colors[a] == c;
      step(active);
    }
  }
  // one driver iteration (MultiRobotExample.cpp:223-258): the active local agents take
  // Agent::iterate(true), all others iterate(false); collective over the processes
  void step(const std::vector<uint8_t> &active) {
    const bool any = std::find(active.begin(), active.end(), (uint8_t)1) != active.end();
    if (!acc_) {
      exchange(X_);
      if (any) {
        DCORA_B200_CALL(dcora_problem_update_G(h_->p, X_, H_, active.data()));
        solve(X_, X_, active);
      }
      return;
    }
    // RBCD++ (Agent.cpp:535-596,1158-1214); see dcora_b200/rbcd.py:_accelerated_step
    // N of updateGamma / updateAlpha: the number of agents (reference, nesterov_blocks = 0), a
    // given block count, or the number of colours (-1: a whole colour class moves per iteration)
    const double N = nesterov_blocks_ > 0 ? (double)nesterov_blocks_
                     : nesterov_blocks_ < 0 ? (double)std::max(1, plan_.num_colors)
                                            : (double)A_global_;
    ++it_;
    gamma_ = (1 + std::sqrt(1 + 4 * N * N * gamma_ * gamma_)) / (2 * N);
    alpha_ = 1 / (gamma_ * N);
    const bool restart = (it_ + 1) % restart_interval_ == 0;
    if (restart) {   // X <- XPrev, solved from XPrev; V = Y = X
      exchange(X_);
      if (any) {
        DCORA_B200_CALL(dcora_problem_update_G(h_->p, X_, H_, active.data()));
        solve(X_, X_, active);
      }
      copy(V_, X_); copy(Y_, X_);
      gamma_ = alpha_ = 0.0;
      return;
    }
    // Y = proj((1 - alpha) X + alpha V), every agent; iterate(false): X = Y
    DCORA_B200_CALL(dcora_problem_combine_project(h_->p, 1 - alpha_, X_, alpha_, V_, nullptr, Y_));
    copy(T_, Y_);
    exchange(Y_);
    if (any) {
      DCORA_B200_CALL(dcora_problem_update_G(h_->p, Y_, H_, active.data()));
      solve(Y_, T_, active);
    }
    // V = proj(V + gamma (Xn - Y))
    DCORA_B200_CALL(dcora_problem_combine(h_->p, 1.0, T_, -1.0, Y_, W_));
    DCORA_B200_CALL(dcora_problem_combine_project(h_->p, 1.0, V_, gamma_, W_, nullptr, V_));
    copy(X_, T_);
  }
  // global cost and gradient norm (MultiRobotExample.cpp:264-281); collective
  void evaluate(double *cost, double *gradnorm) {
    exchange(X_);
    DCORA_B200_CALL(dcora_problem_update_G(h_->p, X_, H_, nullptr));
    const int A = plan_.a1 - plan_.a0;
    std::vector<double> quad(A), lin(A), gn(A);
    DCORA_B200_CALL(dcora_problem_cost_terms(h_->p, X_, quad.data(), lin.data()));
    DCORA_B200_CALL(dcora_problem_rgradnorm(h_->p, X_, gn.data()));
    double v[2] = {0.0, 0.0};
    for (int a = 0; a < A; ++a) { v[0] += quad[a] + 0.5 * lin[a]; v[1] += gn[a] * gn[a]; }
    if (plan_.world > 1) {
      cudaStream_t st = (cudaStream_t)dcora_problem_stream(h_->p);
      DCORA_B200_CUDA(cudaMemcpyAsync(W_, v, sizeof(v), cudaMemcpyHostToDevice, st));
      DCORA_B200_NCCL(ncclAllReduce(W_, W_, 2, ncclDouble, ncclSum, comm_, st));
      DCORA_B200_CUDA(cudaMemcpyAsync(v, W_, sizeof(v), cudaMemcpyDeviceToHost, st));
      DCORA_B200_CUDA(cudaStreamSynchronize(st));
    }
    if (cost) *cost = v[0];
    if (gradnorm) *gradnorm = std::sqrt(v[1]);
  }
  dcora_problem *handle() const { return h_->p; }

 private:
  void copy(double *dst, const double *src) {
    DCORA_B200_CUDA(cudaMemcpy(dst, src, sizeof(double) * (size_t)plan_.K * r_, cudaMemcpyDeviceToDevice));
  }
  void solve(const double *X0, double *Xout, const std::vector<uint8_t> &active) {
    std::vector<dcora_ropt_result> res(plan_.a1 - plan_.a0);
    DCORA_B200_CALL(dcora_rtr_solve(h_->p, X0, Xout, &params_, active.data(), res.data()));
    for (uint8_t f : active) updates_ += f ? 1 : 0;
  }
  unsigned r_, d_;
  int A_global_;
  ncclComm_t comm_;
  bool acc_;
  unsigned restart_interval_;
  int nesterov_blocks_;
  RankPlan plan_;
  std::shared_ptr<detail::Handle> h_;
  dcora_ropt_params params_;
  double *X_ = nullptr, *V_ = nullptr, *Y_ = nullptr, *T_ = nullptr, *W_ = nullptr;
  double *send_ = nullptr, *H_ = nullptr;
  double gamma_ = 0.0, alpha_ = 0.0;
  unsigned it_ = 0;
  int64_t updates_ = 0;
};

}  // namespace DCORA
#endif  // DCORA_B200_DCORA_DIST_HPP_
