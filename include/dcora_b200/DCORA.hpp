// DCORA.hpp -- C++ host mirror of the reference's solver-path classes over the
// dcora_b200 C-ABI (include/dcora_b200.h).
//
// Same names, argument meaning and error behaviour as adthoms/dcora
// (citations are path:line in the reference checkout):
//   DCORA::Graph              include/DCORA/Graph.h:57-251 (data-matrix subset)
//   DCORA::QuadraticProblem   include/DCORA/QuadraticProblem.h:40-185
//   DCORA::QuadraticOptimizer include/DCORA/QuadraticOptimizer.h:25-147
//   DCORA::LiftedSEManifold / LiftedRAManifold
//                             include/DCORA/manifold/LiftedManifold.h:27-119
//   DCORA::ROptParameters / ROPTResult   include/DCORA/DCORA_types.h:152-233
//   DCORA::fastVerification / read_g2o_file   include/DCORA/DCORA_utils.h
// The bodies contain no arithmetic of the hot path: they marshal into the
// C-ABI, whose kernels run on the GPU.  There is no CPU fallback -- without a
// CUDA device every compute call aborts like the reference's glog CHECKs do.
//
// Dense type: a maintainer of the reference compiles this header with
// -DDCORA_B200_USE_EIGEN, which makes DCORA::Matrix = Eigen::MatrixXd (column
// major, include/DCORA/DCORA_types.h:34) exactly as in the reference.  Eigen is
// not part of this repository's image, so by default a minimal column-major
// stand-in with the same accessor subset (rows/cols/data/operator()) is used.
#ifndef DCORA_B200_DCORA_HPP_
#define DCORA_B200_DCORA_HPP_

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../dcora_b200.h"

#ifdef DCORA_B200_USE_EIGEN
#include <Eigen/Dense>
#endif

// glog-style precondition check: the reference aborts the process
// (e.g. src/QuadraticProblem.cpp:39-40, src/Graph.cpp:31).
#define DCORA_B200_CHECK(cond, msg)                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "Check failed: %s (%s) at %s:%d\n", #cond,            \
                   std::string(msg).c_str(), __FILE__, __LINE__);                \
      std::abort();                                                              \
    }                                                                            \
  } while (0)
// hard C-ABI errors (<0) abort with the library's message; soft failures (>0)
// are returned to the caller as the reference's bools
#define DCORA_B200_CALL(expr)                                                    \
  ([&]() {                                                                       \
    const int rc__ = (expr);                                                     \
    if (rc__ < 0) {                                                              \
      std::fprintf(stderr, "dcora_b200: %s failed (%d): %s\n", #expr, rc__,      \
                   dcora_last_error());                                          \
      std::abort();                                                              \
    }                                                                            \
    return rc__;                                                                 \
  }())

namespace ROPTLIB {
// ROPTLIB::tCGstatusSet as stored in ROPTResult (DCORA_types.h:226)
enum tCGstatusSet { TR_NEGCURVTURE = 0, TR_EXCREGION, TR_LCON, TR_SCON, TR_MAXITER };
}  // namespace ROPTLIB

namespace DCORA {

#ifdef DCORA_B200_USE_EIGEN
using Matrix = Eigen::MatrixXd;
using Vector = Eigen::VectorXd;
#else
// Column-major dense matrix (stand-in for Eigen::MatrixXd).
class Matrix {
 public:
  Matrix() = default;
  Matrix(size_t rows, size_t cols) : rows_(rows), cols_(cols), v_(rows * cols, 0.0) {}
  static Matrix Zero(size_t rows, size_t cols) { return Matrix(rows, cols); }
  static Matrix Identity(size_t rows, size_t cols) {
    Matrix M(rows, cols);
    for (size_t i = 0; i < std::min(rows, cols); ++i) M(i, i) = 1.0;
    return M;
  }
  size_t rows() const { return rows_; }
  size_t cols() const { return cols_; }
  size_t size() const { return v_.size(); }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(size_t i, size_t j) { return v_[j * rows_ + i]; }
  double operator()(size_t i, size_t j) const { return v_[j * rows_ + i]; }
  double norm() const {
    double s = 0.0;
    for (double x : v_) s += x * x;
    return std::sqrt(s);
  }
  // rows [0, r) as a new matrix / overwrite rows [0, M.rows())
  Matrix topRows(size_t r) const {
    Matrix M(r, cols_);
    for (size_t j = 0; j < cols_; ++j)
      for (size_t i = 0; i < r; ++i) M(i, j) = (*this)(i, j);
    return M;
  }
  void setTopRows(const Matrix &M) {
    for (size_t j = 0; j < cols_; ++j)
      for (size_t i = 0; i < M.rows(); ++i) (*this)(i, j) = M(i, j);
  }

 private:
  size_t rows_ = 0, cols_ = 0;
  std::vector<double> v_;
};

class Vector {
 public:
  Vector() = default;
  explicit Vector(size_t n) : v_(n, 0.0) {}
  size_t size() const { return v_.size(); }
  size_t rows() const { return v_.size(); }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  double &operator()(size_t i) { return v_[i]; }
  double operator()(size_t i) const { return v_[i]; }
  double norm() const {
    double s = 0.0;
    for (double x : v_) s += x * x;
    return std::sqrt(s);
  }

 private:
  std::vector<double> v_;
};
#endif

// Row-major sparse matrix with int32 indices: the storage of
// Eigen::SparseMatrix<double, Eigen::RowMajor> (DCORA_types.h:36) after
// makeCompressed(), which is what crosses the C-ABI.
struct SparseMatrix {
  int32_t rows = 0, cols = 0;
  std::vector<int32_t> rowptr, col;
  std::vector<double> val;
  int64_t nonZeros() const { return (int64_t)val.size(); }

  struct Triplet {
    int32_t r, c;
    double v;
  };
  // Eigen's setFromTriplets: duplicates are summed.
  static SparseMatrix fromTriplets(int32_t rows, int32_t cols, std::vector<Triplet> t) {
    std::sort(t.begin(), t.end(), [](const Triplet &a, const Triplet &b) {
      return a.r != b.r ? a.r < b.r : a.c < b.c;
    });
    SparseMatrix S;
    S.rows = rows;
    S.cols = cols;
    S.rowptr.assign(rows + 1, 0);
    for (size_t i = 0; i < t.size();) {
      size_t j = i;
      double s = 0.0;
      while (j < t.size() && t[j].r == t[i].r && t[j].c == t[i].c) s += t[j++].v;
      S.col.push_back(t[i].c);
      S.val.push_back(s);
      S.rowptr[t[i].r + 1]++;
      i = j;
    }
    for (int32_t r = 0; r < rows; ++r) S.rowptr[r + 1] += S.rowptr[r];
    return S;
  }
};

enum class GraphType { PoseGraph, RangeAidedSLAMGraph };  // DCORA_types.h:59

// ---------------------------------------------------------------- parameters
class ROptParameters {  // DCORA_types.h:152-200
 public:
  enum class ROptMethod { RTR, RGD };
  ROptParameters()
      : method(ROptMethod::RTR), verbose(false), gradnorm_tol(1e-2), RGD_stepsize(1e-3),
        RGD_use_preconditioner(true), RTR_iterations(3), RTR_tCG_iterations(50),
        RTR_initial_radius(100) {}
  ROptMethod method;
  bool verbose;
  double gradnorm_tol;
  double RGD_stepsize;
  bool RGD_use_preconditioner;
  int RTR_iterations;
  int RTR_tCG_iterations;
  double RTR_initial_radius;
  // The reference hard-codes Solver->TimeBound = 5.0 (QuadraticOptimizer.cpp:252);
  // exposed here with the same default, <= 0 disables it.
  double time_bound_s = 5.0;
  // Stiefel retraction of the RTR step: qf (ROPTLIB) or polar.
  int retraction = DCORA_RETRACT_QF;
};

struct ROPTResult {  // DCORA_types.h:203-233
  ROPTResult(bool suc = false, double f0 = 0, double gn0 = 0, double fStar = 0,
             double gnStar = 0, double ms = 0)
      : success(suc), fInit(f0), gradNormInit(gn0), fOpt(fStar), gradNormOpt(gnStar),
        elapsedMs(ms), tCGStatus(ROPTLIB::TR_MAXITER) {}
  bool success;
  double fInit, gradNormInit, fOpt, gradNormOpt, elapsedMs;
  ROPTLIB::tCGstatusSet tCGStatus;
  int outerIterations = 0, innerIterations = 0;  // extra: iteration counts
};

// -------------------------------------------------------------- measurements
struct RelativePosePoseMeasurement {  // include/DCORA/Measurements.h (pose-pose edge)
  unsigned r1 = 0, p1 = 0, r2 = 0, p2 = 0;
  unsigned d = 3;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major d x d
  double t[3] = {0, 0, 0};
  double kappa = 1, tau = 1, weight = 1;
};

// Range measurement between two translation variables (pose or landmark), as the
// reference's RangeMeasurement (include/DCORA/Measurements.h): unit sphere k of
// the lifted problem is r_k = (t_src - t_dst) / ||.|| (src/DCORA_utils.cpp:1136-1137).
struct RangeMeasurement {
  unsigned p1 = 0, p2 = 0;            // source / destination index
  bool landmark1 = false, landmark2 = false;  // endpoint is a landmark (else a pose)
  double range = 1, precision = 1, weight = 1;
};

namespace detail {
// Quaternion -> rotation matrix WITHOUT normalising the quaternion: the
// reference calls Eigen::Quaterniond(w,x,y,z).toRotationMatrix()
// (DCORA_utils.cpp:331), which assumes a unit quaternion, so file quaternions
// with ~8 digits give slightly non-orthogonal R -- reproduced here on purpose.
inline void quat_to_rot(double qx, double qy, double qz, double qw, double *R) {
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
inline double trace_inv3(const double *A) {  // trace(inv(A)), A symmetric 3x3 row-major
  const double c00 = A[4] * A[8] - A[5] * A[7], c11 = A[0] * A[8] - A[2] * A[6],
               c22 = A[0] * A[4] - A[1] * A[3];
  const double det = A[0] * c00 - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                     A[2] * (A[3] * A[7] - A[4] * A[6]);
  return (c00 + c11 + c22) / det;
}
}  // namespace detail

// read_g2o_file (src/DCORA_utils.cpp:179-375): translation precision
// tau = d / tr(Itt^-1), rotation precision kappa = I33 (2-D) or
// 3 / (2 tr(Irr^-1)) (3-D).  The dimension is also inferred from EDGE lines, so
// files without VERTEX lines (data/kitti_00.g2o) load.
inline std::vector<RelativePosePoseMeasurement> read_g2o_file(const std::string &filename,
                                                              size_t *num_poses) {
  std::vector<RelativePosePoseMeasurement> out;
  std::ifstream in(filename);
  DCORA_B200_CHECK(in.good(), "cannot open " + filename);
  std::string line, tok;
  size_t maxid = 0;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    if (!(ss >> tok)) continue;
    if (tok == "VERTEX_SE2" || tok == "VERTEX_SE3:QUAT") continue;
    RelativePosePoseMeasurement m;
    if (tok == "EDGE_SE2") {
      double dx, dy, dth, I11, I12, I13, I22, I23, I33;
      ss >> m.p1 >> m.p2 >> dx >> dy >> dth >> I11 >> I12 >> I13 >> I22 >> I23 >> I33;
      m.d = 2;
      m.R[0] = std::cos(dth); m.R[1] = -std::sin(dth); m.R[2] = std::sin(dth); m.R[3] = std::cos(dth);
      m.t[0] = dx; m.t[1] = dy;
      const double det = I11 * I22 - I12 * I12;
      m.tau = 2.0 / ((I11 + I22) / det);
      m.kappa = I33;
    } else if (tok == "EDGE_SE3:QUAT") {
      double v[28];
      ss >> m.p1 >> m.p2;
      for (double &x : v) ss >> x;
      m.d = 3;
      m.t[0] = v[0]; m.t[1] = v[1]; m.t[2] = v[2];
      detail::quat_to_rot(v[3], v[4], v[5], v[6], m.R);
      const double *I = v + 7;  // upper triangle of the 6x6 information matrix, row by row
      const double tran[9] = {I[0], I[1], I[2], I[1], I[6], I[7], I[2], I[7], I[11]};
      const double rot[9] = {I[15], I[16], I[17], I[16], I[18], I[19], I[17], I[19], I[20]};
      m.tau = 3.0 / detail::trace_inv3(tran);
      m.kappa = 3.0 / (2.0 * detail::trace_inv3(rot));
    } else {
      DCORA_B200_CHECK(false, "unrecognized g2o record " + tok);
    }
    DCORA_B200_CHECK(!ss.fail(), "malformed g2o line: " + line);
    maxid = std::max<size_t>(maxid, std::max(m.p1, m.p2));
    out.push_back(m);
  }
  if (num_poses) *num_poses = out.empty() ? 0 : maxid + 1;
  return out;
}

// ------------------------------------------------------------------- Graph
// The part of DCORA::Graph the solver path consumes (include/DCORA/Graph.h):
// dimensions, the quadratic matrix Q, the linear matrix G and the
// preconditioner regularisation.  Q is either assembled from pose-pose
// measurements whose endpoints are all local (constructQuadraticCostTermPGO,
// src/Graph.cpp:579-683) or injected (setQuadraticMatrix) by a caller that
// already owns the reference's Graph.
class Graph {
 public:
  Graph(unsigned id, unsigned r, unsigned d, GraphType graphType = GraphType::PoseGraph)
      : id_(id), r_(r), d_(d), type_(graphType) {
    DCORA_B200_CHECK(r >= d, "relaxation rank must be >= dimension");  // Graph.cpp:31
  }
  unsigned d() const { return d_; }
  unsigned r() const { return r_; }
  unsigned n() const { return n_; }
  unsigned l() const { return l_; }
  unsigned b() const { return b_; }
  unsigned k() const { return (d_ + 1) * n_ + l_ + b_; }
  GraphType graphType() const { return type_; }

  void setMeasurements(const std::vector<RelativePosePoseMeasurement> &measurements) {
    DCORA_B200_CHECK(type_ == GraphType::PoseGraph, "pose-pose measurements only");
    meas_ = measurements;
    ranges_.clear();
    n_ = 0; l_ = 0; b_ = 0;
    for (const auto &m : meas_) {
      DCORA_B200_CHECK(m.d == d_, "measurement dimension mismatch");
      n_ = std::max(n_, std::max(m.p1, m.p2) + 1);
    }
    clearDataMatrices();
  }
  // Range-aided graph: pose-pose edges + range edges (one unit sphere per range,
  // in the order given); `num_landmarks` landmark positions.  Column ordering
  // [Y1..Yn | r1..rl | p1..pn | L1..Lb] (LiftedVariable.h:148-152).
  void setMeasurements(const std::vector<RelativePosePoseMeasurement> &measurements,
                       const std::vector<RangeMeasurement> &ranges, unsigned num_landmarks = 0) {
    DCORA_B200_CHECK(type_ == GraphType::RangeAidedSLAMGraph, "range-aided graph expected");
    meas_ = measurements;
    ranges_ = ranges;
    n_ = 0;
    for (const auto &m : meas_) {
      DCORA_B200_CHECK(m.d == d_, "measurement dimension mismatch");
      n_ = std::max(n_, std::max(m.p1, m.p2) + 1);
    }
    for (const auto &rg : ranges_) {
      if (!rg.landmark1) n_ = std::max(n_, rg.p1 + 1); else num_landmarks = std::max(num_landmarks, rg.p1 + 1);
      if (!rg.landmark2) n_ = std::max(n_, rg.p2 + 1); else num_landmarks = std::max(num_landmarks, rg.p2 + 1);
    }
    l_ = (unsigned)ranges_.size();
    b_ = num_landmarks;
    clearDataMatrices();
  }
  // column of the lifted variable holding rotation column c / the translation of pose i,
  // unit sphere e, landmark j (SE: interleaved; RA: DCORA_utils.cpp:2156-2165)
  unsigned rotationColumn(unsigned i, unsigned c) const {
    return type_ == GraphType::PoseGraph ? i * (d_ + 1) + c : i * d_ + c;
  }
  unsigned translationColumn(unsigned i) const {
    return type_ == GraphType::PoseGraph ? i * (d_ + 1) + d_ : d_ * n_ + l_ + i;
  }
  unsigned unitSphereColumn(unsigned e) const { return d_ * n_ + e; }
  unsigned landmarkColumn(unsigned j) const { return d_ * n_ + l_ + n_ + j; }
  // Inject data matrices built elsewhere (e.g. by the reference's own Graph).
  void setQuadraticMatrix(unsigned n, unsigned l, unsigned b, SparseMatrix Q) {
    n_ = n; l_ = l; b_ = b;
    DCORA_B200_CHECK((unsigned)Q.rows == k() && (unsigned)Q.cols == k(), "Q must be k x k");
    Q_ = std::make_shared<SparseMatrix>(std::move(Q));
    version_++;
  }
  void setLinearMatrix(const Matrix &G) {
    DCORA_B200_CHECK(G.rows() == r_ && G.cols() == k(), "G must be r x k");
    G_ = std::make_shared<Matrix>(G);
  }
  void clearDataMatrices() {  // Graph.cpp:552-555
    Q_.reset();
    G_.reset();
    version_++;
  }
  bool constructDataMatrices() {  // Graph.cpp:544-550
    if (!Q_) constructQ();
    return true;
  }
  const SparseMatrix &quadraticMatrix() {  // Graph.cpp:523-528
    if (!Q_) constructQ();
    return *Q_;
  }
  bool hasLinearMatrix() const { return (bool)G_; }
  const Matrix &linearMatrix() {  // Graph.cpp:535-540
    if (!G_) G_ = std::make_shared<Matrix>(Matrix::Zero(r_, k()));
    return *G_;
  }
  // Graph::constructPreconditioner's regularisation (Graph.cpp:1901-1960):
  // 0.1 for pose graphs, lambda_max(Q)/(1e6-1) for range-aided graphs.
  double preconditionerRegularization() const {
    return type_ == GraphType::PoseGraph ? 0.1 : DCORA_REG_AUTO_RA;
  }
  unsigned long version() const { return version_; }

 private:
  // Q = A^T Omega A with, per edge i -> j, T = [R t; 0 1] and
  // Omega = diag(w kappa, .., w tau):  Q_ii += T Omega T^T, Q_ij -= T Omega,
  // Q_ji -= Omega T^T, Q_jj += Omega   (Graph.cpp:579-683).  Range edge k between
  // translation variables u (source) and v (destination) with range rho and
  // weight w = weight * precision adds w c c^T on [u, v, r_k] with
  // c = [-1, +1, rho]  (Q22, Q23, Q33 of Graph.cpp:1137-1141).
  void constructQ() {
    const unsigned dh = d_ + 1;
    std::vector<SparseMatrix::Triplet> trip;
    trip.reserve(meas_.size() * 4 * dh * dh + ranges_.size() * 9);
    std::vector<double> T(dh * dh), TO(dh * dh), om(dh);
    auto col = [&](unsigned pose, unsigned a) {
      return (int32_t)(a < d_ ? rotationColumn(pose, a) : translationColumn(pose));
    };
    for (const auto &m : meas_) {
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
    for (size_t e = 0; e < ranges_.size(); ++e) {
      const RangeMeasurement &rg = ranges_[e];
      const int32_t idx[3] = {
          (int32_t)(rg.landmark1 ? landmarkColumn(rg.p1) : translationColumn(rg.p1)),
          (int32_t)(rg.landmark2 ? landmarkColumn(rg.p2) : translationColumn(rg.p2)),
          (int32_t)unitSphereColumn((unsigned)e)};
      const double c[3] = {-1.0, 1.0, rg.range};
      const double w = rg.weight * rg.precision;
      for (int a = 0; a < 3; ++a)
        for (int b2 = 0; b2 < 3; ++b2) trip.push_back({idx[a], idx[b2], w * c[a] * c[b2]});
    }
    Q_ = std::make_shared<SparseMatrix>(SparseMatrix::fromTriplets(k(), k(), std::move(trip)));
  }

  unsigned id_, r_, d_;
  unsigned n_ = 0, l_ = 0, b_ = 0;
  GraphType type_;
  std::vector<RelativePosePoseMeasurement> meas_;
  std::vector<RangeMeasurement> ranges_;
  std::shared_ptr<SparseMatrix> Q_;
  std::shared_ptr<Matrix> G_;
  unsigned long version_ = 0;
};

namespace detail {
// RAII owner of a dcora_problem handle.
struct Handle {
  dcora_problem *p = nullptr;
  Handle() = default;
  Handle(const Handle &) = delete;
  Handle &operator=(const Handle &) = delete;
  ~Handle() {
    if (p) dcora_problem_destroy(p);
  }
};

// Device handle of one agent's problem (num_agents = 1) from CSR Q (+ G).
inline std::shared_ptr<Handle> make_handle(unsigned r, unsigned d, unsigned n, unsigned l,
                                           unsigned b, GraphType type, const SparseMatrix &Q,
                                           const Matrix *G, double reg, int device) {
  dcora_agent_desc a;
  a.n = (int32_t)n; a.l = (int32_t)l; a.b = (int32_t)b;
  a.rowptr = Q.rowptr.data(); a.col = Q.col.data(); a.val = Q.val.data();
  a.G = G ? G->data() : nullptr;
  a.precond_reg = reg;
  auto h = std::make_shared<Handle>();
  DCORA_B200_CALL(dcora_problem_create(
      &a, 1, (int32_t)r, (int32_t)d,
      type == GraphType::PoseGraph ? DCORA_ORDER_SE : DCORA_ORDER_RA, device, &h->p));
  return h;
}

inline SparseMatrix empty_matrix(unsigned k) {
  SparseMatrix Q;
  Q.rows = Q.cols = (int32_t)k;
  Q.rowptr.assign(k + 1, 0);
  return Q;
}
}  // namespace detail

// ---------------------------------------------------------------- manifolds
// LiftedSEManifold / LiftedRAManifold (LiftedManifold.h:27-119): the product
// manifold (Stiefel(r,d) x R^r)^n, resp. Stiefel^n x Oblique(r,l) x R^{r x n} x
// R^{r x b}.  ROPTLIB objects are gone; the two operations the reference uses
// are served by the K2 kernels.
class LiftedManifoldBase {
 public:
  // projectToSEMatrix / projectToRAMatrix (DCORA_utils.cpp:2201-2220)
  Matrix project(const Matrix &M) const {
    check(M);
    Matrix out(M.rows(), M.cols());
    DCORA_B200_CALL(dcora_problem_project_manifold(h_->p, M.data(), out.data()));
    return out;
  }
  // Projection onto the tangent space at X (LiftedManifold.cpp:37-41,104-108)
  Matrix projectToTangentSpace(const Matrix &X, const Matrix &V) const {
    check(X);
    check(V);
    Matrix out(V.rows(), V.cols());
    DCORA_B200_CALL(dcora_problem_project_tangent(h_->p, X.data(), V.data(), out.data()));
    return out;
  }

 protected:
  LiftedManifoldBase(unsigned r, unsigned d, unsigned n, unsigned l, unsigned b, GraphType t,
                     int device)
      : r_(r), k_((d + 1) * n + l + b) {
    h_ = detail::make_handle(r, d, n, l, b, t, detail::empty_matrix(k_), nullptr,
                             DCORA_REG_NONE, device);
  }
  void check(const Matrix &M) const {
    DCORA_B200_CHECK(M.rows() == r_ && M.cols() == k_, "matrix must be r x k");
  }
  unsigned r_, k_;
  std::shared_ptr<detail::Handle> h_;
};

class LiftedSEManifold : public LiftedManifoldBase {
 public:
  LiftedSEManifold(unsigned r, unsigned d, unsigned n, int device = 0)
      : LiftedManifoldBase(r, d, n, 0, 0, GraphType::PoseGraph, device) {}
};

class LiftedRAManifold : public LiftedManifoldBase {
 public:
  LiftedRAManifold(unsigned r, unsigned d, unsigned n, unsigned l, unsigned b, int device = 0)
      : LiftedManifoldBase(r, d, n, l, b, GraphType::RangeAidedSLAMGraph, device) {}
};

// ---------------------------------------------------------- QuadraticProblem
// f(X) = 0.5 <Q, X^T X> + <X, G>   (QuadraticProblem.h:31-40)
class QuadraticProblem {
 public:
  explicit QuadraticProblem(const std::shared_ptr<Graph> &graph, int device = 0)
      : graph_(graph), device_(device) {
    DCORA_B200_CHECK(graph_ != nullptr, "graph");
    use_se_manifold_ = graph_->graphType() == GraphType::PoseGraph;
    graph_->constructDataMatrices();
    h_ = detail::make_handle(graph_->r(), graph_->d(), graph_->n(), graph_->l(), graph_->b(),
                             graph_->graphType(), graph_->quadraticMatrix(),
                             graph_->hasLinearMatrix() ? &graph_->linearMatrix() : nullptr,
                             graph_->preconditionerRegularization(), device);
  }
  // Same Q and factor at another relaxation rank (the staircase re-creates the
  // Graph per rank, MultiRobotExample.cpp:172-177, but Q does not depend on r).
  QuadraticProblem(const QuadraticProblem &other, unsigned r) : device_(other.device_) {
    graph_ = std::make_shared<Graph>(0, r, other.dimension(), other.graph_->graphType());
    graph_->setQuadraticMatrix(other.num_poses(), other.num_unit_spheres(),
                               other.num_landmarks(), other.graph_->quadraticMatrix());
    use_se_manifold_ = other.use_se_manifold_;
    h_ = std::make_shared<detail::Handle>();
    DCORA_B200_CALL(dcora_problem_clone_rank(other.h_->p, (int32_t)r, &h_->p));
  }

  unsigned dimension() const { return graph_->d(); }
  unsigned relaxation_rank() const { return graph_->r(); }
  unsigned num_poses() const { return graph_->n(); }
  unsigned num_unit_spheres() const { return graph_->l(); }
  unsigned num_landmarks() const { return graph_->b(); }
  unsigned problem_dimension() const { return graph_->k(); }
  bool useSEManifold() const { return use_se_manifold_; }

  // Graph::setNeighborStates -> constructG: replace the linear term
  void setLinearMatrix(const Matrix &G) {
    check(G);
    graph_->setLinearMatrix(G);
    DCORA_B200_CALL(dcora_problem_set_G(h_->p, G.data()));
  }

  double f(const Matrix &Y) const {  // QuadraticProblem.cpp:38-44
    check(Y);
    double v = 0.0;
    DCORA_B200_CALL(dcora_problem_f(h_->p, Y.data(), &v));
    return v;
  }
  Matrix EucGrad(const Matrix &X) const {  // :53-59
    return unary(dcora_problem_egrad, X);
  }
  Matrix EucHessianEta(const Matrix &V) const {  // :61-68 (the Hessian is constant)
    return unary(dcora_problem_hessvec, V);
  }
  Matrix PreConditioner(const Matrix &X, const Matrix &inVec) const {  // :70-84
    check(X);
    check(inVec);
    Matrix out(inVec.rows(), inVec.cols());
    DCORA_B200_CALL(dcora_problem_precond(h_->p, X.data(), inVec.data(), out.data()));
    return out;
  }
  Matrix RieGrad(const Matrix &Y) const { return unary(dcora_problem_rgrad, Y); }  // :86-119
  double RieGradNorm(const Matrix &Y) const {                                       // :121-123
    check(Y);
    double v = 0.0;
    DCORA_B200_CALL(dcora_problem_rgradnorm(h_->p, Y.data(), &v));
    return v;
  }
  Matrix Retract(const Matrix &Y, const Matrix &V) const {  // :125-136
    check(Y);
    check(V);
    Matrix out(Y.rows(), Y.cols());
    DCORA_B200_CALL(dcora_problem_retract(h_->p, Y.data(), V.data(), out.data(), DCORA_RETRACT_QF));
    return out;
  }
  // :138-234.  *this is the rank-r problem, Xopt the (r-1) x k critical point.
  bool escapeSaddle(const Matrix &Xopt, double theta, const Vector &v, double gradient_tolerance,
                    double preconditioned_gradient_tolerance, Matrix *X,
                    bool isSecondOrder = false) {
    DCORA_B200_CHECK(X != nullptr, "X");
    DCORA_B200_CHECK(Xopt.rows() + 1 == relaxation_rank() && Xopt.cols() == problem_dimension(),
                     "Xopt must be (r-1) x k");
    DCORA_B200_CHECK((size_t)v.size() == problem_dimension(), "v must have k entries");
    Matrix out(relaxation_rank(), problem_dimension());
    const int rc = DCORA_B200_CALL(dcora_problem_escape_saddle(
        h_->p, Xopt.data(), theta, v.data(), gradient_tolerance,
        preconditioned_gradient_tolerance, isSecondOrder ? 1 : 0, DCORA_RETRACT_QF, out.data()));
    *X = out;
    return rc == DCORA_OK;
  }

  dcora_problem *handle() const { return h_->p; }
  const std::shared_ptr<Graph> &graph() const { return graph_; }

 private:
  template <typename Fn>
  Matrix unary(Fn fn, const Matrix &X) const {
    check(X);
    Matrix out(X.rows(), X.cols());
    DCORA_B200_CALL(fn(h_->p, X.data(), out.data()));
    return out;
  }
  void check(const Matrix &Y) const {  // QuadraticProblem.cpp:39-40
    DCORA_B200_CHECK(Y.rows() == relaxation_rank(), "rows != r");
    DCORA_B200_CHECK(Y.cols() == problem_dimension(), "cols != k");
  }
  std::shared_ptr<Graph> graph_;
  bool use_se_manifold_ = true;
  int device_ = 0;
  std::shared_ptr<detail::Handle> h_;
};

// -------------------------------------------------------- QuadraticOptimizer
class QuadraticOptimizer {  // QuadraticOptimizer.h:25-147
 public:
  explicit QuadraticOptimizer(QuadraticProblem *p, ROptParameters params = ROptParameters())
      : problem_(p), params_(params) {}
  void setProblem(QuadraticProblem *p) { problem_ = p; }
  void setVerbose(bool v) { params_.verbose = v; }
  void setAlgorithm(ROptParameters::ROptMethod alg) { params_.method = alg; }
  void setRGDStepsize(double s) { params_.RGD_stepsize = s; }
  void setRTRIterations(int iter) { params_.RTR_iterations = iter; }
  void setGradientNormTolerance(double tol) { params_.gradnorm_tol = tol; }
  void setRTRInitialRadius(double radius) { params_.RTR_initial_radius = radius; }
  void setRTRtCGIterations(int iter) { params_.RTR_tCG_iterations = iter; }
  ROPTResult getOptResult() const { return result_; }

  Matrix optimize(const Matrix &Y) {  // QuadraticOptimizer.cpp:28-50
    DCORA_B200_CHECK(problem_ != nullptr, "problem");
    DCORA_B200_CHECK(Y.rows() == problem_->relaxation_rank() &&
                         Y.cols() == problem_->problem_dimension(), "Y must be r x k");
    dcora_ropt_params cp;
    dcora_ropt_params_default(&cp);
    cp.method = params_.method == ROptParameters::ROptMethod::RTR ? 0 : 1;
    cp.verbose = params_.verbose;
    cp.gradnorm_tol = params_.gradnorm_tol;
    cp.RGD_stepsize = params_.RGD_stepsize;
    cp.RGD_use_preconditioner = params_.RGD_use_preconditioner;
    cp.RTR_iterations = params_.RTR_iterations;
    cp.RTR_tCG_iterations = params_.RTR_tCG_iterations;
    cp.RTR_initial_radius = params_.RTR_initial_radius;
    cp.time_bound_s = params_.time_bound_s;
    cp.retraction = params_.retraction;
    dcora_ropt_result cr;
    Matrix out(Y.rows(), Y.cols());
    DCORA_B200_CALL(dcora_rtr_solve(problem_->handle(), Y.data(), out.data(), &cp, nullptr, &cr));
    result_ = ROPTResult(cr.success != 0, cr.fInit, cr.gradNormInit, cr.fOpt, cr.gradNormOpt,
                         cr.elapsedMs);
    result_.tCGStatus = (ROPTLIB::tCGstatusSet)cr.tCGStatus;
    result_.outerIterations = cr.outer_iterations;
    result_.innerIterations = cr.inner_iterations;
    return out;
  }

 private:
  QuadraticProblem *problem_;
  ROptParameters params_;
  ROPTResult result_;
};

// ----------------------------------------------------------------- AgentBatch
// All agents that live on one GPU as ONE batched handle: the device-side state
// behind a set of DCORA::Agent objects of one process (include/DCORA/Agent.h).
// Built from the data matrix of the whole (pose-graph) problem and a contiguous
// partition of the poses over agents (the rule of examples/MultiRobotExample.cpp:
// 56-118): agent a owns the diagonal block Q[a,a] -- what its own Graph builds from
// its private edges and the owned endpoints of its shared edges (Graph.cpp:579-622)
// -- and the off-diagonal blocks become the coupling that turns the neighbours'
// public poses into the linear term G (Graph::constructLinearCostTermPGO, :685-822).
// One RBCD step of a set of mutually non-adjacent agents is
//     updateLinearTerms(X, active);  optimize(&X, active, params);
// i.e. Agent::updateNeighborStates + Agent::iterate(true) (Agent.cpp:535-596,844-906)
// for every active agent at once.
class AgentBatch {
 public:
  AgentBatch(const SparseMatrix &Qglobal, const std::vector<unsigned> &poses_per_agent,
             unsigned r, unsigned d, double precond_reg = 0.1, int device = 0)
      : r_(r), d_(d) {
    const unsigned dh = d + 1;
    const int A = (int)poses_per_agent.size();
    coff_.assign(A + 1, 0);
    for (int a = 0; a < A; ++a) coff_[a + 1] = coff_[a] + (int64_t)dh * poses_per_agent[a];
    DCORA_B200_CHECK(coff_[A] == Qglobal.rows && Qglobal.rows == Qglobal.cols,
                     "partition does not cover the data matrix");
    std::vector<int32_t> agent_of_col(Qglobal.rows);
    for (int a = 0; a < A; ++a)
      for (int64_t c = coff_[a]; c < coff_[a + 1]; ++c) agent_of_col[c] = a;
    // split Q into the agents' diagonal blocks and the cross-agent remainder
    std::vector<SparseMatrix> diag(A);
    for (int a = 0; a < A; ++a) {
      diag[a].rows = diag[a].cols = (int32_t)(coff_[a + 1] - coff_[a]);
      diag[a].rowptr.assign(diag[a].rows + 1, 0);
    }
    cross_.rows = cross_.cols = Qglobal.rows;
    cross_.rowptr.assign(Qglobal.rows + 1, 0);
    adjacency_.assign(A, std::vector<int>());
    for (int32_t row = 0; row < Qglobal.rows; ++row) {
      const int a = agent_of_col[row];
      const int32_t lr = (int32_t)(row - coff_[a]);
      for (int32_t p = Qglobal.rowptr[row]; p < Qglobal.rowptr[row + 1]; ++p) {
        const int32_t c = Qglobal.col[p];
        const int b = agent_of_col[c];
        if (b == a) {
          diag[a].col.push_back((int32_t)(c - coff_[a]));
          diag[a].val.push_back(Qglobal.val[p]);
        } else {
          cross_.col.push_back(c);
          cross_.val.push_back(Qglobal.val[p]);
          if (std::find(adjacency_[a].begin(), adjacency_[a].end(), b) == adjacency_[a].end())
            adjacency_[a].push_back(b);
        }
      }
      diag[a].rowptr[lr + 1] = (int32_t)diag[a].col.size();
      cross_.rowptr[row + 1] = (int32_t)cross_.col.size();
    }
    std::vector<dcora_agent_desc> desc(A);
    for (int a = 0; a < A; ++a) {
      desc[a].n = (int32_t)poses_per_agent[a]; desc[a].l = 0; desc[a].b = 0;
      desc[a].rowptr = diag[a].rowptr.data();
      desc[a].col = diag[a].col.data();
      desc[a].val = diag[a].val.data();
      desc[a].G = nullptr;
      desc[a].precond_reg = precond_reg;
    }
    h_ = std::make_shared<detail::Handle>();
    DCORA_B200_CALL(dcora_problem_create(desc.data(), A, (int32_t)r, (int32_t)d, DCORA_ORDER_SE,
                                         device, &h_->p));
    DCORA_B200_CALL(dcora_problem_set_coupling(h_->p, nullptr, 0, cross_.rowptr.data(),
                                               cross_.col.data(), cross_.val.data()));
  }
  int numAgents() const { return (int)coff_.size() - 1; }
  int64_t problemDimension() const { return coff_.back(); }
  // agents sharing at least one measurement with agent a
  const std::vector<int> &neighbors(int a) const { return adjacency_[a]; }
  // Greedy colouring of the agent graph (largest degree first): agents of one colour share no
  // measurement and may take their RBCD step at the same time.
  std::vector<int> coloring() const {
    const int A = numAgents();
    std::vector<int> order(A), color(A, -1);
    for (int a = 0; a < A; ++a) order[a] = a;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      return adjacency_[x].size() > adjacency_[y].size();
    });
    for (int a : order) {
      int c = 0;
      for (bool clash = true; clash; ++c) {
        clash = false;
        for (int b : adjacency_[a]) clash = clash || color[b] == c;
        if (!clash) break;
      }
      color[a] = c;
    }
    return color;
  }
  // G_a = sum over the neighbours of a of X_nbr * (cross block)  for the active agents
  void updateLinearTerms(const Matrix &X, const std::vector<uint8_t> &active) {
    check(X);
    DCORA_B200_CALL(dcora_problem_update_G(h_->p, X.data(), nullptr, mask(active)));
  }
  // QuadraticOptimizer::optimize of every active agent, in place
  std::vector<ROPTResult> optimize(Matrix *X, const std::vector<uint8_t> &active,
                                   const ROptParameters &params = ROptParameters()) {
    check(*X);
    dcora_ropt_params cp;
    dcora_ropt_params_default(&cp);
    cp.method = params.method == ROptParameters::ROptMethod::RTR ? 0 : 1;
    cp.gradnorm_tol = params.gradnorm_tol;
    cp.RGD_stepsize = params.RGD_stepsize;
    cp.RGD_use_preconditioner = params.RGD_use_preconditioner;
    cp.RTR_iterations = params.RTR_iterations;
    cp.RTR_tCG_iterations = params.RTR_tCG_iterations;
    cp.RTR_initial_radius = params.RTR_initial_radius;
    cp.time_bound_s = params.time_bound_s;
    cp.retraction = params.retraction;
    std::vector<dcora_ropt_result> cr(numAgents());
    DCORA_B200_CALL(dcora_rtr_solve(h_->p, X->data(), X->data(), &cp, mask(active), cr.data()));
    std::vector<ROPTResult> out;
    for (const auto &c : cr) {
      ROPTResult res(c.success != 0, c.fInit, c.gradNormInit, c.fOpt, c.gradNormOpt, c.elapsedMs);
      res.tCGStatus = (ROPTLIB::tCGstatusSet)c.tCGStatus;
      res.outerIterations = c.outer_iterations;
      res.innerIterations = c.inner_iterations;
      out.push_back(res);
    }
    return out;
  }
  // cost of the partitioned problem and norm of its Riemannian gradient: the centralised
  // evaluation the drivers do every iteration (MultiRobotExample.cpp:264-281), blockwise.
  // The linear terms must be current for ALL agents (updateLinearTerms(X, {})).
  void evaluate(const Matrix &X, double *cost, double *gradnorm) {
    check(X);
    const int A = numAgents();
    std::vector<double> quad(A), lin(A), gn(A);
    DCORA_B200_CALL(dcora_problem_cost_terms(h_->p, X.data(), quad.data(), lin.data()));
    DCORA_B200_CALL(dcora_problem_rgradnorm(h_->p, X.data(), gn.data()));
    double f = 0.0, g2 = 0.0;
    for (int a = 0; a < A; ++a) { f += quad[a] + 0.5 * lin[a]; g2 += gn[a] * gn[a]; }
    if (cost) *cost = f;
    if (gradnorm) *gradnorm = std::sqrt(g2);
  }
  dcora_problem *handle() const { return h_->p; }

 private:
  const uint8_t *mask(const std::vector<uint8_t> &active) const {
    if (active.empty()) return nullptr;
    DCORA_B200_CHECK((int)active.size() == numAgents(), "one flag per agent");
    return active.data();
  }
  void check(const Matrix &X) const {
    DCORA_B200_CHECK(X.rows() == r_ && (int64_t)X.cols() == coff_.back(), "X must be r x K");
  }
  unsigned r_, d_;
  std::vector<int64_t> coff_;
  SparseMatrix cross_;
  std::vector<std::vector<int>> adjacency_;
  std::shared_ptr<detail::Handle> h_;
};

// ------------------------------------------------------------ certification
// fastVerification (DCORA_utils.cpp:1713-1735) on the certificate of X:
// S = Q - Lambda(X) (constructDualCertificateMatrix{PGO,RASLAM}, :1898-1982) is
// applied implicitly on the device instead of being assembled and handed over
// as a SparseMatrix -- the only signature change on this path.  Returns true
// iff S + eta I is positive semidefinite; *theta = x^T S x for the computed
// minimum eigenvector x (always available here, also when the test passes).
struct CertificateInfo {
  double lambda_min = 0.0, residual = 0.0;
  int matvecs = 0;
  bool converged = false;
};
inline bool fastVerification(const QuadraticProblem &problem, const Matrix &X, double eta,
                             double *theta, Vector *x, CertificateInfo *info = nullptr,
                             int max_matvecs = 20000, double tol = 1e-9) {
  DCORA_B200_CHECK(X.rows() == problem.relaxation_rank() && X.cols() == problem.problem_dimension(),
                   "X must be r x k");
  dcora_certify_result res;
  Vector v(problem.problem_dimension());
  DCORA_B200_CALL(dcora_certify(problem.handle(), X.data(), eta, max_matvecs, tol, &res, v.data()));
  if (theta) *theta = res.theta;
  if (x) *x = v;
  if (info) {
    info->lambda_min = res.lambda_min;
    info->residual = res.residual;
    info->matvecs = res.matvecs;
    info->converged = res.converged != 0;
  }
  return res.is_psd != 0;
}

}  // namespace DCORA
#endif  // DCORA_B200_DCORA_HPP_
