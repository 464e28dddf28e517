// Rounding of a lifted solution to SE(d) / the range-aided product manifold at rank d.
//
// Replaces (src/DCORA_utils.cpp)
//   projectToRotationGroup          :1661-1675  SVD, det-fix for SO(d)
//   projectSolutionRASLAM           :1984-2031  rank-d truncated SVD of X^T, reflection vote,
//                                               rotations to SO(d), unit spheres normalised
//   alignLiftedTrajectoryToFrame    :2262-2289  R0^T X, rotations to SO(d), first / anchor
//                                               translation to the origin
// The thin SVD of X^T (K x r) is taken through the r x r Gram matrix X X^T = V S^2 V^T
// (device reduction + a tiny host eigen-decomposition): U_d S_d = X^T V_d, i.e. the rank-d
// truncation is V_d^T X.  Everything per column / per pose runs in one kernel each.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "solver.cuh"

namespace dcora {

namespace {

// Nearest rotation (in the reference's sense: U diag(1, .., 1, det(U V^T)) V^T with singular
// values descending) of a D x D matrix M[row][col].  The two leading singular pairs come from
// a Jacobi eigen-decomposition of M^T M; the last pair is completed by cross products, which
// gives det = +1 by construction and stays accurate for (near-)singular M.
template <int D>
__host__ __device__ inline void project_rotation(const double (&M)[D][D], double (&Rm)[D][D]) {
  double A[D][D], V[D][D];
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double s = 0.0;
      for (int k = 0; k < D; ++k) s += M[k][i] * M[k][j];
      A[i][j] = s;
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j)
        if (i != j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < D - 1; ++p)
      for (int q = p + 1; q < D; ++q) {
        const double apq = A[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < D; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < D; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < D; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  // order of the eigenvalues, descending
  int ord[D];
  for (int i = 0; i < D; ++i) ord[i] = i;
  for (int i = 0; i < D; ++i)
    for (int j = i + 1; j < D; ++j)
      if (A[ord[j]][ord[j]] > A[ord[i]][ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
  double v[D][D], u[D][D];   // [pair][component]
  for (int pr = 0; pr < D - 1; ++pr) {
    for (int k = 0; k < D; ++k) v[pr][k] = V[k][ord[pr]];
    double nu = 0.0;
    for (int i = 0; i < D; ++i) {
      double s = 0.0;
      for (int k = 0; k < D; ++k) s += M[i][k] * v[pr][k];
      u[pr][i] = s;
    }
    for (int e = 0; e < pr; ++e) {   // re-orthogonalise against the leading u
      double h = 0.0;
      for (int i = 0; i < D; ++i) h += u[e][i] * u[pr][i];
      for (int i = 0; i < D; ++i) u[pr][i] -= h * u[e][i];
    }
    for (int i = 0; i < D; ++i) nu += u[pr][i] * u[pr][i];
    nu = 1.0 / sqrt(nu);
    for (int i = 0; i < D; ++i) u[pr][i] *= nu;
  }
  if (D == 3) {
    const int l = D - 1;
    u[l][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
    u[l][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
    u[l][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    v[l][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1];
    v[l][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2];
    v[l][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
  } else {
    u[1][0] = -u[0][1]; u[1][1] = u[0][0];
    v[1][0] = -v[0][1]; v[1][1] = v[0][0];
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double s = 0.0;
      for (int pr = 0; pr < D; ++pr) s += u[pr][i] * v[pr][j];
      Rm[i][j] = s;
    }
}

template <int D>
__host__ __device__ inline double det(const double (&M)[D][D]) {
  if (D == 2) return M[0][0] * M[1][1] - M[0][1] * M[1][0];
  return M[0][0] * (M[1][1] * M[2 % D][2 % D] - M[1][2 % D] * M[2 % D][1]) -
         M[0][1] * (M[1][0] * M[2 % D][2 % D] - M[1][2 % D] * M[2 % D][0]) +
         M[0][2 % D] * (M[1][0] * M[2 % D][1] - M[1][1] * M[2 % D][0]);
}

// partial Gram matrices: part[chunk][i * r + j] = sum over the chunk's columns of X[i] X[j]
__global__ void __launch_bounds__(128)
gram_kernel(const Chunk *__restrict__ chunks, int r, const double *__restrict__ X,
            double *__restrict__ part) {
  __shared__ double sm[4];
  const Chunk ch = chunks[blockIdx.x];
  for (int i = 0; i < r; ++i)
    for (int j = i; j < r; ++j) {
      double acc = 0.0;
      for (int c = threadIdx.x; c < ch.ncols; c += 128) {
        const double *x = X + (size_t)(ch.col0 + c) * r;
        acc = fma(x[i], x[j], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0)
        part[(size_t)blockIdx.x * (MAX_RANK * MAX_RANK) + i * r + j] = sm[0] + sm[1] + sm[2] + sm[3];
      __syncthreads();
    }
}

struct RoundArgs {
  double W[3 * MAX_RANK];   // D x r: row i of the rank-d frame (V_d^T, or R0^T)
  double t0[3];             // subtracted from every translation-like column
  int flip;                 // negate the last row (the reflection vote)
  int project;              // 1: rotations to SO(d), spheres normalised; 0: linear map only
};

// OUT (d x K, internal layout) = W X per column, then (project) the per-block projections.
// counts[chunk] = number of rotation blocks of the chunk with positive determinant.
template <int D>
__global__ void __launch_bounds__(128)
round_kernel(const Chunk *__restrict__ chunks, int r, const double *__restrict__ X,
             const RoundArgs A, double *__restrict__ OUT, int *__restrict__ counts) {
  constexpr int DH = D + 1;
  __shared__ int cnt_sm;
  if (threadIdx.x == 0) cnt_sm = 0;
  __syncthreads();
  const Chunk ch = chunks[blockIdx.x];
  int npos = 0;
  if (ch.kind == KIND_POSE) {
    for (int pi = threadIdx.x; pi < ch.ncols / DH; pi += 128) {
      const int c0 = ch.col0 + pi * DH;
      double B[DH][D];   // [column of the pose block][row]
      for (int a = 0; a < DH; ++a) {
        const double *x = X + (size_t)(c0 + a) * r;
        for (int i = 0; i < D; ++i) {
          double s = 0.0;
          for (int q = 0; q < r; ++q) s += A.W[i * r + q] * x[q];
          B[a][i] = (A.flip && i == D - 1) ? -s : s;
        }
      }
      double Mr[D][D], Rr[D][D];
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) Mr[i][j] = B[j][i];
      if (det<D>(Mr) > 0) ++npos;
      if (A.project) {
        project_rotation<D>(Mr, Rr);
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) B[j][i] = Rr[i][j];
      }
      for (int i = 0; i < D; ++i) B[D][i] -= A.t0[i];
      for (int a = 0; a < DH; ++a)
        for (int i = 0; i < D; ++i) OUT[(size_t)(c0 + a) * D + i] = B[a][i];
    }
  } else {
    for (int c = threadIdx.x; c < ch.ncols; c += 128) {
      const double *x = X + (size_t)(ch.col0 + c) * r;
      double y[D];
      double nrm = 0.0;
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int q = 0; q < r; ++q) s += A.W[i * r + q] * x[q];
        y[i] = (A.flip && i == D - 1) ? -s : s;
        nrm += y[i] * y[i];
      }
      if (ch.kind == KIND_SPHERE) {
        if (A.project) { nrm = 1.0 / sqrt(nrm); for (int i = 0; i < D; ++i) y[i] *= nrm; }
      } else {
        for (int i = 0; i < D; ++i) y[i] -= A.t0[i];
      }
      for (int i = 0; i < D; ++i) OUT[(size_t)(ch.col0 + c) * D + i] = y[i];
    }
  }
  if (counts) {
    atomicAdd(&cnt_sm, npos);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = cnt_sm;
  }
}

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (host, n <= 16)
void jacobi_small(int n, std::vector<double> &Am, std::vector<double> &V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  auto at = [&](int i, int j) -> double & { return Am[(size_t)i * n + j]; };
  for (int sweep = 0; sweep < 80; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? dg : off) += at(i, j) * at(i, j);
    if (off <= 1e-34 * (dg + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = at(p, q);
        if (std::fabs(apq) < 1e-300) continue;
        const double tau = (at(q, q) - at(p, p)) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = at(k, p), akq = at(k, q);
          at(k, p) = c * akp - s * akq; at(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = at(p, k), aqk = at(q, k);
          at(p, k) = c * apk - s * aqk; at(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

int run_round(dcora_problem *p, const double *x, const RoundArgs &A, double *out, int *npos_out) {
  DeviceMatrices &M = *p->M;
  int *d_counts = nullptr;
  if (npos_out) DCORA_CUDA(cudaMalloc(&d_counts, sizeof(int) * std::max(1, M.nchunks)));
  if (M.d == 3)
    round_kernel<3><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, x, A, out, d_counts);
  else
    round_kernel<2><<<M.nchunks, 128, 0, p->stream>>>(M.chunks, p->r, x, A, out, d_counts);
  p->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && npos_out) {
    std::vector<int> h(M.nchunks);
    e = cudaMemcpyAsync(h.data(), d_counts, sizeof(int) * M.nchunks, cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    long long s = 0;
    for (int v : h) s += v;
    *npos_out = (int)s;
  }
  cudaFree(d_counts);
  if (e != cudaSuccess) {
    set_error(std::string("rounding: ") + cudaGetErrorString(e));
    return DCORA_ERR_CUDA;
  }
  return DCORA_OK;
}

}  // namespace
}  // namespace dcora

using namespace dcora;

extern "C" {

int dcora_project_to_rotation_group(const double *Mx, int32_t d, double *out) {
  if (!Mx || !out || (d != 2 && d != 3)) return DCORA_ERR_ARG;
  // column-major d x d in and out
  if (d == 3) {
    double Mm[3][3], Rr[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Mm[i][j] = Mx[j * 3 + i];
    project_rotation<3>(Mm, Rr);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[j * 3 + i] = Rr[i][j];
  } else {
    double Mm[2][2], Rr[2][2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) Mm[i][j] = Mx[j * 2 + i];
    project_rotation<2>(Mm, Rr);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) out[j * 2 + i] = Rr[i][j];
  }
  return DCORA_OK;
}

int dcora_problem_round_solution(dcora_problem *p, const double *X, double *Xd) {
  if (!p || !X || !Xd) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  const int r = p->r, d = M.d;
  double *x = p->vec(WS_IO0), *o = p->vec(WS_IO1);
  if (!x || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, r));
  // Gram matrix X X^T
  double *d_part = nullptr;
  DCORA_CUDA(cudaMalloc(&d_part, sizeof(double) * (size_t)std::max(1, M.nchunks) * MAX_RANK * MAX_RANK));
  gram_kernel<<<M.nchunks, 128, 0, p->stream>>>(M.chunks, r, x, d_part);
  p->launches++;
  std::vector<double> hp((size_t)M.nchunks * MAX_RANK * MAX_RANK);
  cudaError_t e = cudaMemcpyAsync(hp.data(), d_part, sizeof(double) * hp.size(),
                                  cudaMemcpyDeviceToHost, p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  cudaFree(d_part);
  DCORA_CUDA(e);
  std::vector<double> G((size_t)r * r, 0.0), V;
  for (int c = 0; c < M.nchunks; ++c)
    for (int i = 0; i < r; ++i)
      for (int j = i; j < r; ++j) G[(size_t)i * r + j] += hp[(size_t)c * MAX_RANK * MAX_RANK + i * r + j];
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < i; ++j) G[(size_t)i * r + j] = G[(size_t)j * r + i];
  jacobi_small(r, G, V);
  std::vector<int> ord(r);
  for (int i = 0; i < r; ++i) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return G[(size_t)a * r + a] > G[(size_t)b * r + b]; });
  RoundArgs A;
  std::memset(&A, 0, sizeof(A));
  for (int i = 0; i < d; ++i)
    for (int q = 0; q < r; ++q) A.W[i * r + q] = V[(size_t)q * r + ord[i]];
  // reflection vote (DCORA_utils.cpp:2011-2021): count the positive determinants first
  A.project = 0;
  int npos = 0;
  DCORA_TRY(run_round(p, x, A, o, &npos));
  A.flip = npos < M.np / 2 ? 1 : 0;
  A.project = 1;
  DCORA_TRY(run_round(p, x, A, o, nullptr));
  return export_matrix(p, o, Xd, d);
}

int dcora_problem_align_trajectory(dcora_problem *p, const double *X, const double *Tw0,
                                   int32_t is_global, double *Td) {
  if (!p || !X || !Tw0 || !Td) return DCORA_ERR_ARG;
  DCORA_CUDA(cudaSetDevice(p->M->device));
  DeviceMatrices &M = *p->M;
  const int r = p->r, d = M.d;
  double *x = p->vec(WS_IO0), *o = p->vec(WS_IO1);
  if (!x || !o) return DCORA_ERR_ALLOC;
  DCORA_TRY(import_matrix(p, X, x, r));
  RoundArgs A;
  std::memset(&A, 0, sizeof(A));
  // W = R0^T (d x r); Tw0 is r x (d+1) column-major: rotation columns then the translation
  for (int i = 0; i < d; ++i)
    for (int q = 0; q < r; ++q) A.W[i * r + q] = Tw0[(size_t)i * r + q];
  // ta: the anchor's translation (global) or the first pose's (local), rotated by R0^T
  std::vector<double> ta(r);
  if (is_global) {
    for (int q = 0; q < r; ++q) ta[q] = Tw0[(size_t)d * r + q];
  } else {
    // first translation column of the batch: internal column d of the first pose block
    DCORA_CUDA(cudaMemcpyAsync(ta.data(), x + (size_t)d * r, sizeof(double) * r,
                               cudaMemcpyDeviceToHost, p->stream));
    DCORA_CUDA(cudaStreamSynchronize(p->stream));
  }
  if (!is_global) {
    // the reference takes ta from the ALREADY rotated trajectory and rotates it once more
    // (t0 = R0^T (R0^T t_first), DCORA_utils.cpp:2275-2279), which needs r == d
    if (r != d) {
      set_error("dcora_problem_align_trajectory: local alignment needs r == d");
      return DCORA_ERR_ARG;
    }
    std::vector<double> tb(d);
    for (int i = 0; i < d; ++i) {
      double s = 0.0;
      for (int q = 0; q < r; ++q) s += A.W[i * r + q] * ta[q];
      tb[i] = s;
    }
    for (int i = 0; i < d; ++i) ta[i] = tb[i];
  }
  for (int i = 0; i < d; ++i) {
    double s = 0.0;
    for (int q = 0; q < r; ++q) s += A.W[i * r + q] * ta[q];
    A.t0[i] = s;
  }
  A.project = 1;
  DCORA_TRY(run_round(p, x, A, o, nullptr));
  return export_matrix(p, o, Td, d);
}

}  // extern "C"
