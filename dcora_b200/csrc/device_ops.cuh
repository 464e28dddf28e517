// Device functions of the hot-path kernels, shared by the stand-alone kernels
// (kernels.cu: one launch per operation over the whole batch) and the per-agent
// cluster kernel (rtr_cluster.cu: the whole RTR/tCG solve of an agent in one launch).
//
// Thread mapping shared by every pose operation: a pose block (r x (d+1),
// column-major, contiguous) is owned by a group of DH=d+1 adjacent lanes, lane
// `a` of the group owns column `a` (r contiguous doubles).  A warp therefore
// covers 32/DH consecutive pose blocks = one contiguous run of memory, so
// loads/stores of X, V, OUT are coalesced, and the (d+1)x(d+1) Q blocks
// (128 B for d=3) are read as one line per group.
#pragma once
#include "kernels.cuh"

namespace dcora {
namespace {

template <int DH>
struct Map {
  static constexpr int PPW = 32 / DH;  // pose blocks per warp
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum (fixed tree); result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *smem /* NT/32 */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += smem[i];
  }
  __syncthreads();
  return t;
}

__device__ __forceinline__ bool agent_on(const int *flags, int agent, int mask_flag) {
  return mask_flag < 0 || flags[agent * NF + mask_flag] != 0;
}

// ---------------------------------------------------------------- K1 SpMM
// Y[:, block i] = sum_j X[:, block j] * Q[j-block, i-block]  (+ G)
// (Round-1 variant study -- wide 128/256-bit loads, register and cp.async
// software pipelines, streaming several rows per lane group -- is recorded in
// profiles/experiments/r01_k1_variants.md: every variant lost to this plain
// gather loop, which keeps 40 warps per SM resident at 48 registers.)
// bval holds, for output block row i and input block j, the (d+1)x(d+1) block
// [a][b] = Q[(j,b),(i,a)], row-major: lane a reads its row (DH contiguous
// doubles; 32 B for d=3), the DH lanes of a pose read the whole 128-byte block.
// One output pose block row i, lane a = output column a of the block.
template <int R, int DH, bool NC>
__device__ __forceinline__ void bsr_row(int i, int a, const int *__restrict__ browptr,
                                        const int *__restrict__ bcol,
                                        const double *__restrict__ bval,
                                        const int *__restrict__ prowcol,
                                        const double *__restrict__ X,
                                        const double *__restrict__ G, double *__restrict__ Y) {
  double acc[R];
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
  const int p0 = browptr[i], p1 = browptr[i + 1];
  for (int p = p0; p < p1; ++p) {
    const int c = __ldg(bcol + p);
    const double *qp = bval + (size_t)p * (DH * DH) + a * DH;
    double qv[DH];
    if (DH == 4) {
      const double2 q01 = __ldg(reinterpret_cast<const double2 *>(qp));
      const double2 q23 = __ldg(reinterpret_cast<const double2 *>(qp) + 1);
      qv[0] = q01.x; qv[1] = q01.y; qv[2] = q23.x; qv[DH - 1] = q23.y;
    } else {
#pragma unroll
      for (int b = 0; b < DH; ++b) qv[b] = __ldg(qp + b);
    }
    const double *xj = X + (size_t)c * R;
#pragma unroll
    for (int b = 0; b < DH; ++b) {
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = fma(qv[b], (NC ? __ldg(xj + b * R + q) : xj[b * R + q]), acc[q]);
    }
  }
  const size_t o = (size_t)(prowcol[i] + a) * R;
  if (G != nullptr) {
#pragma unroll
    for (int q = 0; q < R; ++q) Y[o + q] = acc[q] + G[o + q];
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q) Y[o + q] = acc[q];
  }
}

// scalar remainder rows (unit spheres / landmarks and their couplings)
template <int R, bool NC>
__device__ __forceinline__ void rem_row_apply(int s, const int *__restrict__ rem_row,
                                              const int *__restrict__ rem_rowptr,
                                              const int *__restrict__ rem_col,
                                              const double *__restrict__ rem_val,
                                              const uint8_t *__restrict__ rem_assign,
                                              const double *__restrict__ X,
                                              const double *__restrict__ G,
                                              double *__restrict__ Y) {
  double acc[R];
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
  for (int p = rem_rowptr[s]; p < rem_rowptr[s + 1]; ++p) {
    const double v = rem_val[p];
    const double *x = X + (size_t)rem_col[p] * R;
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = fma(v, (NC ? __ldg(x + q) : x[q]), acc[q]);
  }
  const size_t o = (size_t)rem_row[s] * R;
  if (rem_assign[s]) {
#pragma unroll
    for (int q = 0; q < R; ++q) Y[o + q] = acc[q] + (G ? G[o + q] : 0.0);
  } else {
#pragma unroll
    for (int q = 0; q < R; ++q) Y[o + q] += acc[q];
  }
}

// ------------------------------------------------------------ K2 manifold
template <int R, int D>
__device__ __forceinline__ void load_block(const double *base, double (&B)[D][R]) {
#pragma unroll
  for (int b = 0; b < D; ++b)
#pragma unroll
    for (int q = 0; q < R; ++q) B[b][q] = base[b * R + q];
}

template <int R>
__device__ __forceinline__ double dotR(const double (&x)[R], const double (&y)[R]) {
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < R; ++q) s = fma(x[q], y[q], s);
  return s;
}

// Column a of  V - Y sym(Y^T V)   (Stiefel tangent projection,
// DCORA_utils.cpp:2033-2041; testManifold.cpp:354-390)
template <int R, int D>
__device__ __forceinline__ void stiefel_proj_col(const double (&Y)[D][R],
                                                 const double (&V)[D][R], int a,
                                                 double (&out)[R]) {
  double ya[R], va[R];
#pragma unroll
  for (int b = 0; b < D; ++b)
    if (b == a) {
#pragma unroll
      for (int q = 0; q < R; ++q) { ya[q] = Y[b][q]; va[q] = V[b][q]; }
    }
#pragma unroll
  for (int q = 0; q < R; ++q) out[q] = va[q];
#pragma unroll
  for (int b = 0; b < D; ++b) {
    const double s = 0.5 * (dotR<R>(Y[b], va) + dotR<R>(ya, V[b]));
#pragma unroll
    for (int q = 0; q < R; ++q) out[q] = fma(-s, Y[b][q], out[q]);
  }
}

// In-place thin QR of an R x D block by twice-iterated modified Gram-Schmidt:
// A <- Q with diag(Rfac) > 0 (ROPTLIB Stiefel qf retraction).  Optionally
// returns Rfac (upper triangular, Rf[j][b] = q_j . a_b).
template <int R, int D>
__device__ __forceinline__ void mgs2(double (&A)[D][R], double (*Rf)[D]) {
#pragma unroll
  for (int b = 0; b < D; ++b) {
    double rcol[D];
#pragma unroll
    for (int j = 0; j < D; ++j) rcol[j] = 0.0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        if (j < b) {
          const double h = dotR<R>(A[j], A[b]);
          rcol[j] += h;
#pragma unroll
          for (int q = 0; q < R; ++q) A[b][q] = fma(-h, A[j][q], A[b][q]);
        }
      }
    }
    const double nrm = sqrt(dotR<R>(A[b], A[b]));
    rcol[b] = nrm;
    const double inv = 1.0 / nrm;
#pragma unroll
    for (int q = 0; q < R; ++q) A[b][q] *= inv;
    if (Rf) {
#pragma unroll
      for (int j = 0; j < D; ++j) Rf[j][b] = rcol[j];
    }
  }
}

// inverse-transpose of a DxD matrix (D = 2, 3) via the cofactor matrix
template <int D>
__device__ __forceinline__ void inv_transpose(const double (&M)[D][D], double (&T)[D][D]) {
  if constexpr (D == 2) {
    const double id = 1.0 / (M[0][0] * M[1][1] - M[0][1] * M[1][0]);
    T[0][0] = M[1][1] * id; T[0][1] = -M[1][0] * id;
    T[1][0] = -M[0][1] * id; T[1][1] = M[0][0] * id;
  } else {
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double c10 = M[0][2] * M[2][1] - M[0][1] * M[2][2];
    const double c11 = M[0][0] * M[2][2] - M[0][2] * M[2][0];
    const double c12 = M[0][1] * M[2][0] - M[0][0] * M[2][1];
    const double c20 = M[0][1] * M[1][2] - M[0][2] * M[1][1];
    const double c21 = M[0][2] * M[1][0] - M[0][0] * M[1][2];
    const double c22 = M[0][0] * M[1][1] - M[0][1] * M[1][0];
    const double id = 1.0 / (M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02);
    // inverse = C^T / det  =>  inverse^T = C / det
    T[0][0] = c00 * id; T[0][1] = c01 * id; T[0][2] = c02 * id;
    T[1][0] = c10 * id; T[1][1] = c11 * id; T[1][2] = c12 * id;
    T[2][0] = c20 * id; T[2][1] = c21 * id; T[2][2] = c22 * id;
  }
}

// Polar factor U V^T of an R x D block (projectToStiefelManifold,
// DCORA_utils.cpp:1677-1683): A = Q Rf (MGS2), polar(A) = Q polar(Rf), and
// polar(Rf) by Higham's scaled Newton iteration X <- (g X + X^-T / g)/2.
template <int R, int D>
__device__ __forceinline__ void polar_block(double (&A)[D][R]) {
  double Rf[D][D];
  mgs2<R, D>(A, Rf);
  double Xk[D][D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) Xk[i][j] = (i <= j) ? Rf[i][j] : 0.0;
  for (int it = 0; it < 40; ++it) {
    double T[D][D];
    inv_transpose<D>(Xk, T);
    double nx = 0.0, nt = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) { nx = fma(Xk[i][j], Xk[i][j], nx); nt = fma(T[i][j], T[i][j], nt); }
    const double g = sqrt(sqrt(nt / nx));  // Frobenius scaling
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const double nv = 0.5 * (g * Xk[i][j] + T[i][j] / g);
        const double e = nv - Xk[i][j];
        diff = fma(e, e, diff);
        Xk[i][j] = nv;
      }
    if (diff <= 1e-30 * nx) break;
  }
  // A <- Q * Xk
  double Bm[D][R];
#pragma unroll
  for (int b = 0; b < D; ++b)
#pragma unroll
    for (int q = 0; q < R; ++q) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) s = fma(A[j][q], Xk[j][b], s);
      Bm[b][q] = s;
    }
#pragma unroll
  for (int b = 0; b < D; ++b)
#pragma unroll
    for (int q = 0; q < R; ++q) A[b][q] = Bm[b][q];
}

enum ManifoldOp : int {
  OP_PROJECT = 0,   // OUT = proj_X(V)
  OP_RHESS,         // OUT = proj_X(HV - V*S)
  OP_SYM,           // S = sym(Y^T EG)
  OP_RETRACT_QF,    // OUT = qf(X + V)
  OP_RETRACT_POLAR, // OUT = polar(X + V)
  OP_PROJ_MANIFOLD, // OUT = polar(M) / normalised
  OP_GRAD           // AUX = EG = V + HV(G); OUT = proj_X(EG); S = sym(Y^T EG)
};

// One chunk of a manifold operation, executed by NT cooperating threads (`tid` in
// [0, NT)).  Returns this thread's share of the fused dot product (when `want_dot`):
// <V, OUT> (OP_RHESS: <delta, H delta>), <OUT, OUT> (OP_GRAD: <gf, gf>) or <OUT, AUX>
// (OP_PROJECT after K5: <z, r>).  SROW (OP_PROJECT only): when non-null, row c of V is
// read from V + SROW[c] * R instead of V + c * R (the solution panel of K5 is in factor
// row order).
template <int R, int DH, int OP, int NT>
__device__ __forceinline__ double manifold_chunk(const Chunk ch, int tid,
                                                 const double *__restrict__ X,
                                                 const double *__restrict__ V,
                                                 const double *__restrict__ HV, const double *S,
                                                 double *__restrict__ OUT,
                                                 double *__restrict__ AUX, bool want_dot,
                                                 const int *__restrict__ SROW = nullptr) {
  constexpr int D = DH - 1;
  constexpr int PPW = Map<DH>::PPW;
  double dsum = 0.0;
  if (ch.kind == KIND_POSE) {
    const int nposes = ch.ncols / DH;
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = lane / DH, a = lane - sub * DH;
    for (int base = warp * PPW; base < nposes; base += (NT / 32) * PPW) {
      const int pi = base + sub;
      if (sub >= PPW || pi >= nposes) continue;
      const size_t pb = (size_t)(ch.col0 + pi * DH) * R;  // pose block base
      const size_t o = pb + (size_t)a * R;
      if (OP == OP_SYM) {
        if (a < D) {
          double Yb[D][R], Gb[D][R];
          load_block<R, D>(X + pb, Yb);
          load_block<R, D>(V + pb, Gb);
#pragma unroll
          for (int b = 0; b < D; ++b) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c)
              if (c == a) s = 0.5 * (dotR<R>(Yb[b], Gb[c]) + dotR<R>(Yb[c], Gb[b]));
            OUT[pb + a * D + b] = s;  // S[b][a] (symmetric)
          }
        }
        continue;
      }
      if (a == D) {  // translation column: Euclidean factor
        if (OP == OP_PROJECT) {
          const size_t vo = SROW ? (size_t)(SROW[ch.col0 + pi * DH] + a) * R : o;
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double v = V[vo + q];
            OUT[o + q] = v;
            if (want_dot) dsum = fma(v, AUX[o + q], dsum);
          }
        } else if (OP == OP_RHESS) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double h = HV[o + q];
            OUT[o + q] = h;
            dsum = fma(V[o + q], h, dsum);
          }
        } else if (OP == OP_GRAD) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double e = HV ? V[o + q] + HV[o + q] : V[o + q];
            AUX[o + q] = e;
            OUT[o + q] = e;
            dsum = fma(e, e, dsum);
          }
        } else if (OP == OP_PROJ_MANIFOLD) {
#pragma unroll
          for (int q = 0; q < R; ++q) OUT[o + q] = X[o + q];
        } else {
#pragma unroll
          for (int q = 0; q < R; ++q) OUT[o + q] = X[o + q] + V[o + q];
        }
        continue;
      }
      double Yb[D][R];
      load_block<R, D>(X + pb, Yb);
      double out[R];
      if (OP == OP_GRAD) {
        double Gb[D][R];
        load_block<R, D>(V + pb, Gb);
        if (HV != nullptr) {
#pragma unroll
          for (int b = 0; b < D; ++b)
#pragma unroll
            for (int q = 0; q < R; ++q) Gb[b][q] += HV[pb + b * R + q];
        }
        stiefel_proj_col<R, D>(Yb, Gb, a, out);
#pragma unroll
        for (int b = 0; b < D; ++b) {
          double sv = 0.0;
#pragma unroll
          for (int c = 0; c < D; ++c)
            if (c == a) {
              sv = 0.5 * (dotR<R>(Yb[b], Gb[c]) + dotR<R>(Yb[c], Gb[b]));
#pragma unroll
              for (int q = 0; q < R; ++q) if (b == 0) AUX[o + q] = Gb[c][q];
            }
          const_cast<double *>(S)[pb + a * D + b] = sv;  // S[b][a] (symmetric)
        }
#pragma unroll
        for (int q = 0; q < R; ++q) dsum = fma(out[q], out[q], dsum);
      } else if (OP == OP_PROJECT) {
        double Vb[D][R];
        load_block<R, D>(V + (SROW ? (size_t)SROW[ch.col0 + pi * DH] * R : pb), Vb);
        stiefel_proj_col<R, D>(Yb, Vb, a, out);
        if (want_dot) {
#pragma unroll
          for (int q = 0; q < R; ++q) dsum = fma(out[q], AUX[o + q], dsum);
        }
      } else if (OP == OP_RHESS) {
        double Vb[D][R], Wb[D][R];
        load_block<R, D>(V + pb, Vb);
        load_block<R, D>(HV + pb, Wb);
        double Sm[D][D];
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
          for (int b = 0; b < D; ++b) Sm[b][c] = S[pb + c * D + b];
        // W = HV - V*S  (all columns; needed by the projection)
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
          for (int b = 0; b < D; ++b)
#pragma unroll
            for (int q = 0; q < R; ++q) Wb[c][q] = fma(-Sm[b][c], Vb[b][q], Wb[c][q]);
        stiefel_proj_col<R, D>(Yb, Wb, a, out);
#pragma unroll
        for (int q = 0; q < R; ++q) dsum = fma(V[o + q], out[q], dsum);
      } else if (OP == OP_PROJ_MANIFOLD) {
        polar_block<R, D>(Yb);
#pragma unroll
        for (int b = 0; b < D; ++b)
          if (b == a) {
#pragma unroll
            for (int q = 0; q < R; ++q) out[q] = Yb[b][q];
          }
      } else {
        double Vb[D][R];
        load_block<R, D>(V + pb, Vb);
#pragma unroll
        for (int b = 0; b < D; ++b)
#pragma unroll
          for (int q = 0; q < R; ++q) Yb[b][q] += Vb[b][q];
        if (OP == OP_RETRACT_QF) mgs2<R, D>(Yb, nullptr);
        else polar_block<R, D>(Yb);
#pragma unroll
        for (int b = 0; b < D; ++b)
          if (b == a) {
#pragma unroll
            for (int q = 0; q < R; ++q) out[q] = Yb[b][q];
          }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) OUT[o + q] = out[q];
    }
  } else {
  // unit-sphere (oblique) and Euclidean (landmark) columns: one thread each
  for (int c = tid; c < ch.ncols; c += NT) {
    const size_t o = (size_t)(ch.col0 + c) * R;
    if (ch.kind == KIND_EUCLID) {
      if (OP == OP_SYM) continue;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        double v;
        if (OP == OP_PROJECT) {
          v = V[(SROW ? (size_t)SROW[ch.col0 + c] * R : o) + q];
          if (want_dot) dsum = fma(v, AUX[o + q], dsum);
        }
        else if (OP == OP_RHESS) { v = HV[o + q]; dsum = fma(V[o + q], v, dsum); }
        else if (OP == OP_GRAD) {
          v = HV ? V[o + q] + HV[o + q] : V[o + q];
          AUX[o + q] = v;
          dsum = fma(v, v, dsum);
        }
        else if (OP == OP_PROJ_MANIFOLD) v = X[o + q];
        else v = X[o + q] + V[o + q];
        OUT[o + q] = v;
      }
      continue;
    }
    double y[R];
#pragma unroll
    for (int q = 0; q < R; ++q) y[q] = X[o + q];
    if (OP == OP_SYM) {
      double g[R];
#pragma unroll
      for (int q = 0; q < R; ++q) g[q] = V[o + q];
      OUT[o] = dotR<R>(y, g);
    } else if (OP == OP_GRAD) {
      double g[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        g[q] = HV ? V[o + q] + HV[o + q] : V[o + q];
        AUX[o + q] = g[q];
      }
      const double t = dotR<R>(y, g);
      const_cast<double *>(S)[o] = t;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double gq = fma(-t, y[q], g[q]);
        OUT[o + q] = gq;
        dsum = fma(gq, gq, dsum);
      }
    } else if (OP == OP_PROJECT || OP == OP_RHESS) {
      double w[R];
      if (OP == OP_PROJECT) {
        const size_t vo = SROW ? (size_t)SROW[ch.col0 + c] * R : o;
#pragma unroll
        for (int q = 0; q < R; ++q) w[q] = V[vo + q];
      } else {
        const double s = S[o];
#pragma unroll
        for (int q = 0; q < R; ++q) w[q] = fma(-s, V[o + q], HV[o + q]);
      }
      const double t = dotR<R>(y, w);  // v - y (y^T v), DCORA_utils.cpp:2043-2051
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const double oq = fma(-t, y[q], w[q]);
        OUT[o + q] = oq;
        if (OP == OP_RHESS) dsum = fma(V[o + q], oq, dsum);
        if (OP == OP_PROJECT && want_dot) dsum = fma(oq, AUX[o + q], dsum);
      }
    } else {
      if (OP != OP_PROJ_MANIFOLD) {
#pragma unroll
        for (int q = 0; q < R; ++q) y[q] += V[o + q];
      }
      const double inv = 1.0 / sqrt(dotR<R>(y, y));
#pragma unroll
      for (int q = 0; q < R; ++q) OUT[o + q] = y[q] * inv;
    }
  }
  }
  return dsum;
}


}  // namespace
}  // namespace dcora
