// Hand-written sm_100a kernels of the DCORA hot path: K1 Q-SpMM, K2 manifold
// ops, and the BLAS-1 pieces of K3 (tCG).  fp64, CUDA cores only: the path is
// HBM-bandwidth-bound sparse work, tensor cores are not used.
//
// Thread mapping shared by every pose kernel: a pose block (r x (d+1),
// column-major, contiguous) is owned by a group of DH=d+1 adjacent lanes, lane
// `a` of the group owns column `a` (r contiguous doubles).  A warp therefore
// covers 32/DH consecutive pose blocks = one contiguous run of memory, so
// loads/stores of X, V, OUT are coalesced, and the (d+1)x(d+1) Q blocks
// (128 B for d=3) are read as one line per group.
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "solver.cuh"

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
template <int R, int DH>
__global__ void __launch_bounds__(128)
bsr_spmm_kernel(int np, const int *__restrict__ browptr,
                const int *__restrict__ bcol, const double *__restrict__ bval,
                const int *__restrict__ prowcol, const int *__restrict__ pose_agent,
                const int *__restrict__ flags, int mask_flag,
                const int2 *__restrict__ tiles,
                const double *__restrict__ X, const double *__restrict__ G,
                double *__restrict__ Y) {
  constexpr int PPW = Map<DH>::PPW;
  const int lane = threadIdx.x & 31;
  const int sub = lane / DH, a = lane - sub * DH;
  // a CTA covers 4*PPW consecutive pose rows: of the whole batch, or of one
  // tile of the active agents (tiles never straddle agents)
  int i, iend;
  if (tiles != nullptr) {
    const int2 t = tiles[blockIdx.x];
    i = t.x + (threadIdx.x >> 5) * PPW + sub;
    iend = t.y;
  } else {
    i = ((blockIdx.x * 128 + threadIdx.x) >> 5) * PPW + sub;
    iend = np;
  }
  if (sub >= PPW || i >= iend) return;
  if (mask_flag >= 0 && flags[pose_agent[i] * NF + mask_flag] == 0) return;
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
      for (int q = 0; q < R; ++q) acc[q] = fma(qv[b], __ldg(xj + b * R + q), acc[q]);
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
template <int R>
__global__ void __launch_bounds__(128)
rem_spmm_kernel(int nrows, const int *__restrict__ rem_row,
                const int *__restrict__ rem_rowptr, const int *__restrict__ rem_col,
                const double *__restrict__ rem_val,
                const uint8_t *__restrict__ rem_assign,
                const int *__restrict__ col_agent_of_row,
                const int *__restrict__ flags, int mask_flag,
                const int *__restrict__ list,
                const double *__restrict__ X, const double *__restrict__ G,
                double *__restrict__ Y) {
  int s = blockIdx.x * 128 + threadIdx.x;
  if (s >= nrows) return;
  if (list != nullptr) s = list[s];
  if (mask_flag >= 0 && flags[col_agent_of_row[s] * NF + mask_flag] == 0) return;
  double acc[R];
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
  for (int p = rem_rowptr[s]; p < rem_rowptr[s + 1]; ++p) {
    const double v = rem_val[p];
    const double *x = X + (size_t)rem_col[p] * R;
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = fma(v, x[q], acc[q]);
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

template <int R, int DH, int OP>
__global__ void __launch_bounds__(128)
manifold_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
                int mask_flag, const int *__restrict__ list, const double *__restrict__ X,
                const double *__restrict__ V, const double *__restrict__ HV,
                const double *S, double *__restrict__ OUT,
                double *__restrict__ AUX, double *__restrict__ partials) {
  // `partials` (optional, OP_RHESS / OP_GRAD / OP_PROJECT): the chunk's share of
  // <V, OUT> (OP_RHESS: <delta, H delta>), <OUT, OUT> (OP_GRAD: <gf, gf>) or
  // <OUT, AUX> (OP_PROJECT after K5: <z, r>) goes to partial slot 0 -- the fused
  // form of a following dots_kernel launch.
  constexpr int D = DH - 1;
  constexpr int PPW = Map<DH>::PPW;
  __shared__ double red_sm[4];
  const int cid = list ? list[blockIdx.x] : blockIdx.x;
  const Chunk ch = chunks[cid];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  double dsum = 0.0;
  if (ch.kind == KIND_POSE) {
    const int nposes = ch.ncols / DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / DH, a = lane - sub * DH;
    for (int base = warp * PPW; base < nposes; base += 4 * PPW) {
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
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double v = V[o + q];
            OUT[o + q] = v;
            if (partials != nullptr) dsum = fma(v, AUX[o + q], dsum);
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
        load_block<R, D>(V + pb, Vb);
        stiefel_proj_col<R, D>(Yb, Vb, a, out);
        if (partials != nullptr) {
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
  for (int c = threadIdx.x; c < ch.ncols; c += 128) {
    const size_t o = (size_t)(ch.col0 + c) * R;
    if (ch.kind == KIND_EUCLID) {
      if (OP == OP_SYM) continue;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        double v;
        if (OP == OP_PROJECT) { v = V[o + q]; if (partials != nullptr) dsum = fma(v, AUX[o + q], dsum); }
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
#pragma unroll
        for (int q = 0; q < R; ++q) w[q] = V[o + q];
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
        if (OP == OP_PROJECT && partials != nullptr) dsum = fma(oq, AUX[o + q], dsum);
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
  if ((OP == OP_RHESS || OP == OP_GRAD || OP == OP_PROJECT) && partials != nullptr) {
    const double t = block_sum<128>(dsum, red_sm);
    if (threadIdx.x == 0) partials[(size_t)cid * NPART + 0] = t;
  }
}

// ------------------------------------------------------------ BLAS-1 (K3)
struct DotArgs {
  const double *x[NPART];
  const double *y[NPART];
  int n;
};

__global__ void __launch_bounds__(128)
dots_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
            int mask_flag, const int *__restrict__ list, int R, DotArgs args,
            double *__restrict__ partials) {
  __shared__ double sm[4];
  const int cid = list ? list[blockIdx.x] : blockIdx.x;
  const Chunk ch = chunks[cid];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const size_t o = (size_t)ch.col0 * R;
  const int len = ch.ncols * R;
  for (int s = 0; s < args.n; ++s) {
    const double *x = args.x[s] + o, *y = args.y[s] + o;
    double acc = 0.0;
    for (int i = threadIdx.x; i < len; i += 128) acc = fma(x[i], y[i], acc);
    const double t = block_sum<128>(acc, sm);
    if (threadIdx.x == 0) partials[(size_t)cid * NPART + s] = t;
  }
}

__global__ void __launch_bounds__(128)
axpby_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
             int mask_flag, const int *__restrict__ list, int R,
             const double *__restrict__ scal, int slot_a,
             double ca, const double *__restrict__ x, int slot_b, double cb,
             const double *__restrict__ y, double *__restrict__ out) {
  const Chunk ch = chunks[list ? list[blockIdx.x] : blockIdx.x];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const double a = slot_a >= 0 ? ca * scal[ch.agent * NS + slot_a] : ca;
  const double b = slot_b >= 0 ? cb * scal[ch.agent * NS + slot_b] : cb;
  const size_t o = (size_t)ch.col0 * R;
  const int len = ch.ncols * R;
  if (y != nullptr) {
    for (int i = threadIdx.x; i < len; i += 128) out[o + i] = fma(a, x[o + i], b * y[o + i]);
  } else {
    for (int i = threadIdx.x; i < len; i += 128) out[o + i] = a * x[o + i];
  }
}

// tCG start: r = gf, eta = 0, Heta = 0 in one pass
__global__ void __launch_bounds__(128)
tcg_begin_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
                 int mask_flag, const int *__restrict__ list, int R,
                 const double *__restrict__ gf, double *__restrict__ r,
                 double *__restrict__ eta, double *__restrict__ Heta) {
  const Chunk ch = chunks[list ? list[blockIdx.x] : blockIdx.x];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const size_t o = (size_t)ch.col0 * R;
  const int len = ch.ncols * R;
  for (int i = threadIdx.x; i < len; i += 128) {
    r[o + i] = gf[o + i];
    eta[o + i] = 0.0;
    Heta[o + i] = 0.0;
  }
}

// tCG step: eta += c_eta delta, Heta += c_eta Hd, r += c_r Hd, partial 0 = <r, r>
// (c_eta = scal[S_CETA], c_r = scal[S_CR] of the agent)
__global__ void __launch_bounds__(128)
tcg_update_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
                  int mask_flag, const int *__restrict__ list, int R,
                  const double *__restrict__ scal, int slot_eta, int slot_r,
                  const double *__restrict__ delta, const double *__restrict__ Hd,
                  double *__restrict__ eta, double *__restrict__ Heta, double *__restrict__ r,
                  double *__restrict__ partials) {
  __shared__ double sm[4];
  const int cid = list ? list[blockIdx.x] : blockIdx.x;
  const Chunk ch = chunks[cid];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const double ce = scal[ch.agent * NS + slot_eta], cr = scal[ch.agent * NS + slot_r];
  const size_t o = (size_t)ch.col0 * R;
  const int len = ch.ncols * R;
  double acc = 0.0;
  for (int i = threadIdx.x; i < len; i += 128) {
    const double h = Hd[o + i];
    eta[o + i] = fma(1.0, eta[o + i], ce * delta[o + i]);
    Heta[o + i] = fma(1.0, Heta[o + i], ce * h);
    const double rn = fma(1.0, r[o + i], cr * h);
    r[o + i] = rn;
    acc = fma(rn, rn, acc);
  }
  const double t = block_sum<128>(acc, sm);
  if (threadIdx.x == 0) partials[(size_t)cid * NPART + 0] = t;
}

// reference <-> internal column permutation (gather)
__global__ void __launch_bounds__(256)
permute_kernel(int64_t K, int R, const int *__restrict__ perm,
               const double *__restrict__ in, double *__restrict__ out,
               int to_internal) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= K * R) return;
  const int64_t c = t / R;
  const int q = (int)(t - c * R);
  const int64_t rc = perm[c];  // internal col c <-> reference col rc
  if (to_internal) out[c * R + q] = in[rc * R + q];
  else out[rc * R + q] = in[c * R + q];
}

// internal -> reference write-back of the columns of the listed chunks only
__global__ void __launch_bounds__(128)
export_chunks_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
                     int mask_flag, const int *__restrict__ list, int R,
                     const int *__restrict__ perm, const double *__restrict__ in,
                     double *__restrict__ out) {
  const Chunk ch = chunks[list ? list[blockIdx.x] : blockIdx.x];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const int len = ch.ncols * R;
  for (int i = threadIdx.x; i < len; i += 128) {
    const int c = ch.col0 + i / R, q = i - (i / R) * R;
    const size_t rc = perm ? (size_t)perm[c] : (size_t)c;
    out[rc * R + q] = in[(size_t)c * R + q];
  }
}

// scal[a*NS + dst[s]] = sum over agent a's chunks of partial slot s
struct ReduceArgs { int n; int dst[NPART]; };
__global__ void __launch_bounds__(128)
reduce_scalars_kernel(const int *__restrict__ agent_chunk0,
                      const double *__restrict__ partials, ReduceArgs args,
                      double *__restrict__ scal) {
  __shared__ double sm[4];
  const int a = blockIdx.x;
  const int c0 = agent_chunk0[a], c1 = agent_chunk0[a + 1];
  for (int s = 0; s < args.n; ++s) {
    double acc = 0.0;
    for (int c = c0 + threadIdx.x; c < c1; c += 128) acc += partials[(size_t)c * NPART + s];
    const double t = block_sum<128>(acc, sm);
    if (threadIdx.x == 0) scal[a * NS + args.dst[s]] = t;
  }
}

}  // namespace

// ------------------------------------------------------------- dispatchers
int launch_spmm_mat(dcora_problem *p, const SpmmMat &A, const double *X,
                    const double *G, double *Y, int mask_flag) {
  DeviceMatrices &M = *p->M;
  const ActiveSet *as = active_set(p, mask_flag);
  if (A.np > 0 && (!as || as->n_tiles > 0)) {
    DCORA_DISPATCH_RD(p->r, M.d, {
      const int ppc = spmm_rows_per_cta(M.d);
      const int grid = as ? as->n_tiles : (A.np + ppc - 1) / ppc;
      bsr_spmm_kernel<R, DH><<<grid, 128, 0, p->stream>>>(
          A.np, A.browptr, A.bcol, A.bval, A.prowcol, A.pose_agent, p->flags,
          mask_flag, as ? as->tiles : nullptr, X, G, Y);
    })
    DCORA_LAUNCH_CHECK(p);
  }
  const int *rlist = as ? (A.coupling ? as->crem_rows : as->rem_rows) : nullptr;
  const int64_t nrows = as ? (A.coupling ? as->n_crem : as->n_rem) : A.nrem_rows;
  if (nrows > 0) {
    DCORA_DISPATCH_R(p->r, {
      const int grid = (int)((nrows + 127) / 128);
      rem_spmm_kernel<R><<<grid, 128, 0, p->stream>>>(
          (int)nrows, A.rem_row, A.rem_rowptr, A.rem_col, A.rem_val,
          A.rem_assign, A.rem_agent, p->flags, mask_flag, rlist, X, G, Y);
    })
    DCORA_LAUNCH_CHECK(p);
  }
  return DCORA_OK;
}

int launch_spmm(dcora_problem *p, const double *X, const double *G, double *Y,
                int mask_flag) {
  DeviceMatrices &M = *p->M;
  SpmmMat A;
  A.np = M.np; A.browptr = M.browptr; A.bcol = M.bcol; A.bval = M.bval;
  A.prowcol = M.prowcol; A.pose_agent = M.pose_agent;
  A.nrem_rows = M.nrem_rows; A.rem_row = M.rem_row; A.rem_rowptr = M.rem_rowptr;
  A.rem_col = M.rem_col; A.rem_val = M.rem_val; A.rem_assign = M.rem_assign;
  A.rem_agent = M.rem_agent;
  return launch_spmm_mat(p, A, X, G, Y, mask_flag);
}

template <int OP>
static int launch_manifold(dcora_problem *p, const double *X, const double *V,
                           const double *HV, const double *S, double *OUT,
                           int mask_flag, double *AUX = nullptr, double *partials = nullptr) {
  DeviceMatrices &M = *p->M;
  if (M.nchunks == 0) return DCORA_OK;
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  DCORA_DISPATCH_RD(p->r, M.d, {
    if (R < DH - 1) {
      set_error("manifold op needs r >= d");
      return DCORA_ERR_ARG;
    }
    manifold_kernel<R, DH, OP><<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
        M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, X, V, HV, S, OUT, AUX,
        partials);
  })
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_project_tangent(dcora_problem *p, const double *X, const double *V,
                           double *OUT, int mask_flag, const double *dot_with) {
  // dot_with (optional): also leaves <OUT, dot_with> in partial slot 0
  return launch_manifold<OP_PROJECT>(p, X, V, nullptr, nullptr, OUT, mask_flag,
                                     const_cast<double *>(dot_with),
                                     dot_with ? p->partials : nullptr);
}
int launch_rhess_correct(dcora_problem *p, const double *X, const double *S,
                         const double *V, const double *HV, double *OUT,
                         int mask_flag, bool with_dot) {
  return launch_manifold<OP_RHESS>(p, X, V, HV, S, OUT, mask_flag, nullptr,
                                   with_dot ? p->partials : nullptr);
}
// EG = XQ + G, gf = proj_X(EG), S = sym(X^T EG), partial 0 = <gf, gf>: the
// gradient stage of RTR in one pass (replaces axpby + project + sym + dots)
int launch_gradient_fused(dcora_problem *p, const double *X, const double *XQ,
                          const double *G, double *EG, double *gf, double *S,
                          int mask_flag) {
  return launch_manifold<OP_GRAD>(p, X, XQ, G, S, gf, mask_flag, EG, p->partials);
}
int launch_sym_ytg(dcora_problem *p, const double *X, const double *EG, double *S,
                   int mask_flag) {
  return launch_manifold<OP_SYM>(p, X, EG, nullptr, nullptr, S, mask_flag);
}
int launch_retract(dcora_problem *p, const double *X, const double *V, double *OUT,
                   int kind, int mask_flag) {
  if (kind == DCORA_RETRACT_POLAR)
    return launch_manifold<OP_RETRACT_POLAR>(p, X, V, nullptr, nullptr, OUT, mask_flag);
  return launch_manifold<OP_RETRACT_QF>(p, X, V, nullptr, nullptr, OUT, mask_flag);
}
int launch_project_manifold(dcora_problem *p, const double *Mx, double *OUT,
                            int mask_flag) {
  return launch_manifold<OP_PROJ_MANIFOLD>(p, Mx, nullptr, nullptr, nullptr, OUT, mask_flag);
}

int launch_dots(dcora_problem *p, int n, const double *const *x,
                const double *const *y, int mask_flag) {
  DeviceMatrices &M = *p->M;
  DotArgs args;
  args.n = n;
  for (int s = 0; s < n; ++s) { args.x[s] = x[s]; args.y[s] = y[s]; }
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  dots_kernel<<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
      M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, p->r, args, p->partials);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_axpby(dcora_problem *p, double *out, int slot_a, double ca,
                 const double *x, int slot_b, double cb, const double *y,
                 int mask_flag) {
  DeviceMatrices &M = *p->M;
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  axpby_kernel<<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
      M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, p->r, p->scal, slot_a, ca, x,
      slot_b, cb, y, out);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_tcg_begin(dcora_problem *p, const double *gf, double *r, double *eta, double *Heta,
                     int mask_flag) {
  DeviceMatrices &M = *p->M;
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  tcg_begin_kernel<<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
      M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, p->r, gf, r, eta, Heta);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_tcg_update(dcora_problem *p, int slot_eta, int slot_r, const double *delta,
                      const double *Hd, double *eta, double *Heta, double *r, int mask_flag) {
  DeviceMatrices &M = *p->M;
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  tcg_update_kernel<<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
      M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, p->r, p->scal, slot_eta, slot_r,
      delta, Hd, eta, Heta, r, p->partials);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_copy(dcora_problem *p, double *out, const double *x, int mask_flag) {
  return launch_axpby(p, out, -1, 1.0, x, -1, 0.0, nullptr, mask_flag);
}

int launch_permute_cols(dcora_problem *p, const int *perm, int64_t ncols,
                        const double *in, double *out, int r, bool to_internal) {
  const int64_t tot = ncols * r;
  if (tot == 0) return DCORA_OK;
  const int grid = (int)((tot + 255) / 256);
  permute_kernel<<<grid, 256, 0, p->stream>>>(ncols, r, perm, in, out,
                                              to_internal ? 1 : 0);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int export_active_columns(dcora_problem *p, const double *src, double *dst, int mask_flag) {
  DeviceMatrices &M = *p->M;
  const ActiveSet *as = active_set(p, mask_flag);
  if (as && as->n_chunks == 0) return DCORA_OK;
  export_chunks_kernel<<<as ? as->n_chunks : M.nchunks, 128, 0, p->stream>>>(
      M.chunks, p->flags, mask_flag, as ? as->chunks : nullptr, p->r,
      M.identity_perm ? nullptr : M.perm, src, dst);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

int launch_permute(dcora_problem *p, const double *in, double *out, int r,
                   bool to_internal) {
  return launch_permute_cols(p, p->M->perm, p->M->K, in, out, r, to_internal);
}

int launch_reduce_scalars(dcora_problem *p, int nslots, const int *dst_slots) {
  DeviceMatrices &M = *p->M;
  ReduceArgs args;
  args.n = nslots;
  for (int s = 0; s < nslots; ++s) args.dst[s] = dst_slots[s];
  reduce_scalars_kernel<<<M.A, 128, 0, p->stream>>>(M.agent_chunk0, p->partials, args,
                                                    p->scal);
  DCORA_LAUNCH_CHECK(p);
  return DCORA_OK;
}

}  // namespace dcora
