// K5: preconditioner application  OUT = proj_X( (Q + reg I)^-1 V )  as two
// level-scheduled sparse triangular solves on the host-built Cholesky factor
// (replaces graph_->preconditioner()->solve(INVEC^T)^T, QuadraticProblem.cpp:79).
//
// One CTA per agent walks that agent's level schedule; rows of a level are
// independent and are spread over the CTA's warps, the r right-hand sides are
// processed together (the solution is kept as a (row, r) row-major panel, i.e.
// the same memory shape as every other vector).  Levels are separated by
// __syncthreads() only -- agents never wait for each other.
#include "kernels.cuh"

namespace dcora {

void DeviceFactor::release() {
  cudaFree(rowptr); cudaFree(colidx); cudaFree(val);
  cudaFree(t_rowptr); cudaFree(t_colidx); cudaFree(t_val);
  cudaFree(inv_diag); cudaFree(perm);
  cudaFree(fwd_rows); cudaFree(bwd_rows);
  cudaFree(fwd_lvlptr); cudaFree(bwd_lvlptr);
  cudaFree(fwd_lvl0); cudaFree(bwd_lvl0);
  *this = DeviceFactor();
}

namespace {

constexpr int TRSV_THREADS = 512;

// mode 0: forward  L y = P v   (rhs gathered from V through perm)
// mode 1: backward L^T x = y   (in place on the panel, result scattered to OUT)
template <int R, int MODE>
__global__ void __launch_bounds__(TRSV_THREADS)
sptrsv_kernel(const int *__restrict__ flags, int mask_flag,
              const int *__restrict__ rowptr, const int *__restrict__ colidx,
              const double *__restrict__ val, const double *__restrict__ inv_diag,
              const int *__restrict__ perm, const int *__restrict__ rows,
              const int *__restrict__ lvlptr, const int *__restrict__ lvl0,
              const double *__restrict__ V, double *panel, double *OUT) {
  const int a = blockIdx.x;
  if (mask_flag >= 0 && flags[a * NF + mask_flag] == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TRSV_THREADS / 32;
  const int l0 = lvl0[a], l1 = lvl0[a + 1];  // this agent's levels: lvlptr[l0..l1]
  for (int lv = l0; lv < l1 - 1; ++lv) {
    const int r0 = lvlptr[lv], r1 = lvlptr[lv + 1];
    for (int ri = r0 + warp; ri < r1; ri += NW) {
      const int row = rows[ri];
      double acc[R];
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = 0.0;
      const int p0 = rowptr[row], p1 = rowptr[row + 1];
      for (int p = p0 + lane; p < p1; p += 32) {
        const double v = __ldg(val + p);
        const double *x = panel + (size_t)__ldg(colidx + p) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = fma(v, __ldcg(x + q), acc[q]);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
      if (lane < R) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (q == lane) s = acc[q];
        const double id = inv_diag[row];
        if (MODE == 0) {
          const double b = V[(size_t)perm[row] * R + lane];
          panel[(size_t)row * R + lane] = (b - s) * id;
        } else {
          const double b = __ldcg(panel + (size_t)row * R + lane);
          const double xv = (b - s) * id;
          panel[(size_t)row * R + lane] = xv;
          OUT[(size_t)perm[row] * R + lane] = xv;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

int upload_factor(DeviceMatrices *M, const std::vector<HostFactor> &HF,
                  const HostLayout &H) {
  DeviceFactor &F = M->F;
  const int A = H.A;
  const int64_t K = H.K;
  int64_t nnzL = 0;
  for (int a = 0; a < A; ++a) nnzL += (int64_t)HF[a].Lx.size();
  if (nnzL - K >= ((int64_t)1 << 31)) {
    set_error("upload_factor: factor too large for int32 indexing");
    return DCORA_ERR_ARG;
  }
  const int64_t noff = nnzL - K;  // strictly lower entries
  std::vector<int> rowptr(K + 1, 0), t_rowptr(K + 1, 0), colidx(noff), t_colidx(noff);
  std::vector<double> val(noff), t_val(noff), inv_diag(K);
  std::vector<int> perm(K), fwd_rows(K), bwd_rows(K);
  std::vector<int> fwd_lvlptr, bwd_lvlptr, fwd_lvl0(A + 1, 0), bwd_lvl0(A + 1, 0);
  // counts
  for (int a = 0; a < A; ++a) {
    const HostFactor &f = HF[a];
    const int64_t c0 = H.coff[a];
    for (int j = 0; j < f.k; ++j) {
      perm[c0 + j] = (int)c0 + f.perm[j];
      inv_diag[c0 + j] = 1.0 / f.Lx[f.Lp[j]];
      t_rowptr[c0 + j + 1] = f.Lp[j + 1] - f.Lp[j] - 1;
      for (int p = f.Lp[j] + 1; p < f.Lp[j + 1]; ++p) rowptr[c0 + f.Li[p] + 1]++;
    }
  }
  for (int64_t i = 0; i < K; ++i) { rowptr[i + 1] += rowptr[i]; t_rowptr[i + 1] += t_rowptr[i]; }
  {
    std::vector<int> fill(rowptr.begin(), rowptr.end() - 1);
    for (int a = 0; a < A; ++a) {
      const HostFactor &f = HF[a];
      const int64_t c0 = H.coff[a];
      for (int j = 0; j < f.k; ++j) {
        int tq = t_rowptr[c0 + j];
        for (int p = f.Lp[j] + 1; p < f.Lp[j + 1]; ++p) {
          const int i = f.Li[p];
          const int w = fill[c0 + i]++;
          colidx[w] = (int)c0 + j;  val[w] = f.Lx[p];       // L(i,j), row i
          t_colidx[tq] = (int)c0 + i; t_val[tq] = f.Lx[p];  // L^T(j,i), row j
          ++tq;
        }
      }
    }
  }
  // level schedules per agent
  std::vector<int> level(K);
  for (int a = 0; a < A; ++a) {
    const int64_t c0 = H.coff[a];
    const int k = HF[a].k;
    for (int dir = 0; dir < 2; ++dir) {
      const std::vector<int> &rp = dir == 0 ? rowptr : t_rowptr;
      const std::vector<int> &ci = dir == 0 ? colidx : t_colidx;
      std::vector<int> &rows = dir == 0 ? fwd_rows : bwd_rows;
      std::vector<int> &lvlptr = dir == 0 ? fwd_lvlptr : bwd_lvlptr;
      std::vector<int> &lvl0 = dir == 0 ? fwd_lvl0 : bwd_lvl0;
      int nlev = 0;
      for (int t = 0; t < k; ++t) {
        const int64_t i = dir == 0 ? c0 + t : c0 + k - 1 - t;
        int lv = 0;
        for (int p = rp[i]; p < rp[i + 1]; ++p) lv = std::max(lv, level[ci[p]] + 1);
        level[i] = lv;
        nlev = std::max(nlev, lv + 1);
      }
      std::vector<int> cnt(nlev + 1, 0);
      for (int t = 0; t < k; ++t) cnt[level[c0 + t] + 1]++;
      for (int q = 0; q < nlev; ++q) cnt[q + 1] += cnt[q];
      lvl0[a] = (int)lvlptr.size();
      for (int q = 0; q <= nlev; ++q) lvlptr.push_back((int)c0 + cnt[q]);
      std::vector<int> fill(cnt.begin(), cnt.end() - 1);
      for (int t = 0; t < k; ++t) rows[c0 + fill[level[c0 + t]]++] = (int)c0 + t;
    }
  }
  fwd_lvl0[A] = (int)fwd_lvlptr.size();
  bwd_lvl0[A] = (int)bwd_lvlptr.size();

  auto up_i = [&](int **dst, const std::vector<int> &src) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(int));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice);
  };
  auto up_d = [&](double **dst, const std::vector<double> &src) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(double));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  DCORA_CUDA(up_i(&F.rowptr, rowptr));
  DCORA_CUDA(up_i(&F.colidx, colidx));
  DCORA_CUDA(up_d(&F.val, val));
  DCORA_CUDA(up_i(&F.t_rowptr, t_rowptr));
  DCORA_CUDA(up_i(&F.t_colidx, t_colidx));
  DCORA_CUDA(up_d(&F.t_val, t_val));
  DCORA_CUDA(up_d(&F.inv_diag, inv_diag));
  DCORA_CUDA(up_i(&F.perm, perm));
  DCORA_CUDA(up_i(&F.fwd_rows, fwd_rows));
  DCORA_CUDA(up_i(&F.bwd_rows, bwd_rows));
  DCORA_CUDA(up_i(&F.fwd_lvlptr, fwd_lvlptr));
  DCORA_CUDA(up_i(&F.bwd_lvlptr, bwd_lvlptr));
  DCORA_CUDA(up_i(&F.fwd_lvl0, fwd_lvl0));
  DCORA_CUDA(up_i(&F.bwd_lvl0, bwd_lvl0));
  F.present = true;
  F.nnzL = nnzL;
  F.A = A;
  return DCORA_OK;
}

int launch_precond(dcora_problem *p, const double *X, const double *V, double *OUT,
                   int mask_flag) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  if (!F.present) {
    // Graph::hasPreconditioner() false: the reference only projects
    // (QuadraticProblem.cpp:78-83)
    return launch_project_tangent(p, X, V, OUT, mask_flag);
  }
  double *panel = p->vec(WS_TRSV_PANEL);
  double *sol = p->vec(WS_TRSV_SOL);
  if (!panel || !sol) return DCORA_ERR_ALLOC;
  DCORA_DISPATCH_R(p->r, {
    sptrsv_kernel<R, 0><<<M.A, TRSV_THREADS, 0, p->stream>>>(
        p->flags, mask_flag, F.rowptr, F.colidx, F.val, F.inv_diag, F.perm,
        F.fwd_rows, F.fwd_lvlptr, F.fwd_lvl0, V, panel, nullptr);
    sptrsv_kernel<R, 1><<<M.A, TRSV_THREADS, 0, p->stream>>>(
        p->flags, mask_flag, F.t_rowptr, F.t_colidx, F.t_val, F.inv_diag, F.perm,
        F.bwd_rows, F.bwd_lvlptr, F.bwd_lvl0, nullptr, panel, sol);
  })
  p->launches += 2;
  DCORA_CUDA(cudaGetLastError());
  return launch_project_tangent(p, X, sol, OUT, mask_flag);
}

}  // namespace dcora
