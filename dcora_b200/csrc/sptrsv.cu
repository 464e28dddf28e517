// K5: preconditioner application  OUT = proj_X( (Q + reg I)^-1 V )  as a BLOCK
// triangular solve on the host-built Cholesky factor (replaces
// graph_->preconditioner()->solve(INVEC^T)^T, src/QuadraticProblem.cpp:79).
//
// Schedule (trsv_host.cpp): the factor is grouped by the nested-dissection
// blocks; per level of the (log-depth) block elimination tree
//   phase A  t_B = rhs_B - M[B, solved] y     sparse rows: (d+1)x(d+1) pose
//                                             blocks + scalar remainder
//   phase B  y_B = D_B t_B                    dense packed triangular inverse
// forward with M = L, then backward with M = L^T, in ONE launch.
//
// One thread-block CLUSTER per agent (up to 8 CTAs x 512 threads = 128 warps
// on 8 SMs of one GPC): the tasks of a phase are spread over all warps of the
// cluster and phases are separated by the hardware cluster barrier, so agents
// never wait for each other and an agent is not limited to one SM's load
// bandwidth.  The r right-hand sides are processed together; the solution
// panel is (factor row, r) row-major, the same memory shape as every vector.
#include <cooperative_groups.h>

#include "kernels.cuh"
#include "trsv.h"

namespace cg = cooperative_groups;

namespace dcora {

void DeviceFactor::release() {
  for (int d = 0; d < 2; ++d) {
    DeviceTrsvDir &T = dir[d];
    cudaFree(T.ta_row); cudaFree(T.ta_bptr); cudaFree(T.bcol); cudaFree(T.bval);
    cudaFree(T.srowptr); cudaFree(T.scol); cudaFree(T.sval);
    cudaFree(T.sb_row0); cudaFree(T.sb_w); cudaFree(T.sb_direct); cudaFree(T.sb_off);
    cudaFree(T.lb_row0); cudaFree(T.lb_w); cudaFree(T.lb_direct); cudaFree(T.lb_off);
    cudaFree(T.lr_blk); cudaFree(T.lr_i); cudaFree(T.dense);
    cudaFree(T.levels); cudaFree(T.lvl0);
  }
  cudaFree(perm);
  *this = DeviceFactor();
}

namespace {

constexpr int TRSV_THREADS = 512;
constexpr int TRSV_NW = TRSV_THREADS / 32;

template <int CL>
__device__ __forceinline__ void phase_sync() {
  if constexpr (CL == 1) __syncthreads();
  else cg::this_cluster().sync();
}

// Walks one direction of one agent's schedule.  MODE 0: forward (rhs gathered
// from V through perm), MODE 1: backward (rhs = forward solution in Y; the
// result overwrites Y in place and is scattered to OUT through perm).
template <int R, int DH, int CL, int MODE>
__device__ __forceinline__ void trsv_walk(const DeviceTrsvDir &S, int agent, int gwarp,
                                          int nwarps, const int *__restrict__ perm,
                                          const double *__restrict__ V, double *T,
                                          double *Y, double *OUT) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, a = lane & 3;
  const int l0 = S.lvl0[agent], l1 = S.lvl0[agent + 1];
  for (int lv = l0; lv < l1; ++lv) {
    const TrsvLevel L = S.levels[lv];
    // ------------------------------------------------------------ phase A
    if (L.ta1 > L.ta0) {
      const int gpt = L.gpt;          // groups per task
      const int tpw = 8 / gpt;        // tasks per warp
      const int gg = g & (gpt - 1);   // my group's index within the task
      for (int base = L.ta0 + gwarp * tpw; base < L.ta1; base += nwarps * tpw) {
        const int t = base + g / gpt;
        const bool valid = t < L.ta1;
        // every lane of the warp takes part in the shuffles below
        int rr = 0, p0 = 0, p1 = 0;
        if (valid) { rr = __ldg(S.ta_row + t); p0 = __ldg(S.ta_bptr + t); p1 = __ldg(S.ta_bptr + t + 1); }
        const bool pose = rr >= 0;
        const int r0 = pose ? rr : ~rr;
        double acc[R];
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = 0.0;
        const bool rowlane = valid && (pose ? a < DH : a == 0);
        if (rowlane) {
          for (int p = p0 + gg; p < p1; p += gpt) {
            const int c = __ldg(S.bcol + p);
            const double *qp = S.bval + (size_t)p * (DH * DH) + a * DH;
            double qv[DH];
            if constexpr (DH == 4) {
              const double2 q01 = __ldg(reinterpret_cast<const double2 *>(qp));
              const double2 q23 = __ldg(reinterpret_cast<const double2 *>(qp) + 1);
              qv[0] = q01.x; qv[1] = q01.y; qv[2] = q23.x; qv[DH - 1] = q23.y;
            } else {
#pragma unroll
              for (int b = 0; b < DH; ++b) qv[b] = __ldg(qp + b);
            }
            const double *yj = Y + (size_t)c * R;
#pragma unroll
            for (int b = 0; b < DH; ++b) {
#pragma unroll
              for (int q = 0; q < R; ++q) acc[q] = fma(qv[b], yj[b * R + q], acc[q]);
            }
          }
          const int s0 = __ldg(S.srowptr + r0 + a), s1 = __ldg(S.srowptr + r0 + a + 1);
          for (int s = s0 + gg; s < s1; s += gpt) {
            const double v = __ldg(S.sval + s);
            const double *yj = Y + (size_t)__ldg(S.scol + s) * R;
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q] = fma(v, yj[q], acc[q]);
          }
        }
        for (int o = 4; o < 4 * gpt; o <<= 1) {
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        }
        if (rowlane && gg == 0) {
          const size_t o = (size_t)(r0 + a) * R;
          const double *rhs = MODE == 0 ? V + (size_t)__ldg(perm + r0 + a) * R : Y + o;
#pragma unroll
          for (int q = 0; q < R; ++q) T[o + q] = rhs[q] - acc[q];
        }
      }
      phase_sync<CL>();
    }
    // ------------------------------------------------------------ phase B
    // small blocks: one warp per block, lane = row, matrix packed by columns
    for (int s = L.sb0 + gwarp; s < L.sb1; s += nwarps) {
      const int r0 = __ldg(S.sb_row0 + s), w = __ldg(S.sb_w + s);
      const double *M = S.dense + __ldg(S.sb_off + s);
      const bool direct = MODE == 0 && __ldg(S.sb_direct + s) != 0;
      double tv[R], acc[R];
#pragma unroll
      for (int q = 0; q < R; ++q) { tv[q] = 0.0; acc[q] = 0.0; }
      if (lane < w) {
        const double *src = direct ? V + (size_t)__ldg(perm + r0 + lane) * R
                                   : T + (size_t)(r0 + lane) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) tv[q] = src[q];
      }
      int off = 0;
      for (int j = 0; j < w; ++j) {
        double m = 0.0;
        if (MODE == 0) {            // column j holds rows j..w-1
          if (lane >= j && lane < w) m = __ldg(M + off + (lane - j));
          off += w - j;
        } else {                    // column j holds rows 0..j
          if (lane <= j) m = __ldg(M + off + lane);
          off += j + 1;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = fma(m, __shfl_sync(0xffffffffu, tv[q], j), acc[q]);
      }
      if (lane < w) {
        const size_t o = (size_t)(r0 + lane) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) Y[o + q] = acc[q];
        if (MODE == 1) {
          const size_t oo = (size_t)__ldg(perm + r0 + lane) * R;
#pragma unroll
          for (int q = 0; q < R; ++q) OUT[oo + q] = acc[q];
        }
      }
    }
    // rows of large blocks: one warp per row, matrix packed by rows
    for (int t = L.lr0 + gwarp; t < L.lr1; t += nwarps) {
      const int lb = __ldg(S.lr_blk + t), i = __ldg(S.lr_i + t);
      const int r0 = __ldg(S.lb_row0 + lb), w = __ldg(S.lb_w + lb);
      const bool direct = MODE == 0 && __ldg(S.lb_direct + lb) != 0;
      const double *M = S.dense + __ldg(S.lb_off + lb);
      int j0, j1;
      if (MODE == 0) { M += (int64_t)i * (i + 1) / 2; j0 = 0; j1 = i + 1; }
      else { M += (int64_t)i * w - (int64_t)i * (i - 1) / 2; j0 = i; j1 = w; }
      double acc[R];
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = 0.0;
      for (int j = j0 + lane; j < j1; j += 32) {
        const double m = __ldg(M + (j - j0));
        const double *src = direct ? V + (size_t)__ldg(perm + r0 + j) * R
                                   : T + (size_t)(r0 + j) * R;
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = fma(m, src[q], acc[q]);
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
      if (lane < R) {
        double sres = 0.0;
#pragma unroll
        for (int q = 0; q < R; ++q)
          if (q == lane) sres = acc[q];
        Y[(size_t)(r0 + i) * R + lane] = sres;
        if (MODE == 1) OUT[(size_t)__ldg(perm + r0 + i) * R + lane] = sres;
      }
    }
    phase_sync<CL>();
  }
}

template <int R, int DH, int CL>
__global__ void __launch_bounds__(TRSV_THREADS)
trsv_block_kernel(const int *__restrict__ flags, int mask_flag, DeviceTrsvDir fwd,
                  DeviceTrsvDir bwd, const int *__restrict__ perm,
                  const double *__restrict__ V, double *T, double *Y, double *OUT) {
  const int agent = blockIdx.x / CL;
  if (mask_flag >= 0 && flags[agent * NF + mask_flag] == 0) return;  // whole cluster
  const int crank = blockIdx.x - agent * CL;
  const int gwarp = crank * TRSV_NW + (threadIdx.x >> 5);
  constexpr int nwarps = CL * TRSV_NW;
  trsv_walk<R, DH, CL, 0>(fwd, agent, gwarp, nwarps, perm, V, T, Y, OUT);
  trsv_walk<R, DH, CL, 1>(bwd, agent, gwarp, nwarps, perm, V, T, Y, OUT);
}

template <typename T>
cudaError_t up(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc(dst, std::max<size_t>(1, src.size()) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src.empty()) return cudaSuccess;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

template <typename T>
void append_shift(std::vector<T> &dst, const std::vector<T> &src, T shift) {
  const size_t n0 = dst.size();
  dst.resize(n0 + src.size());
  for (size_t i = 0; i < src.size(); ++i) dst[n0 + i] = src[i] + shift;
}

}  // namespace

int upload_factor(DeviceMatrices *M, const std::vector<HostFactor> &HF,
                  std::vector<AgentTrsv> &AT, const HostLayout &H) {
  DeviceFactor &F = M->F;
  const int A = H.A;
  const int64_t K = H.K;
  int64_t nnzL = 0;
  for (int a = 0; a < A; ++a) nnzL += HF[a].nnz;
  std::vector<int> perm(K);
  for (int a = 0; a < A; ++a)
    for (int j = 0; j < HF[a].k; ++j) perm[H.coff[a] + j] = (int)H.coff[a] + HF[a].perm[j];
  DCORA_CUDA(up(&F.perm, perm));
  int64_t stored = 0;
  int max_levels = 0;
  for (int d = 0; d < 2; ++d) {
    // concatenate the agents' schedules with global row / array offsets
    std::vector<int> ta_row, ta_bptr(1, 0), bcol, srowptr(1, 0), scol;
    std::vector<double> bval, sval, dense;
    std::vector<int> sb_row0, sb_w, sb_direct, lb_row0, lb_w, lb_direct, lr_blk, lr_i;
    std::vector<long long> sb_off, lb_off;
    std::vector<TrsvLevel> levels;
    std::vector<int> lvl0(A + 1, 0);
    for (int a = 0; a < A; ++a) {
      AgentTrsvDir &S = AT[a].dir[d];
      const int c0 = (int)H.coff[a];
      if ((int64_t)bcol.size() + (int64_t)S.bcol.size() >= ((int64_t)1 << 31) / 16 * 16 ||
          (int64_t)scol.size() + (int64_t)S.scol.size() >= ((int64_t)1 << 31)) {
        set_error("upload_factor: factor too large for int32 indexing");
        return DCORA_ERR_ARG;
      }
      const int t0 = (int)ta_row.size(), b0 = (int)bcol.size(), s0 = (int)scol.size();
      const int sb0 = (int)sb_row0.size(), lb0 = (int)lb_row0.size(), lr0 = (int)lr_blk.size();
      const int64_t d0 = (int64_t)dense.size();
      for (int v : S.ta_row) ta_row.push_back(v >= 0 ? v + c0 : ~((~v) + c0));
      for (size_t i = 1; i < S.ta_bptr.size(); ++i) ta_bptr.push_back(S.ta_bptr[i] + b0);
      append_shift(bcol, S.bcol, c0);
      bval.insert(bval.end(), S.bval.begin(), S.bval.end());
      for (size_t i = 1; i < S.srowptr.size(); ++i) srowptr.push_back(S.srowptr[i] + s0);
      append_shift(scol, S.scol, c0);
      sval.insert(sval.end(), S.sval.begin(), S.sval.end());
      append_shift(sb_row0, S.sb_row0, c0);
      sb_w.insert(sb_w.end(), S.sb_w.begin(), S.sb_w.end());
      sb_direct.insert(sb_direct.end(), S.sb_direct.begin(), S.sb_direct.end());
      for (int64_t v : S.sb_off) sb_off.push_back((long long)(v + d0));
      append_shift(lb_row0, S.lb_row0, c0);
      lb_w.insert(lb_w.end(), S.lb_w.begin(), S.lb_w.end());
      lb_direct.insert(lb_direct.end(), S.lb_direct.begin(), S.lb_direct.end());
      for (int64_t v : S.lb_off) lb_off.push_back((long long)(v + d0));
      append_shift(lr_blk, S.lr_blk, lb0);
      lr_i.insert(lr_i.end(), S.lr_i.begin(), S.lr_i.end());
      dense.insert(dense.end(), S.dense.begin(), S.dense.end());
      for (TrsvLevel L : S.levels) {
        L.ta0 += t0; L.ta1 += t0; L.sb0 += sb0; L.sb1 += sb0; L.lr0 += lr0; L.lr1 += lr0;
        levels.push_back(L);
      }
      lvl0[a + 1] = (int)levels.size();
      max_levels = std::max(max_levels, (int)S.levels.size());
      S = AgentTrsvDir();  // release host memory early
    }
    if ((int64_t)srowptr.size() != K + 1) {
      set_error("upload_factor: inconsistent schedule");
      return DCORA_ERR_ARG;
    }
    stored += (int64_t)bval.size() + (int64_t)sval.size() + (int64_t)dense.size();
    DeviceTrsvDir &T = F.dir[d];
    DCORA_CUDA(up(&T.ta_row, ta_row));   DCORA_CUDA(up(&T.ta_bptr, ta_bptr));
    DCORA_CUDA(up(&T.bcol, bcol));       DCORA_CUDA(up(&T.bval, bval));
    DCORA_CUDA(up(&T.srowptr, srowptr)); DCORA_CUDA(up(&T.scol, scol));
    DCORA_CUDA(up(&T.sval, sval));
    DCORA_CUDA(up(&T.sb_row0, sb_row0)); DCORA_CUDA(up(&T.sb_w, sb_w));
    DCORA_CUDA(up(&T.sb_direct, sb_direct)); DCORA_CUDA(up(&T.sb_off, sb_off));
    DCORA_CUDA(up(&T.lb_row0, lb_row0)); DCORA_CUDA(up(&T.lb_w, lb_w));
    DCORA_CUDA(up(&T.lb_direct, lb_direct)); DCORA_CUDA(up(&T.lb_off, lb_off));
    DCORA_CUDA(up(&T.lr_blk, lr_blk));   DCORA_CUDA(up(&T.lr_i, lr_i));
    DCORA_CUDA(up(&T.dense, dense));
    DCORA_CUDA(up(&T.levels, levels));   DCORA_CUDA(up(&T.lvl0, lvl0));
  }
  F.present = true;
  F.nnzL = nnzL;
  F.stored_values = stored / 2;
  F.A = A;
  F.max_levels = max_levels;
  // cluster size: enough CTAs to spread one agent's factor over several SMs,
  // a single CTA for tiny agents (level barriers become __syncthreads)
  const double per_agent = (double)nnzL / std::max(1, A);
  F.cluster = per_agent < 2.0e4 ? 1 : per_agent < 1.0e5 ? 2 : per_agent < 4.0e5 ? 4 : 8;
  return DCORA_OK;
}

template <int R, int DH, int CL>
static int launch_trsv(dcora_problem *p, const double *V, double *T, double *Y,
                       double *OUT, int mask_flag) {
  DeviceMatrices &M = *p->M;
  const DeviceFactor &F = M.F;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(M.A * CL));
  cfg.blockDim = dim3(TRSV_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = p->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  DCORA_CUDA(cudaLaunchKernelEx(&cfg, trsv_block_kernel<R, DH, CL>,
                                (const int *)p->flags, mask_flag, F.dir[0], F.dir[1],
                                (const int *)F.perm, V, T, Y, OUT));
  p->launches++;
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
  double *tmp = p->vec(WS_TRSV_TMP);
  double *sol = p->vec(WS_TRSV_SOL);
  if (!panel || !sol || !tmp) return DCORA_ERR_ALLOC;
  DCORA_DISPATCH_RD(p->r, M.d, {
    switch (F.cluster) {
      case 1: DCORA_TRY((launch_trsv<R, DH, 1>(p, V, tmp, panel, sol, mask_flag))); break;
      case 2: DCORA_TRY((launch_trsv<R, DH, 2>(p, V, tmp, panel, sol, mask_flag))); break;
      case 4: DCORA_TRY((launch_trsv<R, DH, 4>(p, V, tmp, panel, sol, mask_flag))); break;
      default: DCORA_TRY((launch_trsv<R, DH, 8>(p, V, tmp, panel, sol, mask_flag))); break;
    }
  })
  return launch_project_tangent(p, X, sol, OUT, mask_flag);
}

}  // namespace dcora
