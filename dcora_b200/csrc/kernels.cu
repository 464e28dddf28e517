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

#include "device_ops.cuh"
#include "solver.cuh"

namespace dcora {

namespace {

// ---------------------------------------------------------------- K1 SpMM
// Y[:, block i] = sum_j X[:, block j] * Q[j-block, i-block]  (+ G)
// (Round-1 variant study -- wide 128/256-bit loads, register and cp.async
// software pipelines, streaming several rows per lane group -- is recorded in
// profiles/experiments/r01_k1_variants.md: every variant lost to this plain
// gather loop, which keeps 40 warps per SM resident at 48 registers.)
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
  bsr_row<R, DH, true>(i, a, browptr, bcol, bval, prowcol, X, G, Y);
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
  rem_row_apply<R, true>(s, rem_row, rem_rowptr, rem_col, rem_val, rem_assign, X, G, Y);
}

// ------------------------------------------------------------ K2 manifold
template <int R, int DH, int OP>
__global__ void __launch_bounds__(128)
manifold_kernel(const Chunk *__restrict__ chunks, const int *__restrict__ flags,
                int mask_flag, const int *__restrict__ list, const double *__restrict__ X,
                const double *__restrict__ V, const double *__restrict__ HV,
                const double *S, double *__restrict__ OUT,
                double *__restrict__ AUX, double *__restrict__ partials) {
  // `partials` (optional, OP_RHESS / OP_GRAD / OP_PROJECT): the chunk's fused dot product
  // goes to partial slot 0 -- the fused form of a following dots_kernel launch.
  __shared__ double red_sm[4];
  const int cid = list ? list[blockIdx.x] : blockIdx.x;
  const Chunk ch = chunks[cid];
  if (!agent_on(flags, ch.agent, mask_flag)) return;
  const double dsum = manifold_chunk<R, DH, OP, 128>(ch, threadIdx.x, X, V, HV, S, OUT, AUX,
                                                      partials != nullptr);
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
