// K3/K4 as ONE launch: the whole Riemannian trust-region solve (RTR outer loop, truncated
// CG, preconditioner, retraction, cost / gradient evaluations) of every selected agent, one
// thread-block CLUSTER per agent.
//
// Replaces QuadraticOptimizer::optimize -> ROPTLIB RTRNewton::Run / tCG_TR
// (src/QuadraticOptimizer.cpp:28-108,234-280) for batches of agents -- the RBCD colour
// step -- where the per-operation launches of solver.cu are latency bound: a solve is
// ~10^2 dependent short kernels plus a host read-back per tCG iteration.  Here an agent's
// cluster walks the same algorithm with the same device functions (device_ops.cuh,
// trsv_dev.cuh):
//   * vector operations run chunk by chunk (64 poses / 512 scalar columns), a chunk is
//     owned by one 256-thread group, so chains of chunk-local operations (SpMM rows ->
//     Riemannian Hessian correction -> dot product) need no cluster barrier;
//   * dot products: per-chunk partial sums in global memory, a cluster barrier, then
//     every CTA adds the agent's partials in the same fixed order -- all CTAs hold
//     identical trust-region scalars in shared memory and branch identically;
//   * K5 (trsv_run) is inlined: its tape starts streaming while the right-hand side is
//     still being permuted, and the tangent projection + <z, r> read its solution panel
//     directly.
// The host launches once and reads the per-agent results once.
//
// Measured (config 5, 3906-pose agents, 4 CTAs x 16 warps per agent): 5.8 ms per colour
// step against 3.8 ms for the per-operation path -- at 64 warps per agent the chunk passes
// are latency bound and the whole-GPU kernels hide that latency better; for agents that fit
// one CTA it is the other way round (the small-problem tests run 25 % faster), so the kernel
// is built for single-CTA agents only.
//
// This header holds the kernel; rtr_cluster_inst.cu instantiates it for ONE relaxation rank per
// translation unit (-DDCORA_CLUSTER_R=r: the ranks compile in parallel), rtr_cluster.cu is the
// host side.
#pragma once
#include <mutex>

#include "device_ops.cuh"
#include "solver.cuh"
#include "trsv_dev.cuh"

namespace dcora {

struct ClusterArgs {
  // data matrix (BSR over pose blocks + scalar remainder)
  const int *browptr, *bcol, *prowcol;
  const double *bval;
  const int *rem_row, *rem_rowptr, *rem_col;
  const uint8_t *rem_assign;
  const double *rem_val;
  const Chunk *chunks;
  const int *agent_chunk0, *chunk_pose0;
  const int2 *chunk_rem;
  // Cholesky factor tapes
  int has_factor;
  const unsigned long long *tape;
  const long long *tape_off;
  const int *tape_nrec;
  const long long *arow0;
  const int *frow_col, *col_frow;
  // batch state
  const int *agent_list;
  int *flags;
  double *scal, *partials;
  double *x1, *x2, *EG, *gf, *S, *eta, *delta, *Hd, *r, *z, *t0, *xq, *Heta, *T, *Xp;
  const double *G;
  // parameters (ROptParameters, DCORA_types.h:152-201)
  int max_outer, max_inner, single_mode, retraction;
  double tol, delta0, max_delta;
  unsigned long long time_bound_ns;
};

constexpr int CLUSTER_MIN_RANK = 2, CLUSTER_MAX_RANK = 8;   // ranks the kernel is built for

namespace {

struct Ctl {
  double f1, f2, ngf, delta, maxdelta, zr, dpd, epd, epe, alpha, beta, nr0, ceta, cr, finit,
      gninit, rho;
  int outer_on, tcg_on, status, ninner, nouter, accept, end;
};

constexpr int GT = 256;       // threads of a chunk group
constexpr int MAX_DOTS = 4;

template <int R, int DH, int CL, int NW>
__global__ void __launch_bounds__(NW * 32)
rtr_cluster_kernel(const ClusterArgs A) {
  constexpr int NT = NW * 32;
  constexpr int NG = NT / GT;           // chunk groups per CTA
  constexpr int RECW = TapeRec<DH>::W;
  constexpr int PPW = Map<DH>::PPW;
  static_assert(NT % GT == 0, "CTA = whole chunk groups");
  const int slot = blockIdx.x / CL;
  const int agent = A.agent_list ? A.agent_list[slot] : slot;
  const int crank = blockIdx.x - slot * CL;
  const int tid = threadIdx.x;
  const int warp = uniform32(tid >> 5), lane = tid & 31;
  const int gi = tid / GT, gt = tid - gi * GT;       // chunk group of this thread
  const int gid = crank * NG + gi;                   // ... within the cluster
  constexpr int NGC = CL * NG;
  const int c0 = A.agent_chunk0[agent], c1 = A.agent_chunk0[agent + 1];
  const int npass = (c1 - c0 + NGC - 1) / NGC;

  extern __shared__ __align__(128) unsigned long long smem_raw[];
  unsigned long long *ring = smem_raw + (size_t)warp * RING_SLOTS * RECW;
  unsigned long long *mbar = smem_raw + (size_t)NW * RING_SLOTS * RECW + warp * RING_SLOTS;
  double *red = reinterpret_cast<double *>(smem_raw + (size_t)NW * RING_SLOTS * (RECW + 1));
  double *bc = red + MAX_DOTS * NW;                  // broadcast totals
  __shared__ Ctl st;

  TapeRing<RECW> rd;
  rd.setup(ring, mbar);
  const size_t wi = (size_t)agent * (CL * NW) + crank * NW + warp;
  const unsigned long long *my_tape = nullptr;
  int my_nrec = 0;
  if (A.has_factor) {
    my_tape = A.tape + uniform64((unsigned long long)A.tape_off[wi]);
    my_nrec = uniform32(A.tape_nrec[wi]);
  }
  unsigned long long *no_trace = nullptr;

  auto csync = [&]() { phase_sync<CL>(); };

  // ---- per-chunk reductions: `n` thread-local values -> partial slots 0..n-1 of chunk c
  auto chunk_reduce = [&](int c, bool valid, int n, const double (&v)[MAX_DOTS]) {
#pragma unroll
    for (int s = 0; s < MAX_DOTS; ++s)
      if (s < n) {
        const double w = warp_sum(v[s]);
        if (lane == 0) red[s * NW + warp] = w;
      }
    __syncthreads();
    if (gt == 0 && valid) {
#pragma unroll
      for (int s = 0; s < MAX_DOTS; ++s)
        if (s < n) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < GT / 32; ++w) t += red[s * NW + gi * (GT / 32) + w];
          A.partials[(size_t)c * NPART + s] = t;
        }
    }
    __syncthreads();
  };
  // ---- the agent's totals of partial slots 0..n-1 (after a cluster barrier): identical
  // in every CTA (same data, same order)
  auto agent_totals = [&](int n, double (&out)[MAX_DOTS]) {
    double v[MAX_DOTS];
#pragma unroll
    for (int s = 0; s < MAX_DOTS; ++s) v[s] = 0.0;
    for (int c = c0 + tid; c < c1; c += NT)
#pragma unroll
      for (int s = 0; s < MAX_DOTS; ++s)
        if (s < n) v[s] += A.partials[(size_t)c * NPART + s];
#pragma unroll
    for (int s = 0; s < MAX_DOTS; ++s)
      if (s < n) {
        const double w = warp_sum(v[s]);
        if (lane == 0) red[s * NW + warp] = w;
      }
    __syncthreads();
    if (tid < MAX_DOTS && tid < n) {
      double t = 0.0;
      for (int w = 0; w < NW; ++w) t += red[tid * NW + w];
      bc[tid] = t;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < MAX_DOTS; ++s) out[s] = s < n ? bc[s] : 0.0;
    __syncthreads();
  };

  // ---- Y = V * Q on the columns of chunk c (group-local; V complete everywhere)
  auto spmm_chunk = [&](int c, const Chunk &ch, bool valid, const double *V, double *Y) {
    if (valid && ch.kind == KIND_POSE) {
      const int i0 = A.chunk_pose0[c], nposes = ch.ncols / DH;
      const int wg = gt >> 5, sub = lane / DH, a = lane - sub * DH;
      for (int base = wg * PPW; base < nposes; base += (GT / 32) * PPW) {
        const int pi = base + sub;
        if (sub < PPW && pi < nposes)
          bsr_row<R, DH, false>(i0 + pi, a, A.browptr, A.bcol, A.bval, A.prowcol, V, nullptr, Y);
      }
    }
    __syncthreads();   // remainder rows of pose columns accumulate onto the BSR result
    if (valid) {
      const int2 rr = A.chunk_rem[c];
      for (int s = rr.x + gt; s < rr.y; s += GT)
        rem_row_apply<R, false>(s, A.rem_row, A.rem_rowptr, A.rem_col, A.rem_val, A.rem_assign,
                                V, nullptr, Y);
    }
    __syncthreads();
  };

  // ---- K5 + tangent projection: zout = proj_x1((Q + reg I)^-1 v), partial 0 = <zout, dotw>
  // (v complete everywhere on entry; the caller adds the cluster barrier before the total)
  auto precondition = [&](const double *v, double *zout, const double *dotw) {
    if (A.has_factor) {
      rd.start(my_tape, my_nrec);                      // the tape streams from here on
      const long long f0 = A.arow0[agent], f1 = A.arow0[agent + 1];
      trsv_permute_rows<R, false>(f0, f1, crank * NT + tid, CL * NT, A.frow_col, v, A.T);
      csync();
      trsv_run<R, DH, CL>(rd, A.T, A.Xp, (int)f0, no_trace);
      rd.drain();
    }
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      const bool valid = c < c1;
      const Chunk ch = A.chunks[valid ? c : c0];
      double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
      if (valid)
        d[0] = manifold_chunk<R, DH, OP_PROJECT, GT>(
            ch, gt, A.x1, A.has_factor ? A.Xp : v, nullptr, nullptr, zout,
            const_cast<double *>(dotw), true, A.has_factor ? A.col_frow : nullptr);
      chunk_reduce(c, valid, 1, d);
    }
  };

  // ---- cost terms (+ gradient stage) at x: xq = x Q; partials 0: <xq, x>, 1: <x, G>,
  // with_grad: EG = xq + G, gf = proj_x(EG), S = sym(x^T EG), partial 2: <gf, gf>
  auto cost_terms = [&](const double *x, bool with_grad) {
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      const bool valid = c < c1;
      const Chunk ch = A.chunks[valid ? c : c0];
      spmm_chunk(c, ch, valid, x, A.xq);
      double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
      if (valid) {
        const size_t o = (size_t)ch.col0 * R;
        const int len = ch.ncols * R;
        for (int i = gt; i < len; i += GT) {
          const double xv = x[o + i];
          d[0] = fma(A.xq[o + i], xv, d[0]);
          if (A.G) d[1] = fma(xv, A.G[o + i], d[1]);
        }
        if (with_grad)
          d[2] = manifold_chunk<R, DH, OP_GRAD, GT>(ch, gt, x, A.xq, A.G, A.S, A.gf, A.EG, true);
      }
      chunk_reduce(c, valid, with_grad ? 3 : 2, d);
    }
  };

  // =====================================================================
  if (tid == 0) {
    st.delta = A.delta0;
    st.maxdelta = A.max_delta;
    st.outer_on = 0; st.tcg_on = 0; st.status = DCORA_TCG_MAXITER;
    st.ninner = 0; st.nouter = 0; st.accept = 0; st.end = 0;
    st.f2 = 0.0; st.rho = 0.0;
  }
  const unsigned long long t_start = globaltimer_ns();
  double tot[MAX_DOTS];
  // initial statistics: fInit, gradNormInit (QuadraticOptimizer.cpp:30-31)
  cost_terms(A.x1, true);
  csync();
  agent_totals(3, tot);
  if (tid == 0) {
    st.f1 = 0.5 * tot[0] + tot[1];
    st.ngf = sqrt(tot[2]);
    st.finit = st.f1;
    st.gninit = st.ngf;
    // QuadraticOptimizer::trustRegion early return (:54)
    st.outer_on = !(st.ngf < A.tol);
  }
  __syncthreads();

  for (int it = 0; it < A.max_outer && st.outer_on; ++it) {
    // ---- tCG (SolversTR::tCG_TR), eta0 = 0: r = gf, eta = 0, Heta = 0
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      if (c < c1) {
        const Chunk ch = A.chunks[c];
        const size_t o = (size_t)ch.col0 * R;
        const int len = ch.ncols * R;
        for (int i = gt; i < len; i += GT) {
          A.r[o + i] = A.gf[o + i];
          A.eta[o + i] = 0.0;
          A.Heta[o + i] = 0.0;
        }
      }
    }
    csync();                                   // r complete
    precondition(A.r, A.z, A.r);               // z, partial 0 = <z, r>
    csync();
    agent_totals(1, tot);
    if (tid == 0) {
      st.zr = tot[0]; st.dpd = tot[0]; st.epe = 0.0; st.epd = 0.0;
      st.nr0 = st.ngf;
      st.tcg_on = 1; st.end = 0;
      st.status = DCORA_TCG_MAXITER;
    }
    // delta = -z
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      if (c < c1) {
        const Chunk ch = A.chunks[c];
        const size_t o = (size_t)ch.col0 * R;
        const int len = ch.ncols * R;
        for (int i = gt; i < len; i += GT) A.delta[o + i] = -A.z[o + i];
      }
    }
    __syncthreads();
    for (int j = 0; j < A.max_inner && st.tcg_on; ++j) {
      csync();                                 // delta complete
      // Hd = proj(delta Q - delta sym(Y^T EG)), partial 0 = <delta, Hd>
      for (int k = 0; k < npass; ++k) {
        const int c = c0 + gid + k * NGC;
        const bool valid = c < c1;
        const Chunk ch = A.chunks[valid ? c : c0];
        spmm_chunk(c, ch, valid, A.delta, A.t0);
        double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
        if (valid)
          d[0] = manifold_chunk<R, DH, OP_RHESS, GT>(ch, gt, A.x1, A.delta, A.t0, A.S, A.Hd,
                                                     nullptr, true);
        chunk_reduce(c, valid, 1, d);
      }
      csync();
      agent_totals(1, tot);
      if (tid == 0) {
        const double d_Hd = tot[0];
        const double z_r = st.zr, d_Pd = st.dpd, e_Pd = st.epd, e_Pe = st.epe;
        const double Delta = st.delta;
        const double alpha = z_r / d_Hd;
        const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
        if (d_Hd <= 0.0 || e_Pe_new >= Delta * Delta) {
          const double tau =
              (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
          st.ceta = tau;
          st.cr = 0.0;
          st.status = d_Hd <= 0.0 ? DCORA_TCG_NEGCURVTURE : DCORA_TCG_EXCREGION;
          st.end = 1;
        } else {
          st.epe = e_Pe_new;
          st.alpha = alpha;
          st.ceta = alpha;
          st.cr = alpha;
          st.end = 0;
        }
      }
      __syncthreads();
      // eta += c delta, Heta += c Hd, r += c_r Hd, partial 0 = <r, r>
      {
        const double ce = st.ceta, cr = st.cr;
        for (int k = 0; k < npass; ++k) {
          const int c = c0 + gid + k * NGC;
          const bool valid = c < c1;
          double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
          if (valid) {
            const Chunk ch = A.chunks[c];
            const size_t o = (size_t)ch.col0 * R;
            const int len = ch.ncols * R;
            for (int i = gt; i < len; i += GT) {
              const double h = A.Hd[o + i];
              A.eta[o + i] = fma(1.0, A.eta[o + i], ce * A.delta[o + i]);
              A.Heta[o + i] = fma(1.0, A.Heta[o + i], ce * h);
              const double rn = fma(1.0, A.r[o + i], cr * h);
              A.r[o + i] = rn;
              d[0] = fma(rn, rn, d[0]);
            }
          }
          chunk_reduce(c, valid, 1, d);
        }
      }
      csync();                                 // r complete, partials visible
      agent_totals(1, tot);
      if (tid == 0) {
        st.ninner += 1;
        int on = 1;
        if (st.end) {
          on = 0;
        } else {
          const double nr = sqrt(tot[0]), nr0 = st.nr0;
          const double theta = 1.0, kappa = 0.1;
          const double pw = pow(nr0, theta);
          if (nr <= nr0 * (pw < kappa ? pw : kappa)) {
            st.status = kappa < pw ? DCORA_TCG_LCON : DCORA_TCG_SCON;
            on = 0;
          } else if (j + 1 >= A.max_inner) {
            on = 0;  // status stays TR_MAXITER
          }
        }
        st.tcg_on = on;
      }
      __syncthreads();
      if (!st.tcg_on) break;
      precondition(A.r, A.z, A.r);             // z, partial 0 = <z, r>
      csync();
      agent_totals(1, tot);
      if (tid == 0) {
        const double zr = tot[0];
        const double beta = zr / st.zr;
        st.zr = zr;
        st.beta = beta;
        st.epd = beta * (st.epd + st.alpha * st.dpd);
        st.dpd = zr + beta * beta * st.dpd;
      }
      __syncthreads();
      {
        const double beta = st.beta;           // delta = -z + beta delta
        for (int k = 0; k < npass; ++k) {
          const int c = c0 + gid + k * NGC;
          if (c < c1) {
            const Chunk ch = A.chunks[c];
            const size_t o = (size_t)ch.col0 * R;
            const int len = ch.ncols * R;
            for (int i = gt; i < len; i += GT) A.delta[o + i] = fma(-1.0, A.z[o + i], beta * A.delta[o + i]);
          }
        }
      }
      __syncthreads();
    }
    // ---- candidate, model decrease, rho (SolversTR::Run)
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      if (c < c1) {
        const Chunk ch = A.chunks[c];
        if (A.retraction == DCORA_RETRACT_POLAR)
          manifold_chunk<R, DH, OP_RETRACT_POLAR, GT>(ch, gt, A.x1, A.eta, nullptr, nullptr, A.x2,
                                                      nullptr, false);
        else
          manifold_chunk<R, DH, OP_RETRACT_QF, GT>(ch, gt, A.x1, A.eta, nullptr, nullptr, A.x2,
                                                   nullptr, false);
      }
    }
    csync();                                   // x2 complete
    // xq = x2 Q, partials 0: <xq, x2>, 1: <x2, G>, 2: <eta, gf>, 3: <eta, H eta>
    for (int k = 0; k < npass; ++k) {
      const int c = c0 + gid + k * NGC;
      const bool valid = c < c1;
      const Chunk ch = A.chunks[valid ? c : c0];
      spmm_chunk(c, ch, valid, A.x2, A.xq);
      double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
      if (valid) {
        const size_t o = (size_t)ch.col0 * R;
        const int len = ch.ncols * R;
        for (int i = gt; i < len; i += GT) {
          const double xv = A.x2[o + i], ev = A.eta[o + i];
          d[0] = fma(A.xq[o + i], xv, d[0]);
          if (A.G) d[1] = fma(xv, A.G[o + i], d[1]);
          d[2] = fma(ev, A.gf[o + i], d[2]);
          d[3] = fma(ev, A.Heta[o + i], d[3]);
        }
      }
      chunk_reduce(c, valid, 4, d);
    }
    csync();
    agent_totals(4, tot);
    if (tid == 0) {
      st.f2 = 0.5 * tot[0] + tot[1];
      const double f1 = st.f1, f2 = st.f2;
      const double rho = (f1 - f2) / (-(tot[2] + 0.5 * tot[3]));
      st.rho = rho;
      double Delta = st.delta;
      const int stt = st.status;
      if (rho > 0.75) {
        if (stt == DCORA_TCG_EXCREGION || stt == DCORA_TCG_NEGCURVTURE)
          Delta = fmin(2.0 * Delta, st.maxdelta);
      } else if (rho < 0.25) {
        Delta *= 0.25;
      }
      const int acc = (rho > 0.1) ||
                      (fabs(f1 - f2) / (fabs(f1) + 1.0) < 1.4901161193847656e-08 && f2 < f1);
      if (A.single_mode) {
        // QuadraticOptimizer.cpp:254-273: radius /= 4 after a rejection and
        // maximum_Delta = initial_Delta = radius for the next Run()
        if (!acc) { Delta = st.maxdelta * 0.25; st.maxdelta = Delta; }
      }
      st.delta = Delta;
      st.accept = acc;
      st.nouter += 1;
    }
    __syncthreads();
    if (st.accept) {
      // x1 = x2; EG, gf, S at the new point from xq = x2 Q; partial 0 = <gf, gf>
      for (int k = 0; k < npass; ++k) {
        const int c = c0 + gid + k * NGC;
        const bool valid = c < c1;
        const Chunk ch = A.chunks[valid ? c : c0];
        double d[MAX_DOTS] = {0.0, 0.0, 0.0, 0.0};
        if (valid) {
          const size_t o = (size_t)ch.col0 * R;
          const int len = ch.ncols * R;
          for (int i = gt; i < len; i += GT) A.x1[o + i] = A.x2[o + i];
        }
        __syncthreads();
        if (valid)
          d[0] = manifold_chunk<R, DH, OP_GRAD, GT>(ch, gt, A.x1, A.xq, A.G, A.S, A.gf, A.EG, true);
        chunk_reduce(c, valid, 1, d);
      }
    }
    // the time bound is decided by one thread and published before the barrier, so that
    // every CTA of the cluster takes the same branch
    if (crank == 0 && tid == 0)
      A.flags[agent * NF + F_END] =
          (A.time_bound_ns != 0ull && globaltimer_ns() - t_start > A.time_bound_ns) ? 1 : 0;
    csync();
    agent_totals(1, tot);
    if (tid == 0) {
      int on = 1;
      if (st.accept) {
        st.f1 = st.f2;
        st.ngf = sqrt(tot[0]);
        if (st.ngf < A.tol) on = 0;
        if (A.single_mode) on = 0;
      }
      if (st.nouter >= A.max_outer) on = 0;
      if (A.flags[agent * NF + F_END]) on = 0;
      st.outer_on = on;
    }
    __syncthreads();
  }
  // ---- results
  if (crank == 0 && tid == 0) {
    double *s = A.scal + agent * NS;
    int *f = A.flags + agent * NF;
    s[S_F1] = st.f1; s[S_F2] = st.f2; s[S_NGF] = st.ngf; s[S_DELTA] = st.delta;
    s[S_MAXDELTA] = st.maxdelta; s[S_RHO] = st.rho; s[S_FINIT] = st.finit;
    s[S_GNINIT] = st.gninit; s[S_TOL] = A.tol;
    f[F_OUTER] = 0; f[F_TCG] = 0; f[F_STEP] = 0; f[F_END] = 0;
    f[F_STATUS] = st.status; f[F_NINNER] = st.ninner; f[F_NOUTER] = st.nouter;
    f[F_ACCEPT] = st.accept;
  }
  csync();   // no CTA of the cluster exits while another may still address it
}

template <int R, int DH, int CL, int NW>
cudaError_t cluster_set_attrs(int device, int smem_bytes) {
  static std::mutex mtx;
  static unsigned long long done = 0;  // bit d: set on device d
  std::lock_guard<std::mutex> lock(mtx);
  if (device < 64 && (done & (1ull << device))) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(rtr_cluster_kernel<R, DH, CL, NW>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess && device < 64) done |= 1ull << device;
  return e;
}

template <int R, int DH, int CL, int NW>
int launch_cluster(dcora_problem *p, const ClusterArgs &args, int n_agents) {
  constexpr int smem_bytes =
      NW * RING_SLOTS * (TapeRec<DH>::W + 1) * 8 + (MAX_DOTS * NW + MAX_DOTS) * 8;
  DCORA_CUDA((cluster_set_attrs<R, DH, CL, NW>(p->M->device, smem_bytes)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_agents * CL));
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = p->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  DCORA_CUDA(cudaLaunchKernelEx(&cfg, rtr_cluster_kernel<R, DH, CL, NW>, args));
  p->launches++;
  return DCORA_OK;
}

// Built for agents small enough for ONE CTA (8 or 16 warps, trsv_warps_for): there the
// per-operation launches of solver.cu are purely latency bound and one launch per solve
// wins.  Larger agents (>= 1500 columns) need more warps than a cluster of 4 SMs hides
// latency with and stay on the whole-GPU kernels.
template <int R, int DH>
int launch_cluster_shape(dcora_problem *p, const ClusterArgs &args, int n_agents, int nw) {
  if (nw <= 8) return launch_cluster<R, DH, 1, 8>(p, args, n_agents);
  return launch_cluster<R, DH, 1, 16>(p, args, n_agents);
}

}  // namespace
}  // namespace dcora
